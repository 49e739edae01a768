"""Functional kernels: aggregation (`gar`) and neural-network (`nn`) ops, each with a native
sm_100a implementation and a library/torch reference."""
