"""Framework-wide defaults.

Mirrors the tunables of the reference (`config.py:38-66`): job names, training
defaults, evaluation/checkpoint/summary cadence, service-thread idle delay.
B200-specific knobs (bucket sizes, SM carve-outs, symmetric-memory sizes) live
here too so that there is a single place to look.
"""

# --- cluster ----------------------------------------------------------------
default_ps_job_name = "ps"
default_wk_job_name = "workers"
default_ev_job_name = "eval"

# --- training ---------------------------------------------------------------
default_max_step = 10000
default_learning_rate = 1e-3
default_end_learning_rate = 1e-4
default_decay_step = 10000
default_decay_rate = 0.96

# --- evaluation / checkpoint / summary --------------------------------------
default_evaluation_file_name = "eval"
default_evaluation_delta = -1
default_evaluation_period = 10.0
default_checkpoint_base_name = "model"
default_checkpoint_delta = -1
default_checkpoint_period = 120.0
default_summary_delta = -1
default_summary_period = 30.0
checkpoint_keep = 5  # tf.train.Saver default retention in the reference

# Delay (s) of the polling loop of the eval/checkpoint/summary services.
thread_idle_delay = 1.0

# --- B200 -------------------------------------------------------------------
cuda_arch_flags = ("-gencode", "arch=compute_100a,code=sm_100a")
sm_count = 148
# CTAs used by the persistent fused aggregation kernel (one per SM).
gar_ctas = 148
gar_threads = 512
# Chunk granularity of the lossy-transport emulation (bytes): the reference's
# UDP path ships 65 000-byte datagrams (mpi_rendezvous_mgr.patch:600-620).
udp_chunk_bytes = 65000
