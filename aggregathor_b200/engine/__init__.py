"""Step builder / training engine (the role of the reference's `graph.py`)."""

from .schedules import learning_rates, build, LearningRate            # noqa: F401
from .optimizers import optimizers, OptimizerSpec                      # noqa: F401
from .flat import FlatLayout, regularization                           # noqa: F401
