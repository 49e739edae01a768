"""Model zoo: static layer graphs built from `models.core` modules."""

from .core import Context, Model  # noqa: F401
from .nets_factory import networks_map, get_network, default_image_size  # noqa: F401
