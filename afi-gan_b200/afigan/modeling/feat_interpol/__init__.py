from .generator_rdb import Generator  # noqa: F401
from .feature_patch_discriminator import Discriminator  # noqa: F401
