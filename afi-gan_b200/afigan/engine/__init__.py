from .stage1 import Stage1Step  # noqa: F401
from .stage2 import stage2_discriminator_losses, stage2_generator_losses, nearest_half  # noqa: F401
