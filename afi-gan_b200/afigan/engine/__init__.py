from .stage1 import FeaturePrefetcher, Stage1Step  # noqa: F401
from .stage2 import Stage2Step, nearest_half, stage2_discriminator_losses, stage2_generator_losses  # noqa: F401
from .checkpoint import (AFCheckpointer, align_by_suffix, convert_afi_names, load_extractor_into_detector, load_generator_into_extractor,  # noqa: F401
                         remain_only_afi_names, strip_module_prefix)
from .lr import warmup_multistep_lr  # noqa: F401
from .dataset_mapper import PairedScaleMapper, resize_image, shortest_edge_size  # noqa: F401
