from .config import get_cfg  # noqa: F401
from .defaults import add_afigan_config  # noqa: F401
