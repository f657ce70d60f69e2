"""detectron2.layers.batch_norm stand-in (test infrastructure): the reference's bifpn_layers/wrappers.py imports get_norm from here."""
from . import get_norm  # noqa: F401
