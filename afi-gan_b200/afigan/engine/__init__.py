from .stage1 import Stage1Step  # noqa: F401
