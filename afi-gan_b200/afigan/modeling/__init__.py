from .feat_interpol import Generator, Discriminator  # noqa: F401
