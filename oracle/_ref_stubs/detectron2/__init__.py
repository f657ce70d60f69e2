"""Minimal stand-in for the detectron2 names the reference hot-path modules import.

Test infrastructure only: detectron2 and fvcore are not installed in this
image, and the two reference files generator_rdb.py / feature_patch_discriminator.py import nothing
else from them. Semantics follow upstream detectron2: Conv2d = nn.Conv2d + optional norm + optional
activation, bias defaults to True even when a norm is given.
"""
