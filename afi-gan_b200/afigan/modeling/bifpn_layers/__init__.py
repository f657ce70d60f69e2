"""Layer wrappers of the BiFPN neck: drop-in for reference afigan/modeling/bifpn_layers/{wrappers,activations}.py (same class names, constructor
arguments, parameter names and arithmetic).  The pointwise (1x1) half of `SeparableConv2d` and every 1x1 `Conv2d` run through the library's
tcgen05 GEMM engine on CUDA (`functional.conv1x1_autograd`); the depthwise 3x3 conv, max-pool and normalisation are HBM-bound library ops."""
from .wrappers import Conv2d, MaxPool2d, MemoryEfficientSwish, SeparableConv2d, Swish  # noqa: F401
