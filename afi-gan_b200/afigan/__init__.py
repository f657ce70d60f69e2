"""B200-native drop-in for the hot path of AFI-GAN (afigan.modeling.feat_interpol + the stage-1/2 loss blocks).

Mirrors the reference package layout (afigan.config, afigan.modeling, afigan.engine) for the parts on the
hot path; everything numerical runs in libafigan_b200.so (hand-written sm_100a CUDA) through afigan.native.
"""
__version__ = "0.1.0"
