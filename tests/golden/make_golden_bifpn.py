"""Golden fixture for the BiFPN neck: forward of the UNMODIFIED reference `BiFPN_AFIGAN` (afigan/modeling/backbone/bifpn_sr.py:203-731, with its
own bifpn_layers and its own Generator) on the CPU.  Dev container only (needs /root/reference):  python tests/golden/make_golden_bifpn.py

The reference package cannot be imported as a package (its __init__ chain pulls in detectron2's engine, timm, pycocotools), so the modules this
neck needs are loaded file by file under their real names, with oracle/_ref_stubs standing in for detectron2 / fvcore and an empty stand-in for
the Swin bottom-up (only `build_swint_bifpn_sr_backbone` uses it).  Weights are NOT stored: every tensor of the state dict is re-derived from
its key by `keyed_state` (a seeded generator per key), which tests/test_gpu_bifpn.py applies to this repository's BiFPN_AFIGAN as well -- the two
classes must therefore agree on every key and shape.  Stored: key list + shapes, the five output maps (p5..p7 in full, p3 / p4 as norms +
strided samples) for a 1 x 3 x 128 x 128 input in eval mode."""
import importlib.util
import os
import sys
import types
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def keyed_state(state):
    """Deterministic values for every entry of a state dict, derived from the KEY (so two implementations with equal keys get equal weights)."""
    out = {}
    for k, v in state.items():
        g = torch.Generator().manual_seed(zlib.crc32(k.encode()))
        if k.endswith("num_batches_tracked"):
            out[k] = torch.tensor(3, dtype=v.dtype)
        elif k.endswith("running_var"):
            out[k] = torch.rand(v.shape, generator=g) + 0.5
        elif k.endswith("running_mean"):
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif "_w1" in k or "_w2" in k:
            out[k] = 0.3 + 0.5 * torch.rand(v.shape, generator=g)                      # raw attention weights (no ReLU / normalisation, App. D-9)
        elif v.dim() == 1:
            out[k] = (1.0 + 0.2 * torch.randn(v.shape, generator=g)) if ("norm.weight" in k or k.endswith(".1.weight")) else 0.05 * torch.randn(v.shape, generator=g)
        else:
            fan_in = v[0].numel()
            scale = (0.1 if "srf_module" in k else 1.0) * (2.0 / fan_in) ** 0.5          # the interpolator keeps its x0.1 init scale
            out[k] = scale * torch.randn(v.shape, generator=g)
    return out


def tiny_bottom_up(Backbone):
    class BU(Backbone):
        """Three strided 1x1 'stages' (strides 8, 16, 32; 64 / 96 / 128 channels): just enough of a Backbone for the neck."""

        def __init__(self):
            super().__init__()
            self._out_features = ["s2", "s3", "s4"]
            self._out_feature_strides = {"s2": 8, "s3": 16, "s4": 32}
            self._out_feature_channels = {"s2": 64, "s3": 96, "s4": 128}
            self.c = torch.nn.ModuleList([torch.nn.Conv2d(3, c, 1) for c in (64, 96, 128)])

        def forward(self, x):
            return {k: c(torch.nn.functional.avg_pool2d(x, s)) for k, c, s in zip(self._out_features, self.c, (8, 16, 32))}

    return BU()


def sample(t, n=513):
    f = t.detach().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()
    return f[idx].numpy().copy()


def load_reference_bifpn():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref_stubs"))
    for name, path in (("afigan", "afigan"), ("afigan.modeling", "afigan/modeling"), ("afigan.modeling.backbone", "afigan/modeling/backbone")):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, path)]                     # namespace shells: sub-modules resolve to the reference files, __init__ chains do not run
        sys.modules[name] = m
    swin = types.ModuleType("afigan.modeling.backbone.swin_transformer")
    swin.build_swint_backbone = None
    sys.modules[swin.__name__] = swin
    spec = importlib.util.spec_from_file_location("afigan.modeling.backbone.bifpn_sr", os.path.join(REF, "afigan/modeling/backbone/bifpn_sr.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    mod = load_reference_bifpn()
    from detectron2.modeling.backbone import Backbone
    cfg = types.SimpleNamespace(MODEL=types.SimpleNamespace(AFI_FREEZE=False))
    torch.manual_seed(0)
    net = mod.BiFPN_AFIGAN(tiny_bottom_up(Backbone), ["s2", "s3", "s4"], 256, 3, norm="BN", top_block=mod.LastLevelP6P7(128, 256, "BN"), cfg=cfg)
    net.load_state_dict(keyed_state(net.state_dict()), strict=True)
    net.eval()
    x = torch.randn(1, 3, 128, 128, generator=torch.Generator().manual_seed(77))
    with torch.no_grad():
        out = net(x)
    fx = {"keys": np.array(list(net.state_dict().keys())), "shapes": np.array([str(tuple(v.shape)) for v in net.state_dict().values()])}
    for k, v in out.items():
        fx[f"norm/{k}"] = np.array(float(v.norm()))
        fx[f"shape/{k}"] = np.array(v.shape)
        if k in ("p5", "p6", "p7"):
            fx[f"full/{k}"] = v.numpy()
        else:
            fx[f"sample/{k}"] = sample(v)
        print(k, tuple(v.shape), float(v.norm()), float(v.abs().max()))
    path = os.path.join(HERE, "bifpn_small.npz")
    np.savez_compressed(path, **fx)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", len(fx["keys"]), "state-dict entries")


if __name__ == "__main__":
    main()
