"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/afigan_b200.h
declares (no compute calls without a GPU), and the Python modules expose the reference's state-dict contract."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "afigan_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(afi_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from afigan import native
    names = _header_functions()
    assert len(names) >= 25
    lib = ctypes.CDLL(native.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/afigan_b200.h but not exported"
    assert set(native.EXPORTED_SYMBOLS) == set(names), set(native.EXPORTED_SYMBOLS) ^ set(names)
    assert native.lib().afi_abi_version() == 4


def test_ctypes_structs_match_the_library():
    """The ctypes mirror of every boundary struct has the size the library was compiled with (a field added on one side only shows here,
    on the CPU, instead of as memory corruption on the GPU)."""
    from afigan import native
    lib = native.lib()
    mirrors = [native.View4, native.GParams, native.Lateral, native.GCall, native.DParams, native.DCall, native.GGrads, native.DGrads]
    for which, cls in enumerate(mirrors):
        assert lib.afi_sizeof(which) == ctypes.sizeof(cls), (which, cls.__name__, lib.afi_sizeof(which), ctypes.sizeof(cls))
    assert lib.afi_sizeof(99) == 0


def test_view4_carries_strides_and_the_dtype_tag():
    """ABI 4: a boundary view is (pointer, four element strides, dtype tag); fp32 and bf16 tensors of any layout map onto it without a copy,
    anything else is refused by the binding (the callers cast such tensors to fp32 once: native.boundary)."""
    import torch
    from afigan import native
    x = torch.zeros(2, 8, 3, 5)
    v = native.view4(x)
    assert (v.ptr, v.sn, v.sc, v.sh, v.sw, v.dtype, v.reserved) == (x.data_ptr(), 120, 15, 5, 1, native.DT_F32, 0)
    xc = x.bfloat16().contiguous(memory_format=torch.channels_last)[:, :, :2, :4]          # bf16, NHWC strides, top-left crop
    v = native.view4(xc)
    assert (v.ptr, v.sn, v.sc, v.sh, v.sw, v.dtype) == (xc.data_ptr(), 120, 1, 40, 8, native.DT_BF16)
    for bad in (x.half(), x.double(), x[0]):
        try:
            native.view4(bad)
        except TypeError:
            continue
        raise AssertionError("view4 accepted an unsupported tensor")
    assert native.boundary(x) is x and native.boundary(xc) is xc and native.boundary(x.half()).dtype == torch.float32
    assert ctypes.sizeof(native.View4) == 48


def test_integration_snippet():
    """the ctypes stub printed in INTEGRATION.md must be the struct layout the library was built with -- the block between the
    binding-begin / binding-end markers is executed as is and its own check_binding() is run against the built library."""
    from afigan import native
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"# --- binding-begin\n(.*?)# --- binding-end", doc, flags=re.S)
    assert m, "INTEGRATION.md lost its binding block"
    ns = {}
    exec(m.group(1), ns)
    ns["check_binding"](ctypes.CDLL(native.LIB_PATH))
    for name in ("View4", "GParams", "Lateral", "GCall", "DParams", "DCall", "DGrads"):
        assert ctypes.sizeof(ns[name]) == ctypes.sizeof(getattr(native, name)), name
        assert [f[0] for f in ns[name]._fields_] == [f[0] for f in getattr(native, name)._fields_], name


def test_size_queries_without_gpu():
    from afigan import native
    lib = native.lib()
    assert lib.afi_g_gradacc_bytes(3) >= 7_834_624 * 4
    assert lib.afi_d_gradacc_bytes() >= 15_352_321 * 4
    assert lib.afi_g_packed_bytes(native.PREC_BF16, 3) * 2 == lib.afi_g_packed_bytes(native.PREC_FP32, 3)
    small = lib.afi_g_workspace_bytes(native.PREC_BF16, 2, 7, 11, 3, 0, 0)
    big = lib.afi_g_workspace_bytes(native.PREC_BF16, 2, 7, 11, 3, 0, 1)
    assert 0 < small < big
    assert lib.afi_d_workspace_bytes(native.PREC_FP32, 2, 13, 21, 1) > lib.afi_d_workspace_bytes(native.PREC_BF16, 2, 13, 21, 1)


def test_state_dict_contract_matches_reference_layout():
    from afigan.modeling import Discriminator, Generator
    from oracle import afigan_oracle as O
    torch.manual_seed(0)
    G, D = Generator(n_residual_dense_blocks=3), Discriminator()
    g_sd, d_sd = O.init_states(0)
    assert list(G.state_dict()) == list(g_sd) and list(D.state_dict()) == list(d_sd)
    for k, v in g_sd.items():
        assert torch.equal(G.state_dict()[k], v), k      # same seed => same weights as the reference (bit exact)
    for k, v in d_sd.items():
        assert torch.equal(D.state_dict()[k], v), k
    # optimisers are built over Generators[0] / Discriminators[0] only (stage1_trainer.py:109-114)
    assert sum(p.numel() for p in G.Generators[0].parameters()) == 7_834_624
    assert sum(p.numel() for p in D.Discriminators[0].parameters()) == 15_352_321
    assert D.current_step == 0
    G.load_state_dict(g_sd)
    D.load_state_dict(d_sd)


def test_no_cpu_fallback():
    from afigan.modeling import Discriminator, Generator
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Generator(n_residual_dense_blocks=3)(torch.zeros(1, 256, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Discriminator().Discriminators[0](torch.zeros(1, 256, 4, 4))
    with pytest.raises(ValueError):
        Generator(in_channels=128)
