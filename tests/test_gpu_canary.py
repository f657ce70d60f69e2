"""Guard-band test of the caller-owned workspaces and outputs (compute-sanitizer is closed on the pool).  Every
library call gets EXACTLY the number of workspace bytes its *_workspace_bytes() query reports, followed by a canary band; outputs and
gradient buffers are followed by canaries too.  After forward + backward on ragged shapes, in every operand mode, the canaries must be intact
(a kernel writing past the end of its workspace, output or accumulator shows here) and the results must still match the oracle."""
import ctypes as C

import pytest
import torch

from gpu_util import rel
from oracle import afigan_oracle as O

pytestmark = pytest.mark.gpu
BAND = 1 << 16
PATTERN = 0xA5


def _guarded(nbytes, dev):
    """uint8 buffer of nbytes (rounded up to 256) followed by a canary band; returns (whole buffer, view of the payload)."""
    n = (int(nbytes) + 255) // 256 * 256
    buf = torch.full((n + BAND,), PATTERN, dtype=torch.uint8, device=dev)
    return buf, buf[:n]


def _intact(buf):
    return bool((buf[-BAND:] == PATTERN).all())


@pytest.mark.parametrize("precision", ["split", "bf16", "fp32"])
@pytest.mark.parametrize("shape", [(2, 7, 11, 13, 21), (1, 13, 21, 25, 42), (3, 5, 4, 9, 8), (1, 1, 1, 2, 2)], ids=lambda s: "x".join(map(str, s)))
def test_generator_call_respects_its_buffers(shape, precision):
    from afigan import native as N
    from afigan.functional import g_param_struct
    from afigan.modeling import Generator
    n, h, w, oh, ow = shape
    dev = torch.device("cuda")
    prec = N.PRECISIONS[precision]
    torch.manual_seed(0)
    G = Generator(n_residual_dense_blocks=3, precision=precision).cuda()
    params = G._params()
    lib, ctx = N.lib(), N.context(dev)
    ps = g_param_struct(params, 3)
    packed_all, packed = _guarded(lib.afi_g_packed_bytes(prec, 3), dev)
    N.check(lib.afi_g_pack(ctx, prec, C.byref(ps), packed.data_ptr(), N.stream_ptr()))
    ws_all, ws = _guarded(lib.afi_g_workspace_bytes(prec, n, h, w, 3, 0, 1), dev)
    y_all, y8 = _guarded(n * 256 * oh * ow * 4, dev)
    dx_all, dx8 = _guarded(n * 256 * h * w * 4, dev)
    acc_all, acc = _guarded(lib.afi_g_gradacc_bytes(3), dev)
    acc.zero_()
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(n, 256, h, w, generator=gen).cuda()
    dy = torch.randn(n, 256, oh, ow, generator=gen).cuda()
    call = N.GCall(x=N.view4(x), n=n, h=h, w=w, y=y8.data_ptr(), oh=oh, ow=ow, ws=ws.data_ptr(), ws_bytes=lib.afi_g_workspace_bytes(prec, n, h, w, 3, 0, 1))
    N.check(lib.afi_g_forward(ctx, prec, C.byref(ps), packed.data_ptr(), C.byref(call), 1, 1, N.stream_ptr()))
    call.dy, call.dx = N.view4(dy), dx8.data_ptr()
    N.check(lib.afi_g_backward(ctx, prec, C.byref(ps), packed.data_ptr(), C.byref(call), 1, acc.data_ptr(), N.stream_ptr()))
    torch.cuda.synchronize()
    for name, b in (("packed weights", packed_all), ("workspace", ws_all), ("output", y_all), ("dx", dx_all), ("gradient accumulator", acc_all)):
        assert _intact(b), f"{name}: canary band overwritten"
    y = y8[: n * 256 * oh * ow * 4].view(torch.float32).view(n, 256, oh, ow)
    g_sd, _ = O.init_states(0)
    ref = O.generator_forward(g_sd, x.cpu())[:, :, :oh, :ow]
    assert rel(y, ref) < (1e-5 if precision != "bf16" else 1e-3)


@pytest.mark.parametrize("precision", ["split", "bf16", "fp32"])
@pytest.mark.parametrize("shape", [(2, 13, 21), (1, 25, 42), (3, 3, 5), (1, 1, 1)], ids=lambda s: "x".join(map(str, s)))
def test_discriminator_call_respects_its_buffers(shape, precision):
    from afigan import native as N
    from afigan.functional import d_param_struct
    from afigan.modeling import Discriminator, Generator
    n, h, w = shape
    dev = torch.device("cuda")
    prec = N.PRECISIONS[precision]
    torch.manual_seed(0)
    Generator(n_residual_dense_blocks=3)
    D = Discriminator(precision=precision).cuda()
    stack = D.Discriminators[0]
    lib, ctx = N.lib(), N.context(dev)
    ps = d_param_struct(stack._params(), stack._buffers_list())
    packed_all, packed = _guarded(lib.afi_d_packed_bytes(prec), dev)
    N.check(lib.afi_d_pack(ctx, prec, C.byref(ps), packed.data_ptr(), N.stream_ptr()))
    nws = lib.afi_d_workspace_bytes(prec, n, h, w, 1)
    ws_all, ws = _guarded(nws, dev)
    lg_all, lg8 = _guarded(n * h * w * 4, dev)
    dx_all, dx8 = _guarded(n * 256 * h * w * 4, dev)
    acc_all, acc = _guarded(lib.afi_d_gradacc_bytes(), dev)
    acc.zero_()
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(n, 256, h, w, generator=gen).cuda()
    dl = torch.randn(n, 1, h, w, generator=gen).cuda()
    call = N.DCall(x=N.view4(x), n=n, h=h, w=w, logits=lg8.data_ptr(), ws=ws.data_ptr(), ws_bytes=nws)
    N.check(lib.afi_d_forward(ctx, prec, C.byref(ps), packed.data_ptr(), C.byref(call), 1, 1, 0.1, 1e-5, 1, N.stream_ptr()))
    call.dlogits, call.dx = dl.data_ptr(), dx8.data_ptr()
    N.check(lib.afi_d_backward(ctx, prec, C.byref(ps), packed.data_ptr(), C.byref(call), 1, 1, acc.data_ptr(), N.stream_ptr()))
    torch.cuda.synchronize()
    for name, b in (("packed weights", packed_all), ("workspace", ws_all), ("logits", lg_all), ("dx", dx_all), ("gradient accumulator", acc_all)):
        assert _intact(b), f"{name}: canary band overwritten"
    if n * h * w > 1:           # (a single pixel has zero batch variance: the reference's output is all beta there too, nothing to compare)
        _, d_sd = O.init_states(0)
        ref = O.discriminator_forward(d_sd, x.cpu(), True)
        lg = lg8[: n * h * w * 4].view(torch.float32).view(n, 1, h, w)
        assert rel(lg, ref) < (2e-4 if precision != "bf16" else 3e-2)
