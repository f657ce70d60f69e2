"""torch.autograd glue around the C-ABI: parameter marshalling, packed-weight caching, workspaces."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import native as N

CH = 256


def _u8(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def g_param_struct(tensors: Sequence[Optional[torch.Tensor]], n_rdb: int) -> N.GParams:
    """tensors in state-dict order: head_w, head_b, (rdb r: conv1..conv5 weights)*, post_w, post_b, up_w, up_b, out_w, out_b."""
    s = N.GParams()
    s.n_rdb = n_rdb
    it = iter(tensors)
    s.head_w, s.head_b = N.ptr(next(it)), N.ptr(next(it))
    for r in range(n_rdb):
        for i in range(5):
            s.rdb_w[r][i] = N.ptr(next(it))
    s.post_w, s.post_b = N.ptr(next(it)), N.ptr(next(it))
    s.up_w, s.up_b = N.ptr(next(it)), N.ptr(next(it))
    s.out_w, s.out_b = N.ptr(next(it)), N.ptr(next(it))
    return s


def d_param_struct(params: Sequence[torch.Tensor], buffers: Sequence[Optional[torch.Tensor]]) -> N.DParams:
    """params: (w, b, gamma, beta) x 3, w4, b4 ; buffers: (running_mean, running_var, num_batches_tracked) x 3."""
    s = N.DParams()
    for i in range(3):
        s.w[i], s.b[i], s.gamma[i], s.beta[i] = (N.ptr(params[4 * i + k]) for k in range(4))
        s.running_mean[i], s.running_var[i], s.num_batches_tracked[i] = (N.ptr(buffers[3 * i + k]) for k in range(3))
    s.w[3], s.b[3] = N.ptr(params[12]), N.ptr(params[13])
    return s


def d_grad_struct(grads: Sequence[Optional[torch.Tensor]]) -> N.DGrads:
    s = N.DGrads()
    for i in range(3):
        s.w[i], s.b[i], s.gamma[i], s.beta[i] = (N.ptr(grads[4 * i + k]) for k in range(4))
    s.w[3], s.b[3] = N.ptr(grads[12]), N.ptr(grads[13])
    return s


class PackedWeights:
    """Packed GEMM-layout copy of a module's parameters, refreshed when any parameter version changes."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None
        self.key = None

    def get(self, kind: str, prec: int, params: Sequence[torch.Tensor], struct, n_rdb: int = 0) -> torch.Tensor:
        dev = params[0].device
        key = (kind, prec, dev, tuple((p.data_ptr(), p._version) for p in params))
        if self.buf is None or self.key != key:
            ctx = N.context(dev)
            nbytes = N.lib().afi_g_packed_bytes(prec, n_rdb) if kind == "g" else N.lib().afi_d_packed_bytes(prec)
            if self.buf is None or self.buf.numel() != max(int(nbytes), 16) or self.buf.device != dev:
                self.buf = _u8(nbytes, dev)      # otherwise re-packed IN PLACE: captured inference graphs keep pointing at it
            if kind == "g":
                N.check(N.lib().afi_g_pack(ctx, prec, C.byref(struct), self.buf.data_ptr(), N.stream_ptr()))
            else:
                N.check(N.lib().afi_d_pack(ctx, prec, C.byref(struct), self.buf.data_ptr(), N.stream_ptr()))
            self.key = key
        return self.buf


class AFInterpolatorFn(torch.autograd.Function):
    """Generator.forward (reference generator_rdb.py:123-130) with the stage-1 top-left crop folded in and, optionally, the
    FPN / PAFPN top-down merge fused: y = scale * (conv1x1(lat_x; lat_w, lat_b) + G(x))  (fpn_sr.py:151-157)."""

    @staticmethod
    def forward(ctx, x, lat_x, lat_w, lat_b, holder, prec: int, out_hw, scale: float, grad_enabled: bool, *params):
        if not x.is_cuda:
            raise RuntimeError("AF interpolator: input must live on an sm_100a CUDA device (no CPU fallback)")
        if x.dim() != 4 or x.size(1) != CH:
            raise ValueError(f"AF interpolator expects [N,{CH},H,W], got {tuple(x.shape)}")
        x = N.boundary(x)
        n, _, h, w = x.shape
        oh, ow = out_hw if out_hw is not None else (2 * h, 2 * w)
        n_rdb = holder.n_rdb
        dev = x.device
        ps = g_param_struct(params, n_rdb)
        packed = holder.packed.get("g", prec, params, ps, n_rdb)
        need_bwd = grad_enabled and any(ctx.needs_input_grad)
        lib, actx = N.lib(), N.context(dev)
        lat_c, lat = 0, None
        if lat_x is not None:
            lat_x, lat_w = N.boundary(lat_x), lat_w.float().contiguous()
            lat_c = lat_x.size(1)
            if tuple(lat_x.shape) != (n, lat_c, oh, ow) or tuple(lat_w.shape[:2]) != (CH, lat_c):
                raise ValueError(f"lateral input {tuple(lat_x.shape)} / weight {tuple(lat_w.shape)} do not match the output [{n},{CH},{oh},{ow}]")
            lat = N.Lateral(lat_x=N.view4(lat_x), lat_c=lat_c, lat_w=lat_w.data_ptr(), lat_b=N.ptr(lat_b), scale=scale)
        ws = _u8(lib.afi_g_workspace_bytes(prec, n, h, w, n_rdb, lat_c, int(need_bwd)), dev)
        y = torch.empty((n, CH, oh, ow), dtype=torch.float32, device=dev)
        call = N.GCall(x=N.view4(x), n=n, h=h, w=w, y=y.data_ptr(), oh=oh, ow=ow, ws=ws.data_ptr(), ws_bytes=ws.numel())
        if lat is not None:
            call.lateral = C.pointer(lat)
        N.check(lib.afi_g_forward(actx, prec, C.byref(ps), packed.data_ptr(), C.byref(call), 1, int(need_bwd), N.stream_ptr()))
        ctx.prec, ctx.shape, ctx.n_rdb, ctx.scale, ctx.lat_c = prec, (n, h, w, oh, ow), n_rdb, scale, lat_c
        ctx.ws, ctx.packed, ctx.holder = ws, packed, holder
        ctx.has_lat_b = lat_b is not None
        ctx.save_for_backward(lat_x if lat_x is not None else x.new_empty(0), lat_w if lat_w is not None else x.new_empty(0, dtype=torch.float32),
                              lat_b if lat_b is not None else x.new_empty(0), *params)
        return y

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        lat_x, lat_w, lat_b, *params = ctx.saved_tensors
        n, h, w, oh, ow = ctx.shape
        dev = dy.device
        lib, actx = N.lib(), N.context(dev)
        dy = N.boundary(dy)
        holder = ctx.holder
        want = [bool(ctx.needs_input_grad[9 + i]) for i in range(len(params))]
        deferred = holder.deferred and all(want)
        if deferred:
            # ONE packed accumulator for every call of this backward pass; un-packed into .grad by a callback when the pass ends
            if holder.acc is None or holder.acc.device != dev:
                holder.acc = _u8(lib.afi_g_gradacc_bytes(ctx.n_rdb), dev)
            acc = holder.acc
            task = torch._C._current_graph_task_id()
            if holder.pending is None or holder.pending[4] != task:      # first call of this pass (a pass that raised leaves a stale entry)
                N.check(lib.afi_zero(acc.data_ptr(), acc.numel(), N.stream_ptr()))
                holder.pending = (ctx.prec, ctx.n_rdb, list(params), torch.cuda.current_stream(dev), task)
                torch.autograd.Variable._execution_engine.queue_callback(lambda: _flush_deferred(holder, task))
        else:
            acc = _u8(lib.afi_g_gradacc_bytes(ctx.n_rdb), dev)
            N.check(lib.afi_zero(acc.data_ptr(), acc.numel(), N.stream_ptr()))
        ps = g_param_struct(params, ctx.n_rdb)
        call = N.GCall(n=n, h=h, w=w, oh=oh, ow=ow, dy=N.view4(dy), ws=ctx.ws.data_ptr(), ws_bytes=ctx.ws.numel())
        dx = d_lat_x = d_lat_w = d_lat_b = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((n, CH, h, w), dtype=torch.float32, device=dev)
            call.dx = dx.data_ptr()
        lat = None
        if ctx.lat_c:
            lat = N.Lateral(lat_x=N.view4(lat_x), lat_c=ctx.lat_c, lat_w=lat_w.data_ptr(), lat_b=N.ptr(lat_b) if ctx.has_lat_b else None,
                            scale=ctx.scale)
            call.lateral = C.pointer(lat)
            if ctx.needs_input_grad[1]:
                d_lat_x = torch.empty((n, ctx.lat_c, oh, ow), dtype=torch.float32, device=dev)
                call.lat_dx = d_lat_x.data_ptr()
            if ctx.needs_input_grad[2]:
                d_lat_w = torch.empty_like(lat_w)
                call.lat_gw = d_lat_w.data_ptr()
            if ctx.has_lat_b and ctx.needs_input_grad[3]:
                d_lat_b = torch.empty_like(lat_b)
                call.lat_gb = d_lat_b.data_ptr()
        N.check(lib.afi_g_backward(actx, ctx.prec, C.byref(ps), ctx.packed.data_ptr(), C.byref(call), 1, acc.data_ptr(), N.stream_ptr()))
        if deferred:
            grads = [None] * len(params)
        else:
            grads = [torch.empty_like(p) if w_ else None for p, w_ in zip(params, want)]
            gs = g_param_struct(grads, ctx.n_rdb)
            N.check(lib.afi_g_unpack_grads(actx, ctx.prec, acc.data_ptr(), C.byref(gs), 1.0, 0, N.stream_ptr()))
        ctx.ws = None
        return (dx, d_lat_x, d_lat_w, d_lat_b, None, None, None, None, None, *grads)


def _flush_deferred(holder, task: int) -> None:
    """End-of-backward callback of the deferred weight gradients: p.grad (+)= unpack(accumulator), one launch for all 23 tensors."""
    if holder.pending is None or holder.pending[4] != task:
        return
    prec, n_rdb, params, stream, _ = holder.pending
    holder.pending = None
    lib = N.lib()
    dev = params[0].device
    with torch.no_grad(), torch.cuda.device(dev), torch.cuda.stream(stream):     # the stream the backward calls were issued on
        actx = N.context(dev)
        if all(p.grad is not None for p in params):
            gs = g_param_struct([p.grad for p in params], n_rdb)
            N.check(lib.afi_g_unpack_grads(actx, prec, holder.acc.data_ptr(), C.byref(gs), 1.0, 1, N.stream_ptr()))
        else:
            flat = torch.empty(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
            views, off = [], 0
            for p in params:
                views.append(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
            gs = g_param_struct(views, n_rdb)
            N.check(lib.afi_g_unpack_grads(actx, prec, holder.acc.data_ptr(), C.byref(gs), 1.0, 0, N.stream_ptr()))
            for p, v in zip(params, views):
                p.grad = v if p.grad is None else p.grad + v


class InferenceGraphs:
    """Forward-only interpolator calls replayed from CUDA graphs.  At inference the interpolator runs on small maps, one call per merge
    site (28 per image in a 7-layer BiFPN), and a call is ~22 kernel launches: issued eagerly the HOST is the bottleneck (0.47 ms per
    call at any image size, 13 ms per image, measured); a captured graph replays in tens of microseconds.  One graph per (shapes,
    operand pointers) with static input / output buffers; LRU-bounded."""

    MAX_ENTRIES = 64

    def __init__(self):
        import collections
        self.entries = collections.OrderedDict()

    @staticmethod
    def _issue(holder, prec, params, packed, x, y, ws, lat, fuse_cur, fuse_w):
        n, _, h, w = x.shape
        ps = g_param_struct(params, holder.n_rdb)
        call = N.GCall(x=N.view4(x), n=n, h=h, w=w, y=y.data_ptr(), oh=y.size(2), ow=y.size(3), ws=ws.data_ptr(), ws_bytes=ws.numel())
        keep = None
        if lat is not None:
            lat_x, lat_w, lat_b, scale = lat
            keep = N.Lateral(lat_x=N.view4(lat_x), lat_c=lat_x.size(1), lat_w=lat_w.data_ptr(), lat_b=N.ptr(lat_b), scale=scale)
            call.lateral = C.pointer(keep)
        if fuse_cur is not None:
            call.fuse_cur, call.fuse_w = N.view4(fuse_cur), fuse_w.data_ptr()
        N.check(N.lib().afi_g_forward(N.context(x.device), prec, C.byref(ps), packed.data_ptr(), C.byref(call), 1, 0, N.stream_ptr()))

    def run(self, holder, prec: int, params, x: torch.Tensor, out_hw, lat=None, fuse=None) -> torch.Tensor:
        """lat = (lat_x, lat_w [256, lat_c], lat_b or None, scale); fuse = (cur, weight[2] or None)."""
        if not x.is_cuda:
            raise RuntimeError("AF interpolator: input must live on an sm_100a CUDA device (no CPU fallback)")
        x = N.boundary(x)
        n, c, h, w = x.shape
        if c != CH:
            raise ValueError(f"AF interpolator expects [N,{CH},H,W], got {tuple(x.shape)}")
        oh, ow = out_hw if out_hw is not None else (2 * h, 2 * w)
        dev = x.device
        ps = g_param_struct(params, holder.n_rdb)
        packed = holder.packed.get("g", prec, params, ps, holder.n_rdb)
        key = (prec, dev, n, h, w, oh, ow, packed.data_ptr(), tuple(p.data_ptr() for p in params),
               None if lat is None else (lat[0].size(1), lat[1].data_ptr(), N.ptr(lat[2]) if lat[2] is not None else 0, float(lat[3]), lat[0].dtype),
               None if fuse is None else fuse[0].dtype, x.dtype)
        e = self.entries.get(key)
        if e is None:
            lat_c = 0 if lat is None else lat[0].size(1)
            e = {"x": torch.empty((n, CH, h, w), dtype=x.dtype, device=dev),
                 "y": torch.empty((n, CH, oh, ow), dtype=torch.float32, device=dev),
                 "ws": _u8(N.lib().afi_g_workspace_bytes(prec, n, h, w, holder.n_rdb, lat_c, 0), dev),
                 "lat_x": None if lat is None else torch.empty((n, lat_c, oh, ow), dtype=lat[0].dtype, device=dev),
                 "cur": None if fuse is None else torch.empty((n, CH, oh, ow), dtype=fuse[0].dtype, device=dev),
                 "fw": None if fuse is None else torch.ones(2, dtype=torch.float32, device=dev)}
            e["x"].zero_()
            lat_s = None if lat is None else (e["lat_x"].zero_(), lat[1], lat[2], lat[3])
            args = (holder, prec, params, packed, e["x"], e["y"], e["ws"], lat_s, None if fuse is None else e["cur"].zero_(), e["fw"])
            self._issue(*args)                            # warm-up outside the capture (lazy module state, descriptor caches)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._issue(*args)
            e["graph"] = g
            self.entries[key] = e
            while len(self.entries) > self.MAX_ENTRIES:
                self.entries.popitem(last=False)
        else:
            self.entries.move_to_end(key)
        e["x"].copy_(x)
        if lat is not None:
            e["lat_x"].copy_(lat[0])
        if fuse is not None:
            e["cur"].copy_(fuse[0])
            if fuse[1] is not None:
                e["fw"].copy_(fuse[1].detach().reshape(2))
            else:
                e["fw"].fill_(1.0)
        e["graph"].replay()
        return e["y"].clone()


def afi_bifpn_fuse(x: torch.Tensor, cur: torch.Tensor, weight: Optional[torch.Tensor], holder, prec: int, params) -> torch.Tensor:
    """Forward-only BiFPN fusion site (reference bifpn_sr.py:535-548) in ONE library call:
    weight[0] * cur + weight[1] * (Generators[0](x) + bilinear2x(x))[:, :, :H, :W]  (weight None: plain sum)."""
    if not x.is_cuda:
        raise RuntimeError("AF interpolator: input must live on an sm_100a CUDA device (no CPU fallback)")
    x, cur = N.boundary(x), N.boundary(cur)
    n, _, h, w = x.shape
    oh, ow = cur.shape[2:]
    if tuple(cur.shape[:2]) != (n, CH) or oh > 2 * h or ow > 2 * w:
        raise ValueError(f"BiFPN fusion: cur {tuple(cur.shape)} does not fit the up-sampled top {(n, CH, 2 * h, 2 * w)}")
    dev = x.device
    fw = (weight.detach().float().reshape(2) if weight is not None else torch.ones(2, device=dev)).contiguous()
    ps = g_param_struct(params, holder.n_rdb)
    packed = holder.packed.get("g", prec, params, ps, holder.n_rdb)
    lib, actx = N.lib(), N.context(dev)
    ws = _u8(lib.afi_g_workspace_bytes(prec, n, h, w, holder.n_rdb, 0, 0), dev)
    y = torch.empty((n, CH, oh, ow), dtype=torch.float32, device=dev)
    call = N.GCall(x=N.view4(x), n=n, h=h, w=w, y=y.data_ptr(), oh=oh, ow=ow, ws=ws.data_ptr(), ws_bytes=ws.numel(),
                   fuse_cur=N.view4(cur), fuse_w=fw.data_ptr())
    N.check(lib.afi_g_forward(actx, prec, C.byref(ps), packed.data_ptr(), C.byref(call), 1, 0, N.stream_ptr()))
    return y


class PatchDiscriminatorFn(torch.autograd.Function):
    """Discriminators[0](x) (reference feature_patch_discriminator.py:32-41)."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, holder, prec: int, training: bool, momentum: float, eps: float, buffers, grad_enabled: bool,
                *params: torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError("feature-patch discriminator: input must live on an sm_100a CUDA device (no CPU fallback)")
        if x.dim() != 4 or x.size(1) != CH:
            raise ValueError(f"discriminator expects [N,{CH},H,W], got {tuple(x.shape)}")
        x = N.boundary(x)
        n, _, h, w = x.shape
        dev = x.device
        ps = d_param_struct(params, buffers)
        packed = holder.packed.get("d", prec, params, ps)
        need_bwd = grad_enabled and (any(ctx.needs_input_grad[8:]) or ctx.needs_input_grad[0])
        lib, actx = N.lib(), N.context(dev)
        ws = _u8(lib.afi_d_workspace_bytes(prec, n, h, w, int(need_bwd)), dev)
        logits = torch.empty((n, 1, h, w), dtype=torch.float32, device=dev)
        call = N.DCall(x=N.view4(x), n=n, h=h, w=w, logits=logits.data_ptr(), ws=ws.data_ptr(), ws_bytes=ws.numel())
        N.check(lib.afi_d_forward(actx, prec, C.byref(ps), packed.data_ptr(), C.byref(call), 1, int(training), momentum, eps, int(need_bwd),
                                  N.stream_ptr()))
        ctx.prec, ctx.shape, ctx.ws, ctx.packed, ctx.buffers, ctx.training = prec, (n, h, w), ws, packed, buffers, bool(training)
        ctx.save_for_backward(*params)
        return logits

    @staticmethod
    def backward(ctx, dlogits: torch.Tensor):
        params = ctx.saved_tensors
        n, h, w = ctx.shape
        dev = dlogits.device
        lib, actx = N.lib(), N.context(dev)
        dl = dlogits.float().contiguous()
        acc = _u8(lib.afi_d_gradacc_bytes(), dev)
        N.check(lib.afi_zero(acc.data_ptr(), acc.numel(), N.stream_ptr()))
        ps = d_param_struct(params, ctx.buffers)
        call = N.DCall(n=n, h=h, w=w, dlogits=dl.data_ptr(), ws=ctx.ws.data_ptr(), ws_bytes=ctx.ws.numel())
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((n, CH, h, w), dtype=torch.float32, device=dev)
            call.dx = dx.data_ptr()
        N.check(lib.afi_d_backward(actx, ctx.prec, C.byref(ps), ctx.packed.data_ptr(), C.byref(call), 1, int(ctx.training), acc.data_ptr(),
                                   N.stream_ptr()))
        grads = [torch.empty_like(p) if ctx.needs_input_grad[8 + i] else None for i, p in enumerate(params)]
        gs = d_grad_struct(grads)
        N.check(lib.afi_d_unpack_grads(actx, ctx.prec, acc.data_ptr(), C.byref(gs), 1.0, 0, N.stream_ptr()))
        ctx.ws = None
        return (dx, None, None, None, None, None, None, None, *grads)


def bce_with_logits(logits: torch.Tensor, target: float) -> torch.Tensor:
    """nn.BCEWithLogitsLoss()(logits, full_like(logits, target)) -- forward only (stage-1 G phase: stage1_trainer.py:408)."""
    out = torch.zeros((), dtype=torch.float32, device=logits.device)
    lg = logits.detach().float().contiguous()
    N.check(N.lib().afi_bce_with_logits(lg.data_ptr(), lg.numel(), float(target), out.data_ptr(), None, 0.0, None, 0.0, N.stream_ptr()))
    return out


def conv3x3(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], lrelu: bool, precision: str) -> torch.Tensor:
    """Single 3x3/s1/p1 convolution through the library's GEMM engine (unit tests, kernel benchmarks)."""
    prec = N.PRECISIONS[precision]
    n, cin, h, w = x.shape
    cout = weight.shape[0]
    lib, actx = N.lib(), N.context(x.device)
    ws = _u8(lib.afi_conv3x3_workspace_bytes(prec, n, cin, h, w, cout), x.device)
    y = torch.empty((n, cout, h, w), dtype=torch.float32, device=x.device)
    N.check(lib.afi_conv3x3(actx, prec, N.view4(x), n, cin, h, w, weight.data_ptr(), N.ptr(bias), cout, int(lrelu), y.data_ptr(),
                            ws.data_ptr(), ws.numel(), N.stream_ptr()))
    return y


def conv3x3_backward(x: torch.Tensor, dy: torch.Tensor, weight: torch.Tensor, precision: str, need_dx: bool = True):
    prec = N.PRECISIONS[precision]
    n, cin, h, w = x.shape
    cout = weight.shape[0]
    lib, actx = N.lib(), N.context(x.device)
    ws = _u8(lib.afi_conv3x3_workspace_bytes(prec, n, cin, h, w, cout), x.device)
    dw = torch.empty_like(weight)
    dx = torch.empty(x.shape, dtype=torch.float32, device=x.device) if need_dx else None
    N.check(lib.afi_conv3x3_backward(actx, prec, N.view4(x), N.view4(dy), n, cin, h, w, weight.data_ptr(), cout, dw.data_ptr(), None, N.ptr(dx),
                                     ws.data_ptr(), ws.numel(), N.stream_ptr()))
    return dw, dx


class Conv3x3Fn(torch.autograd.Function):
    """3x3 / stride 1 / pad 1 convolution (+bias) through the library's implicit-GEMM engine with autograd: the necks' output convs
    (reference fpn_sr.py:144-158, pafpn_sr.py:184-193) when they carry no norm."""

    @staticmethod
    def forward(ctx, x, weight, bias, prec: int):
        if not x.is_cuda:
            raise RuntimeError("conv3x3: input must live on an sm_100a CUDA device (no CPU fallback)")
        x, weight = N.boundary(x), weight.float().contiguous()
        n, cin, h, w = x.shape
        cout = weight.shape[0]
        lib, actx = N.lib(), N.context(x.device)
        ws = _u8(lib.afi_conv3x3_workspace_bytes(prec, n, cin, h, w, cout), x.device)
        y = torch.empty((n, cout, h, w), dtype=torch.float32, device=x.device)
        N.check(lib.afi_conv3x3(actx, prec, N.view4(x), n, cin, h, w, weight.data_ptr(), N.ptr(bias), cout, 0, y.data_ptr(), ws.data_ptr(),
                                ws.numel(), N.stream_ptr()))
        ctx.prec, ctx.has_bias = prec, bias is not None
        ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = N.boundary(dy)
        n, cin, h, w = x.shape
        cout = weight.shape[0]
        lib, actx = N.lib(), N.context(x.device)
        ws = _u8(lib.afi_conv3x3_workspace_bytes(ctx.prec, n, cin, h, w, cout), x.device)
        dw = torch.empty_like(weight)
        db = torch.empty(cout, dtype=torch.float32, device=x.device) if ctx.has_bias and ctx.needs_input_grad[2] else None
        dx = torch.empty(x.shape, dtype=torch.float32, device=x.device) if ctx.needs_input_grad[0] else None
        N.check(lib.afi_conv3x3_backward(actx, ctx.prec, N.view4(x), N.view4(dy), n, cin, h, w, weight.data_ptr(), cout, dw.data_ptr(), N.ptr(db),
                                         N.ptr(dx), ws.data_ptr(), ws.numel(), N.stream_ptr()))
        return dx, dw, db, None


class Conv1x1Fn(torch.autograd.Function):
    """1x1 convolution (+bias) through the library's GEMM engine with autograd: the necks' lateral convs incl. the top-level one (reference
    fpn_sr.py:79-81, 144-145), the BiFPN's input laterals and the pointwise half of its depthwise-separable convs (bifpn_sr.py:160-183,
    bifpn_layers/wrappers.py:166-206)."""

    @staticmethod
    def forward(ctx, x, weight, bias, prec: int):
        if not x.is_cuda:
            raise RuntimeError("conv1x1: input must live on an sm_100a CUDA device (no CPU fallback)")
        x, weight = N.boundary(x), weight.float().contiguous()
        n, cin, h, w = x.shape
        cout = weight.shape[0]
        lib, actx = N.lib(), N.context(x.device)
        ws = _u8(lib.afi_conv1x1_workspace_bytes(prec, n, cin, h, w, cout), x.device)
        y = torch.empty((n, cout, h, w), dtype=torch.float32, device=x.device)
        N.check(lib.afi_conv1x1(actx, prec, N.view4(x), n, cin, h, w, weight.data_ptr(), N.ptr(bias), cout, y.data_ptr(), ws.data_ptr(),
                                ws.numel(), N.stream_ptr()))
        ctx.prec, ctx.has_bias = prec, bias is not None
        ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = N.boundary(dy)
        n, cin, h, w = x.shape
        cout = weight.shape[0]
        lib, actx = N.lib(), N.context(x.device)
        ws = _u8(lib.afi_conv1x1_workspace_bytes(ctx.prec, n, cin, h, w, cout), x.device)
        dw = torch.empty_like(weight)
        db = torch.empty(cout, dtype=torch.float32, device=x.device) if ctx.has_bias and ctx.needs_input_grad[2] else None
        dx = torch.empty(x.shape, dtype=torch.float32, device=x.device) if ctx.needs_input_grad[0] else None
        N.check(lib.afi_conv1x1_backward(actx, ctx.prec, N.view4(x), N.view4(dy), n, cin, h, w, weight.data_ptr(), cout, dw.data_ptr(), N.ptr(db),
                                         N.ptr(dx), ws.data_ptr(), ws.numel(), N.stream_ptr()))
        return dx, dw, db, None


def conv1x1_autograd(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], precision: Optional[str] = None) -> torch.Tensor:
    return Conv1x1Fn.apply(x, weight, bias, N.PRECISIONS[precision or N.default_precision()])


class Conv3x3S2Fn(torch.autograd.Function):
    """3x3 / stride 2 / pad 1 convolution (+bias) through the tcgen05 engine: the PANet neck's bottom-up down-sampling convs (reference
    pafpn_sr.py:103-117, 186-193) -- a 9-tap stride-1 implicit GEMM over the four sub-pixel phase views of the input (no wasted products)."""

    @staticmethod
    def forward(ctx, x, weight, bias, prec: int):
        if not x.is_cuda:
            raise RuntimeError("conv3x3s2: input must live on an sm_100a CUDA device (no CPU fallback)")
        x, weight = N.boundary(x), weight.float().contiguous()
        n, cin, h, w = x.shape
        cout = weight.shape[0]
        lib, actx = N.lib(), N.context(x.device)
        ws = _u8(lib.afi_conv3x3s2_workspace_bytes(prec, n, cin, h, w, cout), x.device)
        y = torch.empty((n, cout, (h + 1) // 2, (w + 1) // 2), dtype=torch.float32, device=x.device)
        N.check(lib.afi_conv3x3s2(actx, prec, N.view4(x), n, cin, h, w, weight.data_ptr(), N.ptr(bias), cout, y.data_ptr(), ws.data_ptr(),
                                  ws.numel(), N.stream_ptr()))
        ctx.prec, ctx.has_bias = prec, bias is not None
        ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = N.boundary(dy)
        n, cin, h, w = x.shape
        cout = weight.shape[0]
        lib, actx = N.lib(), N.context(x.device)
        ws = _u8(lib.afi_conv3x3s2_workspace_bytes(ctx.prec, n, cin, h, w, cout), x.device)
        dw = torch.empty_like(weight)
        db = torch.empty(cout, dtype=torch.float32, device=x.device) if ctx.has_bias and ctx.needs_input_grad[2] else None
        dx = torch.empty(x.shape, dtype=torch.float32, device=x.device) if ctx.needs_input_grad[0] else None
        N.check(lib.afi_conv3x3s2_backward(actx, ctx.prec, N.view4(x), N.view4(dy), n, cin, h, w, weight.data_ptr(), cout, dw.data_ptr(),
                                           N.ptr(db), N.ptr(dx), ws.data_ptr(), ws.numel(), N.stream_ptr()))
        return dx, dw, db, None


def conv3x3s2_autograd(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], precision: Optional[str] = None) -> torch.Tensor:
    return Conv3x3S2Fn.apply(x, weight, bias, N.PRECISIONS[precision or N.default_precision()])


def sepconv_eval(x: torch.Tensor, dw_weight: torch.Tensor, pw_weight: torch.Tensor, pw_bias: Optional[torch.Tensor], pre_swish: bool,
                 precision: Optional[str] = None) -> torch.Tensor:
    """Forward-only depthwise-separable conv of the BiFPN neck in ONE library call (reference bifpn_layers/wrappers.py:166-206): the swish in
    front of it, the depthwise 3x3 and the layout conversion are one HBM-bound pass, the pointwise conv (norm already folded into
    pw_weight / pw_bias by the caller) runs on the tcgen05 engine."""
    if not x.is_cuda:
        raise RuntimeError("sepconv: input must live on an sm_100a CUDA device (no CPU fallback)")
    prec = N.PRECISIONS[precision or N.default_precision()]
    x = N.boundary(x)
    n, c, h, w = x.shape
    cout = pw_weight.shape[0]
    lib, actx = N.lib(), N.context(x.device)
    ws = _u8(lib.afi_sepconv_workspace_bytes(prec, n, c, h, w, cout), x.device)
    y = torch.empty((n, cout, h, w), dtype=torch.float32, device=x.device)
    N.check(lib.afi_sepconv(actx, prec, N.view4(x), n, c, h, w, dw_weight.data_ptr(), pw_weight.data_ptr(), N.ptr(pw_bias), cout, int(pre_swish),
                            y.data_ptr(), ws.data_ptr(), ws.numel(), N.stream_ptr()))
    return y


def bifpn_fuse_down(a: torch.Tensor, b: Optional[torch.Tensor], down: torch.Tensor, weight: Optional[torch.Tensor]) -> torch.Tensor:
    """Forward-only bottom-up fusion site of the BiFPN neck (reference bifpn_sr.py:550-564): w[0]*a + w[1]*b + w[2]*pool(down) (or the two-term
    form when b is None), pool = the zero-padded 3x3 / stride-2 max-pool of the reference's MaxPool2d wrapper -- one elementwise pass."""
    if not a.is_cuda:
        raise RuntimeError("bifpn_fuse_down: tensors must live on an sm_100a CUDA device (no CPU fallback)")
    a, down = N.boundary(a), N.boundary(down)
    n, c, h, w = a.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=a.device)
    bv = N.view4(N.boundary(b)) if b is not None else N.View4()
    wt = weight.detach().float().contiguous() if weight is not None else None
    N.check(N.lib().afi_bifpn_fuse_down(N.view4(a), bv, N.view4(down), N.ptr(wt), n, c, h, w, down.size(2), down.size(3), out.data_ptr(),
                                        N.stream_ptr()))
    return out


class FuseActFn(torch.autograd.Function):
    """Top-down BiFPN fusion site with autograd (reference bifpn_sr.py:542-548, swish of :591-594): act(w0 * cur + w1 * up) in one pass."""

    @staticmethod
    def forward(ctx, cur, up, weight, act: bool):
        if not cur.is_cuda:
            raise RuntimeError("bifpn fusion: tensors must live on an sm_100a CUDA device (no CPU fallback)")
        cur, up = N.boundary(cur), N.boundary(up)
        if cur.shape != up.shape or cur.dim() != 4:
            raise ValueError(f"bifpn fusion: shapes {tuple(cur.shape)} / {tuple(up.shape)}")
        n, c, h, w = cur.shape
        wt = weight.detach().float().contiguous() if weight is not None else None
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=cur.device)
        s = torch.empty_like(out) if act else None
        N.check(N.lib().afi_bifpn_fuse_act(N.view4(cur), N.view4(up), N.ptr(wt), int(act), n, c, h, w, N.ptr(s), out.data_ptr(), N.stream_ptr()))
        ctx.act, ctx.has_w = bool(act), weight is not None
        ctx.save_for_backward(cur, up, wt if wt is not None else out.new_empty(0), s if s is not None else out.new_empty(0))
        return out

    @staticmethod
    def backward(ctx, dout):
        cur, up, wt, s = ctx.saved_tensors
        dout = N.boundary(dout)
        n, c, h, w = cur.shape
        dev = cur.device
        d_cur = torch.empty((n, c, h, w), dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        d_up = torch.empty((n, c, h, w), dtype=torch.float32, device=dev) if ctx.needs_input_grad[1] else None
        d_w = torch.empty(2, dtype=torch.float32, device=dev) if (ctx.has_w and ctx.needs_input_grad[2]) else None
        N.check(N.lib().afi_bifpn_fuse_act_backward(N.view4(dout), N.ptr(s) if ctx.act else None, N.view4(cur), N.view4(up),
                                                    N.ptr(wt) if ctx.has_w else None, int(ctx.act), n, c, h, w, N.ptr(d_cur), N.ptr(d_up),
                                                    N.ptr(d_w), N.stream_ptr()))
        return d_cur, d_up, d_w, None


def bifpn_fuse_act(cur: torch.Tensor, up: torch.Tensor, weight: Optional[torch.Tensor], swish: bool) -> torch.Tensor:
    return FuseActFn.apply(cur, up, weight, bool(swish))


def conv3x3_autograd(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], precision: Optional[str] = None) -> torch.Tensor:
    return Conv3x3Fn.apply(x, weight, bias, N.PRECISIONS[precision or N.default_precision()])
