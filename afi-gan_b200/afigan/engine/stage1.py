"""The stage-1 AFI-GAN training step on pre-extracted features: AFIGAN_Trainer.run_step of the reference
(afigan/engine/stage1_trainer.py:305-435) minus the guide-model forward and the logging plumbing.

Per level l in p2..p6:
  D phase  tr = crop(G(lr_l)).detach(); d_l = BCE(D0(hr_l), 1) + BCE(D0(tr), 0)   (real BEFORE fake, :349-350)
           D.zero_grad(); sum(d_l).backward(); SGD(D)
  G phase  tr = crop(G(lr_l)); adv = BCE(D0(tr).detach(), 1) [no gradient, :399-408]; _ = D0(hr_l) (fake BEFORE real)
           g_l = 1e-3 * adv + L1(tr, hr_l); G.zero_grad(); sum(g_l).backward(); SGD(G)

This fast path drives the C-ABI directly (no autograd graph, no per-level host syncs, losses stay on the device),
keeps packed gradient accumulators across the five levels, and all-reduces ONE flat gradient buffer per optimiser
(the DDP semantics the reference intended: stage1_trainer.py:80-89).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from .. import native as N
from ..functional import _u8, d_grad_struct, d_param_struct, g_param_struct
from .sync import FlatGradSync

CH = 256


class _StepBase:
    """State and helpers shared by the fused stage-1 step and the grouped stage-2 loss block: the discriminator side (flat gradient buffer,
    momentum, packed weights, per-call workspaces, grouped forward / backward over level x {real, fake} calls), the loss slots and the
    multi-tensor SGD."""

    def _init_common(self, dev, lr, momentum, weight_decay, weight_decay_norm, precision, process_group, overlap, overlap_comm):
        import os
        if dev.type != "cuda":
            raise RuntimeError("the fused training steps need the modules on an sm_100a CUDA device (no CPU fallback)")
        self.dev = dev
        self.lr, self.momentum, self.wd, self.wd_norm = lr, momentum, weight_decay, weight_decay_norm
        self.precision = precision
        self.prec = N.PRECISIONS[self.precision]
        self.pg = process_group
        self.lib, self.ctx = N.lib(), N.context(dev)
        self.steps_done = 0
        self._mom_restored = False
        # Gradient all-reduces off the critical path: D's runs while the generator's backward pass (which does not read D) computes, G's
        # while the G phase's discriminator forwards do; only the optimiser update waits for its collective.
        self.overlap_comm = overlap_comm and os.environ.get("AFIGAN_OVERLAP_COMM", "1") != "0"
        self._ws: Dict[tuple, torch.Tensor] = {}
        self._bufs: Dict[tuple, torch.Tensor] = {}
        self._sgd_tabs: Dict[int, tuple] = {}
        self.losses = torch.zeros(4, 8, dtype=torch.float32, device=dev)   # rows: d_loss, g_loss, adv, content ; cols: levels
        self._tmp = torch.zeros(64, dtype=torch.float32, device=dev)
        # Two side streams: the discriminator calls of a phase are issued as two groups (real / fake) whose HBM-bound passes (BatchNorm
        # apply / backward, reductions, head) overlap the other group's tensor-bound GEMMs -- the persistent GEMM CTAs leave enough
        # registers / shared memory per SM for the small elementwise CTAs to co-reside.
        self.overlap = overlap
        self.streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)] if overlap else []
        self.g_params: List[torch.nn.Parameter] = []
        self.n_rdb = 0

    def _init_d(self, D, distributed):
        self.D = D
        self.Dstack = D.Discriminators[0]
        self.d_params: List[torch.nn.Parameter] = self.Dstack._params()
        # flat gradient buffer: param.grad are views, so one all-reduce per optimiser covers every parameter
        self.d_sync = FlatGradSync(self.d_params, self.pg, distributed)
        self.d_flat, self.d_grads = self.d_sync.flat, self.d_sync.views
        self.distributed, self.world = self.d_sync.enabled, self.d_sync.world
        self.d_mom = [torch.zeros_like(p) for p in self.d_params]
        self.d_acc = _u8(self.lib.afi_d_gradacc_bytes(), self.dev)
        self.d_packed = _u8(self.lib.afi_d_packed_bytes(self.prec), self.dev)

    def _d_update(self):
        """D's optimiser step (one multi-tensor launch; BatchNorm affine parameters without weight decay) + re-layout of its GEMM operands."""
        self._sgd(self.d_params, self.d_grads, self.d_mom, [i % 4 >= 2 and i < 12 for i in range(14)])
        self._pack_d()
        self.Dstack._native.packed.key = None      # the module's own packed copy (autograd path) is stale now

    def _param_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.g_params + self.d_params)

    def _refresh_packed(self):
        """Re-pack the GEMM-layout weight copies when a parameter changed behind this object's back (load_state_dict, checkpoint interop, an
        external optimiser, a broadcast): the in-library SGD updates do not bump torch's version counters and re-pack explicitly."""
        key = self._param_key()
        if key != self._pack_key:
            if self.g_params:
                self._pack_g()
            self._pack_d()
            self._pack_key = key

    def load_optimizer_state(self, g_momentum: Sequence[torch.Tensor], d_momentum: Sequence[torch.Tensor], steps_done: int):
        """Resume: restored SGD momentum buffers (state-dict order) and the iteration count (detectron2 checkpoints carry both)."""
        for dst, src in zip(getattr(self, "g_mom", []), g_momentum):
            dst.copy_(src)
        for dst, src in zip(self.d_mom, d_momentum):
            dst.copy_(src)
        self.steps_done = int(steps_done)
        self._mom_restored = True
        if self.distributed:
            if hasattr(self, "g_sync"):
                self.g_sync.broadcast_parameters(self.g_mom)
            self.d_sync.broadcast_parameters(self.d_mom)

    def optimizer_state_dict(self) -> Dict[str, object]:
        """What a trainer checkpoints next to the two model files (detectron2 saves `optimizer` / `iteration` alongside `model`): the SGD
        momentum buffers in state-dict order and the number of completed steps.  `AFCheckpointer(G, ..., optimizer=step)` stores it."""
        return {"g_momentum": [m.detach().clone() for m in getattr(self, "g_mom", [])], "d_momentum": [m.detach().clone() for m in self.d_mom],
                "steps_done": self.steps_done, "lr": self.lr, "momentum": self.momentum, "weight_decay": self.wd, "weight_decay_norm": self.wd_norm}

    state_dict = optimizer_state_dict            # so that AFCheckpointer treats the step object as a checkpointable

    def load_state_dict(self, state: Dict[str, object]) -> None:
        self.load_optimizer_state(state.get("g_momentum", []), state["d_momentum"], int(state["steps_done"]))

    def _ws_for(self, kind: str, n: int, h: int, w: int, save: bool, tag="") -> torch.Tensor:
        key = (kind, n, h, w, save, tag)
        if key not in self._ws:
            if kind == "g":
                nb = self.lib.afi_g_workspace_bytes(self.prec, n, h, w, self.n_rdb, 0, int(save))
            else:
                nb = self.lib.afi_d_workspace_bytes(self.prec, n, h, w, int(save))
            self._ws[key] = _u8(nb, self.dev)
        return self._ws[key]

    def _buf(self, name: str, shape) -> torch.Tensor:
        key = (name, tuple(shape))
        if key not in self._bufs:
            self._bufs[key] = torch.empty(shape, dtype=torch.float32, device=self.dev)
        return self._bufs[key]

    def _ds(self):
        return d_param_struct(self.d_params, self.Dstack._buffers_list())

    def _pack_d(self):
        ps = self._ds()
        N.check(self.lib.afi_d_pack(self.ctx, self.prec, C.byref(ps), self.d_packed.data_ptr(), N.stream_ptr()))

    def _d_calls(self, xs, tags, save: bool, dlogits=None, stats_only=None, staged: bool = False):
        calls = (N.DCall * len(xs))()
        logits = []
        for i, (x, tag) in enumerate(zip(xs, tags)):
            n, _, h, w = x.shape
            # one workspace per (level, real / fake) serves both phases of a step (a forward-only call fits into the workspace of a
            # forward + backward call); `staged`: it still holds this very input in the library's layout from the D phase
            ws = self._ws_for("d", n, h, w, True, tag)
            calls[i].input_staged = int(staged)
            lg = self._buf("logit" + tag, (n, 1, h, w))
            c = calls[i]
            # a call whose output nobody reads (the reference's dead D0(hr) of the G phase) runs for its BatchNorm statistics only
            c.x, c.n, c.h, c.w, c.logits = N.view4(x), n, h, w, (None if stats_only and stats_only[i] else lg.data_ptr())
            c.ws, c.ws_bytes = ws.data_ptr(), ws.numel()
            if dlogits is not None:
                c.dlogits = dlogits[i].data_ptr()
            logits.append(lg)
        return calls, logits

    def _sub(self, calls, idxs):
        sub = (N.DCall * len(idxs))()
        for j, i in enumerate(idxs):
            sub[j] = calls[i]
        return sub

    def _d_phase(self, xs, tags, targets, loss_rows, save: bool, backward: bool, staged: bool = False):
        """Discriminator calls `xs` (reference order), BCE against `targets[i]` accumulated into losses[loss_rows[i], level]; with
        backward=True also d(BCE)/d(params) into the packed accumulator.  Calls with even / odd index (real / fake) form the two groups."""
        lib = self.lib
        calls, logits = self._d_calls(xs, tags, save, stats_only=[r is None and not backward for r in loss_rows], staged=staged)
        ps = self._ds()
        bn = self.Dstack[0][0].norm
        mom = 0.1 if bn.momentum is None else bn.momentum
        dls = [self._buf(f"dlogit{i}", tuple(lg.shape)) if backward else None for i, lg in enumerate(logits)]
        lp = self.losses.data_ptr()
        groups = [list(range(0, len(xs), 2)), list(range(1, len(xs), 2))] if self.overlap else [list(range(len(xs)))]
        cur = torch.cuda.current_stream()
        keep = []
        for gi, idxs in enumerate(groups):
            st = self.streams[gi] if self.overlap else cur
            if self.overlap:
                st.wait_stream(cur)
            with torch.cuda.stream(st):
                sp = C.c_void_p(st.cuda_stream)
                for o in range(0, len(idxs), N.MAX_CALLS):
                    part = idxs[o:o + N.MAX_CALLS]
                    sub = self._sub(calls, part)
                    keep.append(sub)
                    N.check(lib.afi_d_forward(self.ctx, self.prec, C.byref(ps), self.d_packed.data_ptr(), sub, len(part), 2, mom, bn.eps,
                                              int(save), sp))
                    for j, i in enumerate(part):
                        row = loss_rows[i]
                        if row is not None:
                            N.check(lib.afi_bce_with_logits(logits[i].data_ptr(), logits[i].numel(), targets[i], self._tmp[i:].data_ptr(),
                                                            lp + 4 * (row * 8 + i // 2), 1.0, N.ptr(dls[i]), 1.0, sp))
                        if backward:
                            sub[j].dlogits = dls[i].data_ptr()
                    if backward:
                        N.check(lib.afi_d_backward(self.ctx, self.prec, C.byref(ps), self.d_packed.data_ptr(), sub, len(part), 1,
                                                   self.d_acc.data_ptr(), sp))
        if self.overlap:
            for st in self.streams:
                cur.wait_stream(st)
        # BatchNorm running buffers: momentum updates in the reference's call order, after both groups are done
        for o in range(0, len(xs), N.MAX_CALLS):
            k = min(N.MAX_CALLS, len(xs) - o)
            N.check(lib.afi_d_update_running(self.ctx, self.prec, C.byref(ps), C.cast(C.byref(calls, o * C.sizeof(N.DCall)), C.POINTER(N.DCall)),
                                             k, mom, N.stream_ptr()))
        return logits

    def _sgd(self, params, grads, moms, is_norm):
        # parameters are updated in place behind torch's back; the packed GEMM-layout copies are refreshed explicitly by the caller.
        # ONE multi-tensor launch per optimiser; the host-side pointer tables are rebuilt only when a tensor moved.
        first = int(self.steps_done == 0 and not self._mom_restored)
        n = len(params)
        key = tuple(t.data_ptr() for ts in (params, grads, moms) for t in ts) + (self.wd, self.wd_norm)
        tab = self._sgd_tabs.get(id(params))
        if tab is None or tab[0] != key:
            P = (C.c_void_p * n)(*[p.data_ptr() for p in params])
            G = (C.c_void_p * n)(*[g.data_ptr() for g in grads])
            M = (C.c_void_p * n)(*[m.data_ptr() for m in moms])
            cnt = (C.c_longlong * n)(*[p.numel() for p in params])
            wd = (C.c_float * n)(*[self.wd_norm if nrm else self.wd for nrm in is_norm])
            tab = (key, P, G, M, cnt, wd)
            self._sgd_tabs[id(params)] = tab
        _, P, G, M, cnt, wd = tab
        N.check(self.lib.afi_sgd_step_multi(n, P, G, M, cnt, wd, self.lr, self.momentum, 1.0 / self.world, first, N.stream_ptr()))


class Stage1Step(_StepBase):
    def __init__(self, G, D, lr: float = 1e-3, momentum: float = 0.9, weight_decay: float = 1e-4, weight_decay_norm: float = 0.0,
                 precision: Optional[str] = None, process_group=None, distributed: Optional[bool] = None, overlap: bool = True,
                 reuse_g_forward: Optional[bool] = None, overlap_comm: bool = True):
        self.G = G
        g_params = G._params()
        self._init_common(g_params[0].device, lr, momentum, weight_decay, weight_decay_norm, precision or G.precision or N.default_precision(),
                          process_group, overlap, overlap_comm)
        dev = self.dev
        self.g_params = g_params
        self.n_rdb = G.n_residual_dense_blocks
        self.g_sync = FlatGradSync(self.g_params, process_group, distributed)
        self.g_flat, self.g_grads = self.g_sync.flat, self.g_sync.views
        self._init_d(D, distributed)
        self.g_mom = [torch.zeros_like(p) for p in self.g_params]
        self.g_acc = _u8(self.lib.afi_g_gradacc_bytes(self.n_rdb), dev)
        self.g_packed = _u8(self.lib.afi_g_packed_bytes(self.prec, self.n_rdb), dev)
        # The reference evaluates G(lr) twice per step -- detached for the D phase (:341-346), with a graph for the G phase (:390-395) --
        # with the SAME generator weights (G's optimiser only steps at the end of the G phase) and the same input, and G has no
        # normalisation, dropout or other state: the two evaluations are bit-identical.  One forward (kept for the backward) serves both.
        # AFIGAN_REUSE_G_FORWARD=0 restores the literal second evaluation.
        if reuse_g_forward is None:
            import os
            reuse_g_forward = os.environ.get("AFIGAN_REUSE_G_FORWARD", "1") != "0"
        self.reuse_g_forward = reuse_g_forward
        self._pack_key = None
        self._refresh_packed()

    # ---- helpers --------------------------------------------------------------------------------------
    def _gs(self):
        return g_param_struct(self.g_params, self.n_rdb)

    def _pack_g(self):
        ps = self._gs()
        N.check(self.lib.afi_g_pack(self.ctx, self.prec, C.byref(ps), self.g_packed.data_ptr(), N.stream_ptr()))

    # ---- grouped calls: ONE kernel launch per layer covers all pyramid levels (and real + fake for the discriminator)
    def _g_calls(self, lr_feats, hr_feats, save: bool, tag: str, dys=None):
        calls = (N.GCall * len(lr_feats))()
        outs = []
        for l, (lo, hi) in enumerate(zip(lr_feats, hr_feats)):
            n, _, h, w = lo.shape
            oh, ow = min(2 * h, hi.size(2)), min(2 * w, hi.size(3))          # _reshape_stage1: crop to the element-wise min
            ws = self._ws_for("g", n, h, w, save, l)
            y = self._buf(f"tr{tag}{l}", (n, CH, oh, ow))
            c = calls[l]
            c.x, c.n, c.h, c.w, c.y, c.oh, c.ow = N.view4(lo), n, h, w, y.data_ptr(), oh, ow
            c.ws, c.ws_bytes = ws.data_ptr(), ws.numel()
            if dys is not None:
                c.dy = N.view4(dys[l])
            outs.append(y)
        return calls, outs

    def _g_forward(self, lr_feats, hr_feats, save: bool, tag: str):
        calls, outs = self._g_calls(lr_feats, hr_feats, save, tag)
        ps = self._gs()
        for i in range(0, len(calls), N.MAX_CALLS):
            k = min(N.MAX_CALLS, len(calls) - i)
            N.check(self.lib.afi_g_forward(self.ctx, self.prec, C.byref(ps), self.g_packed.data_ptr(),
                                           C.cast(C.byref(calls, i * C.sizeof(N.GCall)), C.POINTER(N.GCall)), k, int(save), N.stream_ptr()))
        return outs

    def _g_backward(self, lr_feats, hr_feats, dys, tag: str):
        calls, _ = self._g_calls(lr_feats, hr_feats, True, tag, dys)
        ps = self._gs()
        for i in range(0, len(calls), N.MAX_CALLS):
            k = min(N.MAX_CALLS, len(calls) - i)
            N.check(self.lib.afi_g_backward(self.ctx, self.prec, C.byref(ps), self.g_packed.data_ptr(),
                                            C.cast(C.byref(calls, i * C.sizeof(N.GCall)), C.POINTER(N.GCall)), k, self.g_acc.data_ptr(),
                                            N.stream_ptr()))

    def _allreduce(self, flat: torch.Tensor):
        (self.g_sync if flat is self.g_flat else self.d_sync).all_reduce()

    # ---- the step ---------------------------------------------------------------------------------------
    @torch.no_grad()
    def run_step(self, lr_feats: Sequence[torch.Tensor], hr_feats: Sequence[torch.Tensor], apply_updates: bool = True,
                 timers: Optional[Dict[str, float]] = None):
        """Returns the device tensor `losses` [4, 8]: row 0 d_loss_p*, 1 g_loss_p*, 2 adv_loss_p*, 3 content_loss_p* (cols = levels).
        timers (a dict): filled with the device milliseconds of the step's phases (synchronises; instrumented steps only)."""
        lib, st = self.lib, N.stream_ptr
        nl = len(lr_feats)
        assert nl == len(hr_feats) and nl <= 8
        self._refresh_packed()
        self.losses.zero_()
        lp = self.losses.data_ptr()
        marks = []

        def mark(name):
            if timers is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))

        def slot(row, col):
            return lp + 4 * (row * 8 + col)

        overlap_comm = self.overlap_comm and self.distributed and self.reuse_g_forward and timers is None
        mark("start")
        # ------------------------------ D phase (stage1_trainer.py:334-381)
        N.check(lib.afi_zero(self.d_acc.data_ptr(), self.d_acc.numel(), st()))
        # G(lr).detach(), cropped
        trs = self._g_forward(lr_feats, hr_feats, True, "g") if self.reuse_g_forward else self._g_forward(lr_feats, hr_feats, False, "d")
        mark("g_forward")
        xs, tags = [], []
        for l, (tr, hi) in enumerate(zip(trs, hr_feats)):
            xs += [hi[:, :, :tr.size(2), :tr.size(3)], tr]                          # D0(hr) BEFORE D0(tr)  (:349-350)
            tags += [f"r{l}", f"f{l}"]
        # d_loss_l = BCE(D0(hr), 1) + BCE(D0(tr), 0); backward into D only
        self._d_phase(xs, tags, [1.0 if i % 2 == 0 else 0.0 for i in range(len(xs))], [0] * len(xs), True, True)
        gs = d_grad_struct(self.d_grads)
        N.check(lib.afi_d_unpack_grads(self.ctx, self.prec, self.d_acc.data_ptr(), C.byref(gs), 1.0, 0, st()))
        mark("d_forward_backward")

        def d_update():
            if apply_updates:
                self._d_update()

        def g_backward_pass():
            # L1(tr, hr) and its gradient; backward through G (reads the activations the ONE forward pass kept; does not read D)
            N.check(lib.afi_zero(self.g_acc.data_ptr(), self.g_acc.numel(), st()))
            dtrs = []
            for l, (tr, hi) in enumerate(zip(trs, hr_feats)):
                dtr = self._buf(f"dtr{l}", tuple(tr.shape))
                hi_c = hi[:, :, :tr.size(2), :tr.size(3)]
                N.check(lib.afi_l1_loss(N.view4(tr), N.view4(hi_c), tr.size(0), CH, tr.size(2), tr.size(3), self._tmp[40 + l:].data_ptr(),
                                        slot(3, l), 1.0, dtr.data_ptr(), 1.0, st()))
                dtrs.append(dtr)
            self._g_backward(lr_feats, hr_feats, dtrs, "g")
            gsg = g_param_struct(self.g_grads, self.n_rdb)
            N.check(lib.afi_g_unpack_grads(self.ctx, self.prec, self.g_acc.data_ptr(), C.byref(gsg), 1.0, 0, st()))

        def g_phase_d_forwards(trs_):
            xs_, tags_ = [], []
            for l, (tr, hi) in enumerate(zip(trs_, hr_feats)):
                xs_ += [tr, hi[:, :, :tr.size(2), :tr.size(3)]]                     # D0(tr) BEFORE D0(hr) here (:399-400)
                tags_ += [f"f{l}", f"r{l}"]
            # adv = BCE(D0(tr).detach(), 1): no gradient (:399); D0(hr) is dead compute kept for its BN running-stat side effect (:400)
            # (with one G forward per step the G phase feeds the discriminator the tensors the D phase already staged)
            self._d_phase(xs_, tags_, [1.0] * len(xs_), [2 if i % 2 == 0 else None for i in range(len(xs_))], False, False,
                          staged=self.reuse_g_forward)

        if overlap_comm:
            # Same values as the literal order below (the generator's backward does not depend on D's update, the G phase's discriminator
            # forwards do not depend on G's gradients); only the ISSUE order changes so that each all-reduce has compute to hide behind.
            self.d_sync.start()
            g_backward_pass()
            self.g_sync.start()
            self.d_sync.finish()
            d_update()
            g_phase_d_forwards(trs)
            self.g_sync.finish()
        else:
            self._allreduce(self.d_flat)
            d_update()
            mark("d_allreduce_sgd_pack")
            # ------------------------------ G phase (stage1_trainer.py:384-433)
            if not self.reuse_g_forward:
                trs = self._g_forward(lr_feats, hr_feats, True, "g")
            g_phase_d_forwards(trs)
            mark("g_phase_d_forwards")
            g_backward_pass()
            mark("l1_g_backward")
            self._allreduce(self.g_flat)
        self.losses[1, :nl] = 1e-3 * self.losses[2, :nl] + self.losses[3, :nl]
        if apply_updates:
            self._sgd(self.g_params, self.g_grads, self.g_mom, [False] * len(self.g_params))
            self._pack_g()
            self.G._native.packed.key = None
            self.steps_done += 1
        mark("g_allreduce_sgd_pack")
        if timers is not None:
            torch.cuda.synchronize(self.dev)
            for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
                timers[name] = timers.get(name, 0.0) + a.elapsed_time(b)
        return self.losses

    def metrics(self, n_levels: int = 5) -> Dict[str, float]:
        """Host copy of the per-level losses under the reference's metric names (stage1_trainer.py:357,409)."""
        v = self.losses.cpu()
        out = {}
        for l in range(n_levels):
            out[f"d_loss_p{l + 2}"] = float(v[0, l])
            out[f"g_loss_p{l + 2}"] = float(v[1, l])
            out[f"adv_loss_p{l + 2}"] = float(v[2, l])
            out[f"content_loss_p{l + 2}"] = float(v[3, l])
        return out


class FeaturePrefetcher:
    """Host -> device staging of the step's feature pyramids (pinned host tensors) on a dedicated copy stream, double-buffered, so the H2D
    copy of batch i+1 overlaps the compute of batch i (what a DataLoader with pin_memory + non_blocking copies gives the reference trainer,
    rcnn_only.py:37).  `next()` returns device tensors that are safe to use on the current stream."""

    def __init__(self, device):
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.slots = [None, None]
        self.events = [torch.cuda.Event(), torch.cuda.Event()]
        self.i = 0

    def submit(self, host_tensors):
        """Start copying a batch (list of pinned CPU tensors) into the free slot."""
        k = self.i & 1
        self.stream.wait_stream(torch.cuda.current_stream())          # the slot's previous contents must have been consumed
        with torch.cuda.stream(self.stream):
            if self.slots[k] is None:
                self.slots[k] = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host_tensors]
            for d, h in zip(self.slots[k], host_tensors):
                d.copy_(h, non_blocking=True)
            self.events[k].record(self.stream)
        self.i += 1
        return k

    def get(self, k):
        torch.cuda.current_stream().wait_event(self.events[k])
        return self.slots[k]
