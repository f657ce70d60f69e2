"""Executable model of the CUDA kernel SEQUENCE (test infrastructure).

The library (afi-gan_b200/csrc) evaluates the hot path as a chain of two GEMM primitives over
pixel-major (NHWC) views plus a few elementwise passes:

  conv_taps : out[n,y,x,co]  = sum_taps  W[slab][co][:] . in[view][n, y+dy, x+dx, :]      (zero outside HxW)
  wgrad_taps: dW[slab][co][ci] += sum_p dY[p][co] * X[p + (dy,dx)][ci]

This file states that chain in plain torch so that the decomposition identities the kernels rely on
(transposed conv = 4 phase convs; dgrad = conv with flipped/transposed slabs; deconv dgrad = one
36-tap conv over 4 phase views; dense-block backward with the growth convs' dgrads split by destination and their
weight gradients in one GEMM; train-mode BatchNorm backward in closed form; the head backward's mask from bn3(z3);
the 1024->1 logit conv as 9 per-pixel dot products + a 3x3 shift-sum) are checked against the oracle's
autograd on the CPU (tests/test_kernel_model.py).  csrc/api.cu is a transliteration of the
g_forward / g_backward / d_forward / d_backward functions below.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

S = 0.2  # LeakyReLU slope == residual scale

TAPS9 = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]


def lrelu(t):
    return torch.where(t > 0, t, S * t)


def lmask(a):
    """d lrelu / d pre-activation recovered from the (in-place) activation output: sign is preserved."""
    return torch.where(a > 0, torch.ones_like(a), torch.full_like(a, S))


def shift(v: torch.Tensor, dy: int, dx: int) -> torch.Tensor:
    """v[n, y+dy, x+dx, :] with zeros outside the grid."""
    n, h, w, c = v.shape
    out = torch.zeros_like(v)
    ys, ye = max(0, -dy), min(h, h - dy)
    xs, xe = max(0, -dx), min(w, w - dx)
    if ys < ye and xs < xe:
        out[:, ys:ye, xs:xe] = v[:, ys + dy:ye + dy, xs + dx:xe + dx]
    return out


def conv_taps(views: Sequence[torch.Tensor], taps, w: torch.Tensor, cin: int) -> torch.Tensor:
    """taps: list of (dy, dx, view, slab); w: [nslab, cout, cin]."""
    acc = None
    for dy, dx, vi, sl in taps:
        t = shift(views[vi][..., :cin], dy, dx) @ w[sl].transpose(0, 1)
        acc = t if acc is None else acc + t
    return acc


def wgrad_taps(dy_view: torch.Tensor, x_view: torch.Tensor, taps2, cin: int) -> torch.Tensor:
    """returns [ntaps, cout, cin]; taps2: list of (dy, dx)."""
    co = dy_view.shape[-1]
    g = dy_view.reshape(-1, co)
    return torch.stack([g.transpose(0, 1) @ shift(x_view[..., :cin], dy, dx).reshape(-1, cin) for dy, dx in taps2])


# ---- weight packing (torch layouts -> [slab][cout][cin]) -------------------------------------------------
def pack_fwd(w):      # [co,ci,3,3] -> slab t=(dy+1)*3+(dx+1): [co,ci]
    return w.permute(2, 3, 0, 1).reshape(9, w.shape[0], w.shape[1])


def pack_dgrad(w):    # conv of dOut with taps (e,f): slab[(e+1)*3+(f+1)][ci][co] = w[co][ci][1-e][1-f]
    return w.flip(2, 3).permute(2, 3, 1, 0).reshape(9, w.shape[1], w.shape[0])


def unpack_wgrad(dw):  # [9,co,ci] -> [co,ci,3,3]
    return dw.reshape(3, 3, dw.shape[1], dw.shape[2]).permute(2, 3, 0, 1)


def deconv_ky(a, dy):  # kernel row used by output parity a for input offset dy
    return 2 * (1 - dy) + a


def pack_deconv_fwd(w):   # w [ci,co,6,6] -> [4 phases * 9, co, ci]
    out = []
    for a in (0, 1):
        for b in (0, 1):
            for dy, dx in TAPS9:
                out.append(w[:, :, deconv_ky(a, dy), deconv_ky(b, dx)].transpose(0, 1))
    return torch.stack(out)


def pack_deconv_dgrad(w):  # 36 taps (phase view ab, offset e,f): slab[ci_out][co_in] = w[ci][co][2(1+e)+a][2(1+f)+b]
    out = []
    for a in (0, 1):
        for b in (0, 1):
            for e, f in TAPS9:
                out.append(w[:, :, 2 * (1 + e) + a, 2 * (1 + f) + b])
    return torch.stack(out)


def unpack_deconv_wgrad(dw):  # [36, co, ci] (phase-major, taps (dy,dx)) -> [ci,co,6,6]
    co, ci = dw.shape[1], dw.shape[2]
    out = torch.zeros(ci, co, 6, 6, dtype=dw.dtype)
    i = 0
    for a in (0, 1):
        for b in (0, 1):
            for dy, dx in TAPS9:
                out[:, :, deconv_ky(a, dy), deconv_ky(b, dx)] = dw[i].transpose(0, 1)
                i += 1
    return out


def std_taps(view=0, slab0=0):
    return [(dy, dx, view, slab0 + i) for i, (dy, dx) in enumerate(TAPS9)]


def bilinear2x_nhwc(f):
    def up(t, dim):
        n = t.size(dim)
        idx = torch.arange(n)
        prev = t.index_select(dim, (idx - 1).clamp(min=0))
        nxt = t.index_select(dim, (idx + 1).clamp(max=n - 1))
        o = torch.stack((0.25 * prev + 0.75 * t, 0.75 * t + 0.25 * nxt), dim=dim + 1)
        shape = list(t.shape)
        shape[dim] = 2 * n
        return o.reshape(shape)
    return up(up(f, 1), 2)


# ---- Generator ---------------------------------------------------------------------------------------------
def g_forward(sd: Dict[str, torch.Tensor], x_nchw: torch.Tensor, n_rdb: int = 3):
    """Returns (y_nchw, saved).  Buffers mirror the CUDA workspace: X0, B[r] (384 ch), H1, H2, H3."""
    X0 = x_nchw.permute(0, 2, 3, 1).contiguous()
    N, H, W, C = X0.shape
    B = [torch.zeros(N, H, W, C + 128, dtype=X0.dtype) for _ in range(n_rdb)]
    B[0][..., :C] = lrelu(conv_taps([X0], std_taps(), pack_fwd(sd["Generators.0.0.0.weight"]), C) + sd["Generators.0.0.0.bias"])
    H1 = None
    for r in range(n_rdb):
        p = f"Generators.0.1.RDBs.{r}."
        for i in range(1, 5):
            cin = C + 32 * (i - 1)
            B[r][..., cin:cin + 32] = lrelu(conv_taps([B[r]], std_taps(), pack_fwd(sd[p + f"conv{i}.0.weight"]), cin))
        acc = conv_taps([B[r]], std_taps(), pack_fwd(sd[p + "conv5.weight"]), C + 128)
        if r + 1 < n_rdb:
            B[r + 1][..., :C] = B[r][..., :C] + S * acc
        else:
            H1 = S * S * acc + S * B[r][..., :C] + B[0][..., :C]
    H2 = lrelu(conv_taps([H1], std_taps(), pack_fwd(sd["Generators.0.2.0.weight"]), C) + sd["Generators.0.2.0.bias"])
    H3 = torch.zeros(N, 2 * H, 2 * W, C, dtype=X0.dtype)
    Wd = pack_deconv_fwd(sd["Generators.0.3.0.weight"])
    for ph, (a, b) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        H3[:, a::2, b::2] = lrelu(conv_taps([H2], std_taps(0, 9 * ph), Wd, C) + sd["Generators.0.3.0.bias"])
    Yb = conv_taps([H3], std_taps(), pack_fwd(sd["Generators.0.4.0.weight"]), C) + sd["Generators.0.4.0.bias"]
    y = (Yb + bilinear2x_nhwc(X0)).permute(0, 3, 1, 2).contiguous()
    return y, dict(X0=X0, B=B, H1=H1, H2=H2, H3=H3)


def g_backward(sd, saved, dy_nchw: torch.Tensor, n_rdb: int = 3, need_dx: bool = False):
    """dy_nchw is the gradient wrt the (possibly cropped) output; zero-extended to 2Hx2W."""
    X0, B, H1, H2, H3 = saved["X0"], saved["B"], saved["H1"], saved["H2"], saved["H3"]
    N, H, W, C = X0.shape
    G0 = torch.zeros(N, 2 * H, 2 * W, C, dtype=X0.dtype)
    oh, ow = dy_nchw.shape[2:]
    G0[:, :oh, :ow] = dy_nchw.permute(0, 2, 3, 1)
    grads = {}
    # output conv
    grads["Generators.0.4.0.weight"] = unpack_wgrad(wgrad_taps(G0, H3, TAPS9, C))
    grads["Generators.0.4.0.bias"] = G0.sum((0, 1, 2))
    G1 = conv_taps([G0], std_taps(), pack_dgrad(sd["Generators.0.4.0.weight"]), C) * lmask(H3)
    # transposed conv: wgrad per phase; dgrad = ONE conv with 36 taps over the 4 phase views of G1
    phases = ((0, 0), (0, 1), (1, 0), (1, 1))
    views = [G1[:, a::2, b::2] for a, b in phases]
    dwd = torch.cat([wgrad_taps(v, H2, TAPS9, C) for v in views])
    grads["Generators.0.3.0.weight"] = unpack_deconv_wgrad(dwd)
    grads["Generators.0.3.0.bias"] = G1.sum((0, 1, 2))
    taps36 = [(e, f, ph, ph * 9 + i) for ph in range(4) for i, (e, f) in enumerate(TAPS9)]
    G2 = conv_taps(views, taps36, pack_deconv_dgrad(sd["Generators.0.3.0.weight"]), C) * lmask(H2)
    # post conv
    grads["Generators.0.2.0.weight"] = unpack_wgrad(wgrad_taps(G2, H1, TAPS9, C))
    grads["Generators.0.2.0.bias"] = G2.sum((0, 1, 2))
    dH1 = conv_taps([G2], std_taps(), pack_dgrad(sd["Generators.0.2.0.weight"]), C)
    # RiR: h1 = 0.2 * x3 + h0
    d_out = S * dH1
    for r in reversed(range(n_rdb)):
        p = f"Generators.0.1.RDBs.{r}."
        dc5 = S * d_out
        grads[p + "conv5.weight"] = unpack_wgrad(wgrad_taps(dc5, B[r], TAPS9, C + 128))
        GA = conv_taps([dc5], std_taps(), pack_dgrad(sd[p + "conv5.weight"]), C)      # [.., 384]
        GA[..., :C] += d_out
        # The four masked growth gradients sit side by side in ONE 128-channel buffer GC.  Their dgrads are split by destination:
        #   * the part that reaches the EARLIER growth channels [C, C + 32 (i-1)) feeds the remaining masks: three small convs, in order;
        #   * the part that reaches the block input x [0, C) is one conv with gemm-cin = 128 (all four at once), at the end.
        # Their weight gradients are one [9][128][C + 96] GEMM whose non-causal blocks are ignored.
        GC = torch.zeros(N, H, W, 128, dtype=X0.dtype)
        for i in (4, 3, 2, 1):
            cin = C + 32 * (i - 1)
            GC[..., 32 * (i - 1):32 * i] = GA[..., cin:cin + 32] * lmask(B[r][..., cin:cin + 32])
            if i > 1:
                wd = pack_dgrad(sd[p + f"conv{i}.0.weight"])                    # [9][cin][32]
                GA[..., C:cin] += conv_taps([GC[..., 32 * (i - 1):32 * i]], std_taps(), wd[:, C:cin, :], 32)
        wx = torch.cat([pack_dgrad(sd[p + f"conv{i}.0.weight"])[:, :C, :] for i in (1, 2, 3, 4)], dim=2)   # [9][C][128]
        GA[..., :C] += conv_taps([GC], std_taps(), wx, 128)
        dw_all = wgrad_taps(GC, B[r], TAPS9, C + 96)                             # [9][128][C + 96]
        for i in (1, 2, 3, 4):
            grads[p + f"conv{i}.0.weight"] = unpack_wgrad(dw_all[:, 32 * (i - 1):32 * i, :C + 32 * (i - 1)])
        d_out = GA[..., :C]
    g_head = (d_out + dH1) * lmask(B[0][..., :C])
    grads["Generators.0.0.0.weight"] = unpack_wgrad(wgrad_taps(g_head, X0, TAPS9, C))
    grads["Generators.0.0.0.bias"] = g_head.sum((0, 1, 2))
    dx = None
    if need_dx:
        dX = conv_taps([g_head], std_taps(), pack_dgrad(sd["Generators.0.0.0.weight"]), C)
        dx = dX.permute(0, 3, 1, 2) + bilinear2x_transpose(G0.permute(0, 3, 1, 2))
    return grads, dx


def bilinear2x_transpose(g):
    """Adjoint of the x2 bilinear upsample (align_corners=False) on NCHW g [N,C,2H,2W] -> [N,C,H,W]."""
    def down(t, dim):
        n = t.size(dim) // 2
        ev = t.index_select(dim, torch.arange(0, 2 * n, 2))
        od = t.index_select(dim, torch.arange(1, 2 * n, 2))
        out = 0.75 * ev + 0.75 * od
        # out[i] also receives 0.25*even[i+1] (as 'prev' of i+1) and 0.25*odd[i-1] (as 'next' of i-1); edges clamp onto themselves
        sl = [slice(None)] * t.dim()
        a, b = list(sl), list(sl)
        a[dim], b[dim] = slice(0, n - 1), slice(1, n)
        out[tuple(a)] += 0.25 * ev[tuple(b)]
        out[tuple(b)] += 0.25 * od[tuple(a)]
        first, last = list(sl), list(sl)
        first[dim], last[dim] = slice(0, 1), slice(n - 1, n)
        out[tuple(first)] += 0.25 * ev[tuple(first)]
        out[tuple(last)] += 0.25 * od[tuple(last)]
        return out
    return down(down(g, 2), 3)


# ---- Discriminator -----------------------------------------------------------------------------------------
def d_forward(sd, x_nchw, eps=1e-5):
    A = [x_nchw.permute(0, 2, 3, 1).contiguous()]
    Z, stats = [], []
    for n in range(3):
        p = f"Discriminators.0.{n}.0."
        z = conv_taps([A[n]], std_taps(), pack_fwd(sd[p + "weight"]), A[n].shape[-1]) + sd[p + "bias"]
        M = z.numel() // z.shape[-1]
        s1, s2 = z.sum((0, 1, 2)), (z * z).sum((0, 1, 2))          # bn_stats kernel (double accumulators)
        mean = s1 / M
        var = s2 / M - mean * mean                                  # biased
        rstd = (var + eps).rsqrt()
        Z.append(z)
        stats.append((mean, rstd, var * M / max(M - 1, 1)))
        A.append(lrelu((z - mean) * rstd * sd[p + "norm.weight"] + sd[p + "norm.bias"]))
    w4 = sd["Discriminators.0.3.0.weight"][0]                       # [1024,3,3]
    t9 = A[3] @ w4.reshape(w4.shape[0], 9)                          # per-pixel dot products, tap = ky*3+kx
    logit = sd["Discriminators.0.3.0.bias"] + sum(shift(t9[..., i:i + 1], dy, dx) for i, (dy, dx) in enumerate(TAPS9))
    return logit.permute(0, 3, 1, 2), dict(A=A, Z=Z, stats=stats)


def d_backward(sd, saved, dlogit_nchw):
    A, Z, stats = saved["A"], saved["Z"], saved["stats"]
    g = dlogit_nchw.permute(0, 2, 3, 1)                              # [N,H,W,1]
    grads = {}
    w4 = sd["Discriminators.0.3.0.weight"][0]
    C3 = w4.shape[0]
    gs = torch.cat([shift(g, -dy, -dx) for dy, dx in TAPS9], -1)     # gs[q][tap] = g[q - tap]
    grads["Discriminators.0.3.0.weight"] = torch.stack(
        [(shift(A[3], dy, dx) * g).sum((0, 1, 2)) for dy, dx in TAPS9], -1).reshape(1, C3, 3, 3)
    grads["Discriminators.0.3.0.bias"] = g.sum().reshape(1)
    dA = gs @ w4.reshape(C3, 9).transpose(0, 1)
    for n in (2, 1, 0):
        p = f"Discriminators.0.{n}.0."
        mean, rstd, _ = stats[n]
        gamma = sd[p + "norm.weight"]
        xhat = (Z[n] - mean) * rstd
        if n == 2:
            # layer 3 (tensor-core engine): dy3 is never stored.  Two passes over z3 recompute the 9-tap product gs @ w4^T, and the
            # LeakyReLU mask comes from bn3(z3) (same sign as the stored activation), not from A[3]
            dyv = dA * lmask(xhat * gamma + sd[p + "norm.bias"])
        else:
            dyv = dA * lmask(A[n + 1])
        M = dyv.numel() // dyv.shape[-1]
        s_dy, s_dyx = dyv.sum((0, 1, 2)), (dyv * xhat).sum((0, 1, 2))
        grads[p + "norm.weight"], grads[p + "norm.bias"] = s_dyx, s_dy
        dz = gamma * rstd * (dyv - s_dy / M - xhat * s_dyx / M)
        cin = A[n].shape[-1]
        grads[p + "weight"] = unpack_wgrad(wgrad_taps(dz, A[n], TAPS9, cin))
        grads[p + "bias"] = dz.sum((0, 1, 2))
        if n > 0:
            dA = conv_taps([dz], std_taps(), pack_dgrad(sd[p + "weight"]), dz.shape[-1])
    return grads
