// HBM-bound passes of the AFI-GAN hot path: layout changes, weight (un)packing, BatchNorm statistics /
// apply / backward, the 1024->1 discriminator head, the BCE / L1 losses and the SGD update.
// All kernels are coalesced along the channel (innermost NHWC) dimension, 4 channels per thread,
// reductions are warp/block-reduced before one atomic per channel per block.
#include <string.h>
#include "common.cuh"

namespace afi {

// ---------------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld4(const void* base, long long off, int dt) {
    if (dt == DT_F32) return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
    uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(base) + off);
    float4 r;
    r.x = __uint_as_float(u.x << 16); r.y = __uint_as_float(u.x & 0xffff0000u);
    r.z = __uint_as_float(u.y << 16); r.w = __uint_as_float(u.y & 0xffff0000u);
    return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st4(void* base, long long off, int dt, float4 v) {
    if (dt == DT_F32) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off) = v;
    } else {
        uint2 u; u.x = pack_bf16x2(v.x, v.y); u.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(base) + off) = u;
    }
}
__device__ __forceinline__ float lmaskf(float a, float slope) { return a > 0.f ? 1.f : slope; }

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ double block_reduce_sum(double v) {
    __shared__ double sh[32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    if (warp == 0) for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// decode a flat (pixel, channel-quad) index
struct PixIdx { int n, y, x; };
__device__ __forceinline__ PixIdx decode_pixel(long long p, int H, int W) {
    PixIdx r; r.x = (int)(p % W); long long t = p / W; r.y = (int)(t % H); r.n = (int)(t / H); return r;
}
__device__ __forceinline__ long long voff(const PView& v, PixIdx q) { return q.n * v.sn + q.y * v.sy + q.x * v.sx; }

// ---------------------------------------------------------------------------------------------------
// NCHW fp32 (strided) -> NHWC T
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_nchw_to_nhwc(afi_view4 src, int c, int h, int w, PView dst) {
    __shared__ float tile[32][33];
    int x0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    int n = blockIdx.z / h, y = blockIdx.z % h;
    int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int cc = c0 + ty + 8 * i, xx = x0 + tx;
        float v = 0.f;
        if (cc < c && xx < w) v = vld(src, n * src.sn + cc * src.sc + y * src.sh + xx * src.sw);
        tile[ty + 8 * i][tx] = v;
    }
    __syncthreads();
    T* d = reinterpret_cast<T*>(dst.ptr);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int xx = x0 + ty + 8 * i, cc = c0 + tx;
        if (cc < c && xx < w) d[n * dst.sn + y * dst.sy + xx * dst.sx + cc] = (T)tile[tx][ty + 8 * i];
    }
}
// bf16 destination with an even channel count: 64-channel x 32-pixel tiles, so that the NHWC side moves 4 bytes per lane = 128 bytes per
// warp row (2-byte lanes reach 43 % of the copy bandwidth, profiles/r01_hbm_passes.md)
__global__ void __launch_bounds__(256) k_nchw_to_nhwc_bf16x2(afi_view4 src, int c, int h, int w, PView dst) {
    __shared__ float tile[32][65];      // [x][c]
    const int x0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
    const int n = blockIdx.z / h, y = blockIdx.z % h;
    const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int cc = c0 + ty + 8 * i, xx = x0 + tx;
        float v = 0.f;
        if (cc < c && xx < w) v = vld(src, n * src.sn + cc * src.sc + y * src.sh + xx * src.sw);
        tile[tx][ty + 8 * i] = v;
    }
    __syncthreads();
    bf16* d = reinterpret_cast<bf16*>(dst.ptr);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int xl = ty + 8 * i, xx = x0 + xl, cc = c0 + 2 * tx;
        if (cc < c && xx < w)
            *reinterpret_cast<uint32_t*>(d + n * dst.sn + y * dst.sy + xx * dst.sx + cc) = pack_bf16x2(tile[xl][2 * tx], tile[xl][2 * tx + 1]);
    }
}
// channels_last source (sc == 1: the view is NHWC already): a straight copy / conversion, four channels per thread
template <typename T, typename S>
__global__ void __launch_bounds__(256) k_nhwc_view_to_nhwc(afi_view4 src, int c4, int h, int w, long long total, PView dst) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int cc = (int)(i % c4) * 4;
    const long long p = i / c4;
    const int x = (int)(p % w), y = (int)((p / w) % h), n = (int)(p / ((long long)w * h));
    const S* sp = reinterpret_cast<const S*>(src.ptr) + n * src.sn + y * src.sh + x * src.sw + cc;
    float v[4];
    if constexpr (sizeof(S) == 4) {
        const float4 t = *reinterpret_cast<const float4*>(sp);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        const uint2 t = *reinterpret_cast<const uint2*>(sp);
        v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u); v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
    }
    T* dp = reinterpret_cast<T*>(dst.ptr) + n * dst.sn + y * dst.sy + x * dst.sx + cc;
    if constexpr (sizeof(T) == 4) *reinterpret_cast<float4*>(dp) = make_float4(v[0], v[1], v[2], v[3]);
    else *reinterpret_cast<uint2*>(dp) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
}
template <typename T>
int nchw_to_nhwc(afi_view4 src, int n, int c, int h, int w, PView dst, cudaStream_t st) {
    AFI_REQUIRE(view_dtype_ok(src), "layout conversion: unknown view dtype %d", src.dtype);
    {
        const int ses = src.dtype == AFI_DT_BF16 ? 2 : 4;
        const uintptr_t salign = (uintptr_t)(4 * ses) - 1, dalign = (uintptr_t)(4 * sizeof(T)) - 1;
        if (src.sc == 1 && c % 4 == 0 && src.sn % 4 == 0 && src.sh % 4 == 0 && src.sw % 4 == 0 && ((uintptr_t)src.ptr & salign) == 0 &&
            dst.sn % 4 == 0 && dst.sy % 4 == 0 && dst.sx % 4 == 0 && ((uintptr_t)dst.ptr & dalign) == 0) {
            const long long total = (long long)n * h * w * (c / 4);
            const int grid = (int)cdiv(total, 256);
            if (src.dtype == AFI_DT_BF16) k_nhwc_view_to_nhwc<T, bf16><<<grid, 256, 0, st>>>(src, c / 4, h, w, total, dst);
            else k_nhwc_view_to_nhwc<T, float><<<grid, 256, 0, st>>>(src, c / 4, h, w, total, dst);
            AFI_LAUNCH_CHECK();
            return AFI_OK;
        }
    }
    AFI_REQUIRE((long long)n * h <= 65535, "layout conversion: n * h = %lld exceeds the 65535-block grid limit (split the batch)", (long long)n * h);
    if (dt_of<T>::v == DT_BF16 && c % 2 == 0 && dst.sx % 2 == 0 && dst.sy % 2 == 0 && dst.sn % 2 == 0 && ((uintptr_t)dst.ptr & 3) == 0) {
        dim3 grid(cdiv(w, 32), cdiv(c, 64), n * h), block(32, 8);
        k_nchw_to_nhwc_bf16x2<<<grid, block, 0, st>>>(src, c, h, w, dst);
        AFI_LAUNCH_CHECK();
        return AFI_OK;
    }
    dim3 grid(cdiv(w, 32), cdiv(c, 32), n * h), block(32, 8);
    k_nchw_to_nhwc<T><<<grid, block, 0, st>>>(src, c, h, w, dst);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
template int nchw_to_nhwc<float>(afi_view4, int, int, int, int, PView, cudaStream_t);
template int nchw_to_nhwc<bf16>(afi_view4, int, int, int, int, PView, cudaStream_t);

// ---------------------------------------------------------------------------------------------------
// BiFPN neck at inference (reference bifpn_sr.py:583-729, bifpn_layers/wrappers.py:166-252): the HBM-bound halves of its two building blocks.
//   k_dw3x3_to_nhwc : dst[n,y,x,c] (NHWC T) = sum_{3x3} w[c][ky][kx] * act(src[n,c,y+ky-1,x+kx-1]), act = swish when pre_swish (zero padding =
//                     'static_same' for k3 s1).  The depthwise half of a SeparableConv2d fused with the swish in front of it and with the
//                     layout conversion the pointwise GEMM needs anyway: one pass instead of three.
//   k_bifpn_fuse_down : out = w0 * a + w1 * b [+ w2 * maxpool3x3s2(c)], the bottom-up fusion site; the max-pool ZERO-pads right / bottom and the
//                     padding takes part in the max ('static_same' of the reference's MaxPool2d wrapper).
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_dw3x3_to_nhwc(afi_view4 src, const float* __restrict__ wdw, int c, int h, int w, int pre_swish, PView dst) {
    __shared__ float tile[32][33];
    const int x0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int n = blockIdx.z / h, y = blockIdx.z % h;
    const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int cc = c0 + ty + 8 * i, xx = x0 + tx;
        float acc = 0.f;
        if (cc < c && xx < w) {
            const long long sp = n * src.sn + cc * src.sc;
            const float* wk = wdw + cc * 9;
#pragma unroll
            for (int ky = 0; ky < 3; ky++) {
                const int yy = y + ky - 1;
                if (yy < 0 || yy >= h) continue;
#pragma unroll
                for (int kx = 0; kx < 3; kx++) {
                    const int xs = xx + kx - 1;
                    if (xs < 0 || xs >= w) continue;
                    float v = vld(src, sp + yy * src.sh + xs * src.sw);
                    if (pre_swish) v = v / (1.f + __expf(-v));
                    acc += __ldg(wk + ky * 3 + kx) * v;
                }
            }
        }
        tile[ty + 8 * i][tx] = acc;
    }
    __syncthreads();
    T* d = reinterpret_cast<T*>(dst.ptr);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int xx = x0 + ty + 8 * i, cc = c0 + tx;
        if (cc < c && xx < w) d[n * dst.sn + y * dst.sy + xx * dst.sx + cc] = (T)tile[tx][ty + 8 * i];
    }
}
template <typename T>
int dw3x3_to_nhwc(afi_view4 src, const float* wdw, int n, int c, int h, int w, int pre_swish, PView dst, cudaStream_t st) {
    AFI_REQUIRE(view_dtype_ok(src), "dw3x3: unknown view dtype %d", src.dtype);
    AFI_REQUIRE((long long)n * h <= 65535, "dw3x3_to_nhwc: n * h = %lld exceeds the grid limit", (long long)n * h);
    dim3 grid(cdiv(w, 32), cdiv(c, 32), n * h), block(32, 8);
    k_dw3x3_to_nhwc<T><<<grid, block, 0, st>>>(src, wdw, c, h, w, pre_swish, dst);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
template int dw3x3_to_nhwc<float>(afi_view4, const float*, int, int, int, int, int, PView, cudaStream_t);
template int dw3x3_to_nhwc<bf16>(afi_view4, const float*, int, int, int, int, int, PView, cudaStream_t);

__global__ void __launch_bounds__(256) k_bifpn_fuse_down(afi_view4 a, afi_view4 b, afi_view4 dn, const float* __restrict__ wts, int nw, int c, int h, int w,
                                                         int dh, int dw_, float* __restrict__ out, long long total) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % w), y = (int)((i / w) % h), cc = (int)((i / ((long long)w * h)) % c), n = (int)(i / ((long long)w * h * c));
    // max over the 3x3 window at stride 2 of `dn` padded with ZEROS on the right / bottom (F.pad(x, (0, 1, 0, 1)) then max_pool2d(3, 2))
    const long long dp = n * dn.sn + cc * dn.sc;
    float m = -3.402823466e38f;
#pragma unroll
    for (int ky = 0; ky < 3; ky++)
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
            const int yy = 2 * y + ky, xx = 2 * x + kx;
            const float v = (yy < dh && xx < dw_) ? vld(dn, dp + yy * dn.sh + xx * dn.sw) : 0.f;
            m = fmaxf(m, v);
        }
    float r;
    const float va = vld(a, n * a.sn + cc * a.sc + y * a.sh + x * a.sw);
    if (b.ptr) {
        const float vb = vld(b, n * b.sn + cc * b.sc + y * b.sh + x * b.sw);
        r = wts ? wts[0] * va + wts[1] * vb + wts[2] * m : va + vb + m;
    } else {
        r = wts ? wts[0] * va + wts[1] * m : va + m;
    }
    (void)nw;
    out[i] = r;
}
int bifpn_fuse_down(afi_view4 a, afi_view4 b, afi_view4 dn, const float* wts, int nw, int n, int c, int h, int w, int dh, int dw_, float* out,
                    cudaStream_t st) {
    const long long total = (long long)n * c * h * w;
    if (total == 0) return AFI_OK;
    AFI_REQUIRE(view_dtype_ok(a) && view_dtype_ok(b) && view_dtype_ok(dn), "bifpn_fuse_down: unknown view dtype");
    k_bifpn_fuse_down<<<cdiv(total, 256), 256, 0, st>>>(a, b, dn, wts, nw, c, h, w, dh, dw_, out, total);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}

// ---------------------------------------------------------------------------------------------------
// NHWC T -> contiguous NCHW fp32, fused with the bilinear x2 skip (generator_rdb.py:125,130), the lateral
// add and the merge scale (fpn_sr.py:154-157), and the top-left crop (stage1_trainer.py:437-443).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilin_coord(int o, int size, int& i0, int& i1, float& l0, float& l1) {
    // area_pixel_compute_source_index(scale=0.5, align_corners=False) as ATen's upsample_bilinear2d does
    float s = (o + 0.5f) * 0.5f - 0.5f;
    if (s < 0.f) s = 0.f;
    i0 = (int)s;
    i1 = i0 + (i0 < size - 1 ? 1 : 0);
    l1 = s - (float)i0;
    l0 = 1.f - l1;
}
template <typename T>
__global__ void k_nhwc_to_nchw(PView a, PView lat, afi_view4 skip, int sh, int sw_, float scale, int c, int oh, int ow,
                               float* __restrict__ dst, afi_view4 add, const float* __restrict__ fw) {
    __shared__ float tile[32][33];
    int x0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    int n = blockIdx.z / oh, y = blockIdx.z % oh;
    int tx = threadIdx.x, ty = threadIdx.y;
    const T* ap = reinterpret_cast<const T*>(a.ptr);
    const T* lp = reinterpret_cast<const T*>(lat.ptr);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int xx = x0 + ty + 8 * i, cc = c0 + tx;
        float v = 0.f;
        if (cc < c && xx < ow) {
            v = (float)ap[n * a.sn + y * a.sy + xx * a.sx + cc];
            if (lp) v += (float)lp[n * lat.sn + y * lat.sy + xx * lat.sx + cc];
        }
        tile[ty + 8 * i][tx] = v;   // [x][c]
    }
    __syncthreads();
    int y0, y1; float ly0, ly1;
    bilin_coord(y, sh, y0, y1, ly0, ly1);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int cc = c0 + ty + 8 * i, xx = x0 + tx;
        if (cc < c && xx < ow) {
            float v = tile[tx][ty + 8 * i];
            if (skip.ptr) {
                int xa, xb; float lx0, lx1;
                bilin_coord(xx, sw_, xa, xb, lx0, lx1);
                const long long s = n * skip.sn + cc * skip.sc;
                float v00 = vld(skip, s + y0 * skip.sh + xa * skip.sw), v01 = vld(skip, s + y0 * skip.sh + xb * skip.sw);
                float v10 = vld(skip, s + y1 * skip.sh + xa * skip.sw), v11 = vld(skip, s + y1 * skip.sh + xb * skip.sw);
                v += ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
            }
            v *= scale;
            if (fw) v = fw[0] * vld(add, n * add.sn + cc * add.sc + y * add.sh + xx * add.sw) + fw[1] * v;
            dst[(((long long)n * c + cc) * oh + y) * ow + xx] = v;
        }
    }
}
template <typename T>
int nhwc_to_nchw(PView a, PView lat, afi_view4 skip, int skip_h, int skip_w, float scale, int n, int c, int oh, int ow,
                 float* dst, cudaStream_t st, const afi_view4* add, const float* fw) {
    AFI_REQUIRE((long long)n * oh <= 65535, "layout conversion: n * h = %lld exceeds the 65535-block grid limit (split the batch)", (long long)n * oh);
    AFI_REQUIRE(view_dtype_ok(skip) && (!add || view_dtype_ok(*add)), "layout conversion: unknown view dtype");
    dim3 grid(cdiv(ow, 32), cdiv(c, 32), n * oh), block(32, 8);
    afi_view4 addv; memset(&addv, 0, sizeof(addv));
    if (add && fw) addv = *add; else fw = nullptr;
    // (a 64-channel-tile variant with 4-byte loads, as in k_nchw_to_nhwc_bf16x2, measured SLOWER here: 0.27 vs 0.23 ms per step -- the
    //  bilinear skip makes the output side the heavy one)
    k_nhwc_to_nchw<T><<<grid, block, 0, st>>>(a, lat, skip, skip_h, skip_w, scale, c, oh, ow, dst, addv, fw);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
template int nhwc_to_nchw<float>(PView, PView, afi_view4, int, int, float, int, int, int, int, float*, cudaStream_t, const afi_view4*, const float*);
template int nhwc_to_nchw<bf16>(PView, PView, afi_view4, int, int, float, int, int, int, int, float*, cudaStream_t, const afi_view4*, const float*);

// ---------------------------------------------------------------------------------------------------
// Top-down fusion site of the BiFPN neck WITH autograd (reference bifpn_sr.py:542-548 + the swish of :591-594):
//   s = w0 * cur + w1 * up,  out = act ? s * sigmoid(s) : s          (w = NULL: plain sum)
// One elementwise pass forward (five torch kernels otherwise) and one backward:
//   ds = dout * (act ? sig(s) * (1 + s * (1 - sig(s))) : 1),  d_cur = w0 * ds,  d_up = w1 * ds,  d_w += (sum ds * cur, sum ds * up)
// Views carry their own strides / dtype; s, out and the gradients are contiguous fp32 NCHW.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fuse_act_fwd(afi_view4 cur, afi_view4 up, const float* __restrict__ wts, int act, int c, int h, int w,
                                                      long long total, float* __restrict__ s_out, float* __restrict__ out) {
    const float w0 = wts ? wts[0] : 1.f, w1 = wts ? wts[1] : 1.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % w), y = (int)((i / w) % h), cc = (int)((i / ((long long)w * h)) % c), n = (int)(i / ((long long)w * h * c));
        const float s = w0 * vld(cur, n * cur.sn + cc * cur.sc + y * cur.sh + x * cur.sw) + w1 * vld(up, n * up.sn + cc * up.sc + y * up.sh + x * up.sw);
        if (s_out) s_out[i] = s;
        out[i] = act ? s / (1.f + __expf(-s)) : s;
    }
}
__global__ void __launch_bounds__(256) k_fuse_act_bwd(afi_view4 dout, const float* __restrict__ s_in, afi_view4 cur, afi_view4 up,
                                                      const float* __restrict__ wts, int act, int c, int h, int w, long long total,
                                                      float* __restrict__ d_cur, float* __restrict__ d_up, float* d_w) {
    __shared__ float red[2][8];
    const float w0 = wts ? wts[0] : 1.f, w1 = wts ? wts[1] : 1.f;
    float a0 = 0.f, a1 = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % w), y = (int)((i / w) % h), cc = (int)((i / ((long long)w * h)) % c), n = (int)(i / ((long long)w * h * c));
        float ds = vld(dout, n * dout.sn + cc * dout.sc + y * dout.sh + x * dout.sw);
        if (act) {
            const float s = s_in[i], sg = 1.f / (1.f + __expf(-s));
            ds *= sg * (1.f + s * (1.f - sg));
        }
        if (d_cur) d_cur[i] = w0 * ds;
        if (d_up) d_up[i] = w1 * ds;
        if (d_w) {
            a0 += ds * vld(cur, n * cur.sn + cc * cur.sc + y * cur.sh + x * cur.sw);
            a1 += ds * vld(up, n * up.sn + cc * up.sc + y * up.sh + x * up.sw);
        }
    }
    if (d_w) {      // uniform branch
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); }
        if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a0; red[1][threadIdx.x >> 5] = a1; }
        __syncthreads();
        if (threadIdx.x < 2) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < 8; j++) t += red[threadIdx.x][j];
            atomicAdd(d_w + threadIdx.x, t);
        }
    }
}
extern "C" int afi_bifpn_fuse_act(afi_view4 cur, afi_view4 up, const float* weights, int act, int n, int c, int h, int w, float* s_out, float* out,
                                  void* stream) {
    const long long total = (long long)n * c * h * w;
    AFI_REQUIRE(cur.ptr && up.ptr && out && total > 0, "afi_bifpn_fuse_act: bad argument");
    AFI_REQUIRE(view_dtype_ok(cur) && view_dtype_ok(up), "afi_bifpn_fuse_act: unknown view dtype");
    long long grid = cdiv(total, 256 * 4); if (grid > 148 * 16) grid = 148 * 16; if (grid < 1) grid = 1;
    k_fuse_act_fwd<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(cur, up, weights, act, c, h, w, total, act ? s_out : nullptr, out);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
extern "C" int afi_bifpn_fuse_act_backward(afi_view4 dout, const float* s, afi_view4 cur, afi_view4 up, const float* weights, int act, int n, int c,
                                           int h, int w, float* d_cur, float* d_up, float* d_w, void* stream) {
    const long long total = (long long)n * c * h * w;
    AFI_REQUIRE(dout.ptr && total > 0 && (!act || s), "afi_bifpn_fuse_act_backward: bad argument");
    AFI_REQUIRE(!d_w || (cur.ptr && up.ptr), "afi_bifpn_fuse_act_backward: the weight gradient needs both operands");
    AFI_REQUIRE(view_dtype_ok(dout) && view_dtype_ok(cur) && view_dtype_ok(up), "afi_bifpn_fuse_act_backward: unknown view dtype");
    cudaStream_t st = (cudaStream_t)stream;
    if (d_w) AFI_CUDA(cudaMemsetAsync(d_w, 0, 2 * sizeof(float), st));
    long long grid = cdiv(total, 256 * 4); if (grid > 148 * 8) grid = 148 * 8; if (grid < 1) grid = 1;
    k_fuse_act_bwd<<<(int)grid, 256, 0, st>>>(dout, s, cur, up, weights, act, c, h, w, total, d_cur, d_up, d_w);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}

// ---------------------------------------------------------------------------------------------------
// Input gradient of the interpolator: dx[n,c,i,j] = dXb[n,i,j,c] (head-conv dgrad, NHWC fp32) + bilinear2x^T(dy)[n,c,i,j].
// Adjoint of the x2 bilinear upsample (align_corners=False): x[i] feeds out[2i-1] (1/4), out[2i], out[2i+1] (3/4), out[2i+2] (1/4);
// the edge clamps fold the missing neighbour's 1/4 onto out[0] / out[2H-1] (weight 1).  dy is only defined on the oh x ow crop.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float bilin_adj_w(int i, int o, int size) {
    if ((i == 0 && o == 0) || (i == size - 1 && o == 2 * size - 1)) return 1.f;
    return (o == 2 * i || o == 2 * i + 1) ? 0.75f : 0.25f;
}
__global__ void k_g_input_grad(PView dxb, afi_view4 dy, float dy_scale, int c, int h, int w, int oh, int ow, float* __restrict__ dst) {
    __shared__ float tile[32][33];
    int x0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    int n = blockIdx.z / h, y = blockIdx.z % h;
    int tx = threadIdx.x, ty = threadIdx.y;
    const float* ap = reinterpret_cast<const float*>(dxb.ptr);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int xx = x0 + ty + 8 * i, cc = c0 + tx;
        tile[ty + 8 * i][tx] = (cc < c && xx < w) ? ap[n * dxb.sn + y * dxb.sy + xx * dxb.sx + cc] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int cc = c0 + ty + 8 * i, xx = x0 + tx;
        if (cc < c && xx < w) {
            float v = tile[tx][ty + 8 * i];
            const long long g = n * dy.sn + cc * dy.sc;
            float acc = 0.f;
#pragma unroll
            for (int a = -1; a <= 2; a++) {
                int oy = 2 * y + a;
                if (oy < 0 || oy >= oh || oy >= 2 * h) continue;
                float wy = bilin_adj_w(y, oy, h);
#pragma unroll
                for (int b = -1; b <= 2; b++) {
                    int ox = 2 * xx + b;
                    if (ox < 0 || ox >= ow || ox >= 2 * w) continue;
                    acc += wy * bilin_adj_w(xx, ox, w) * vld(dy, g + oy * dy.sh + ox * dy.sw);
                }
            }
            dst[(((long long)n * c + cc) * h + y) * w + xx] = v + dy_scale * acc;
        }
    }
}
int g_input_grad(PView dxb, afi_view4 dy, float dy_scale, int n, int c, int h, int w, int oh, int ow, float* dst, cudaStream_t st) {
    AFI_REQUIRE(view_dtype_ok(dy), "input gradient: unknown view dtype %d", dy.dtype);
    AFI_REQUIRE((long long)n * h <= 65535, "g_input_grad: n * h = %lld exceeds the 65535-block grid limit (split the batch)", (long long)n * h);
    dim3 grid(cdiv(w, 32), cdiv(c, 32), n * h), block(32, 8);
    k_g_input_grad<<<grid, block, 0, st>>>(dxb, dy, dy_scale, c, h, w, oh, ow, dst);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}

// ---------------------------------------------------------------------------------------------------
// dst = scale * (a + b) * lrelu'(mask)
// ---------------------------------------------------------------------------------------------------
__global__ void k_ew_combine(PView dst, int dst_dt, PView a, int a_dt, PView b, int b_dt, PView mask, int mask_dt,
                             float mslope, float scale, int H, int W, int cq, long long total) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    int c = (int)(i % cq) * 4;
    PixIdx q = decode_pixel(i / cq, H, W);
    float4 v = ld4(a.ptr, voff(a, q) + c, a_dt);
    if (b.ptr) { float4 t = ld4(b.ptr, voff(b, q) + c, b_dt); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    if (mask.ptr) {
        float4 m = ld4(mask.ptr, voff(mask, q) + c, mask_dt);
        v.x *= lmaskf(m.x, mslope); v.y *= lmaskf(m.y, mslope); v.z *= lmaskf(m.z, mslope); v.w *= lmaskf(m.w, mslope);
    }
    st4(dst.ptr, voff(dst, q) + c, dst_dt, v);
}
int ew_combine(PView dst, int dst_dt, PView a, int a_dt, PView b, int b_dt, PView mask, int mask_dt, float mask_slope,
               float scale, int n, int h, int w, int c, cudaStream_t st) {
    long long total = (long long)n * h * w * (c / 4);
    if (total == 0) return AFI_OK;
    k_ew_combine<<<cdiv(total, 256), 256, 0, st>>>(dst, dst_dt, a, a_dt, b, b_dt, mask, mask_dt, mask_slope, scale, h, w, c / 4, total);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}

// the same op over up to AFI_MAX_PROB problems (the pyramid levels of a generator backward) in ONE launch: most of these are a few
// hundred KB, i.e. launch-bound on their own
struct EwProb { PView dst, a, b, mask; int H, W; long long total; int block_begin, pad_; };
struct EwGroup { int nprob; EwProb p[AFI_MAX_PROB + 1]; };
__global__ void k_ew_combine_group(const __grid_constant__ EwGroup G, int dst_dt, int a_dt, int b_dt, int mask_dt, float mslope, float scale, int cq) {
    int k = 0;
    while (k + 1 < G.nprob && (int)blockIdx.x >= G.p[k + 1].block_begin) k++;
    const EwProb& e = G.p[k];
    long long i = (long long)(blockIdx.x - e.block_begin) * blockDim.x + threadIdx.x;
    if (i >= e.total) return;
    int c = (int)(i % cq) * 4;
    PixIdx q = decode_pixel(i / cq, e.H, e.W);
    float4 v = ld4(e.a.ptr, voff(e.a, q) + c, a_dt);
    if (e.b.ptr) { float4 t = ld4(e.b.ptr, voff(e.b, q) + c, b_dt); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    if (e.mask.ptr) {
        float4 m = ld4(e.mask.ptr, voff(e.mask, q) + c, mask_dt);
        v.x *= lmaskf(m.x, mslope); v.y *= lmaskf(m.y, mslope); v.z *= lmaskf(m.z, mslope); v.w *= lmaskf(m.w, mslope);
    }
    st4(e.dst.ptr, voff(e.dst, q) + c, dst_dt, v);
}
int ew_combine_group(int nprob, const PView* dst, int dst_dt, const PView* a, int a_dt, const PView* b, int b_dt, const PView* mask, int mask_dt,
                     float mask_slope, float scale, const int* n, const int* h, const int* w, int c, cudaStream_t st) {
    AFI_REQUIRE(nprob >= 1 && nprob <= AFI_MAX_PROB, "ew_combine_group: bad problem count %d", nprob);
    EwGroup G; memset(&G, 0, sizeof(G));
    G.nprob = nprob;
    int blocks = 0;
    for (int k = 0; k < nprob; k++) {
        EwProb& e = G.p[k];
        e.dst = dst[k]; e.a = a[k]; e.b = b ? b[k] : pview_null(); e.mask = mask ? mask[k] : pview_null();
        e.H = h[k]; e.W = w[k]; e.total = (long long)n[k] * h[k] * w[k] * (c / 4); e.block_begin = blocks;
        blocks += cdiv(e.total, 256);
    }
    G.p[nprob].block_begin = blocks;
    if (blocks == 0) return AFI_OK;
    k_ew_combine_group<<<blocks, 256, 0, st>>>(G, dst_dt, a_dt, b_dt, mask_dt, mask_slope, scale, c / 4);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}

// ---------------------------------------------------------------------------------------------------
// per-channel reductions over pixels.  Block = 256 threads = (c/4 channel quads) x (pixel lanes); each block
// owns a chunk of pixels, accumulates in fp32 registers, reduces lanes through shared memory and issues one
// atomic per channel.  MODE 0: sum, sumsq of x (double).  MODE 1: sum dy, sum dy*xhat (double).  MODE 2: sum x (float).
// ---------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void k_col_reduce(PView x, PView z, int dt, const float* __restrict__ mean, const float* __restrict__ rstd,
                             int H, int W, int cq, long long npix, int chunk, double* o0, double* o1, float* of) {
    __shared__ float4 s0[256];
    __shared__ float4 s1[256];
    int lanes = blockDim.x / cq;
    int q4 = threadIdx.x % cq, lane = threadIdx.x / cq;
    long long p0 = (long long)blockIdx.x * chunk, p1 = p0 + chunk;
    if (p1 > npix) p1 = npix;
    float4 a0 = make_float4(0, 0, 0, 0), a1 = make_float4(0, 0, 0, 0);
    float4 mu = make_float4(0, 0, 0, 0), rs = make_float4(1, 1, 1, 1);
    if (MODE == 1 && lane < lanes) { mu = *reinterpret_cast<const float4*>(mean + q4 * 4); rs = *reinterpret_cast<const float4*>(rstd + q4 * 4); }
    if (lane < lanes) {
        for (long long p = p0 + lane; p < p1; p += lanes) {
            PixIdx q = decode_pixel(p, H, W);
            float4 v = ld4(x.ptr, voff(x, q) + q4 * 4, dt);
            if (MODE == 0) {
                a0.x += v.x; a0.y += v.y; a0.z += v.z; a0.w += v.w;
                a1.x += v.x * v.x; a1.y += v.y * v.y; a1.z += v.z * v.z; a1.w += v.w * v.w;
            } else if (MODE == 1) {
                float4 zz = ld4(z.ptr, voff(z, q) + q4 * 4, dt);
                a0.x += v.x; a0.y += v.y; a0.z += v.z; a0.w += v.w;
                a1.x += v.x * (zz.x - mu.x) * rs.x; a1.y += v.y * (zz.y - mu.y) * rs.y;
                a1.z += v.z * (zz.z - mu.z) * rs.z; a1.w += v.w * (zz.w - mu.w) * rs.w;
            } else {
                a0.x += v.x; a0.y += v.y; a0.z += v.z; a0.w += v.w;
            }
        }
    }
    s0[threadIdx.x] = a0; s1[threadIdx.x] = a1;
    __syncthreads();
    if (lane == 0) {
        for (int l = 1; l < lanes; l++) {
            float4 t = s0[l * cq + q4]; a0.x += t.x; a0.y += t.y; a0.z += t.z; a0.w += t.w;
            if (MODE != 2) { float4 u = s1[l * cq + q4]; a1.x += u.x; a1.y += u.y; a1.z += u.z; a1.w += u.w; }
        }
        int c = q4 * 4;
        if (MODE == 2) {
            atomicAdd(of + c, a0.x); atomicAdd(of + c + 1, a0.y); atomicAdd(of + c + 2, a0.z); atomicAdd(of + c + 3, a0.w);
        } else {
            atomicAdd(o0 + c, (double)a0.x); atomicAdd(o0 + c + 1, (double)a0.y); atomicAdd(o0 + c + 2, (double)a0.z); atomicAdd(o0 + c + 3, (double)a0.w);
            if (o1) { atomicAdd(o1 + c, (double)a1.x); atomicAdd(o1 + c + 1, (double)a1.y); atomicAdd(o1 + c + 2, (double)a1.z); atomicAdd(o1 + c + 3, (double)a1.w); }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Dense fast paths.  Every BatchNorm / column-reduction operand of the discriminator is a dense [P][C] buffer, so these
// kernels drop the per-element (n,y,x) decode: a thread owns 8 consecutive channels (one 16-byte bf16 / two 16-byte fp32
// accesses), keeps the per-channel parameters in registers, and walks rows with an unrolled loop so that several
// independent 16-byte loads are in flight per thread (HBM-bound kernels: memory-level parallelism is what matters).
// (Negative result: streaming cache hints -- __ldcs / __stcs on these accesses -- leave the BatchNorm apply passes at 1.38 / 0.69 ms per step
//  and slow the two-operand column reduction from 0.36 to 0.45 ms.)
// ---------------------------------------------------------------------------------------------------
template <typename T> struct V8;
template <> struct V8<float> {
    __device__ static __forceinline__ void ld(const float* p, float* v) {
        float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    __device__ static __forceinline__ void st(float* p, const float* v) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
};
template <> struct V8<bf16> {
    __device__ static __forceinline__ void ld(const bf16* p, float* v) {
        uint4 u = *reinterpret_cast<const uint4*>(p);
        uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; i++) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
    }
    __device__ static __forceinline__ void st(bf16* p, const float* v) {
        uint4 u;
        u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(p) = u;
    }
};
static inline bool is_dense(const PView& v, int h, int w, int c) {
    return v.sx == c && v.sy == (long long)w * c && v.sn == (long long)h * w * c;
}
static inline bool dense_ok(int c) { return c % 8 == 0 && c / 8 <= 256 && 256 % (c / 8) == 0; }
constexpr int DENSE_ROWS = 64;     // rows per block of the elementwise passes
constexpr int REDUCE_ROWS = 256;   // max rows per block of the column reductions (fewer for small tensors: keep >= ~600 blocks)

template <int MODE, typename T>
__global__ void __launch_bounds__(256) k_col_reduce_dense(const T* __restrict__ x, const T* __restrict__ z, long long P, int C,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          double* o0, double* o1, float* of, int rows_per_block) {
    __shared__ float sm0[256 * 8];
    __shared__ float sm1[MODE == 2 ? 8 : 256 * 8];
    const int tpr = C >> 3, rpb = 256 / tpr;
    const int cq = threadIdx.x % tpr, lane_r = threadIdx.x / tpr;
    const int c0 = cq * 8;
    float a0[8], a1[8], mu[8], rs[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { a0[k] = 0.f; a1[k] = 0.f; mu[k] = 0.f; rs[k] = 1.f; }
    if (MODE == 1) {
#pragma unroll
        for (int k = 0; k < 8; k++) { mu[k] = mean[c0 + k]; rs[k] = rstd[c0 + k]; }
    }
    long long r0 = (long long)blockIdx.x * rows_per_block, r1 = r0 + rows_per_block;
    if (r1 > P) r1 = P;
#pragma unroll 4
    for (long long r = r0 + lane_r; r < r1; r += rpb) {
        float v[8];
        V8<T>::ld(x + r * C + c0, v);
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < 8; k++) { a0[k] += v[k]; a1[k] = fmaf(v[k], v[k], a1[k]); }
        } else if (MODE == 1) {
            float zz[8];
            V8<T>::ld(z + r * C + c0, zz);
#pragma unroll
            for (int k = 0; k < 8; k++) { a0[k] += v[k]; a1[k] = fmaf(v[k], (zz[k] - mu[k]) * rs[k], a1[k]); }
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) a0[k] += v[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        sm0[(lane_r * 8 + k) * tpr + cq] = a0[k];
        if (MODE != 2) sm1[(lane_r * 8 + k) * tpr + cq] = a1[k];
    }
    __syncthreads();
    // thread t < C finalises channel t: sum over the rpb row lanes
    for (int c = threadIdx.x; c < C; c += 256) {
        int q = c >> 3, k = c & 7;
        float t0 = 0.f, t1 = 0.f;
        for (int l = 0; l < rpb; l++) {
            t0 += sm0[(l * 8 + k) * tpr + q];
            if (MODE != 2) t1 += sm1[(l * 8 + k) * tpr + q];
        }
        if (MODE == 2) atomicAdd(of + c, t0);
        else { atomicAdd(o0 + c, (double)t0); if (o1) atomicAdd(o1 + c, (double)t1); }
    }
}


// ---------------------------------------------------------------------------------------------------
// Grouped dense passes: ONE launch covers the same BatchNorm pass of up to AFI_MAX_PROB discriminator calls (each call keeps its own
// statistics).  The small pyramid levels are latency-bound as separate launches (10-17 us each for a few MB); grouped they ride along
// with the large ones.  blockIdx.x -> (problem, row block) through a prefix table.
// ---------------------------------------------------------------------------------------------------
struct DenseProbD {
    const void* a; const void* b; void* out;
    long long P;
    const float* mean; const float* rstd;
    double* o0; double* o1;
    int block_begin, rows_per_block;
};
struct DenseGroupD { int nprob, C; DenseProbD p[AFI_MAX_PROB + 1]; };

__device__ __forceinline__ int find_prob(const DenseGroupD& G, int b) {
    int k = 0;
    while (k + 1 < G.nprob && b >= G.p[k + 1].block_begin) k++;
    return k;
}

template <typename T>
__global__ void __launch_bounds__(256) k_bn_apply_group(const __grid_constant__ DenseGroupD G, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float slope) {
    const int k = find_prob(G, blockIdx.x);
    const DenseProbD& q = G.p[k];
    const int C = G.C, tpr = C >> 3, rpb = 256 / tpr;
    const int c0 = (threadIdx.x % tpr) * 8;
    const T* __restrict__ z = reinterpret_cast<const T*>(q.a);
    T* __restrict__ a = reinterpret_cast<T*>(q.out);
    float mu[8], rs[8], ga[8], be[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { mu[i] = q.mean[c0 + i]; rs[i] = q.rstd[c0 + i]; ga[i] = gamma[c0 + i]; be[i] = beta[c0 + i]; }
    long long r0 = (long long)(blockIdx.x - q.block_begin) * q.rows_per_block, r1 = r0 + q.rows_per_block;
    if (r1 > q.P) r1 = q.P;
#pragma unroll 4
    for (long long r = r0 + threadIdx.x / tpr; r < r1; r += rpb) {
        float v[8];
        V8<T>::ld(z + r * C + c0, v);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float t = (v[i] - mu[i]) * rs[i] * ga[i] + be[i];
            v[i] = t > 0.f ? t : t * slope;
        }
        V8<T>::st(a + r * C + c0, v);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_bn_bwd_apply_group(const __grid_constant__ DenseGroupD G, const float* __restrict__ gamma,
                                                            float* dgamma_acc, float* dbeta_acc, int eval_mode) {
    const int k = find_prob(G, blockIdx.x);
    const DenseProbD& q = G.p[k];
    const int C = G.C, tpr = C >> 3, rpb = 256 / tpr;
    if ((int)blockIdx.x == q.block_begin) {     // d gamma = sum dy*xhat, d beta = sum dy: once per call
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            if (dgamma_acc) atomicAdd(dgamma_acc + c, (float)q.o1[c]);
            if (dbeta_acc) atomicAdd(dbeta_acc + c, (float)q.o0[c]);
        }
    }
    const int c0 = (threadIdx.x % tpr) * 8;
    T* __restrict__ dy = reinterpret_cast<T*>(q.out);
    const T* __restrict__ z = reinterpret_cast<const T*>(q.b);
    const float inv_m = 1.f / (float)q.P;
    float mu[8], rs[8], gr[8], m1[8], m2[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        mu[i] = q.mean[c0 + i]; rs[i] = q.rstd[c0 + i]; gr[i] = gamma[c0 + i] * rs[i];
        // eval mode: the statistics are constants (running buffers), so dz = gamma * rstd * dy without the two mean terms
        m1[i] = eval_mode ? 0.f : (float)q.o0[c0 + i] * inv_m; m2[i] = eval_mode ? 0.f : (float)q.o1[c0 + i] * inv_m;
    }
    long long r0 = (long long)(blockIdx.x - q.block_begin) * q.rows_per_block, r1 = r0 + q.rows_per_block;
    if (r1 > q.P) r1 = q.P;
#pragma unroll 4
    for (long long r = r0 + threadIdx.x / tpr; r < r1; r += rpb) {
        float g[8], zz[8];
        V8<T>::ld(dy + r * C + c0, g);
        V8<T>::ld(z + r * C + c0, zz);
#pragma unroll
        for (int i = 0; i < 8; i++) g[i] = gr[i] * (g[i] - m1[i] - (zz[i] - mu[i]) * rs[i] * m2[i]);
        V8<T>::st(dy + r * C + c0, g);
    }
}

// MODE 0: o0 += sum x, o1 += sum x^2.  MODE 1: o0 += sum dy, o1 += sum dy*xhat.
template <int MODE, typename T>
__global__ void __launch_bounds__(256) k_col_reduce_group(const __grid_constant__ DenseGroupD G) {
    __shared__ float sm0[256 * 8];
    __shared__ float sm1[256 * 8];
    const int k = find_prob(G, blockIdx.x);
    const DenseProbD& q = G.p[k];
    const int C = G.C, tpr = C >> 3, rpb = 256 / tpr;
    const int cq = threadIdx.x % tpr, lane_r = threadIdx.x / tpr;
    const int c0 = cq * 8;
    const T* __restrict__ x = reinterpret_cast<const T*>(q.a);
    const T* __restrict__ z = reinterpret_cast<const T*>(q.b);
    float a0[8], a1[8], mu[8], rs[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a0[i] = 0.f; a1[i] = 0.f; mu[i] = 0.f; rs[i] = 1.f; }
    if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 8; i++) { mu[i] = q.mean[c0 + i]; rs[i] = q.rstd[c0 + i]; }
    }
    long long r0 = (long long)(blockIdx.x - q.block_begin) * q.rows_per_block, r1 = r0 + q.rows_per_block;
    if (r1 > q.P) r1 = q.P;
#pragma unroll 4
    for (long long r = r0 + lane_r; r < r1; r += rpb) {
        float v[8];
        V8<T>::ld(x + r * C + c0, v);
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) { a0[i] += v[i]; a1[i] = fmaf(v[i], v[i], a1[i]); }
        } else {
            float zz[8];
            V8<T>::ld(z + r * C + c0, zz);
#pragma unroll
            for (int i = 0; i < 8; i++) { a0[i] += v[i]; a1[i] = fmaf(v[i], (zz[i] - mu[i]) * rs[i], a1[i]); }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) { sm0[(lane_r * 8 + i) * tpr + cq] = a0[i]; sm1[(lane_r * 8 + i) * tpr + cq] = a1[i]; }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        int qd = c >> 3, i = c & 7;
        float t0 = 0.f, t1 = 0.f;
        for (int l = 0; l < rpb; l++) { t0 += sm0[(l * 8 + i) * tpr + qd]; t1 += sm1[(l * 8 + i) * tpr + qd]; }
        atomicAdd(q.o0 + c, (double)t0);
        atomicAdd(q.o1 + c, (double)t1);
    }
}

// per call: mean / rstd / unbiased var from the sums (training) -- grid (C/256, ncalls)
struct FinalizeGroup { int ncalls; const double* sum[AFI_MAX_PROB]; const double* sumsq[AFI_MAX_PROB]; long long count[AFI_MAX_PROB];
                       float* mean[AFI_MAX_PROB]; float* rstd[AFI_MAX_PROB]; float* var[AFI_MAX_PROB]; };
__global__ void k_bn_finalize_group(const __grid_constant__ FinalizeGroup F, int c, float eps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (i >= c) return;
    double cnt = (double)F.count[k];
    double m = F.sum[k][i] / cnt;
    double var = F.sumsq[k][i] / cnt - m * m;
    if (var < 0) var = 0;
    F.mean[k][i] = (float)m;
    F.rstd[k][i] = (float)(1.0 / sqrt(var + (double)eps));
    F.var[k][i] = (float)(cnt > 1 ? var * cnt / (cnt - 1) : var);
}
struct ZeroGroup { int n; void* p[AFI_MAX_PROB]; };
__global__ void k_zero_group(const __grid_constant__ ZeroGroup Z, int words) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < words) reinterpret_cast<uint32_t*>(Z.p[blockIdx.y])[i] = 0u;
}

static void dense_group_blocks(DenseGroupD& G, int max_rows, int target_blocks) {
    long long total = 0;
    for (int k = 0; k < G.nprob; k++) total += G.p[k].P;
    long long rows = total / target_blocks;
    if (rows > max_rows) rows = max_rows;
    if (rows < 8) rows = 8;
    int b = 0;
    for (int k = 0; k < G.nprob; k++) {
        G.p[k].block_begin = b;
        G.p[k].rows_per_block = (int)rows;
        b += (int)((G.p[k].P + rows - 1) / rows);
    }
    G.p[G.nprob].block_begin = b;
}

int dense_group_ok(int c) { return dense_ok(c) ? 1 : 0; }

int bn_apply_lrelu_group(int nprob, const PView* z, const PView* a, int dt, const float* const* mean, const float* const* rstd,
                         const float* gamma, const float* beta, float slope, const long long* P, int c, cudaStream_t st) {
    DenseGroupD G; memset(&G, 0, sizeof(G));
    G.nprob = nprob; G.C = c;
    for (int k = 0; k < nprob; k++) { G.p[k].a = z[k].ptr; G.p[k].out = a[k].ptr; G.p[k].P = P[k]; G.p[k].mean = mean[k]; G.p[k].rstd = rstd[k]; }
    dense_group_blocks(G, DENSE_ROWS, 2368);
    int grid = G.p[nprob].block_begin;
    if (grid == 0) return AFI_OK;
    if (dt == DT_F32) k_bn_apply_group<float><<<grid, 256, 0, st>>>(G, gamma, beta, slope);
    else k_bn_apply_group<bf16><<<grid, 256, 0, st>>>(G, gamma, beta, slope);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
int bn_bwd_apply_group(int nprob, const PView* dy, const PView* z, int dt, const float* const* mean, const float* const* rstd, const float* gamma,
                       double* const* s_dy, double* const* s_dyx, float* dgamma_acc, float* dbeta_acc, const long long* P, int c, int eval_mode,
                       cudaStream_t st) {
    DenseGroupD G; memset(&G, 0, sizeof(G));
    G.nprob = nprob; G.C = c;
    for (int k = 0; k < nprob; k++) {
        G.p[k].out = dy[k].ptr; G.p[k].b = z[k].ptr; G.p[k].P = P[k]; G.p[k].mean = mean[k]; G.p[k].rstd = rstd[k];
        G.p[k].o0 = s_dy[k]; G.p[k].o1 = s_dyx[k];
    }
    dense_group_blocks(G, DENSE_ROWS, 2368);
    int grid = G.p[nprob].block_begin;
    if (grid == 0) return AFI_OK;
    if (dt == DT_F32) k_bn_bwd_apply_group<float><<<grid, 256, 0, st>>>(G, gamma, dgamma_acc, dbeta_acc, eval_mode);
    else k_bn_bwd_apply_group<bf16><<<grid, 256, 0, st>>>(G, gamma, dgamma_acc, dbeta_acc, eval_mode);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
int col_reduce_group(int mode, int nprob, const PView* x, const PView* z, int dt, const float* const* mean, const float* const* rstd,
                     double* const* o0, double* const* o1, const long long* P, int c, cudaStream_t st) {
    DenseGroupD G; memset(&G, 0, sizeof(G));
    G.nprob = nprob; G.C = c;
    for (int k = 0; k < nprob; k++) {
        G.p[k].a = x[k].ptr; G.p[k].b = z ? z[k].ptr : nullptr; G.p[k].P = P[k];
        G.p[k].mean = mean ? mean[k] : nullptr; G.p[k].rstd = rstd ? rstd[k] : nullptr; G.p[k].o0 = o0[k]; G.p[k].o1 = o1[k];
    }
    dense_group_blocks(G, REDUCE_ROWS, 1184);
    int grid = G.p[nprob].block_begin;
    if (grid == 0) return AFI_OK;
    if (mode == 0) { if (dt == DT_F32) k_col_reduce_group<0, float><<<grid, 256, 0, st>>>(G); else k_col_reduce_group<0, bf16><<<grid, 256, 0, st>>>(G); }
    else { if (dt == DT_F32) k_col_reduce_group<1, float><<<grid, 256, 0, st>>>(G); else k_col_reduce_group<1, bf16><<<grid, 256, 0, st>>>(G); }
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
int bn_finalize_group(int ncalls, const double* const* sum, const double* const* sumsq, const long long* count, int c, float eps,
                      float* const* mean, float* const* rstd, float* const* var, cudaStream_t st) {
    FinalizeGroup F; memset(&F, 0, sizeof(F));
    F.ncalls = ncalls;
    for (int k = 0; k < ncalls; k++) { F.sum[k] = sum[k]; F.sumsq[k] = sumsq[k]; F.count[k] = count[k]; F.mean[k] = mean[k]; F.rstd[k] = rstd[k]; F.var[k] = var[k]; }
    dim3 grid(cdiv(c, 256), ncalls);
    k_bn_finalize_group<<<grid, 256, 0, st>>>(F, c, eps);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
int zero_group(int n, void* const* ptrs, size_t bytes, cudaStream_t st) {
    ZeroGroup Z; memset(&Z, 0, sizeof(Z));
    Z.n = n;
    for (int k = 0; k < n; k++) Z.p[k] = ptrs[k];
    int words = (int)(bytes / 4);
    dim3 grid(cdiv(words, 256), n);
    k_zero_group<<<grid, 256, 0, st>>>(Z, words);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}

static int col_reduce_launch(int mode, PView x, PView z, int dt, const float* mean, const float* rstd, int n, int h, int w, int c,
                             double* o0, double* o1, float* of, cudaStream_t st) {
    AFI_REQUIRE(c % 4 == 0 && c / 4 <= 256 && 256 % (c / 4) == 0, "col_reduce: unsupported channel count %d", c);
    long long npix = (long long)n * h * w;
    if (npix == 0) return AFI_OK;
    if (is_dense(x, h, w, c) && dense_ok(c) && (mode != 1 || is_dense(z, h, w, c))) {
        int rows = (int)(npix / 592);
        if (rows > REDUCE_ROWS) rows = REDUCE_ROWS;
        if (rows < 8) rows = 8;
        int grid = cdiv(npix, rows);
#define AFI_CR(M, T) k_col_reduce_dense<M, T><<<grid, 256, 0, st>>>((const T*)x.ptr, (const T*)z.ptr, npix, c, mean, rstd, o0, o1, of, rows)
        if (dt == DT_F32) { if (mode == 0) AFI_CR(0, float); else if (mode == 1) AFI_CR(1, float); else AFI_CR(2, float); }
        else { if (mode == 0) AFI_CR(0, bf16); else if (mode == 1) AFI_CR(1, bf16); else AFI_CR(2, bf16); }
#undef AFI_CR
        AFI_LAUNCH_CHECK();
        return AFI_OK;
    }
    int chunk = 128;   // pixels per block: short fp32 partial sums (<= 128 terms per lane) keep BN statistics at ~1e-7
    int grid = cdiv(npix, chunk);
    if (mode == 0) k_col_reduce<0><<<grid, 256, 0, st>>>(x, z, dt, mean, rstd, h, w, c / 4, npix, chunk, o0, o1, of);
    else if (mode == 1) k_col_reduce<1><<<grid, 256, 0, st>>>(x, z, dt, mean, rstd, h, w, c / 4, npix, chunk, o0, o1, of);
    else k_col_reduce<2><<<grid, 256, 0, st>>>(x, z, dt, mean, rstd, h, w, c / 4, npix, chunk, o0, o1, of);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
int col_sum_f32(PView x, int dt, int n, int h, int w, int c, float* out, cudaStream_t st) {
    return col_reduce_launch(2, x, pview_null(), dt, nullptr, nullptr, n, h, w, c, nullptr, nullptr, out, st);
}

// ---------------------------------------------------------------------------------------------------
// BatchNorm finalize / apply / backward-apply
// ---------------------------------------------------------------------------------------------------
__global__ void k_bn_finalize(const double* sum, const double* sumsq, long long count, int c, float eps, float momentum,
                              int training, float* mean, float* rstd, float* var_unb, float* rmean, float* rvar, long long* nbt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    if (training) {
        double m = sum[i] / (double)count;
        double var = sumsq[i] / (double)count - m * m;
        if (var < 0) var = 0;
        double unb = count > 1 ? var * (double)count / (double)(count - 1) : var;
        mean[i] = (float)m;
        rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
        if (var_unb) var_unb[i] = (float)unb;
        if (rmean) rmean[i] = (1.f - momentum) * rmean[i] + momentum * (float)m;
        if (rvar) rvar[i] = (1.f - momentum) * rvar[i] + momentum * (float)unb;
        if (i == 0 && nbt) *nbt += 1;
    } else {
        mean[i] = rmean[i];
        rstd[i] = 1.f / sqrtf(rvar[i] + eps);
    }
}
int bn_finalize(const double* sum, const double* sumsq, long long count, int c, float eps, float momentum, int training,
                float* mean, float* rstd, float* var_unb, float* running_mean, float* running_var, long long* nbt, cudaStream_t st) {
    if (!training) AFI_REQUIRE(running_mean && running_var, "bn_finalize: eval mode needs running statistics");
    k_bn_finalize<<<cdiv(c, 256), 256, 0, st>>>(sum, sumsq, count, c, eps, momentum, training, mean, rstd, var_unb, running_mean, running_var, nbt);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
// Deferred running-statistics update: applies the momentum EMA of up to AFI_MAX_PROB calls IN CALL ORDER (so concurrent streams
// may evaluate the calls themselves in any order): running = (1 - m) * running + m * batch_stat, num_batches_tracked += ncalls.
struct BnRunArgs { int ncalls; const float* mean[AFI_MAX_PROB]; const float* var[AFI_MAX_PROB]; };
__global__ void k_bn_update_running(BnRunArgs a, int c, float momentum, float* rmean, float* rvar, long long* nbt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    float m = rmean ? rmean[i] : 0.f, v = rvar ? rvar[i] : 0.f;
    for (int k = 0; k < a.ncalls; k++) {
        m = (1.f - momentum) * m + momentum * a.mean[k][i];
        v = (1.f - momentum) * v + momentum * a.var[k][i];
    }
    if (rmean) rmean[i] = m;
    if (rvar) rvar[i] = v;
    if (i == 0 && nbt) *nbt += a.ncalls;
}
int bn_update_running(int ncalls, const float* const* mean, const float* const* var, int c, float momentum, float* rmean, float* rvar,
                      long long* nbt, cudaStream_t st) {
    BnRunArgs a;
    a.ncalls = ncalls;
    for (int k = 0; k < ncalls; k++) { a.mean[k] = mean[k]; a.var[k] = var[k]; }
    k_bn_update_running<<<cdiv(c, 256), 256, 0, st>>>(a, c, momentum, rmean, rvar, nbt);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}


// ---------------------------------------------------------------------------------------------------
// discriminator head: Conv2d 1024 -> 1, 3x3 (feature_patch_discriminator.py:40-41).  GEMV-like, HBM-bound:
// forward = 9 per-pixel dot products (one read of a3) + a 3x3 shift-sum; backward fuses the weight gradient,
// the input gradient and the LeakyReLU mask of layer 3 in one pass over a3.
// ---------------------------------------------------------------------------------------------------
__global__ void k_dhead_dots(PView a3, int dt, const float* __restrict__ w4, int H, int W, int c, long long npix, float* __restrict__ t9) {
    extern __shared__ float ws[];   // [9][c]
    for (int i = threadIdx.x; i < 9 * c; i += blockDim.x) { int cc = i / 9, t = i % 9; ws[t * c + cc] = w4[i]; }
    __syncthreads();
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (long long p = (long long)blockIdx.x * nwarp + warp; p < npix; p += (long long)gridDim.x * nwarp) {
        PixIdx q = decode_pixel(p, H, W);
        long long base = voff(a3, q);
        float acc[9];
#pragma unroll
        for (int t = 0; t < 9; t++) acc[t] = 0.f;
        for (int cc = lane * 4; cc < c; cc += 128) {
            float4 v = ld4(a3.ptr, base + cc, dt);
#pragma unroll
            for (int t = 0; t < 9; t++) {
                float4 wv = *reinterpret_cast<const float4*>(ws + t * c + cc);
                acc[t] += v.x * wv.x + v.y * wv.y + v.z * wv.z + v.w * wv.w;
            }
        }
#pragma unroll
        for (int t = 0; t < 9; t++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int t = 0; t < 9; t++) t9[p * 9 + t] = acc[t];
        }
    }
}
__global__ void k_dhead_stencil(const float* __restrict__ t9, const float* __restrict__ b4, int H, int W, long long npix, float* __restrict__ logits) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= npix) return;
    PixIdx q = decode_pixel(p, H, W);
    float v = b4 ? b4[0] : 0.f;
#pragma unroll
    for (int t = 0; t < 9; t++) {
        int yy = q.y + t / 3 - 1, xx = q.x + t % 3 - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v += t9[(((long long)q.n * H + yy) * W + xx) * 9 + t];
    }
    logits[p] = v;
}
int dhead_forward(PView a3, int dt, const float* w4, const float* b4, int n, int h, int w, int c, float* t9, float* logits, cudaStream_t st) {
    long long npix = (long long)n * h * w;
    if (npix == 0) return AFI_OK;
    AFI_REQUIRE(c % 128 == 0, "dhead: channels must be a multiple of 128");
    size_t smem = (size_t)9 * c * sizeof(float);
    AFI_CUDA(cudaFuncSetAttribute(k_dhead_dots, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (int)((npix + 7) / 8); if (grid > 148 * 8) grid = 148 * 8;
    k_dhead_dots<<<grid, 256, smem, st>>>(a3, dt, w4, h, w, c, npix, t9);
    AFI_LAUNCH_CHECK();
    k_dhead_stencil<<<cdiv(npix, 256), 256, 0, st>>>(t9, b4, h, w, npix, logits);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}

// block = c/4 threads (256 for c = 1024); each thread owns 4 channels and keeps its 4x9 weight-gradient partials in registers
__global__ void k_dhead_backward(PView a3, int dt, const float* __restrict__ w4, const float* __restrict__ g, int H, int W, int c,
                                 long long npix, int chunk, float* dw4_acc, float* db4_acc, PView dy3, float slope) {
    __shared__ float g9s[9];
    int c0 = threadIdx.x * 4;
    float wr[4][9], dw[4][9];
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int t = 0; t < 9; t++) { wr[k][t] = w4[(c0 + k) * 9 + t]; dw[k][t] = 0.f; }
    float db = 0.f;
    long long p0 = (long long)blockIdx.x * chunk, p1 = p0 + chunk;
    if (p1 > npix) p1 = npix;
    for (long long p = p0; p < p1; p++) {
        PixIdx q = decode_pixel(p, H, W);
        __syncthreads();
        if (threadIdx.x < 9) {   // g9[t] = g[q - tap_t]
            int t = threadIdx.x;
            int yy = q.y - (t / 3 - 1), xx = q.x - (t % 3 - 1);
            g9s[t] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? g[((long long)q.n * H + yy) * W + xx] : 0.f;
        }
        __syncthreads();
        float g9[9];
#pragma unroll
        for (int t = 0; t < 9; t++) g9[t] = g9s[t];
        long long o = voff(a3, q) + c0;
        float4 a = ld4(a3.ptr, o, dt);
        float av[4] = {a.x, a.y, a.z, a.w}, r[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float s = 0.f;
#pragma unroll
            for (int t = 0; t < 9; t++) { s += g9[t] * wr[k][t]; dw[k][t] += av[k] * g9[t]; }
            r[k] = s * lmaskf(av[k], slope);
        }
        st4(dy3.ptr, voff(dy3, q) + c0, dt, make_float4(r[0], r[1], r[2], r[3]));
        if (threadIdx.x == 0) db += g9[4];
    }
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int t = 0; t < 9; t++) atomicAdd(dw4_acc + (c0 + k) * 9 + t, dw[k][t]);
    if (threadIdx.x == 0 && db4_acc) atomicAdd(db4_acc, db);
}
int dhead_backward(PView a3, int dt, const float* w4, const float* g, int n, int h, int w, int c, float* dw4_acc, float* db4_acc,
                   PView dy3, cudaStream_t st) {
    long long npix = (long long)n * h * w;
    if (npix == 0) return AFI_OK;
    AFI_REQUIRE(c % 4 == 0 && c / 4 <= 1024, "dhead_backward: unsupported channel count");
    int chunk = (int)((npix + 148 * 4 - 1) / (148 * 4)); if (chunk < 16) chunk = 16;
    k_dhead_backward<<<cdiv(npix, chunk), c / 4, 0, st>>>(a3, dt, w4, g, h, w, c, npix, chunk, dw4_acc, db4_acc, dy3, 0.2f);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}

// ---------------------------------------------------------------------------------------------------
// discriminator head on the tensor-core engine (bf16 mode): the 1024 -> 1 3x3 conv is evaluated as a 1x1 conv to 16
// channels (t9[p][t] = <a3[p], w4[:, t]>, taps 9..15 zero) followed by a 3x3 shift-sum, and its backward as a 16 -> 1024
// 1x1 conv of g9[q][t] = g[q - tap_t] plus a K = pixels weight-gradient GEMM.  Both stream a3 exactly once (HBM-bound).
// ---------------------------------------------------------------------------------------------------
__global__ void k_dhead_pack_tc(const float* __restrict__ w4, int c, bf16* __restrict__ fwd, bf16* __restrict__ bwd) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 16 * c) return;
    int t = i / c, cc = i % c;
    float v = t < 9 ? w4[cc * 9 + t] : 0.f;
    fwd[t * c + cc] = (bf16)v;       // [cout = 16][cin = c]
    bwd[cc * 16 + t] = (bf16)v;      // [cout = c][cin = 16]
}
int dhead_pack_tc(const float* w4, int c, void* fwd, void* bwd, cudaStream_t st) {
    k_dhead_pack_tc<<<cdiv(16 * c, 256), 256, 0, st>>>(w4, c, (bf16*)fwd, (bf16*)bwd);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
__global__ void k_dhead_build_g9(const float* __restrict__ g, int H, int W, long long npix, bf16* __restrict__ g9, float* __restrict__ g9f) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= npix * 16) return;
    int t = (int)(i & 15);
    PixIdx q = decode_pixel(i >> 4, H, W);
    float v = 0.f;
    if (t < 9) {
        int yy = q.y - (t / 3 - 1), xx = q.x - (t % 3 - 1);
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = g[((long long)q.n * H + yy) * W + xx];
    }
    g9[i] = (bf16)v;
    if (g9f && t < 12) g9f[(i >> 4) * 12 + t] = v;     // fp32 copy [P][12] for the fused head-backward passes
}
int dhead_build_g9(const float* g, int n, int h, int w, void* g9, float* g9f, cudaStream_t st) {
    long long npix = (long long)n * h * w;
    k_dhead_build_g9<<<cdiv(npix * 16, 256), 256, 0, st>>>(g, h, w, npix, (bf16*)g9, g9f);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
__global__ void k_dhead_stencil16(const float* __restrict__ t9, const float* __restrict__ b4, int H, int W, long long npix, float* __restrict__ logits) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= npix) return;
    PixIdx q = decode_pixel(p, H, W);
    float v = b4 ? b4[0] : 0.f;
#pragma unroll
    for (int t = 0; t < 9; t++) {
        int yy = q.y + t / 3 - 1, xx = q.x + t % 3 - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v += t9[(((long long)q.n * H + yy) * W + xx) * 16 + t];
    }
    logits[p] = v;
}
int dhead_stencil16(const float* t9, const float* b4, int n, int h, int w, float* logits, cudaStream_t st) {
    long long npix = (long long)n * h * w;
    k_dhead_stencil16<<<cdiv(npix, 256), 256, 0, st>>>(t9, b4, h, w, npix, logits);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
template <typename T> struct V4;
template <> struct V4<float> {
    __device__ static __forceinline__ void ld(const float* p, float* v) { float4 a = *reinterpret_cast<const float4*>(p); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; }
    __device__ static __forceinline__ void st(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct V4<bf16> {
    __device__ static __forceinline__ void ld(const bf16* p, float* v) {
        uint2 u = *reinterpret_cast<const uint2*>(p);
        v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u); v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
    }
    __device__ static __forceinline__ void st(bf16* p, const float* v) {
        uint2 u; u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
        *reinterpret_cast<uint2*>(p) = u;
    }
};
// ---------------------------------------------------------------------------------------------------
// Discriminator head backward fused with layer 3's train-mode BatchNorm backward, grouped over calls, in TWO passes that never
// materialise dy3:   dy3[q][c] = (sum_t g9[q][t] * w4[c][t]) * lrelu'(bn3(z3[q][c]))      (9 FMAs per element, recomputed in both)
//   PASS 1: s_dy[c] += dy3, s_dyx[c] += dy3 * xhat                      reads z3 only          (2 B / element in bf16)
//   PASS 2: dz3 = gamma * rstd * (dy3 - s_dy/M - xhat * s_dyx/M)        reads z3, writes dz3   (4 B / element)
// versus 12 B / element for "write dy3 (reading a3 and z3), then a BatchNorm-backward pass over dy3 and z3".  The LeakyReLU mask is
// re-derived from z3 with the SAME fp32 expression bn_apply uses, so it equals the sign of the stored activation.
// A thread owns 4 channels (its 4 x 9 head weights stay in registers) and keeps ROWS independent 8-byte loads in flight.
// ---------------------------------------------------------------------------------------------------

// ---------------------------------------------------------------------------------------------------
// The same two passes with the 9-tap product on the tensor cores (bf16 operand mode): dy[16 rows x 8 channels] = g9[16 x 16] * w4^T[16 x 8]
// is one warp-level mma.sync.m16n8k16 (bf16 x bf16 -> fp32) instead of 9 x 128 FMAs, which turns the passes from FMA-bound into
// HBM-bound.  The columns of the four MMAs a warp issues per 16-row block are PERMUTED (column j of MMA nb <-> channel
// (j/2)*8 + nb*2 + j%2 of the warp's 32) so that the accumulator fragment of a thread is 8 CONSECUTIVE channels of two rows: z3 / dz3
// move as 16-byte vectors.  Block = 8 warps = 256 channels (blockIdx.y) x a range of rows; g9 is the bf16 [P][16] tensor the dW4 GEMM uses.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int PASS>
__global__ void __launch_bounds__(256, 2) k_dhead_bwd_mma(const __grid_constant__ DenseGroupD G, const float* __restrict__ w4,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float* dgamma_acc, float* dbeta_acc, float slope, int eval_mode) {
    constexpr int UNROLL = 2;            // 16-row blocks in flight per warp: 4 x 16 B of z3 per thread
    const int kp = find_prob(G, blockIdx.x);
    const DenseProbD& pr = G.p[kp];
    const int C = G.C;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = lane & 3, rl = lane >> 2;
    const int cw = blockIdx.y * 256 + warp * 32;     // first channel of this warp
    const int c0 = cw + q * 8;                       // this thread's 8 channels
    const bf16* __restrict__ g9 = reinterpret_cast<const bf16*>(pr.a);
    const bf16* __restrict__ z = reinterpret_cast<const bf16*>(pr.b);
    bf16* __restrict__ dz = reinterpret_cast<bf16*>(pr.out);
    if (PASS == 2 && (int)blockIdx.x == pr.block_begin) {     // d gamma = sum dy*xhat, d beta = sum dy: once per call and channel
        const int c = blockIdx.y * 256 + threadIdx.x;
        if (dgamma_acc) atomicAdd(dgamma_acc + c, (float)pr.o1[c]);
        if (dbeta_acc) atomicAdd(dbeta_acc + c, (float)pr.o0[c]);
    }
    // B fragments: MMA nb, column n = rl  <->  channel cw + (rl/2)*8 + nb*2 + rl%2; rows k = 2q, 2q+1 (b0) and 2q+8, 2q+9 (b1) = taps
    uint32_t bfr[4][2];
#pragma unroll
    for (int nb = 0; nb < 4; nb++) {
        const float* wc = w4 + (size_t)(cw + (rl >> 1) * 8 + nb * 2 + (rl & 1)) * 9;
        const int k0 = 2 * q;
        bfr[nb][0] = pack_bf16x2(wc[k0], wc[k0 + 1]);                    // taps 0..7
        bfr[nb][1] = q == 0 ? pack_bf16x2(wc[8], 0.f) : 0u;              // tap 8; taps 9..15 do not exist
    }
    float mu[8], rs[8], ga[8], be[8], a0[8], a1[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        mu[k] = pr.mean[c0 + k]; rs[k] = pr.rstd[c0 + k]; ga[k] = gamma[c0 + k]; be[k] = beta[c0 + k];
        if (PASS == 1) { a0[k] = 0.f; a1[k] = 0.f; }
        else {
            const float inv_m = 1.f / (float)pr.P;
            a0[k] = eval_mode ? 0.f : (float)pr.o0[c0 + k] * inv_m; a1[k] = eval_mode ? 0.f : (float)pr.o1[c0 + k] * inv_m;
        }
    }
    const long long r0 = (long long)(blockIdx.x - pr.block_begin) * pr.rows_per_block;
    long long r1 = r0 + pr.rows_per_block;
    if (r1 > pr.P) r1 = pr.P;
    for (long long rb = r0; rb < r1; rb += 16 * UNROLL) {
        uint4 zv[UNROLL][2];
        uint32_t af[UNROLL][4];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const long long r = rb + u * 16 + h * 8 + rl;
                zv[u][h] = make_uint4(0, 0, 0, 0);
                af[u][h] = 0; af[u][h + 2] = 0;
                if (r < r1) {
                    zv[u][h] = *reinterpret_cast<const uint4*>(z + r * C + c0);
                    const uint32_t* gp = reinterpret_cast<const uint32_t*>(g9 + r * 16);
                    af[u][h] = __ldg(gp + q); af[u][h + 2] = __ldg(gp + q + 4);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            float c[4][4];
#pragma unroll
            for (int nb = 0; nb < 4; nb++) { c[nb][0] = c[nb][1] = c[nb][2] = c[nb][3] = 0.f; mma_bf16_16816(c[nb], af[u], bfr[nb][0], bfr[nb][1]); }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const long long r = rb + u * 16 + h * 8 + rl;
                const uint32_t zw[4] = {zv[u][h].x, zv[u][h].y, zv[u][h].z, zv[u][h].w};
                float o[8];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const float zz = (k & 1) ? __uint_as_float(zw[k >> 1] & 0xffff0000u) : __uint_as_float(zw[k >> 1] << 16);
                    float d = c[k >> 1][(k & 1) + 2 * h];
                    const float xh = (zz - mu[k]) * rs[k];
                    const float y = (zz - mu[k]) * rs[k] * ga[k] + be[k];      // == bn_apply's expression: sign(y) is the stored activation's sign
                    d *= y > 0.f ? 1.f : slope;
                    if (PASS == 1) { if (r < r1) { a0[k] += d; a1[k] = fmaf(d, xh, a1[k]); } }
                    else o[k] = ga[k] * rs[k] * (d - a0[k] - xh * a1[k]);
                }
                if (PASS == 2 && r < r1) {
                    uint4 ov;
                    ov.x = pack_bf16x2(o[0], o[1]); ov.y = pack_bf16x2(o[2], o[3]); ov.z = pack_bf16x2(o[4], o[5]); ov.w = pack_bf16x2(o[6], o[7]);
                    *reinterpret_cast<uint4*>(dz + r * C + c0) = ov;
                }
            }
        }
    }
    if (PASS == 1) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) { a0[k] += __shfl_xor_sync(0xffffffffu, a0[k], o); a1[k] += __shfl_xor_sync(0xffffffffu, a1[k], o); }
        }
        if (rl == 0) {
#pragma unroll
            for (int k = 0; k < 8; k++) { atomicAdd(pr.o0 + c0 + k, (double)a0[k]); atomicAdd(pr.o1 + c0 + k, (double)a1[k]); }
        }
    }
}
int dhead_backward_group_mma(int pass, int nprob, const void* const* g9, const PView* z3, const PView* dz3, const float* w4,
                             const float* const* mean, const float* const* rstd, const float* gamma, const float* beta, double* const* s_dy,
                             double* const* s_dyx, float* dgamma_acc, float* dbeta_acc, const long long* P, int c, int eval_mode, cudaStream_t st) {
    AFI_REQUIRE(c % 256 == 0, "dhead_backward_group_mma: C must be a multiple of 256");
    DenseGroupD G; memset(&G, 0, sizeof(G));
    G.nprob = nprob; G.C = c;
    for (int k = 0; k < nprob; k++) {
        G.p[k].a = g9[k]; G.p[k].b = z3[k].ptr; G.p[k].out = dz3[k].ptr; G.p[k].P = P[k]; G.p[k].mean = mean[k]; G.p[k].rstd = rstd[k];
        G.p[k].o0 = s_dy[k]; G.p[k].o1 = s_dyx[k];
    }
    // rows per block: a multiple of 32 (two 16-row MMA blocks per iteration); ~600 row blocks x C/256 channel slabs
    long long total = 0;
    for (int k = 0; k < nprob; k++) total += P[k];
    long long rows = (total / 600 + 31) / 32 * 32;
    if (rows > 512) rows = 512;
    if (rows < 32) rows = 32;
    int b = 0;
    for (int k = 0; k < nprob; k++) { G.p[k].block_begin = b; G.p[k].rows_per_block = (int)rows; b += (int)((P[k] + rows - 1) / rows); }
    G.p[nprob].block_begin = b;
    if (b == 0) return AFI_OK;
    dim3 grid(b, c / 256);
    if (pass == 1) k_dhead_bwd_mma<1><<<grid, 256, 0, st>>>(G, w4, gamma, beta, dgamma_acc, dbeta_acc, 0.2f, eval_mode);
    else k_dhead_bwd_mma<2><<<grid, 256, 0, st>>>(G, w4, gamma, beta, dgamma_acc, dbeta_acc, 0.2f, eval_mode);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}

__global__ void k_sum_f32(const float* __restrict__ x, long long n, float* out) {
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) acc += (double)x[i];
    acc = block_reduce_sum(acc);
    if (threadIdx.x == 0) atomicAdd(out, (float)acc);
}
int sum_f32(const float* x, long long n, float* out, cudaStream_t st) {
    int grid = cdiv(n, 1024); if (grid > 296) grid = 296;
    k_sum_f32<<<grid, 256, 0, st>>>(x, n, out);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
__global__ void k_dhead_unpack_tc(const float* __restrict__ acc, int c, float* __restrict__ dst, float scale, int accumulate) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 9 * c) return;
    int cc = i / 9, t = i % 9;
    float v = scale * acc[t * c + cc];
    dst[i] = accumulate ? dst[i] + v : v;
}
int dhead_unpack_tc(const float* acc, int c, float* dst, float scale, int accumulate, cudaStream_t st) {
    k_dhead_unpack_tc<<<cdiv(9 * c, 256), 256, 0, st>>>(acc, c, dst, scale, accumulate);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}

// ---------------------------------------------------------------------------------------------------
// weight packing: torch layouts -> GEMM slabs
//   KN: dst[slab][k = gemm-cin][n = gemm-cout]   (CUDA-core engine, N contiguous)
//   NK: dst[slab][n = gemm-cout][k = gemm-cin]   (tensor-core engine, K contiguous = UMMA K-major B operand)
// forward conv : w[co][ci][ky][kx], slab = ky*3+kx, gemm-cin = ci, gemm-cout = co
// dgrad conv   : slab t=(e+1)*3+(f+1) <- w[co][ci][1-e][1-f], gemm-cin = co, gemm-cout = ci
// deconv fwd   : w[ci][co][6][6], slab = phase*9 + t, phase = a*2+b, t=(dy+1)*3+(dx+1) <- [2(1-dy)+a][2(1-dx)+b]
// deconv dgrad : slab = phase*9 + t, t=(e+1)*3+(f+1) <- w[ci][co][2(1+e)+a][2(1+f)+b], gemm-cin = co, gemm-cout = ci
// 1x1 forward  : w[co][ci] (lateral conv of the FPN merge), one slab (mode kind 4)
// ---------------------------------------------------------------------------------------------------
// destination element types of the packers: float, bf16, or Split3 = three bf16 planes hi + mid + lo (AFI_PREC_SPLIT)
struct Split3 { bf16 v; };
__device__ __forceinline__ void split3(float x, bf16& hi, bf16& mid, bf16& lo) {
    hi = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(hi);       // exact in fp32
    mid = __float2bfloat16_rn(r1);
    lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
}
template <typename T> __device__ __forceinline__ void pack_put(T* dst, long long i, float v, long long) { dst[i] = (T)v; }
template <> __device__ __forceinline__ void pack_put<Split3>(Split3* dst, long long i, float v, long long pstride) {
    bf16* d = reinterpret_cast<bf16*>(dst);
    bf16 h, m, l;
    split3(v, h, m, l);
    d[i] = h; d[i + pstride] = m; d[i + 2 * pstride] = l;
}
__device__ __forceinline__ long long pack_total_dev(int co, int ci, int mode) {
    int kind = mode >> 1;
    int slabs = (kind == 4 || kind == 5) ? 1 : (kind >= 2 ? 36 : 9);
    return (long long)slabs * co * ci;
}
template <typename T>
__device__ __forceinline__ void pack_one(const float* __restrict__ w, int co, int ci, int mode, T* __restrict__ dst, long long i, long long pstride = 0) {
    int nk = mode & 1, kind = mode >> 1;
    int gk = (kind == 0 || kind == 2 || kind == 4) ? ci : co;   // gemm-cin
    int gn = (kind == 0 || kind == 2 || kind == 4) ? co : ci;   // gemm-cout   (kinds 1, 3, 5 are dgrads: roles swapped)
    int inner = (int)(i % (nk ? gk : gn));
    long long t2 = i / (nk ? gk : gn);
    int outer = (int)(t2 % (nk ? gn : gk));
    int slab = (int)(t2 / (nk ? gn : gk));
    int k = nk ? inner : outer, n = nk ? outer : inner;
    float v;
    if (kind == 0) {            // fwd: k = ci, n = co
        v = w[((long long)n * ci + k) * 9 + slab];
    } else if (kind == 1) {     // dgrad: k = co, n = ci
        v = w[((long long)k * ci + n) * 9 + (8 - slab)];
    } else if (kind == 4) {     // 1x1 forward: w[co][ci], k = ci, n = co
        v = w[(long long)n * ci + k];
    } else if (kind == 5) {     // 1x1 dgrad: k = co, n = ci
        v = w[(long long)k * ci + n];
    } else {
        int ph = slab / 9, t = slab % 9, a = ph >> 1, b = ph & 1, d0 = t / 3 - 1, d1 = t % 3 - 1;
        if (kind == 2) {        // deconv fwd: k = ci, n = co
            int ky = 2 * (1 - d0) + a, kx = 2 * (1 - d1) + b;
            v = w[(((long long)k * co + n) * 6 + ky) * 6 + kx];
        } else {                // deconv dgrad: k = co, n = ci
            int ky = 2 * (1 + d0) + a, kx = 2 * (1 + d1) + b;
            v = w[(((long long)n * co + k) * 6 + ky) * 6 + kx];
        }
    }
    pack_put<T>(dst, i, v, pstride);
}
template <typename T>
__global__ void k_pack(const float* __restrict__ w, int co, int ci, int mode, T* __restrict__ dst, long long total) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < total) pack_one<T>(w, co, ci, mode, dst, i, total);
}
static inline long long pack_total(int co, int ci, int mode) {
    int kind = mode >> 1;
    int slabs = (kind == 4 || kind == 5) ? 1 : (kind >= 2 ? 36 : 9);
    return (long long)slabs * co * ci;
}
int pack_weights(const float* w, int co, int ci, int mode, void* dst, int dst_dt, cudaStream_t st) {
    long long total = pack_total(co, ci, mode);
    if (dst_dt == DT_F32) k_pack<float><<<cdiv(total, 256), 256, 0, st>>>(w, co, ci, mode, (float*)dst, total);
    else if (dst_dt == DT_BF16X3) k_pack<Split3><<<cdiv(total, 256), 256, 0, st>>>(w, co, ci, mode, (Split3*)dst, total);
    else k_pack<bf16><<<cdiv(total, 256), 256, 0, st>>>(w, co, ci, mode, (bf16*)dst, total);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
// all the re-layouts of one module in ONE launch (the generator has 38 of them, most far too small to fill a launch of their own)
struct PackGroup { int njobs; int block_begin[AFI_MAX_PACK + 1]; PackJob j[AFI_MAX_PACK]; };
// 3x3 forward / dgrad layouts (incl. dgrad sub-blocks): one thread per (gemm-cout, gemm-cin) pair reads its NINE taps as 36 contiguous
// bytes and scatters them to the nine slabs (the element-wise version re-reads every 32-byte sector of the weights nine times)
template <typename T>
__device__ __forceinline__ void pack_taps9(const PackJob& q, long long u) {
    const int nk = q.mode & 1, dgrad = (q.mode >> 1) == 1;
    T* dst = reinterpret_cast<T*>(q.dst);
    float v[9];
    if (q.sub == 2) {
        // forward sub-block (NK only): gemm-cin range [n0, n0 + ncnt) of this conv, all of its couts, into rows [koff, koff + co) of a
        // [ktot rows][ncnt columns] slab (several convs that read the same input share one packed operand: N = sum of their couts)
        const int kk = (int)(u % q.ncnt), n = (int)(u / q.ncnt);
        const float* src = q.w + ((long long)n * q.ci + q.n0 + kk) * 9;
#pragma unroll
        for (int t = 0; t < 9; t++) v[t] = src[t];
        const long long slab_sz = (long long)q.ktot * q.ncnt, off = (long long)(q.koff + n) * q.ncnt + kk;
#pragma unroll
        for (int t = 0; t < 9; t++) pack_put<T>(dst, (long long)t * slab_sz + off, v[t], q.pstride);
        return;
    }
    const int gk = q.sub ? q.co : (dgrad ? q.co : q.ci);                 // extent of this job's gemm-cin range
    const int gn = q.sub ? q.ncnt : (dgrad ? q.ci : q.co);               // ... gemm-cout range
    const int I = nk ? gk : gn, O = nk ? gn : gk;                        // inner / outer index of a slab
    const int inner = (int)(u % I), outer = (int)(u / I);
    const int k = nk ? inner : outer, n = nk ? outer : inner;
    const float* src = q.w + (dgrad ? ((long long)k * q.ci + (q.sub ? q.n0 : 0) + n) : ((long long)n * q.ci + k)) * 9;
#pragma unroll
    for (int t = 0; t < 9; t++) v[t] = src[t];
    // destination strides: full slabs [O][I], or the [koff, koff + co) gemm-cin slice of a [ncnt][ktot] (NK) / [ktot][ncnt] (KN) slab
    const long long slab_sz = q.sub ? (long long)q.ncnt * q.ktot : (long long)O * I;
    const long long off = q.sub ? (nk ? (long long)n * q.ktot + q.koff + k : (long long)(q.koff + k) * q.ncnt + n) : (long long)outer * I + inner;
#pragma unroll
    for (int t = 0; t < 9; t++) pack_put<T>(dst, (long long)t * slab_sz + off, v[dgrad ? 8 - t : t], q.pstride);
}
template <typename T>
__global__ void k_pack_group(const __grid_constant__ PackGroup G) {
    int k = 0;
    while (k + 1 < G.njobs && (int)blockIdx.x >= G.block_begin[k + 1]) k++;
    const PackJob& q = G.j[k];
    long long i = (long long)(blockIdx.x - G.block_begin[k]) * 1024 + threadIdx.x;
    const int kind = q.mode >> 1;
    if (kind <= 1) {
        const long long units = q.sub ? (long long)q.co * q.ncnt : (long long)q.co * q.ci;
#pragma unroll
        for (int u = 0; u < 4; u++, i += 256)
            if (i < units) pack_taps9<T>(q, i);
        return;
    }
    const long long total = pack_total_dev(q.co, q.ci, q.mode);
#pragma unroll
    for (int u = 0; u < 4; u++, i += 256)
        if (i < total) pack_one<T>(q.w, q.co, q.ci, q.mode, reinterpret_cast<T*>(q.dst), i, q.pstride);
}
int pack_weights_group(int njobs, const PackJob* jobs, int dst_dt, cudaStream_t st) {
    AFI_REQUIRE(njobs >= 0 && njobs <= AFI_MAX_PACK, "pack_weights_group: %d jobs (max %d)", njobs, AFI_MAX_PACK);
    if (njobs == 0) return AFI_OK;
    PackGroup G; memset(&G, 0, sizeof(G));
    G.njobs = njobs;
    int b = 0;
    for (int k = 0; k < njobs; k++) {
        G.j[k] = jobs[k]; G.block_begin[k] = b;
        // plane stride of a three-plane destination = the whole packed operand of this job
        G.j[k].pstride = jobs[k].sub ? (long long)9 * jobs[k].ncnt * jobs[k].ktot : pack_total(jobs[k].co, jobs[k].ci, jobs[k].mode);
        if (jobs[k].sub == 1) AFI_REQUIRE((jobs[k].mode >> 1) == 1, "pack_weights_group: sub-block packing (sub = 1) is a dgrad mode");
        if (jobs[k].sub == 2) AFI_REQUIRE(jobs[k].mode == 1, "pack_weights_group: forward sub-block packing (sub = 2) takes the NK forward mode");
        const int kind = jobs[k].mode >> 1;
        const long long work = kind <= 1 ? (jobs[k].sub ? (long long)jobs[k].co * jobs[k].ncnt : (long long)jobs[k].co * jobs[k].ci)
                                         : pack_total(jobs[k].co, jobs[k].ci, jobs[k].mode);
        b += cdiv(work, 1024);
    }
    G.block_begin[njobs] = b;
    if (dst_dt == DT_F32) k_pack_group<float><<<b, 256, 0, st>>>(G);
    else if (dst_dt == DT_BF16X3) k_pack_group<Split3><<<b, 256, 0, st>>>(G);
    else k_pack_group<bf16><<<b, 256, 0, st>>>(G);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}

// ---------------------------------------------------------------------------------------------------
// AFI_PREC_SPLIT operand staging: fp32 pixel-major views -> dense bf16 planes [3][pixels][cpad].  HBM-bound: 4 B read + 6 B written
// per element; one thread per 8 channels (32 B in, 3 x 16 B out); every view of every problem of a grouped GEMM in ONE launch.
// ---------------------------------------------------------------------------------------------------
int split_cpad(int c) { return c <= 64 ? c : (c + 63) / 64 * 64; }
size_t split_planes_bytes(long long pixels, int channels) { return ((size_t)pixels * split_cpad(channels) * 2 * 3 + 255) / 256 * 256; }
struct SplitGroup { int njobs; int block_begin[AFI_MAX_SPLIT + 1]; SplitJob j[AFI_MAX_SPLIT]; };
__global__ void __launch_bounds__(256) k_split3_group(const __grid_constant__ SplitGroup G) {
    int k = 0;
    while (k + 1 < G.njobs && (int)blockIdx.x >= G.block_begin[k + 1]) k++;
    const SplitJob& q = G.j[k];
    const int cpad = q.c <= 64 ? q.c : (q.c + 63) / 64 * 64;
    const int cq = cpad >> 3;
    const long long P = (long long)q.n * q.h * q.w, units = P * cq;
    const long long pstride = P * cpad;
    bf16* d = reinterpret_cast<bf16*>(q.dst);
    for (long long u = (long long)(blockIdx.x - G.block_begin[k]) * 1024 + threadIdx.x, e = u + 1024; u < e && u < units; u += 256) {
        const int c0 = (int)(u % cq) * 8;
        const long long pix = u / cq;
        const int x = (int)(pix % q.w), y = (int)((pix / q.w) % q.h), n = (int)(pix / ((long long)q.w * q.h));
        float v[8];
        if (c0 < q.c) {
            const float4* s4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(q.src.ptr) + n * q.src.sn + y * q.src.sy + x * q.src.sx + c0);
            const float4 a = s4[0], b = s4[1];
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = 0.f;
        }
        uint32_t ph[4], pm[4], pl[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            bf16 h0, m0, l0, h1, m1, l1;
            split3(v[2 * i], h0, m0, l0);
            split3(v[2 * i + 1], h1, m1, l1);
            ph[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            pm[i] = (uint32_t)__bfloat16_as_ushort(m0) | ((uint32_t)__bfloat16_as_ushort(m1) << 16);
            pl[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        const long long o = pix * cpad + c0;
        *reinterpret_cast<uint4*>(d + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        *reinterpret_cast<uint4*>(d + pstride + o) = make_uint4(pm[0], pm[1], pm[2], pm[3]);
        if (q.nplanes > 2) *reinterpret_cast<uint4*>(d + 2 * pstride + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
}
int split3_group(int njobs, const SplitJob* jobs, cudaStream_t st) {
    AFI_REQUIRE(njobs >= 0 && njobs <= AFI_MAX_SPLIT, "split3_group: %d jobs (max %d)", njobs, AFI_MAX_SPLIT);
    SplitGroup G; memset(&G, 0, sizeof(G));
    int b = 0, nj = 0;
    for (int k = 0; k < njobs; k++) {
        const SplitJob& q = jobs[k];
        const long long P = (long long)q.n * q.h * q.w;
        if (P == 0) continue;
        AFI_REQUIRE(q.c % 8 == 0 && q.src.sx % 4 == 0 && q.src.sy % 4 == 0 && q.src.sn % 4 == 0 && ((uintptr_t)q.src.ptr & 15) == 0,
                    "split3_group: view %d is not 16-byte aligned per pixel (c %d)", k, q.c);
        G.j[nj] = q; G.block_begin[nj] = b;
        b += cdiv(P * (split_cpad(q.c) / 8), 1024);
        nj++;
    }
    if (nj == 0) return AFI_OK;
    G.njobs = nj; G.block_begin[nj] = b;
    k_split3_group<<<b, 256, 0, st>>>(G);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}

// packed fp32 weight gradient -> torch layout.  Conv: dst[co][ci][t]; deconv: dst[ci][co][ky][kx] (forward tap mapping).
__global__ void k_unpack_wgrad(const float* __restrict__ packed, int co, int ci, int nk, int deconv, float* __restrict__ dst,
                               float scale, int accumulate, long long total) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    int slab, c_o, c_i;
    if (!deconv) {
        slab = (int)(i % 9); long long t = i / 9; c_i = (int)(t % ci); c_o = (int)(t / ci);
    } else {
        int kx = (int)(i % 6); long long t = i / 6; int ky = (int)(t % 6); t /= 6; c_o = (int)(t % co); c_i = (int)(t / co);
        int a = ky & 1, b = kx & 1, d0 = 1 - (ky >> 1), d1 = 1 - (kx >> 1);
        slab = (a * 2 + b) * 9 + (d0 + 1) * 3 + (d1 + 1);
    }
    long long src = nk ? ((long long)slab * co + c_o) * ci + c_i : ((long long)slab * ci + c_i) * co + c_o;
    float v = packed[src] * scale;
    dst[i] = accumulate ? dst[i] + v : v;
}
int unpack_wgrad(const float* packed, int co, int ci, int layout_nk, int deconv, float* dst, float scale, int accumulate, cudaStream_t st) {
    long long total = (long long)co * ci * (deconv ? 36 : 9);
    k_unpack_wgrad<<<cdiv(total, 256), 256, 0, st>>>(packed, co, ci, layout_nk, deconv, dst, scale, accumulate, total);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
__global__ void k_unpack_wgrad_sub(const float* __restrict__ packed, int CO, int CI, int co_off, int co, int ci, int nk, float* __restrict__ dst,
                                   float scale, int accumulate, long long total) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    int slab = (int)(i % 9); long long t = i / 9; int c_i = (int)(t % ci); int c_o = (int)(t / ci);
    long long src = nk ? ((long long)slab * CO + co_off + c_o) * CI + c_i : ((long long)slab * CI + c_i) * CO + co_off + c_o;
    float v = packed[src] * scale;
    dst[i] = accumulate ? dst[i] + v : v;
}
int unpack_wgrad_sub(const float* packed, int CO, int CI, int co_off, int co, int ci, int layout_nk, float* dst, float scale, int accumulate,
                     cudaStream_t st) {
    long long total = (long long)co * ci * 9;
    k_unpack_wgrad_sub<<<cdiv(total, 256), 256, 0, st>>>(packed, CO, CI, co_off, co, ci, layout_nk, dst, scale, accumulate, total);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
// all the gradient un-layouts of one module in ONE launch (the generator has 23 gradient tensors: one launch instead of 45)
struct UnpackGroup { int njobs; int block_begin[AFI_MAX_UNPACK + 1]; UnpackJob j[AFI_MAX_UNPACK]; };
__global__ void __launch_bounds__(256) k_unpack_group(const __grid_constant__ UnpackGroup G, float scale, int accumulate) {
    int k = 0;
    while (k + 1 < G.njobs && (int)blockIdx.x >= G.block_begin[k + 1]) k++;
    const UnpackJob& q = G.j[k];
    long long i = (long long)(blockIdx.x - G.block_begin[k]) * 1024 + threadIdx.x;
#pragma unroll
    for (int u = 0; u < 4; u++, i += 256) {
        if (i >= q.total) return;
        long long src;
        if (q.kind == UNPACK_COPY) src = i;
        else if (q.kind == UNPACK_HEAD_TC) { const int cc = (int)(i / 9), t = (int)(i % 9); src = (long long)t * q.ci + cc; }
        else {
            int slab, c_o, c_i;
            if (q.kind == UNPACK_DECONV) {
                int kx = (int)(i % 6); long long t = i / 6; int ky = (int)(t % 6); t /= 6; c_o = (int)(t % q.co); c_i = (int)(t / q.co);
                int a = ky & 1, b = kx & 1, d0 = 1 - (ky >> 1), d1 = 1 - (kx >> 1);
                slab = (a * 2 + b) * 9 + (d0 + 1) * 3 + (d1 + 1);
            } else {
                slab = (int)(i % 9); long long t = i / 9; c_i = (int)(t % q.ci); c_o = (int)(t / q.ci);
            }
            c_o += q.co_off;
            src = q.nk ? ((long long)slab * q.CO + c_o) * q.CI + c_i : ((long long)slab * q.CI + c_i) * q.CO + c_o;
        }
        const float v = q.src[src] * scale;
        q.dst[i] = accumulate ? q.dst[i] + v : v;
    }
}
int unpack_group(int njobs, const UnpackJob* jobs, float scale, int accumulate, cudaStream_t st) {
    AFI_REQUIRE(njobs >= 0 && njobs <= AFI_MAX_UNPACK, "unpack_group: %d jobs (max %d)", njobs, AFI_MAX_UNPACK);
    UnpackGroup G; memset(&G, 0, sizeof(G));
    int b = 0, nj = 0;
    for (int k = 0; k < njobs; k++) {
        if (!jobs[k].dst || jobs[k].total == 0) continue;
        G.j[nj] = jobs[k]; G.block_begin[nj] = b;
        b += cdiv(jobs[k].total, 1024);
        nj++;
    }
    if (nj == 0) return AFI_OK;
    G.njobs = nj; G.block_begin[nj] = b;
    k_unpack_group<<<b, 256, 0, st>>>(G, scale, accumulate);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
// 1x1 weight gradient: packed [ci][co] (KN) or [co][ci] (NK) -> torch [co][ci]
__global__ void k_unpack_1x1(const float* __restrict__ packed, int co, int ci, int nk, float* __restrict__ dst, float scale, int accumulate) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)co * ci) return;
    int c_i = (int)(i % ci), c_o = (int)(i / ci);
    float v = scale * (nk ? packed[i] : packed[(long long)c_i * co + c_o]);
    dst[i] = accumulate ? dst[i] + v : v;
}
int unpack_1x1(const float* packed, int co, int ci, int layout_nk, float* dst, float scale, int accumulate, cudaStream_t st) {
    k_unpack_1x1<<<cdiv((long long)co * ci, 256), 256, 0, st>>>(packed, co, ci, layout_nk, dst, scale, accumulate);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
__global__ void k_axpby(const float* __restrict__ src, float* __restrict__ dst, long long n, float scale, int accumulate) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = accumulate ? dst[i] + scale * src[i] : scale * src[i];
}
int axpby_f32(const float* src, float* dst, long long n, float scale, int accumulate, cudaStream_t st) {
    if (n == 0) return AFI_OK;
    k_axpby<<<cdiv(n, 256), 256, 0, st>>>(src, dst, n, scale, accumulate);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}

// ---------------------------------------------------------------------------------------------------
// losses and the optimiser step
// ---------------------------------------------------------------------------------------------------
__global__ void k_bce(const float* __restrict__ x, long long count, float target, float* loss_out, float* loss_sum, float weight,
                      float* __restrict__ dlogits, float gscale) {
    double acc = 0.0;
    float inv = 1.f / (float)count;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        float v = x[i];
        float l = fmaxf(v, 0.f) - v * target + log1pf(expf(-fabsf(v)));
        acc += (double)l;
        if (dlogits) dlogits[i] = gscale * (1.f / (1.f + expf(-v)) - target) * inv;
    }
    acc = block_reduce_sum(acc);
    if (threadIdx.x == 0) {
        float m = (float)(acc / (double)count);
        if (loss_out) atomicAdd(loss_out, m);
        if (loss_sum) atomicAdd(loss_sum, weight * m);
    }
}
extern "C" int afi_bce_with_logits(const float* logits, long long count, float target, float* loss_out, float* loss_sum,
                                   float weight, float* dlogits, float grad_scale, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(count > 0, "bce: empty logits");
    if (loss_out) AFI_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
    int grid = cdiv(count, 256); if (grid > 592) grid = 592;
    k_bce<<<grid, 256, 0, st>>>(logits, count, target, loss_out, loss_sum, weight, dlogits, grad_scale);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}

__global__ void k_l1(afi_view4 a, afi_view4 b, int c, int h, int w, long long total, float* loss_out, float* loss_sum, float weight,
                     float* __restrict__ da, float gscale) {
    double acc = 0.0;
    float ginv = gscale / (float)total;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int x = (int)(i % w); long long t = i / w; int y = (int)(t % h); t /= h; int cc = (int)(t % c); int n = (int)(t / c);
        float d = vld(a, n * a.sn + cc * a.sc + y * a.sh + x * a.sw) - vld(b, n * b.sn + cc * b.sc + y * b.sh + x * b.sw);
        acc += (double)fabsf(d);
        if (da) da[i] = d > 0.f ? ginv : (d < 0.f ? -ginv : 0.f);
    }
    acc = block_reduce_sum(acc);
    if (threadIdx.x == 0) {
        float m = (float)(acc / (double)total);
        if (loss_out) atomicAdd(loss_out, m);
        if (loss_sum) atomicAdd(loss_sum, weight * m);
    }
}
// both operands contiguous NCHW: a flat float4 stream, four independent 16-byte loads per operand in flight per thread
__global__ void __launch_bounds__(256) k_l1_dense(const float4* __restrict__ a, const float4* __restrict__ b, long long n4, long long total,
                                                  float* loss_out, float* loss_sum, float weight, float4* __restrict__ da, float gscale) {
    double acc = 0.0;
    const float ginv = gscale / (float)total;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < n4; i0 += 4 * stride) {
        float4 va[4], vb[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { long long i = i0 + u * stride; if (i < n4) { va[u] = a[i]; vb[u] = b[i]; } }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            long long i = i0 + u * stride;
            if (i < n4) {
                float d0 = va[u].x - vb[u].x, d1 = va[u].y - vb[u].y, d2 = va[u].z - vb[u].z, d3 = va[u].w - vb[u].w;
                acc += (double)((fabsf(d0) + fabsf(d1)) + (fabsf(d2) + fabsf(d3)));
                if (da) da[i] = make_float4(d0 > 0.f ? ginv : (d0 < 0.f ? -ginv : 0.f), d1 > 0.f ? ginv : (d1 < 0.f ? -ginv : 0.f),
                                            d2 > 0.f ? ginv : (d2 < 0.f ? -ginv : 0.f), d3 > 0.f ? ginv : (d3 < 0.f ? -ginv : 0.f));
            }
        }
    }
    double accd = block_reduce_sum(acc);
    if (threadIdx.x == 0) {
        float m = (float)(accd / (double)total);
        if (loss_out) atomicAdd(loss_out, m);
        if (loss_sum) atomicAdd(loss_sum, weight * m);
    }
}
static inline bool view4_contiguous(const afi_view4& v, int c, int h, int w) {
    return v.sw == 1 && v.sh == w && v.sc == (long long)h * w && v.sn == (long long)c * h * w;
}
extern "C" int afi_l1_loss(afi_view4 a, afi_view4 b, int n, int c, int h, int w, float* loss_out, float* loss_sum, float weight,
                           float* da, float grad_scale, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    long long total = (long long)n * c * h * w;
    AFI_REQUIRE(total > 0, "l1: empty tensors");
    AFI_REQUIRE(view_dtype_ok(a) && view_dtype_ok(b), "l1: unknown view dtype");
    if (loss_out) AFI_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
    if (total % 4 == 0 && a.dtype == AFI_DT_F32 && b.dtype == AFI_DT_F32 && view4_contiguous(a, c, h, w) && view4_contiguous(b, c, h, w) && ((uintptr_t)a.ptr & 15) == 0 &&
        ((uintptr_t)b.ptr & 15) == 0 && (!da || ((uintptr_t)da & 15) == 0)) {
        long long n4 = total / 4;
        int grid = cdiv(n4, 256 * 4); if (grid > 148 * 8) grid = 148 * 8;
        k_l1_dense<<<grid, 256, 0, st>>>((const float4*)a.ptr, (const float4*)b.ptr, n4, total, loss_out, loss_sum, weight, (float4*)da, grad_scale);
        AFI_LAUNCH_CHECK();
        return AFI_OK;
    }
    int grid = cdiv(total, 256 * 8); if (grid > 148 * 8) grid = 148 * 8;
    k_l1<<<grid, 256, 0, st>>>(a, b, c, h, w, total, loss_out, loss_sum, weight, da, grad_scale);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}

__global__ void k_sgd(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, long long n, float lr, float mom,
                      float wd, float gscale, int first) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float d = gscale * g[i] + wd * p[i];
    float b = first ? d : mom * m[i] + d;
    m[i] = b;
    p[i] = p[i] - lr * b;
}
extern "C" int afi_sgd_step(float* p, const float* g, float* m, long long count, float lr, float momentum, float wd,
                            float grad_scale, int first, void* stream) {
    if (count == 0) return AFI_OK;
    k_sgd<<<cdiv(count, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, count, lr, momentum, wd, grad_scale, first);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
// every parameter of one optimiser in ONE launch: blockIdx.x -> (tensor, 1024-element chunk) through a prefix table
struct SgdGroup { int n; int block_begin[AFI_MAX_SGD + 1]; float* p[AFI_MAX_SGD]; const float* g[AFI_MAX_SGD]; float* m[AFI_MAX_SGD];
                  long long count[AFI_MAX_SGD]; float wd[AFI_MAX_SGD]; };
__global__ void k_sgd_group(const __grid_constant__ SgdGroup G, float lr, float mom, float gscale, int first) {
    int k = 0;
    while (k + 1 < G.n && (int)blockIdx.x >= G.block_begin[k + 1]) k++;
    float* __restrict__ p = G.p[k]; const float* __restrict__ g = G.g[k]; float* __restrict__ m = G.m[k];
    const long long n = G.count[k]; const float wd = G.wd[k];
    long long i = (long long)(blockIdx.x - G.block_begin[k]) * 1024 + threadIdx.x;
#pragma unroll
    for (int u = 0; u < 4; u++, i += 256) {
        if (i < n) {
            float d = gscale * g[i] + wd * p[i];
            float b = first ? d : mom * m[i] + d;
            m[i] = b;
            p[i] = p[i] - lr * b;
        }
    }
}
extern "C" int afi_sgd_step_multi(int n, float* const* p, const float* const* g, float* const* m, const long long* count, const float* wd,
                                  float lr, float momentum, float grad_scale, int first, void* stream) {
    AFI_REQUIRE(n >= 0 && (n == 0 || (p && g && m && count && wd)), "afi_sgd_step_multi: null argument");
    for (int o = 0; o < n; o += AFI_MAX_SGD) {
        SgdGroup G; memset(&G, 0, sizeof(G));
        int b = 0, k = 0;
        for (; k < AFI_MAX_SGD && o + k < n; k++) {
            G.p[k] = p[o + k]; G.g[k] = g[o + k]; G.m[k] = m[o + k]; G.count[k] = count[o + k]; G.wd[k] = wd[o + k];
            G.block_begin[k] = b; b += cdiv(count[o + k], 1024);
        }
        G.n = k; G.block_begin[k] = b;
        if (b == 0) continue;
        k_sgd_group<<<b, 256, 0, (cudaStream_t)stream>>>(G, lr, momentum, grad_scale, first);
        AFI_LAUNCH_CHECK();
    }
    return AFI_OK;
}
extern "C" int afi_zero(void* ptr, size_t bytes, void* stream) {
    if (bytes) AFI_CUDA(cudaMemsetAsync(ptr, 0, bytes, (cudaStream_t)stream));
    return AFI_OK;
}

}  // namespace afi
