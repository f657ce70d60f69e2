// Internal declarations shared by the translation units of libafigan_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/afigan_b200.h"

namespace afi {

typedef __nv_bfloat16 bf16;

// ---- error plumbing ---------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern long long g_launches;
#define AFI_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            afi::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return AFI_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)
#define AFI_LAUNCH_CHECK()                                                                     \
    do {                                                                                       \
        afi::g_launches++;                                                                     \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            afi::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return AFI_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)
#define AFI_TRY(expr)                \
    do {                             \
        int _r = (expr);             \
        if (_r != AFI_OK) return _r; \
    } while (0)
#define AFI_REQUIRE(cond, ...)        \
    do {                              \
        if (!(cond)) {                \
            afi::set_error(__VA_ARGS__); \
            return AFI_ERR_INVALID;   \
        }                             \
    } while (0)

// ---- pixel-major (NHWC) views -------------------------------------------------------------------------
// addr(n, y, x, c) = ptr + n*sn + y*sy + x*sx + c, strides in ELEMENTS of the view's dtype.
struct PView {
    void* ptr;
    long long sn, sy, sx;
};
// DT_BF16X3: an fp32 value stored as THREE bf16 planes hi + mid + lo (AFI_PREC_SPLIT GEMM operands; plane p of a tensor follows
// plane p - 1 after `plane stride` elements)
enum DType { DT_F32 = 0, DT_BF16 = 1, DT_BF16X3 = 2 };

static inline PView pview(void* p, int h, int w, int cs) {
    PView v; v.ptr = p; v.sx = cs; v.sy = (long long)w * cs; v.sn = (long long)h * w * cs; return v;
}
static inline PView pview_null() { PView v; v.ptr = nullptr; v.sn = v.sy = v.sx = 0; return v; }
// channel-offset sub-view
static inline PView pview_ch(PView v, int coff, int elem_bytes) {
    v.ptr = (char*)v.ptr + (size_t)coff * elem_bytes; return v;
}

struct Tap { int dy, dx, view, slab; };
#define AFI_MAX_TAPS 36

// One GROUPED implicit-GEMM convolution over up to AFI_MAX_PROB independent problems that share weights, taps and the
// epilogue recipe (the pyramid levels of one generator call, or the level x {real, fake} calls of the discriminator):
//   out_p[n,y,x,co] = epi( sum_taps sum_ci in_p[view][n,y+dy,x+dx,ci] * W[slab][.][.] )        for every problem p
// Epilogue order:  v = acc + bias;  v = act ? lrelu(v) : v;  v *= alpha;  v += beta1*r1 + beta2*r2 + accin;
//                  v *= mask > 0 ? 1 : mask_slope;   store (bf16/f32 as out_dt says).
#define AFI_MAX_PROB 10
struct ConvProb {
    int N, H, W, pad_;    // logical pixel grid (output pixels; every input view has the same grid)
    PView in[4];          // dtype = storage dtype T
    PView out, r1, r2;
    PView accin;          // f32
    PView mask;           // dtype T
    // fused per-channel reductions of the stored output v (tensor-core engine only; see ConvArgs.stat_mode)
    double* stat0; double* stat1;
    // AFI_PREC_SPLIT only: caller-owned scratch in which the fp32 input views of THIS problem are split into bf16 planes
    void* sws; size_t sws_bytes;
    PView out2;           // tensor-core engine, ConvArgs.split_col > 0: where the columns below split_col go (dtype ConvArgs.out2_dt)
    // tensor-core engine: extent of input view v when it differs from the problem grid (0 = H / W).  Reads outside a view's own extent
    // are zero (TMA out-of-bounds fill): the sub-pixel phase views of a stride-2 convolution have ceil / floor halves of an odd size.
    int vh[4], vw[4];
};
struct ConvArgs {
    int cin, cout;
    int ntaps, nprob;
    Tap taps[AFI_MAX_TAPS];
    const void* w;        // packed slabs, dtype T.  SIMT engine: [slab][cin][cout]; tensor-core engine: [slab][cout][cin]
    const float* bias;
    int act; float slope;
    float alpha, beta1, beta2, mask_slope;
    int out_dt, r1_dt, r2_dt;
    // tensor-core engine only.  act_post: the activation is applied AFTER alpha / residuals / accin (v = lrelu(alpha * acc + accin)).
    // split_col > 0 (a multiple of 16): output columns [0, split_col) get the LeakyReLU and are stored to p[].out2 (dtype out2_dt);
    // the columns from split_col on are stored raw to p[].out.  One GEMM then serves several consumers of the same input: the dense
    // block's x-part GEMM writes growth channel 1 (activated) into the block buffer and the partial sums of channels 2-4 as fp32.
    int act_post, split_col, out2_dt;
    int nslab;            // slabs of the packed weight operand when the taps do not use them all (0 = highest slab used + 1)
    // 0: none.  1: stat0[c] += sum_p v, stat1[c] += sum_p v^2 (BatchNorm batch statistics of a conv output).
    int stat_mode;
    // nphase = 4 (tensor-core engine only): FOUR convolutions of the same input in one launch -- the sub-pixel phases of the stride-2
    // transposed conv.  N tile p uses the slabs 9 p + tap.slab and stores its cout columns at out + (p >> 1) * out.sy / 2 +
    // (p & 1) * out.sx / 2 (out is the stride-2 view of phase 0 in the 2h x 2w map).  0 = a single convolution.
    int nphase;
    // split-precision operands (set by conv_tc_split, never by callers): 0 = plain bf16; 3 = every input view is three bf16 planes
    // [hi | mid | lo] stacked along the image axis (plane p of image n = image p * N + n of the view), the packed weights are three
    // planes stacked along the slab axis, and the product is evaluated as the six plane pairs whose weight is >= 2^-16 (hi hi, hi mid,
    // mid hi, mid mid, hi lo, lo hi) with fp32 accumulation in TMEM: fp32-accurate products on the bf16 tensor cores.
    int split;
    // split_pairs = 6: all six plane pairs (fp32-accurate products: the forward convs of a training step, whose pre-activation SIGNS
    // decide the LeakyReLU slopes of the backward pass -- an error eps there moves the gradients by ~sqrt(eps));
    // split_pairs = 3: hi hi + hi mid + mid hi only (products accurate to ~2^-17: everything a gradient or a forward-only result depends
    // on LINEARLY -- dgrads, weight gradients, inference); the lo planes are then neither produced nor read.  Set by run_conv / run_wgrad.
    int split_pairs;
    int aux_f32;          // residual / mask operands are fp32 (split mode: activations stay fp32 in HBM)
    ConvProb p[AFI_MAX_PROB];
};

// dW[slab][..] += sum over ALL problems' pixels of dY[p][co] * X[p + tap][ci]   (fp32 atomics into a pre-zeroed accumulator)
// SIMT engine layout: [slab][cin][cout]; tensor-core engine: [slab][cout][cin].
struct WgradProb {
    int N, H, W, pad_;
    PView x;              // T
    PView dy;             // T
    void* sws; size_t sws_bytes;   // AFI_PREC_SPLIT: scratch for the bf16 planes of x and dy
    int xh, xw;                    // tensor-core engine: extent of x when it differs from the dY grid (0 = H / W); reads outside are zero
};
struct WgradArgs {
    int cin, cout;
    int ntaps, nprob;
    Tap taps[9];          // .view unused, .slab = output slab
    float* dw;
    int split, split_pairs;   // see ConvArgs.split / split_pairs (set by run_wgrad / wgrad_tc_split)
    WgradProb p[AFI_MAX_PROB];
};

// optional per-launch CUDA-event profiling of the GEMM kernels (afi_profile_begin/_end; bench.py's roofline leg)
enum ProfKind { PROF_CONV_TC = 0, PROF_WGRAD_TC = 1, PROF_CONV_SIMT = 2, PROF_WGRAD_SIMT = 3, PROF_CONV_PAIR = 4, PROF_CONV_HALO = 5 };
extern bool g_prof_on;
void prof_begin(int kind, double flops, int cin, int cout, long long pixels, cudaStream_t st);
void prof_end(cudaStream_t st);
struct ProfScope {
    cudaStream_t st; bool on;
    ProfScope(int kind, double flops, int cin, int cout, long long pixels, cudaStream_t s) : st(s), on(g_prof_on) {
        if (on) prof_begin(kind, flops, cin, cout, pixels, st);
    }
    ~ProfScope() { if (on) prof_end(st); }
};

void conv_args_init(ConvArgs& a);
void set_std_taps(Tap* taps, int view, int slab0);

// engines
template <typename T> int conv_simt(const ConvArgs& a, cudaStream_t st);
template <typename T> int wgrad_simt(const WgradArgs& a, cudaStream_t st);
int conv_tc(afi_ctx* ctx, const ConvArgs& a, cudaStream_t st);
int wgrad_tc(afi_ctx* ctx, const WgradArgs& a, cudaStream_t st);
// AFI_PREC_SPLIT: fp32 views in / fp32 out.  The input views are split into bf16 planes (p[k].sws scratch), then the tcgen05 engine
// runs the six-pair product.  a.w points at the packed [3 planes][slab][cout][cin] bf16 weights.
int conv_tc_split(afi_ctx* ctx, const ConvArgs& a, cudaStream_t st);
int wgrad_tc_split(afi_ctx* ctx, const WgradArgs& a, cudaStream_t st);
// bytes of split scratch one problem needs: conv (nviews input views of cin channels) / wgrad (x of cin + dy of cout channels)
size_t split_planes_bytes(long long pixels, int channels);
// fp32 strided views -> dense bf16 planes [nplanes = 2 | 3][pixels][cpad] (cpad = split_cpad(c); pad channels zeroed), all jobs in ONE launch
struct SplitJob { PView src; int n, h, w, c; void* dst; int nplanes; };
#define AFI_MAX_SPLIT 40
int split_cpad(int c);
int split3_group(int njobs, const SplitJob* jobs, cudaStream_t st);
int tc_init(afi_ctx* ctx);

// ---- elementwise / layout kernels (elementwise.cu) ------------------------------------------------------
// element `off` of a boundary view as fp32 (the dtype branch is uniform over a launch)
#ifdef __CUDACC__
__device__ __forceinline__ float vld(const afi_view4& v, long long off) {
    return v.dtype == AFI_DT_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(v.ptr)[off]) : reinterpret_cast<const float*>(v.ptr)[off];
}
#endif
static inline bool view_dtype_ok(const afi_view4& v) { return (v.dtype == AFI_DT_F32 || v.dtype == AFI_DT_BF16) && v.reserved == 0; }
template <typename T> int nchw_to_nhwc(afi_view4 src, int n, int c, int h, int w, PView dst, cudaStream_t st);
// BiFPN inference building blocks (elementwise.cu): depthwise 3x3 (+ swish in front) fused with the NCHW -> NHWC conversion; bottom-up fusion site
template <typename T> int dw3x3_to_nhwc(afi_view4 src, const float* wdw, int n, int c, int h, int w, int pre_swish, PView dst, cudaStream_t st);
int bifpn_fuse_down(afi_view4 a, afi_view4 b, afi_view4 dn, const float* wts, int nw, int n, int c, int h, int w, int dh, int dw_, float* out,
                    cudaStream_t st);
// dst[n,c,y,x] (contiguous [n,c,oh,ow]) = scale * ( a[n,y,x,c] (T) [+ lat[n,y,x,c] (T)] [+ bilinear2x(skip)[n,c,y,x]] )
// ... optionally followed by dst = fw[0] * add[n,c,y,x] + fw[1] * dst (fw: two device floats; the BiFPN fusion site)
template <typename T> int nhwc_to_nchw(PView a, PView lat, afi_view4 skip, int skip_h, int skip_w, float scale,
                                       int n, int c, int oh, int ow, float* dst, cudaStream_t st, const afi_view4* add = nullptr,
                                       const float* fw = nullptr);
// dst = scale * (a + b) * mask ; a,b f32 or T views (dtype flags), mask T view (optional), dst T or f32
int ew_combine(PView dst, int dst_dt, PView a, int a_dt, PView b, int b_dt, PView mask, int mask_dt, float mask_slope,
               float scale, int n, int h, int w, int c, cudaStream_t st);
int ew_combine_group(int nprob, const PView* dst, int dst_dt, const PView* a, int a_dt, const PView* b, int b_dt, const PView* mask, int mask_dt,
                     float mask_slope, float scale, const int* n, const int* h, const int* w, int c, cudaStream_t st);
// float accumulate variant for bias gradients: out[c] += sum_p x[p][c]
int col_sum_f32(PView x, int dt, int n, int h, int w, int c, float* out, cudaStream_t st);
// weight re-layout: dst[slab][r][c] from torch [co][ci][k][k] (see pack modes in elementwise.cu)
enum PackMode { PACK_FWD_KN = 0, PACK_FWD_NK = 1, PACK_DGRAD_KN = 2, PACK_DGRAD_NK = 3,
                PACK_DECONV_FWD_KN = 4, PACK_DECONV_FWD_NK = 5, PACK_DECONV_DGRAD_KN = 6, PACK_DECONV_DGRAD_NK = 7,
                PACK_1X1_KN = 8, PACK_1X1_NK = 9, PACK_1X1_DGRAD_KN = 10, PACK_1X1_DGRAD_NK = 11 };
int pack_weights(const float* w, int co, int ci, int mode, void* dst, int dst_dt, cudaStream_t st);
#define AFI_MAX_PACK 96
#define AFI_MAX_SGD 48
// sub != 0 (dgrad modes only): pack the gemm-cout range [n0, n0 + ncnt) of this weight into the gemm-cin slice [koff, koff + co) of a
// destination whose gemm-cin extent is ktot (several convs that read the same gradient buffer share one packed operand)
// pstride (DT_BF16X3 destinations): elements between the three planes of the destination
struct PackJob { const float* w; void* dst; int co, ci, mode, sub; int n0, ncnt, koff, ktot; long long pstride; };
int pack_weights_group(int njobs, const PackJob* jobs, int dst_dt, cudaStream_t st);
// gradient un-layout (fp32): torch-layout grad = [grad +] scale * packed ; layout_nk: packed is [slab][cout][cin]
int unpack_wgrad(const float* packed, int co, int ci, int layout_nk, int deconv, float* dst, float scale, int accumulate, cudaStream_t st);
// the [co][ci] sub-block starting at row co_off of a packed [slab][CO][CI] (nk) / [slab][CI][CO] gradient -> torch [co][ci][3][3]
int unpack_wgrad_sub(const float* packed, int CO, int CI, int co_off, int co, int ci, int layout_nk, float* dst, float scale, int accumulate, cudaStream_t st);
int axpby_f32(const float* src, float* dst, long long n, float scale, int accumulate, cudaStream_t st);
// grouped form of the un-layouts above: dst = [dst +] scale * unpack(src) for every job of a module in one launch.
//   UNPACK_STD: packed [9][CO][CI] (nk) / [9][CI][CO] -> torch [co][ci][3][3], rows co_off .. co_off + co of the packed operand (CO x CI its extents)
//   UNPACK_DECONV: packed 36 slabs -> torch ConvTranspose [ci][co][6][6];  UNPACK_COPY: total floats;  UNPACK_HEAD_TC: [16][ci] -> [ci][9]
enum UnpackKind { UNPACK_STD = 0, UNPACK_DECONV = 1, UNPACK_COPY = 2, UNPACK_HEAD_TC = 3 };
struct UnpackJob { const float* src; float* dst; int kind, nk, co, ci, CO, CI, co_off; long long total; };
#define AFI_MAX_UNPACK 40
int unpack_group(int njobs, const UnpackJob* jobs, float scale, int accumulate, cudaStream_t st);
int unpack_1x1(const float* packed, int co, int ci, int layout_nk, float* dst, float scale, int accumulate, cudaStream_t st);
// dx (contiguous NCHW) = dxb (NHWC fp32) + dy_scale * bilinear2x^T(dy)
int g_input_grad(PView dxb, afi_view4 dy, float dy_scale, int n, int c, int h, int w, int oh, int ow, float* dst, cudaStream_t st);

// BatchNorm helpers (stats buffers: double sum[C], sumsq[C])
int bn_finalize(const double* sum, const double* sumsq, long long count, int c, float eps, float momentum, int training,
                float* mean, float* rstd, float* var_unb, float* running_mean, float* running_var, long long* nbt, cudaStream_t st);
int bn_update_running(int ncalls, const float* const* mean, const float* const* var, int c, float momentum, float* rmean, float* rvar,
                      long long* nbt, cudaStream_t st);
// discriminator head (1024 -> 1 conv): t9[p][tap] = <a3[p], w4[tap]>, logits = b + 3x3 shift-sum of t9
int dhead_forward(PView a3, int dt, const float* w4 /*[c][9] torch layout*/, const float* b4, int n, int h, int w, int c,
                  float* t9, float* logits, cudaStream_t st);
// dW4 += , db4 += , dy3 = (sum_tap g[q-tap] w4[tap]) * lrelu'(a3)
int dhead_backward(PView a3, int dt, const float* w4, const float* g, int n, int h, int w, int c, float* dw4_acc, float* db4_acc,
                   PView dy3, cudaStream_t st);

// grouped dense BatchNorm passes over several discriminator calls (dense [P][C] operands, C/8 dividing 256)
int dense_group_ok(int c);
int bn_apply_lrelu_group(int nprob, const PView* z, const PView* a, int dt, const float* const* mean, const float* const* rstd,
                         const float* gamma, const float* beta, float slope, const long long* P, int c, cudaStream_t st);
int bn_bwd_apply_group(int nprob, const PView* dy, const PView* z, int dt, const float* const* mean, const float* const* rstd, const float* gamma,
                       double* const* s_dy, double* const* s_dyx, float* dgamma_acc, float* dbeta_acc, const long long* P, int c, int eval_mode,
                       cudaStream_t st);
int col_reduce_group(int mode, int nprob, const PView* x, const PView* z, int dt, const float* const* mean, const float* const* rstd,
                     double* const* o0, double* const* o1, const long long* P, int c, cudaStream_t st);
int bn_finalize_group(int ncalls, const double* const* sum, const double* const* sumsq, const long long* count, int c, float eps,
                      float* const* mean, float* const* rstd, float* const* var, cudaStream_t st);
int zero_group(int n, void* const* ptrs, size_t bytes, cudaStream_t st);

// tensor-core formulation of the discriminator head (bf16 mode)
int dhead_pack_tc(const float* w4, int c, void* fwd, void* bwd, cudaStream_t st);
int dhead_build_g9(const float* g, int n, int h, int w, void* g9, float* g9f, cudaStream_t st);
int dhead_stencil16(const float* t9, const float* b4, int n, int h, int w, float* logits, cudaStream_t st);
int dhead_unpack_tc(const float* acc, int c, float* dst, float scale, int accumulate, cudaStream_t st);
int sum_f32(const float* x, long long n, float* out, cudaStream_t st);

// bf16 variant with the 9-tap product on warp-level tensor-core MMAs (g9: the bf16 [P][16] shifted head gradients)
int dhead_backward_group_mma(int pass, int nprob, const void* const* g9, const PView* z3, const PView* dz3, const float* w4,
                             const float* const* mean, const float* const* rstd, const float* gamma, const float* beta, double* const* s_dy,
                             double* const* s_dyx, float* dgamma_acc, float* dbeta_acc, const long long* P, int c, int eval_mode, cudaStream_t st);

template <typename T> struct dt_of;
template <> struct dt_of<float> { static const int v = DT_F32; };
template <> struct dt_of<bf16> { static const int v = DT_BF16; };

}  // namespace afi

struct afi_ctx {
    int device;
    int sm_count;
    void* encode_tiled;   // cuTensorMapEncodeTiled entry point (resolved through cudaGetDriverEntryPoint: no libcuda link dependency)
};
