// C-ABI entry points and the kernel SEQUENCES of the AF interpolator and the feature-patch discriminator.
// The sequences transliterate tests/kernel_model.py (verified against the oracle in fp64); math per
// SURVEY.md App. A, reference generator_rdb.py:15-130 and feature_patch_discriminator.py:18-55.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace afi {

static thread_local char g_err[512] = "";
long long g_launches = 0;
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- per-launch profiling -------------------------------------------------------------------------------
struct ProfRec { int kind, cin, cout; long long pixels; double flops; cudaEvent_t e0, e1; float ms; };
bool g_prof_on = false;
static ProfRec* g_prof = nullptr;
static int g_prof_n = 0, g_prof_cap = 0;
void prof_begin(int kind, double flops, int cin, int cout, long long pixels, cudaStream_t st) {
    if (g_prof_n >= g_prof_cap) { g_prof_on = false; return; }
    ProfRec& r = g_prof[g_prof_n];
    r.kind = kind; r.flops = flops; r.cin = cin; r.cout = cout; r.pixels = pixels; r.ms = 0.f;
    if (!r.e0) { cudaEventCreate(&r.e0); cudaEventCreate(&r.e1); }
    cudaEventRecord(r.e0, st);
}
void prof_end(cudaStream_t st) {
    if (g_prof_n < g_prof_cap) { cudaEventRecord(g_prof[g_prof_n].e1, st); g_prof_n++; }
}

void conv_args_init(ConvArgs& a) {
    memset(&a, 0, sizeof(a));
    a.alpha = 1.f;
    a.slope = 0.2f;
    a.mask_slope = 0.2f;
    a.out_dt = DT_F32;
    a.nprob = 1;
}
void set_std_taps(Tap* taps, int view, int slab0) {
    int i = 0;
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++, i++) { taps[i].dy = dy; taps[i].dx = dx; taps[i].view = view; taps[i].slab = slab0 + i; }
}

static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }
struct Carver {
    char* base; size_t off;
    explicit Carver(void* b) : base((char*)b), off(0) {}
    void* take(size_t bytes) { void* p = base ? base + off : nullptr; off += align_up(bytes); return p; }
};

static inline bool prec_ok(int prec) { return prec == AFI_PREC_FP32 || prec == AFI_PREC_BF16 || prec == AFI_PREC_BF16_SIMT || prec == AFI_PREC_SPLIT; }
// activation storage dtype: AFI_PREC_SPLIT keeps every tensor fp32 in HBM (its elementwise passes are the fp32 mode's); only the GEMM
// operands are staged as bf16 planes
static inline int prec_dt(int prec) { return (prec == AFI_PREC_FP32 || prec == AFI_PREC_SPLIT) ? DT_F32 : DT_BF16; }
static inline size_t dt_size(int dt) { return dt == DT_F32 ? 4 : 2; }
// tcgen05 engine with bf16 storage: the tensor-core formulations of the discriminator head and the fused statistics epilogues
static inline bool prec_tc(int prec) { return prec == AFI_PREC_BF16; }
// tcgen05 GEMM engine (bf16 or split-precision operands): [slab][cout][cin] weight / gradient layouts, four-phase transposed conv
static inline bool prec_nk(int prec) { return prec == AFI_PREC_BF16 || prec == AFI_PREC_SPLIT; }
// packed weights: dtype tag and bytes per logical element (split: three bf16 planes)
static inline int prec_wdt(int prec) { return prec == AFI_PREC_SPLIT ? DT_BF16X3 : prec_dt(prec); }
static inline size_t prec_wes(int prec) { return prec == AFI_PREC_SPLIT ? 6 : dt_size(prec_dt(prec)); }

// AFI_PREC_SPLIT operand scratch of the API call in flight: problem k of a grouped launch stages its operands in the scratch of call
// k0 + k (k0 != 0 only for the single-problem launches issued per call)
// pairs: plane pairs per product for the convs of the call in flight -- 6 for the forward pass of a call that keeps its activations for a
// backward pass (pre-activation signs decide the LeakyReLU slopes of the gradients), 3 for forward-only calls and every backward GEMM
struct SplitScratch { void* p[AFI_MAX_PROB]; size_t n[AFI_MAX_PROB]; int k0; int pairs; };
static thread_local SplitScratch g_ss;
static inline int split_pairs_env(int dflt) {      // AFIGAN_SPLIT_PAIRS=6 forces the six-pair product everywhere (accuracy experiments)
    const char* e = getenv("AFIGAN_SPLIT_PAIRS");
    return (e && atoi(e) == 6) ? 6 : dflt;
}

static int run_conv(afi_ctx* ctx, int prec, const ConvArgs& a, cudaStream_t st) {
    if (prec == AFI_PREC_FP32) return conv_simt<float>(a, st);
    if (prec == AFI_PREC_BF16_SIMT) return conv_simt<bf16>(a, st);
    if (prec == AFI_PREC_SPLIT) {
        ConvArgs b = a;
        for (int k = 0; k < b.nprob; k++) { b.p[k].sws = g_ss.p[g_ss.k0 + k]; b.p[k].sws_bytes = g_ss.n[g_ss.k0 + k]; }
        b.split_pairs = g_ss.pairs == 3 ? 3 : 6;
        return conv_tc_split(ctx, b, st);
    }
    return conv_tc(ctx, a, st);
}
static int run_wgrad(afi_ctx* ctx, int prec, const WgradArgs& a, cudaStream_t st) {
    if (prec == AFI_PREC_FP32) return wgrad_simt<float>(a, st);
    if (prec == AFI_PREC_BF16_SIMT) return wgrad_simt<bf16>(a, st);
    if (prec == AFI_PREC_SPLIT) {
        WgradArgs b = a;
        for (int k = 0; k < b.nprob; k++) { b.p[k].sws = g_ss.p[g_ss.k0 + k]; b.p[k].sws_bytes = g_ss.n[g_ss.k0 + k]; }
        b.split_pairs = split_pairs_env(3);
        return wgrad_tc_split(ctx, b, st);
    }
    return wgrad_tc(ctx, a, st);
}
static int to_nhwc(int prec, afi_view4 src, int n, int c, int h, int w, PView dst, cudaStream_t st) {
    return prec_dt(prec) == DT_F32 ? nchw_to_nhwc<float>(src, n, c, h, w, dst, st) : nchw_to_nhwc<bf16>(src, n, c, h, w, dst, st);
}
static int to_nchw(int prec, PView a, PView lat, afi_view4 skip, int sh, int sw, float scale, int n, int c, int oh, int ow, float* dst,
                   cudaStream_t st, const afi_view4* add = nullptr, const float* fw = nullptr) {
    return prec_dt(prec) == DT_F32 ? nhwc_to_nchw<float>(a, lat, skip, sh, sw, scale, n, c, oh, ow, dst, st, add, fw)
                                 : nhwc_to_nchw<bf16>(a, lat, skip, sh, sw, scale, n, c, oh, ow, dst, st, add, fw);
}
// forward / dgrad pack modes per engine
static inline int pm(int prec, int kind) { return kind * 2 + (prec_nk(prec) ? 1 : 0); }

// standard 3x3 weight gradient over a group of problems (dims / views per problem)
struct Dim3 { int n, h, w; };
static int wgrad_std(afi_ctx* ctx, int prec, int nprob, const Dim3* d, const PView* x, int cin, const PView* dy, int cout, float* dw,
                     cudaStream_t st) {
    WgradArgs g;
    memset(&g, 0, sizeof(g));
    g.cin = cin; g.cout = cout; g.ntaps = 9; g.nprob = nprob;
    set_std_taps(g.taps, 0, 0);
    for (int k = 0; k < nprob; k++) { g.p[k].N = d[k].n; g.p[k].H = d[k].h; g.p[k].W = d[k].w; g.p[k].x = x[k]; g.p[k].dy = dy[k]; }
    g.dw = dw;
    return run_wgrad(ctx, prec, g, st);
}
// common part of a 3x3 conv over a group; the caller fills a.p[k].in/out/... afterwards
static void conv_std(ConvArgs& a, int nprob, const Dim3* d, int cin, int cout, const void* w, int slab0 = 0) {
    conv_args_init(a);
    a.cin = cin; a.cout = cout; a.ntaps = 9; a.nprob = nprob; a.w = w;
    set_std_taps(a.taps, 0, slab0);
    for (int k = 0; k < nprob; k++) { a.p[k].N = d[k].n; a.p[k].H = d[k].h; a.p[k].W = d[k].w; }
}

// =====================================================================================================
// Generator
// =====================================================================================================
static const int C = AFI_CH, GR = 32, CB = AFI_CH + 4 * 32;   // 256, growth 32, dense buffer width 384

struct GPacked {   // element offsets into the packed weight buffer (dtype T).  Dgrad operands of a dense block: rdb_d[r][4] the fusion conv;
                   // rdb_xd[r] = [9][256][128], the part of the four growth convs that reaches the block input x (gemm-cin = the four
                   // 32-channel growth gradients side by side); rdb_cd[r][i], i = 1..3 = [9][32 i][32], the part of growth conv i+1 that
                   // reaches the earlier growth channels
    // Forward operands of a dense block on the tensor-core engines: rdb_fx[r] = [9][128][256], the x part (input channels 0..255) of the four
    // growth convs as ONE N = 128 GEMM; rdb_fc[r][i], i = 1..3 = [9][32][32 i], the part of growth conv i+1 that reads the earlier growth channels
    size_t head_f, head_d, rdb_f[AFI_MAX_RDB][5], rdb_d[AFI_MAX_RDB][5], rdb_xd[AFI_MAX_RDB], rdb_cd[AFI_MAX_RDB][4], rdb_fx[AFI_MAX_RDB],
        rdb_fc[AFI_MAX_RDB][4], post_f, post_d, up_f, up_d, out_f, out_d, total;
};
static GPacked g_packed_layout(int n_rdb) {
    GPacked L; size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 127) / 128 * 128; return r; };
    L.head_f = take(9 * C * C); L.head_d = take(9 * C * C);
    for (int r = 0; r < n_rdb; r++) {
        for (int i = 0; i < 5; i++) {
            size_t n = (size_t)9 * (C + GR * i) * (i < 4 ? GR : C);
            L.rdb_f[r][i] = take(n);
            L.rdb_d[r][i] = i == 4 ? take(n) : 0;
        }
        L.rdb_xd[r] = take((size_t)9 * C * 4 * GR);
        L.rdb_cd[r][0] = 0;
        for (int i = 1; i < 4; i++) L.rdb_cd[r][i] = take((size_t)9 * GR * i * GR);
        L.rdb_fx[r] = take((size_t)9 * 4 * GR * C);
        L.rdb_fc[r][0] = 0;
        for (int i = 1; i < 4; i++) L.rdb_fc[r][i] = take((size_t)9 * GR * GR * i);
    }
    L.post_f = take(9 * C * C); L.post_d = take(9 * C * C);
    L.up_f = take(36 * C * C); L.up_d = take(36 * C * C);
    L.out_f = take(9 * C * C); L.out_d = take(9 * C * C);
    L.total = o;
    return L;
}
struct GGradAcc {  // float offsets.  rdb_g[r]: the four growth convs of dense block r share ONE [9][128][352] accumulator (row block i =
                   // conv i+1; its columns >= 256 + 32 i are non-causal products nobody reads); rdb_w[r][4]: the fusion conv
    size_t head_w, head_b, rdb_g[AFI_MAX_RDB], rdb_w[AFI_MAX_RDB][5], post_w, post_b, up_w, up_b, out_w, out_b, total;
};
static GGradAcc g_gradacc_layout(int n_rdb) {
    GGradAcc L; size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 63) / 64 * 64; return r; };
    L.head_w = take(9 * C * C); L.head_b = take(C);
    for (int r = 0; r < n_rdb; r++) {
        L.rdb_g[r] = take((size_t)9 * (4 * GR) * (C + 3 * GR));
        for (int i = 0; i < 4; i++) L.rdb_w[r][i] = L.rdb_g[r];
        L.rdb_w[r][4] = take((size_t)9 * CB * C);
    }
    L.post_w = take(9 * C * C); L.post_b = take(C);
    L.up_w = take(36 * C * C); L.up_b = take(C);
    L.out_w = take(9 * C * C); L.out_b = take(C);
    L.total = o;
    return L;
}
struct GWs {
    void *X0, *B[AFI_MAX_RDB], *H1, *H2, *H3, *Yb, *LX, *LAT;           // forward (X0..H3 saved for backward)
    void* GP;                                                           // fp32 [P][128]: x-part partial sums of a dense block's growth convs
    void *G0, *G1, *G2, *dH1, *GA[2], *DC5, *GC, *GH, *DXb, *LW;        // backward scratch
    void *LWD, *LWG, *LDX;                                              // lateral backward: dgrad pack, wgrad accumulator, d(lat_x) NHWC
    void* SPL; size_t spl_bytes;                                        // AFI_PREC_SPLIT: bf16 operand planes of the GEMM in flight
    size_t total;
};
static GWs g_ws_layout(void* base, int prec, int n, int h, int w, int n_rdb, int lat_c, int backward) {
    GWs W; memset(&W, 0, sizeof(W));
    Carver cv(base);
    size_t es = dt_size(prec_dt(prec)), wes = prec_wes(prec);
    size_t P = (size_t)n * h * w, P4 = 4 * P;
    W.X0 = cv.take(P * C * es);
    for (int r = 0; r < n_rdb; r++) W.B[r] = cv.take(P * CB * es);
    W.H1 = cv.take(P * C * es); W.H2 = cv.take(P * C * es); W.H3 = cv.take(P4 * C * es);
    W.Yb = cv.take(P4 * C * es);
    W.GP = cv.take(P * 4 * GR * 4);
    if (lat_c > 0) { W.LX = cv.take(P4 * lat_c * es); W.LAT = cv.take(P4 * C * es); W.LW = cv.take((size_t)C * lat_c * wes); }
    if (backward) {
        W.G0 = cv.take(P4 * C * es); W.G1 = cv.take(P4 * C * es); W.G2 = cv.take(P * C * es);
        W.dH1 = cv.take(P * C * 4); W.GA[0] = cv.take(P * CB * 4); W.GA[1] = cv.take(P * CB * 4);
        W.DC5 = cv.take(P * C * es); W.GC = cv.take(P * 4 * GR * es); W.GH = cv.take(P * C * es);
        W.DXb = cv.take(P * C * 4);
        if (lat_c > 0) { W.LWD = cv.take((size_t)C * lat_c * wes); W.LWG = cv.take((size_t)C * lat_c * 4); W.LDX = cv.take(P4 * lat_c * 4); }
    }
    if (prec == AFI_PREC_SPLIT) {
        // the widest GEMM of a call: the output conv's weight gradient (x and dy on the 2h x 2w grid), or the lateral's
        W.spl_bytes = 2 * split_planes_bytes((long long)P4, lat_c > C ? lat_c : C);
        W.SPL = cv.take(W.spl_bytes);
    }
    W.total = cv.off;
    return W;
}

static int g_check(int prec, int n, int h, int w, int n_rdb) {
    AFI_REQUIRE(prec_ok(prec), "unknown precision mode %d", prec);
    AFI_REQUIRE(n_rdb >= 1 && n_rdb <= AFI_MAX_RDB, "n_rdb %d out of range [1,%d]", n_rdb, AFI_MAX_RDB);
    AFI_REQUIRE(n >= 0 && h >= 1 && w >= 1, "bad feature shape n=%d h=%d w=%d", n, h, w);
    return AFI_OK;
}

}  // namespace afi
using namespace afi;

extern "C" {

int afi_abi_version(void) { return AFI_ABI_VERSION; }
size_t afi_sizeof(int which) {
    switch (which) {
        case 0: return sizeof(afi_view4);
        case 1: return sizeof(afi_g_params);
        case 2: return sizeof(afi_lateral);
        case 3: return sizeof(afi_g_call);
        case 4: return sizeof(afi_d_params);
        case 5: return sizeof(afi_d_call);
        case 6: return sizeof(afi_g_grads);
        case 7: return sizeof(afi_d_grads);
        default: return 0;
    }
}
const char* afi_last_error(void) { return g_err; }
long long afi_launch_count(int reset) { long long v = g_launches; if (reset) g_launches = 0; return v; }

int afi_profile_begin(int max_launches) {
    AFI_REQUIRE(max_launches > 0 && max_launches <= (1 << 20), "afi_profile_begin: bad capacity");
    if (max_launches > g_prof_cap) {
        ProfRec* n = (ProfRec*)calloc(max_launches, sizeof(ProfRec));
        if (g_prof) { memcpy(n, g_prof, sizeof(ProfRec) * g_prof_cap); free(g_prof); }
        g_prof = n; g_prof_cap = max_launches;
    }
    g_prof_n = 0; g_prof_on = true;
    return AFI_OK;
}
int afi_profile_end(int* n_launches) {
    g_prof_on = false;
    AFI_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < g_prof_n; i++) AFI_CUDA(cudaEventElapsedTime(&g_prof[i].ms, g_prof[i].e0, g_prof[i].e1));
    if (n_launches) *n_launches = g_prof_n;
    return AFI_OK;
}
int afi_profile_get(int i, int* kind, double* flops, float* ms, int* cin, int* cout, long long* pixels) {
    AFI_REQUIRE(i >= 0 && i < g_prof_n, "afi_profile_get: index out of range");
    const ProfRec& r = g_prof[i];
    if (kind) *kind = r.kind; if (flops) *flops = r.flops; if (ms) *ms = r.ms;
    if (cin) *cin = r.cin; if (cout) *cout = r.cout; if (pixels) *pixels = r.pixels;
    return AFI_OK;
}

int afi_create(afi_ctx** out) {
    AFI_REQUIRE(out != nullptr, "afi_create: null output");
    int dev = 0;
    AFI_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    AFI_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only and has no fallback", dev, prop.major, prop.minor);
        return AFI_ERR_ARCH;
    }
    afi_ctx* c = new afi_ctx();
    c->device = dev; c->sm_count = prop.multiProcessorCount; c->encode_tiled = nullptr;
    int r = tc_init(c);
    if (r != AFI_OK) { delete c; return r; }
    *out = c;
    return AFI_OK;
}
void afi_destroy(afi_ctx* ctx) { delete ctx; }

size_t afi_g_packed_bytes(int prec, int n_rdb) { return g_packed_layout(n_rdb).total * prec_wes(prec); }
size_t afi_g_gradacc_bytes(int n_rdb) { return g_gradacc_layout(n_rdb).total * 4; }
size_t afi_g_workspace_bytes(int prec, int n, int h, int w, int n_rdb, int lat_c, int save_for_backward) {
    return g_ws_layout(nullptr, prec, n, h, w, n_rdb, lat_c, save_for_backward).total;
}

int afi_g_pack(afi_ctx* ctx, int prec, const afi_g_params* p, void* packed, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(ctx && p && packed, "afi_g_pack: null argument");
    AFI_TRY(g_check(prec, 1, 1, 1, p->n_rdb));
    GPacked L = g_packed_layout(p->n_rdb);
    const int dt = prec_wdt(prec); const size_t es = prec_wes(prec);
    char* b = (char*)packed;
    PackJob jobs[AFI_MAX_PACK]; int nj = 0;     // forward + dgrad layouts of every conv: one grouped launch
    auto add = [&](const float* w, int co, int ci, int kind, size_t off) {
        if (nj >= AFI_MAX_PACK) return;
        PackJob& j = jobs[nj++]; memset(&j, 0, sizeof(j)); j.w = w; j.dst = b + off * es; j.co = co; j.ci = ci; j.mode = pm(prec, kind);
    };
    auto add_sub = [&](const float* w, int co, int ci, size_t off, int n0, int ncnt, int koff, int ktot) {
        PackJob& j = jobs[nj++]; memset(&j, 0, sizeof(j)); j.w = w; j.dst = b + off * es; j.co = co; j.ci = ci; j.mode = pm(prec, 1);
        j.sub = 1; j.n0 = n0; j.ncnt = ncnt; j.koff = koff; j.ktot = ktot;
    };
    add(p->head_w, C, C, 0, L.head_f); add(p->head_w, C, C, 1, L.head_d);
    for (int r = 0; r < p->n_rdb; r++)
        for (int i = 0; i < 5; i++) {
            int co = i < 4 ? GR : C, ci = C + GR * i;
            if (i == 4 || !prec_nk(prec)) add(p->rdb_w[r][i], co, ci, 0, L.rdb_f[r][i]);      // (the tensor-core engines read rdb_fx / rdb_fc instead)
            if (i < 4 && prec_nk(prec)) {       // forward sub-blocks: x part of all four convs in one operand; growth part per conv
                PackJob& j = jobs[nj++]; memset(&j, 0, sizeof(j));
                j.w = p->rdb_w[r][i]; j.dst = b + L.rdb_fx[r] * es; j.co = GR; j.ci = ci; j.mode = 1; j.sub = 2; j.n0 = 0; j.ncnt = C; j.koff = GR * i; j.ktot = 4 * GR;
                if (i > 0) {
                    PackJob& k2 = jobs[nj++]; memset(&k2, 0, sizeof(k2));
                    k2.w = p->rdb_w[r][i]; k2.dst = b + L.rdb_fc[r][i] * es; k2.co = GR; k2.ci = ci; k2.mode = 1; k2.sub = 2; k2.n0 = C; k2.ncnt = GR * i; k2.koff = 0; k2.ktot = GR;
                }
            }
            if (i == 4) add(p->rdb_w[r][i], co, ci, 1, L.rdb_d[r][i]);
            else {
                add_sub(p->rdb_w[r][i], GR, ci, L.rdb_xd[r], 0, C, GR * i, 4 * GR);                  // x part: all four convs, one operand
                if (i > 0) add_sub(p->rdb_w[r][i], GR, ci, L.rdb_cd[r][i], C, GR * i, 0, GR);       // growth part of conv i+1
            }
        }
    add(p->post_w, C, C, 0, L.post_f); add(p->post_w, C, C, 1, L.post_d);
    add(p->up_w, C, C, 2, L.up_f); add(p->up_w, C, C, 3, L.up_d);
    add(p->out_w, C, C, 0, L.out_f); add(p->out_w, C, C, 1, L.out_d);
    return pack_weights_group(nj, jobs, dt, st);
}

int afi_g_forward(afi_ctx* ctx, int prec, const afi_g_params* p, const void* packed, const afi_g_call* calls, int ncalls, int save,
                  void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(ctx && p && packed && calls, "afi_g_forward: null argument");
    AFI_REQUIRE(ncalls >= 1 && ncalls <= AFI_MAX_PROB, "afi_g_forward: %d calls per group (max %d)", ncalls, AFI_MAX_PROB);
    const int nr = p->n_rdb, dt = prec_dt(prec);
    const int es = (int)dt_size(dt), wes = (int)prec_wes(prec);
    GWs W[AFI_MAX_PROB];
    Dim3 d1[AFI_MAX_PROB], d2[AFI_MAX_PROB];
    g_ss.k0 = 0; g_ss.pairs = split_pairs_env(save ? 6 : 3);
    for (int k = 0; k < ncalls; k++) {
        const afi_g_call& c = calls[k];
        AFI_REQUIRE(c.x.ptr && c.y && c.ws, "afi_g_forward: call %d has a null pointer", k);
        AFI_TRY(g_check(prec, c.n, c.h, c.w, nr));
        AFI_REQUIRE(c.n >= 1, "afi_g_forward: empty batch");
        AFI_REQUIRE(c.oh >= 1 && c.ow >= 1 && c.oh <= 2 * c.h && c.ow <= 2 * c.w, "afi_g_forward: output crop %dx%d exceeds %dx%d", c.oh, c.ow,
                    2 * c.h, 2 * c.w);
        const int lat_c = c.lateral ? c.lateral->lat_c : 0;
        if (c.lateral) AFI_REQUIRE(lat_c % 32 == 0 && lat_c <= 2048 && c.lateral->lat_x.ptr && c.lateral->lat_w, "afi_g_forward: bad lateral");
        W[k] = g_ws_layout(c.ws, prec, c.n, c.h, c.w, nr, lat_c, save);
        if (W[k].total > c.ws_bytes) { set_error("afi_g_forward: workspace %zu B < required %zu B", c.ws_bytes, W[k].total); return AFI_ERR_WORKSPACE; }
        g_ss.p[k] = W[k].SPL; g_ss.n[k] = W[k].spl_bytes;
        d1[k] = {c.n, c.h, c.w};
        d2[k] = {c.n, 2 * c.h, 2 * c.w};
    }
    GPacked L = g_packed_layout(nr);
    const char* pk = (const char*)packed;
    ConvArgs a;

    for (int k = 0; k < ncalls; k++) AFI_TRY(to_nhwc(prec, calls[k].x, calls[k].n, C, calls[k].h, calls[k].w, pview(W[k].X0, calls[k].h, calls[k].w, C), st));
    // [0] head conv + bias + LeakyReLU -> B0[:, 0:256]                                   generator_rdb.py:91-93
    conv_std(a, ncalls, d1, C, C, pk + L.head_f * wes);
    a.bias = p->head_b; a.act = 1; a.out_dt = dt;
    for (int k = 0; k < ncalls; k++) { a.p[k].in[0] = pview(W[k].X0, d1[k].h, d1[k].w, C); a.p[k].out = pview(W[k].B[0], d1[k].h, d1[k].w, CB); }
    AFI_TRY(run_conv(ctx, prec, a, st));
    // [1] residual-in-residual: dense blocks write their growth channels into slices of one 384-ch buffer   :39-71
    for (int r = 0; r < nr; r++) {
        if (prec_nk(prec)) {
            // Tensor-core engines: c_i = L(W_i^x * x + W_i^c * [c_1 .. c_{i-1}]).  The x parts of all four growth convs are ONE GEMM
            // (K = 2304, N = 128: four N = 32 GEMMs run at a fifth of the rate, each M128 x N32 MMA being bound by its A-operand read);
            // its epilogue stores c_1 (activated) into the block buffer and the partial sums of c_2 .. c_4 as fp32.  The growth parts
            // are three short-K convs (K = 288, 576, 864) that add their partial sum and activate in the epilogue.
            conv_std(a, ncalls, d1, C, 4 * GR, pk + L.rdb_fx[r] * wes);
            a.out_dt = DT_F32; a.split_col = GR; a.out2_dt = dt;
            for (int k = 0; k < ncalls; k++) {
                PView Br = pview(W[k].B[r], d1[k].h, d1[k].w, CB);
                a.p[k].in[0] = Br; a.p[k].out = pview(W[k].GP, d1[k].h, d1[k].w, 4 * GR); a.p[k].out2 = pview_ch(Br, C, es);
            }
            AFI_TRY(run_conv(ctx, prec, a, st));
            for (int i = 1; i < 4; i++) {
                conv_std(a, ncalls, d1, GR * i, GR, pk + L.rdb_fc[r][i] * wes);
                a.act = 1; a.act_post = 1; a.out_dt = dt;
                for (int k = 0; k < ncalls; k++) {
                    PView Br = pview(W[k].B[r], d1[k].h, d1[k].w, CB);
                    a.p[k].in[0] = pview_ch(Br, C, es); a.p[k].accin = pview_ch(pview(W[k].GP, d1[k].h, d1[k].w, 4 * GR), GR * i, 4);
                    a.p[k].out = pview_ch(Br, C + GR * i, es);
                }
                AFI_TRY(run_conv(ctx, prec, a, st));
            }
        } else for (int i = 0; i < 4; i++) {
            conv_std(a, ncalls, d1, C + GR * i, GR, pk + L.rdb_f[r][i] * wes);
            a.act = 1; a.out_dt = dt;
            for (int k = 0; k < ncalls; k++) {
                PView Br = pview(W[k].B[r], d1[k].h, d1[k].w, CB);
                a.p[k].in[0] = Br; a.p[k].out = pview_ch(Br, C + GR * i, es);
            }
            AFI_TRY(run_conv(ctx, prec, a, st));
        }
        conv_std(a, ncalls, d1, CB, C, pk + L.rdb_f[r][4] * wes);
        a.out_dt = dt; a.r1_dt = dt; a.r2_dt = dt;
        if (r + 1 < nr) { a.alpha = 0.2f; a.beta1 = 1.f; }           // x_{r+1} = x_r + 0.2 * conv5
        else { a.alpha = 0.04f; a.beta1 = 0.2f; a.beta2 = 1.f; }      // h1 = 0.2 * (x + 0.2 conv5) + h0            :27-30
        for (int k = 0; k < ncalls; k++) {
            PView Br = pview(W[k].B[r], d1[k].h, d1[k].w, CB);
            a.p[k].in[0] = Br; a.p[k].r1 = Br;
            if (r + 1 < nr) a.p[k].out = pview(W[k].B[r + 1], d1[k].h, d1[k].w, CB);
            else { a.p[k].r2 = pview(W[k].B[0], d1[k].h, d1[k].w, CB); a.p[k].out = pview(W[k].H1, d1[k].h, d1[k].w, C); }
        }
        AFI_TRY(run_conv(ctx, prec, a, st));
    }
    // [2] post conv + bias + LeakyReLU                                                   :97-99
    conv_std(a, ncalls, d1, C, C, pk + L.post_f * wes);
    a.bias = p->post_b; a.act = 1; a.out_dt = dt;
    for (int k = 0; k < ncalls; k++) { a.p[k].in[0] = pview(W[k].H1, d1[k].h, d1[k].w, C); a.p[k].out = pview(W[k].H2, d1[k].h, d1[k].w, C); }
    AFI_TRY(run_conv(ctx, prec, a, st));
    // [3] ConvTranspose2d k6 s2 p2 == four 3x3 sub-pixel convs with interleaved stores      :101-105, App. G
    if (prec_nk(prec)) {      // tensor-core engine: the four phases are the four N tiles of ONE launch (the input tile is staged once)
        conv_std(a, ncalls, d1, C, C, pk + L.up_f * wes, 0);
        a.bias = p->up_b; a.act = 1; a.out_dt = dt; a.nphase = 4;
        for (int k = 0; k < ncalls; k++) {
            const int H2x = d2[k].h, W2x = d2[k].w;
            a.p[k].in[0] = pview(W[k].H2, d1[k].h, d1[k].w, C);
            PView o; o.ptr = W[k].H3;
            o.sx = 2 * C; o.sy = (long long)2 * W2x * C; o.sn = (long long)H2x * W2x * C;
            a.p[k].out = o;
        }
        AFI_TRY(run_conv(ctx, prec, a, st));
    } else for (int ph = 0; ph < 4; ph++) {
        int pa = ph >> 1, pb = ph & 1;
        conv_std(a, ncalls, d1, C, C, pk + L.up_f * wes, 9 * ph);
        a.bias = p->up_b; a.act = 1; a.out_dt = dt;
        for (int k = 0; k < ncalls; k++) {
            const int H2x = d2[k].h, W2x = d2[k].w;
            a.p[k].in[0] = pview(W[k].H2, d1[k].h, d1[k].w, C);
            PView o; o.ptr = (char*)W[k].H3 + ((size_t)pa * W2x + pb) * C * es;
            o.sx = 2 * C; o.sy = (long long)2 * W2x * C; o.sn = (long long)H2x * W2x * C;
            a.p[k].out = o;
        }
        AFI_TRY(run_conv(ctx, prec, a, st));
    }
    // Lateral 1x1 conv of the FPN / PAFPN merge (fpn_sr.py:152), one problem per call.  Tensor-core engines: it runs BEFORE the output
    // conv, into LAT laid out on the 2h x 2w grid, and the output conv adds it in its GEMM epilogue (residual operand) -- the merge
    // `lateral + G(x)` costs no pass of its own.  CUDA-core engines: added by the final layout pass.
    const bool lat_in_epilogue = prec_nk(prec);
    PView latv[AFI_MAX_PROB];
    for (int k = 0; k < ncalls; k++) {
        const afi_g_call& c = calls[k];
        latv[k] = pview_null();
        if (!c.lateral) continue;
        const afi_lateral* lat = c.lateral;
        const int lat_c = lat->lat_c;
        AFI_TRY(to_nhwc(prec, lat->lat_x, c.n, lat_c, c.oh, c.ow, pview(W[k].LX, c.oh, c.ow, lat_c), st));
        AFI_TRY(pack_weights(lat->lat_w, C, lat_c, pm(prec, 4), W[k].LW, prec_wdt(prec), st));
        conv_args_init(a);
        a.cin = lat_c; a.cout = C; a.ntaps = 1; a.nprob = 1;
        a.taps[0].dy = 0; a.taps[0].dx = 0; a.taps[0].view = 0; a.taps[0].slab = 0;
        a.p[0].N = c.n; a.p[0].H = c.oh; a.p[0].W = c.ow;
        a.p[0].in[0] = pview(W[k].LX, c.oh, c.ow, lat_c); a.w = W[k].LW; a.bias = lat->lat_b;
        latv[k] = lat_in_epilogue ? pview(W[k].LAT, d2[k].h, d2[k].w, C) : pview(W[k].LAT, c.oh, c.ow, C);
        a.p[0].out = latv[k]; a.out_dt = dt;
        g_ss.k0 = k;
        AFI_TRY(run_conv(ctx, prec, a, st));
        g_ss.k0 = 0;
    }
    // [4] output conv + bias (+ lateral) on the 2h x 2w grid                                :107-108
    conv_std(a, ncalls, d2, C, C, pk + L.out_f * wes);
    a.bias = p->out_b; a.out_dt = dt; a.r1_dt = dt; a.beta1 = 1.f;
    for (int k = 0; k < ncalls; k++) {
        a.p[k].in[0] = pview(W[k].H3, d2[k].h, d2[k].w, C); a.p[k].out = pview(W[k].Yb, d2[k].h, d2[k].w, C);
        // (outside the oh x ow crop LAT holds whatever the workspace held: those outputs are never read)
        if (lat_in_epilogue && latv[k].ptr) a.p[k].r1 = latv[k];
    }
    AFI_TRY(run_conv(ctx, prec, a, st));
    for (int k = 0; k < ncalls; k++) {
        const afi_g_call& c = calls[k];
        const float scale = c.lateral ? c.lateral->scale : 1.f;
        // y = (branch + bilinear2x(x) [+ lateral]) * scale, cropped to oh x ow            :125,130; stage1_trainer.py:437-443
        // [+ the BiFPN fusion w0 * cur + w1 * y (bifpn_sr.py:535-548) when the call carries one]
        AFI_REQUIRE(!c.fuse_w || (c.fuse_cur.ptr && !save), "afi_g_forward: call %d: the fused BiFPN site is forward-only and needs fuse_cur", k);
        AFI_TRY(to_nchw(prec, pview(W[k].Yb, d2[k].h, d2[k].w, C), lat_in_epilogue ? pview_null() : latv[k], c.x, c.h, c.w, scale, c.n, C, c.oh,
                        c.ow, c.y, st, c.fuse_w ? &c.fuse_cur : nullptr, c.fuse_w));
    }
    return AFI_OK;
}

int afi_g_backward(afi_ctx* ctx, int prec, const afi_g_params* p, const void* packed, const afi_g_call* calls, int ncalls, float* gradacc,
                   void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(ctx && p && packed && calls && gradacc, "afi_g_backward: null argument");
    AFI_REQUIRE(ncalls >= 1 && ncalls <= AFI_MAX_PROB, "afi_g_backward: %d calls per group (max %d)", ncalls, AFI_MAX_PROB);
    const int nr = p->n_rdb, dt = prec_dt(prec);
    const int es = (int)dt_size(dt), wes = (int)prec_wes(prec);
    GWs W[AFI_MAX_PROB];
    Dim3 d1[AFI_MAX_PROB], d2[AFI_MAX_PROB];
    g_ss.k0 = 0; g_ss.pairs = split_pairs_env(3);
    PView X0[AFI_MAX_PROB], H1[AFI_MAX_PROB], H2[AFI_MAX_PROB], H3[AFI_MAX_PROB], G0[AFI_MAX_PROB], G1[AFI_MAX_PROB], G2[AFI_MAX_PROB],
        dH1[AFI_MAX_PROB], DC5[AFI_MAX_PROB], GC[AFI_MAX_PROB], GH[AFI_MAX_PROB], tmpx[AFI_MAX_PROB], tmpy[AFI_MAX_PROB];
    for (int k = 0; k < ncalls; k++) {
        const afi_g_call& c = calls[k];
        AFI_REQUIRE(c.dy.ptr && c.ws, "afi_g_backward: call %d has a null pointer", k);
        AFI_REQUIRE((c.dx != nullptr) == (calls[0].dx != nullptr), "afi_g_backward: either every call of a group or none asks for dx");
        AFI_TRY(g_check(prec, c.n, c.h, c.w, nr));
        W[k] = g_ws_layout(c.ws, prec, c.n, c.h, c.w, nr, c.lateral ? c.lateral->lat_c : 0, 1);
        if (W[k].total > c.ws_bytes) { set_error("afi_g_backward: workspace %zu B < required %zu B", c.ws_bytes, W[k].total); return AFI_ERR_WORKSPACE; }
        g_ss.p[k] = W[k].SPL; g_ss.n[k] = W[k].spl_bytes;
        const int h = c.h, w = c.w, H2x = 2 * h, W2x = 2 * w;
        d1[k] = {c.n, h, w};
        d2[k] = {c.n, H2x, W2x};
        X0[k] = pview(W[k].X0, h, w, C); H1[k] = pview(W[k].H1, h, w, C); H2[k] = pview(W[k].H2, h, w, C); H3[k] = pview(W[k].H3, H2x, W2x, C);
        G0[k] = pview(W[k].G0, H2x, W2x, C); G1[k] = pview(W[k].G1, H2x, W2x, C); G2[k] = pview(W[k].G2, h, w, C);
        dH1[k] = pview(W[k].dH1, h, w, C); DC5[k] = pview(W[k].DC5, h, w, C); GC[k] = pview(W[k].GC, h, w, 4 * GR); GH[k] = pview(W[k].GH, h, w, C);
    }
    GPacked L = g_packed_layout(nr);
    GGradAcc GL = g_gradacc_layout(nr);
    const char* pk = (const char*)packed;
    ConvArgs a;
    int dn[AFI_MAX_PROB], dh[AFI_MAX_PROB], dw_[AFI_MAX_PROB];
    for (int k = 0; k < ncalls; k++) { dn[k] = d1[k].n; dh[k] = d1[k].h; dw_[k] = d1[k].w; }

    // dL/d(branch) on the full 2h x 2w grid: the crop's complement gets zero gradient
    for (int k = 0; k < ncalls; k++) {
        const afi_g_call& c = calls[k];
        AFI_CUDA(cudaMemsetAsync(W[k].G0, 0, (size_t)4 * c.n * c.h * c.w * C * es, st));
        AFI_TRY(to_nhwc(prec, c.dy, c.n, C, c.oh, c.ow, G0[k], st));
        if (c.lateral) {
            // y = scale * (lateral_conv(lat_x) + G(x)): both branches see scale * dy                       fpn_sr.py:152-157
            const afi_lateral* lat = c.lateral;
            const int lat_c = lat->lat_c;
            g_ss.k0 = k;
            if (lat->scale != 1.f)
                AFI_TRY(ew_combine(G0[k], dt, G0[k], dt, pview_null(), 0, pview_null(), 0, 0.2f, lat->scale, c.n, 2 * c.h, 2 * c.w, C, st));
            PView G0c = G0[k];                                   // the oh x ow crop of the 2h x 2w gradient buffer (same strides)
            PView LX = pview(W[k].LX, c.oh, c.ow, lat_c);        // lateral input kept by the forward pass
            Dim3 dc = {c.n, c.oh, c.ow};
            if (c.lat_gw) {
                AFI_CUDA(cudaMemsetAsync(W[k].LWG, 0, (size_t)C * lat_c * 4, st));
                WgradArgs g;
                memset(&g, 0, sizeof(g));
                g.cin = lat_c; g.cout = C; g.ntaps = 1; g.nprob = 1; g.dw = (float*)W[k].LWG;
                g.p[0].N = dc.n; g.p[0].H = dc.h; g.p[0].W = dc.w; g.p[0].x = LX; g.p[0].dy = G0c;
                AFI_TRY(run_wgrad(ctx, prec, g, st));
                AFI_TRY(unpack_1x1((const float*)W[k].LWG, C, lat_c, prec_nk(prec) ? 1 : 0, c.lat_gw, 1.f, 0, st));
            }
            if (c.lat_gb) {
                AFI_CUDA(cudaMemsetAsync(c.lat_gb, 0, C * sizeof(float), st));
                AFI_TRY(col_sum_f32(G0c, dt, dc.n, dc.h, dc.w, C, c.lat_gb, st));
            }
            if (c.lat_dx) {
                AFI_TRY(pack_weights(lat->lat_w, C, lat_c, pm(prec, 5), W[k].LWD, prec_wdt(prec), st));
                conv_args_init(a);
                a.cin = C; a.cout = lat_c; a.ntaps = 1; a.nprob = 1; a.w = W[k].LWD; a.out_dt = DT_F32;
                a.p[0].N = dc.n; a.p[0].H = dc.h; a.p[0].W = dc.w;
                a.p[0].in[0] = G0c; a.p[0].out = pview(W[k].LDX, dc.h, dc.w, lat_c);
                AFI_TRY(run_conv(ctx, prec, a, st));
                afi_view4 none; memset(&none, 0, sizeof(none));
                AFI_TRY(nhwc_to_nchw<float>(pview(W[k].LDX, dc.h, dc.w, lat_c), pview_null(), none, 0, 0, 1.f, dc.n, lat_c, dc.h, dc.w, c.lat_dx, st));
            }
            g_ss.k0 = 0;
        }
    }
    // output conv
    AFI_TRY(wgrad_std(ctx, prec, ncalls, d2, H3, C, G0, C, gradacc + GL.out_w, st));
    for (int k = 0; k < ncalls; k++) AFI_TRY(col_sum_f32(G0[k], dt, d2[k].n, d2[k].h, d2[k].w, C, gradacc + GL.out_b, st));
    conv_std(a, ncalls, d2, C, C, pk + L.out_d * wes);
    a.out_dt = dt;
    for (int k = 0; k < ncalls; k++) { a.p[k].in[0] = G0[k]; a.p[k].mask = H3[k]; a.p[k].out = G1[k]; }
    AFI_TRY(run_conv(ctx, prec, a, st));
    // transposed conv: wgrad per phase, bias, and dgrad as ONE 36-tap conv over the four phase views of G1
    PView ph[AFI_MAX_PROB][4];
    for (int i = 0; i < 4; i++) {
        int pa = i >> 1, pb = i & 1;
        for (int k = 0; k < ncalls; k++) {
            const int H2x = d2[k].h, W2x = d2[k].w;
            ph[k][i].ptr = (char*)W[k].G1 + ((size_t)pa * W2x + pb) * C * es;
            ph[k][i].sx = 2 * C; ph[k][i].sy = (long long)2 * W2x * C; ph[k][i].sn = (long long)H2x * W2x * C;
            tmpy[k] = ph[k][i];
        }
        AFI_TRY(wgrad_std(ctx, prec, ncalls, d1, H2, C, tmpy, C, gradacc + GL.up_w + (size_t)i * 9 * C * C, st));
    }
    for (int k = 0; k < ncalls; k++) AFI_TRY(col_sum_f32(G1[k], dt, d2[k].n, d2[k].h, d2[k].w, C, gradacc + GL.up_b, st));
    conv_args_init(a);
    a.cin = C; a.cout = C; a.ntaps = 36; a.nprob = ncalls; a.w = pk + L.up_d * wes; a.out_dt = dt;
    for (int i = 0; i < 4; i++) set_std_taps(a.taps + 9 * i, i, 9 * i);
    for (int k = 0; k < ncalls; k++) {
        a.p[k].N = d1[k].n; a.p[k].H = d1[k].h; a.p[k].W = d1[k].w;
        for (int i = 0; i < 4; i++) a.p[k].in[i] = ph[k][i];
        a.p[k].mask = H2[k]; a.p[k].out = G2[k];
    }
    AFI_TRY(run_conv(ctx, prec, a, st));
    // post conv
    AFI_TRY(wgrad_std(ctx, prec, ncalls, d1, H1, C, G2, C, gradacc + GL.post_w, st));
    for (int k = 0; k < ncalls; k++) AFI_TRY(col_sum_f32(G2[k], dt, d1[k].n, d1[k].h, d1[k].w, C, gradacc + GL.post_b, st));
    conv_std(a, ncalls, d1, C, C, pk + L.post_d * wes);
    a.out_dt = DT_F32;
    for (int k = 0; k < ncalls; k++) { a.p[k].in[0] = G2[k]; a.p[k].out = dH1[k]; }
    AFI_TRY(run_conv(ctx, prec, a, st));
    // residual-in-residual: h1 = 0.2 * x_nr + h0.  d_out (fp32 view) = 0.2 * dH1 entering the last dense block.
    float d_scale = 0.2f; int cur = 0;
    PView d_out[AFI_MAX_PROB], GA[AFI_MAX_PROB], Br[AFI_MAX_PROB];
    for (int k = 0; k < ncalls; k++) d_out[k] = dH1[k];
    for (int r = nr - 1; r >= 0; r--) {
        for (int k = 0; k < ncalls; k++) {
            const int n = d1[k].n, h = d1[k].h, w = d1[k].w;
            Br[k] = pview(W[k].B[r], h, w, CB);
            GA[k] = pview(W[k].GA[cur], h, w, CB);
            // GA[:, 256:384] = 0 (the conv5 dgrad below accumulates over all 384 channels)
            AFI_CUDA(cudaMemsetAsync(W[k].GA[cur], 0, (size_t)n * h * w * CB * 4, st));
        }
        // dc5 = 0.2 * d_out (GEMM operand in storage dtype); GA[:, 0:256] = d_out -- one grouped launch each over all levels
        AFI_TRY(ew_combine_group(ncalls, DC5, dt, d_out, DT_F32, nullptr, 0, nullptr, 0, 0.2f, 0.2f * d_scale, dn, dh, dw_, C, st));
        AFI_TRY(ew_combine_group(ncalls, GA, DT_F32, d_out, DT_F32, nullptr, 0, nullptr, 0, 0.2f, d_scale, dn, dh, dw_, C, st));
        AFI_TRY(wgrad_std(ctx, prec, ncalls, d1, Br, CB, DC5, C, gradacc + GL.rdb_w[r][4], st));
        conv_std(a, ncalls, d1, C, CB, pk + L.rdb_d[r][4] * wes);
        a.out_dt = DT_F32;
        for (int k = 0; k < ncalls; k++) { a.p[k].in[0] = DC5[k]; a.p[k].accin = GA[k]; a.p[k].out = GA[k]; }
        AFI_TRY(run_conv(ctx, prec, a, st));
        for (int i = 3; i >= 0; i--) {
            int cin_f = C + GR * i;
            // g = GA[:, slice_i] * lrelu'(c_{i+1})
            // (the four masked growth gradients sit side by side in one 128-channel buffer: slice i feeds this dgrad, all four feed ONE
            //  weight-gradient GEMM below)
            PView gdst[AFI_MAX_PROB], gsrc[AFI_MAX_PROB], gmask[AFI_MAX_PROB];
            for (int k = 0; k < ncalls; k++) { gdst[k] = pview_ch(GC[k], GR * i, es); gsrc[k] = pview_ch(GA[k], cin_f, 4); gmask[k] = pview_ch(Br[k], cin_f, es); }
            AFI_TRY(ew_combine_group(ncalls, gdst, dt, gsrc, DT_F32, nullptr, 0, gmask, dt, 0.2f, 1.f, dn, dh, dw_, GR, st));
            if (i > 0) {      // the part of this conv's dgrad that the remaining masks depend on: growth channels [256, 256 + 32 i)
                conv_std(a, ncalls, d1, GR, GR * i, pk + L.rdb_cd[r][i] * wes);
                a.out_dt = DT_F32;
                for (int k = 0; k < ncalls; k++) {
                    a.p[k].in[0] = pview_ch(GC[k], GR * i, es); a.p[k].accin = pview_ch(GA[k], C, 4); a.p[k].out = pview_ch(GA[k], C, 4);
                }
                AFI_TRY(run_conv(ctx, prec, a, st));
            }
        }
        // ... and the part that reaches the block input x, for all four growth convs in ONE GEMM (K = 4 x 32 channels per tap, N = 256)
        conv_std(a, ncalls, d1, 4 * GR, C, pk + L.rdb_xd[r] * wes);
        a.out_dt = DT_F32;
        for (int k = 0; k < ncalls; k++) { a.p[k].in[0] = GC[k]; a.p[k].accin = GA[k]; a.p[k].out = GA[k]; }
        AFI_TRY(run_conv(ctx, prec, a, st));
        // weight gradients of the four growth convs in one GEMM: dW[128 = 4 x 32 couts][352 cins] per tap fills a whole 128-row MMA tile
        // (four separate 32-cout GEMMs use a quarter of it each); the non-causal blocks are computed and ignored
        AFI_TRY(wgrad_std(ctx, prec, ncalls, d1, Br, C + 3 * GR, GC, 4 * GR, gradacc + GL.rdb_g[r], st));
        for (int k = 0; k < ncalls; k++) d_out[k] = GA[k];
        d_scale = 1.f; cur ^= 1;
    }
    // head conv: g_head = (d_out + dH1) * lrelu'(h0)
    {
        PView hmask[AFI_MAX_PROB];
        for (int k = 0; k < ncalls; k++) hmask[k] = pview(W[k].B[0], d1[k].h, d1[k].w, CB);
        AFI_TRY(ew_combine_group(ncalls, GH, dt, d_out, DT_F32, dH1, DT_F32, hmask, dt, 0.2f, 1.f, dn, dh, dw_, C, st));
    }
    AFI_TRY(wgrad_std(ctx, prec, ncalls, d1, X0, C, GH, C, gradacc + GL.head_w, st));
    for (int k = 0; k < ncalls; k++) AFI_TRY(col_sum_f32(GH[k], dt, d1[k].n, d1[k].h, d1[k].w, C, gradacc + GL.head_b, st));
    if (calls[0].dx) {
        // dx = dgrad(head conv)(g_head) + scale * bilinear2x^T(dy)   (needed when the interpolator sits inside a detector: stage 2/3)
        conv_std(a, ncalls, d1, C, C, pk + L.head_d * wes);
        a.out_dt = DT_F32;
        for (int k = 0; k < ncalls; k++) { a.p[k].in[0] = GH[k]; a.p[k].out = pview(W[k].DXb, d1[k].h, d1[k].w, C); }
        AFI_TRY(run_conv(ctx, prec, a, st));
        for (int k = 0; k < ncalls; k++) {
            const afi_g_call& c = calls[k];
            AFI_TRY(g_input_grad(pview(W[k].DXb, c.h, c.w, C), c.dy, c.lateral ? c.lateral->scale : 1.f, c.n, C, c.h, c.w, c.oh, c.ow, c.dx, st));
        }
    }
    (void)tmpx;
    return AFI_OK;
}

int afi_g_unpack_grads(afi_ctx* ctx, int prec, const float* gradacc, const afi_g_grads* g, float scale, int accumulate, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(ctx && gradacc && g, "afi_g_unpack_grads: null argument");
    AFI_TRY(g_check(prec, 1, 1, 1, g->n_rdb));
    GGradAcc GL = g_gradacc_layout(g->n_rdb);
    int nk = prec_nk(prec) ? 1 : 0;
    // every gradient tensor of the module in ONE launch
    UnpackJob jobs[AFI_MAX_UNPACK]; int nj = 0;
    auto wjob = [&](size_t off, float* dst, int co, int ci, int CO, int CI, int co_off, int kind) {
        if (!dst) return;
        UnpackJob& j = jobs[nj++]; memset(&j, 0, sizeof(j));
        j.src = gradacc + off; j.dst = dst; j.kind = kind; j.nk = nk; j.co = co; j.ci = ci; j.CO = CO; j.CI = CI; j.co_off = co_off;
        j.total = (long long)co * ci * (kind == UNPACK_DECONV ? 36 : 9);
    };
    auto cjob = [&](size_t off, float* dst, long long n) {
        if (!dst) return;
        UnpackJob& j = jobs[nj++]; memset(&j, 0, sizeof(j));
        j.src = gradacc + off; j.dst = dst; j.kind = UNPACK_COPY; j.total = n;
    };
    wjob(GL.head_w, g->head_w, C, C, C, C, 0, UNPACK_STD); cjob(GL.head_b, g->head_b, C);
    for (int r = 0; r < g->n_rdb; r++) {
        for (int i = 0; i < 4; i++) wjob(GL.rdb_g[r], g->rdb_w[r][i], GR, C + GR * i, 4 * GR, C + 3 * GR, GR * i, UNPACK_STD);
        wjob(GL.rdb_w[r][4], g->rdb_w[r][4], C, CB, C, CB, 0, UNPACK_STD);
    }
    wjob(GL.post_w, g->post_w, C, C, C, C, 0, UNPACK_STD); cjob(GL.post_b, g->post_b, C);
    wjob(GL.up_w, g->up_w, C, C, C, C, 0, UNPACK_DECONV); cjob(GL.up_b, g->up_b, C);
    wjob(GL.out_w, g->out_w, C, C, C, C, 0, UNPACK_STD); cjob(GL.out_b, g->out_b, C);
    return unpack_group(nj, jobs, scale, accumulate, st);
}

}  // extern "C"

// =====================================================================================================
// Discriminator
// =====================================================================================================
namespace afi {
static const int DC[4] = {256, 512, 1024, 1024};

struct DPacked { size_t f[3], d[3], hf, hb, total; };   // hf / hb: tensor-core head packs [16][1024] / [1024][16]
static DPacked d_packed_layout() {
    DPacked L; size_t o = 0;
    for (int i = 0; i < 3; i++) { L.f[i] = o; o += (size_t)9 * DC[i] * DC[i + 1]; L.d[i] = o; o += (size_t)9 * DC[i] * DC[i + 1]; }
    L.hf = o; o += 16 * DC[3];
    L.hb = o; o += 16 * DC[3];
    L.total = o;
    return L;
}
struct DGradAcc { size_t w[4], b[4], gamma[3], beta[3], total; };
static DGradAcc d_gradacc_layout() {
    DGradAcc L; size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 63) / 64 * 64; return r; };
    for (int i = 0; i < 3; i++) { L.w[i] = take((size_t)9 * DC[i] * DC[i + 1]); L.b[i] = take(DC[i + 1]); L.gamma[i] = take(DC[i + 1]); L.beta[i] = take(DC[i + 1]); }
    L.w[3] = take(16 * DC[3]); L.b[3] = take(1);   // head: torch layout [1024][9] (CUDA-core engine) or [16][1024] (tensor-core engine)
    L.total = o;
    return L;
}
struct DWs {
    void *A[4], *Z[3];
    float *mean[3], *rstd[3], *var[3], *T9;   // per-layer batch mean / 1/sqrt(var+eps) / unbiased var of THIS call
    double* sums;          // [2][1024]
    void *DY[3], *DXb, *G9;
    float* G9F;            // fp32 copy [P][12] of the shifted head gradients
    void* SPL; size_t spl_bytes;   // AFI_PREC_SPLIT: bf16 operand planes of the GEMM in flight
    size_t total;
};
static DWs d_ws_layout(void* base, int prec, int n, int h, int w, int backward) {
    DWs W; memset(&W, 0, sizeof(W));
    Carver cv(base);
    size_t es = dt_size(prec_dt(prec)), P = (size_t)n * h * w;
    W.A[0] = cv.take(P * DC[0] * es);
    for (int i = 0; i < 3; i++) { W.Z[i] = cv.take(P * DC[i + 1] * es); W.A[i + 1] = cv.take(P * DC[i + 1] * es); }
    for (int i = 0; i < 3; i++) { W.mean[i] = (float*)cv.take(1024 * 4); W.rstd[i] = (float*)cv.take(1024 * 4); W.var[i] = (float*)cv.take(1024 * 4); }
    W.T9 = (float*)cv.take(P * 16 * 4);
    W.sums = (double*)cv.take(2 * 1024 * 8);
    if (backward) {
        for (int i = 0; i < 3; i++) W.DY[i] = cv.take(P * DC[i + 1] * es);
        W.DXb = cv.take(P * DC[0] * 4);
        W.G9 = cv.take(P * 16 * 2);
        W.G9F = (float*)cv.take(P * 12 * 4);
    }
    if (prec == AFI_PREC_SPLIT) {      // the widest GEMM: layer 3's weight gradient (x and dy of 1024 channels); forward only: one 1024-ch input
        W.spl_bytes = (backward ? 2 : 1) * split_planes_bytes((long long)P, 1024);
        W.SPL = cv.take(W.spl_bytes);
    }
    W.total = cv.off;
    return W;
}
}  // namespace afi

extern "C" {

size_t afi_d_packed_bytes(int prec) { return d_packed_layout().total * prec_wes(prec); }
size_t afi_d_gradacc_bytes(void) { return d_gradacc_layout().total * 4; }
size_t afi_d_workspace_bytes(int prec, int n, int h, int w, int save_for_backward) { return d_ws_layout(nullptr, prec, n, h, w, save_for_backward).total; }

int afi_d_pack(afi_ctx* ctx, int prec, const afi_d_params* p, void* packed, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(ctx && p && packed && prec_ok(prec), "afi_d_pack: bad argument");
    DPacked L = d_packed_layout();
    const int dt = prec_wdt(prec); const size_t wes = prec_wes(prec);
    PackJob jobs[6];
    for (int i = 0; i < 3; i++) {
        memset(&jobs[2 * i], 0, 2 * sizeof(PackJob));
        jobs[2 * i].w = jobs[2 * i + 1].w = p->w[i];
        jobs[2 * i].dst = (char*)packed + L.f[i] * wes; jobs[2 * i + 1].dst = (char*)packed + L.d[i] * wes;
        jobs[2 * i].co = jobs[2 * i + 1].co = DC[i + 1]; jobs[2 * i].ci = jobs[2 * i + 1].ci = DC[i];
        jobs[2 * i].mode = pm(prec, 0); jobs[2 * i + 1].mode = pm(prec, 1);
    }
    AFI_TRY(pack_weights_group(6, jobs, dt, st));
    if (prec_tc(prec)) AFI_TRY(dhead_pack_tc(p->w[3], DC[3], (char*)packed + L.hf * wes, (char*)packed + L.hb * wes, st));
    return AFI_OK;
}

static int d_calls_check(const char* who, int prec, const afi_d_call* calls, int ncalls, int save, DWs* W, Dim3* d) {
    AFI_REQUIRE(prec_ok(prec) && calls, "%s: bad argument", who);
    AFI_REQUIRE(ncalls >= 1 && ncalls <= AFI_MAX_PROB, "%s: %d calls per group (max %d)", who, ncalls, AFI_MAX_PROB);
    for (int k = 0; k < ncalls; k++) {
        const afi_d_call& c = calls[k];
        AFI_REQUIRE(c.ws && c.n >= 1 && c.h >= 1 && c.w >= 1, "%s: call %d: bad shape n=%d h=%d w=%d or null workspace", who, k, c.n, c.h, c.w);
        W[k] = d_ws_layout(c.ws, prec, c.n, c.h, c.w, save);
        if (W[k].total > c.ws_bytes) { set_error("%s: workspace %zu B < required %zu B", who, c.ws_bytes, W[k].total); return AFI_ERR_WORKSPACE; }
        g_ss.p[k] = W[k].SPL; g_ss.n[k] = W[k].spl_bytes;
        d[k] = {c.n, c.h, c.w};
    }
    g_ss.k0 = 0;
    return AFI_OK;
}

int afi_d_forward(afi_ctx* ctx, int prec, const afi_d_params* p, const void* packed, const afi_d_call* calls, int ncalls, int training,
                  float momentum, float eps, int save, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(ctx && p && packed, "afi_d_forward: null argument");
    DWs W[AFI_MAX_PROB]; Dim3 d[AFI_MAX_PROB];
    AFI_TRY(d_calls_check("afi_d_forward", prec, calls, ncalls, save, W, d));
    g_ss.pairs = split_pairs_env(save ? 6 : 3);
    const int dt = prec_dt(prec); const size_t wes = prec_wes(prec);
    DPacked L = d_packed_layout();
    for (int k = 0; k < ncalls; k++) {
        AFI_REQUIRE(calls[k].x.ptr, "afi_d_forward: call %d has a null input", k);
        // logits == NULL: the call is evaluated for its BatchNorm statistics only (stage1_trainer.py:400 computes D(hr) in the G phase and
        // drops the result: what survives is the running-buffer update), so layer 3's normalise pass and the head are skipped
        AFI_REQUIRE(calls[k].logits || (training && !save), "afi_d_forward: call %d: a statistics-only call (logits == NULL) needs training mode and no backward", k);
        if (!calls[k].input_staged) AFI_TRY(to_nhwc(prec, calls[k].x, d[k].n, DC[0], d[k].h, d[k].w, pview(W[k].A[0], d[k].h, d[k].w, DC[0]), st));
    }
    const bool tc = prec_tc(prec);
    for (int i = 0; i < 3; i++) {
        // Conv2d 3x3 + bias (all calls in one grouped launch) -> per call: BatchNorm with THIS call's batch statistics
        // -> LeakyReLU(0.2)                                                         feature_patch_discriminator.py:36-38
        ConvArgs a;
        conv_std(a, ncalls, d, DC[i], DC[i + 1], (const char*)packed + L.f[i] * wes);
        a.bias = p->b[i]; a.out_dt = dt;
        for (int k = 0; k < ncalls; k++) { a.p[k].in[0] = pview(W[k].A[i], d[k].h, d[k].w, DC[i]); a.p[k].out = pview(W[k].Z[i], d[k].h, d[k].w, DC[i + 1]); }
        void* sums_p[AFI_MAX_PROB]; const double* sum_c[AFI_MAX_PROB]; const double* sq_c[AFI_MAX_PROB]; double* sum_m[AFI_MAX_PROB]; double* sq_m[AFI_MAX_PROB];
        float* mean_p[AFI_MAX_PROB]; float* rstd_p[AFI_MAX_PROB]; float* var_p[AFI_MAX_PROB]; const float* mean_c[AFI_MAX_PROB]; const float* rstd_c[AFI_MAX_PROB];
        const float* var_c[AFI_MAX_PROB]; long long cnt[AFI_MAX_PROB]; PView Zv[AFI_MAX_PROB], Av[AFI_MAX_PROB];
        for (int k = 0; k < ncalls; k++) {
            sums_p[k] = W[k].sums; sum_c[k] = sum_m[k] = W[k].sums; sq_c[k] = sq_m[k] = W[k].sums + 1024;
            mean_p[k] = W[k].mean[i]; rstd_p[k] = W[k].rstd[i]; var_p[k] = W[k].var[i];
            mean_c[k] = W[k].mean[i]; rstd_c[k] = W[k].rstd[i]; var_c[k] = W[k].var[i];
            cnt[k] = (long long)d[k].n * d[k].h * d[k].w;
            Zv[k] = pview(W[k].Z[i], d[k].h, d[k].w, DC[i + 1]); Av[k] = pview(W[k].A[i + 1], d[k].h, d[k].w, DC[i + 1]);
        }
        if (training) AFI_TRY(zero_group(ncalls, sums_p, 2 * 1024 * sizeof(double), st));
        const bool fuse_stats = training && tc && DC[i] >= 512;   // long-K layers hide the extra epilogue work behind the MMAs
        if (fuse_stats) {   // tensor-core engine: the per-channel sum / sum of squares come out of the GEMM epilogue
            a.stat_mode = 1;
            for (int k = 0; k < ncalls; k++) { a.p[k].stat0 = W[k].sums; a.p[k].stat1 = W[k].sums + 1024; }
        }
        AFI_TRY(run_conv(ctx, prec, a, st));
        if (training) {
            // per call: BatchNorm with THIS call's batch statistics (one grouped launch per pass covers all calls)
            if (!fuse_stats) AFI_TRY(col_reduce_group(0, ncalls, Zv, nullptr, dt, nullptr, nullptr, sum_m, sq_m, cnt, DC[i + 1], st));
            AFI_TRY(bn_finalize_group(ncalls, sum_c, sq_c, cnt, DC[i + 1], eps, mean_p, rstd_p, var_p, st));
            // training == 1: running buffers updated here, in call order; training == 2: deferred to afi_d_update_running
            if (training != 2)
                AFI_TRY(bn_update_running(ncalls, mean_c, var_c, DC[i + 1], momentum, p->running_mean[i], p->running_var[i], p->num_batches_tracked[i], st));
        } else {
            for (int k = 0; k < ncalls; k++)
                AFI_TRY(bn_finalize(W[k].sums, W[k].sums + 1024, cnt[k], DC[i + 1], eps, momentum, 0, W[k].mean[i], W[k].rstd[i], W[k].var[i],
                                    p->running_mean[i], p->running_var[i], nullptr, st));
        }
        long long cnt_apply[AFI_MAX_PROB];
        for (int k = 0; k < ncalls; k++) cnt_apply[k] = (i == 2 && !calls[k].logits) ? 0 : cnt[k];
        AFI_TRY(bn_apply_lrelu_group(ncalls, Zv, Av, dt, mean_c, rstd_c, p->gamma[i], p->beta[i], 0.2f, cnt_apply, DC[i + 1], st));
    }
    // Conv2d 1024 -> 1                                                                   feature_patch_discriminator.py:40-41
    if (tc) {
        ConvArgs a;
        conv_args_init(a);
        a.cin = DC[3]; a.cout = 16; a.ntaps = 1; a.nprob = ncalls; a.w = (const char*)packed + L.hf * wes; a.out_dt = DT_F32;
        for (int k = 0; k < ncalls; k++) {
            a.p[k].N = calls[k].logits ? d[k].n : 0; a.p[k].H = d[k].h; a.p[k].W = d[k].w;
            a.p[k].in[0] = pview(W[k].A[3], d[k].h, d[k].w, DC[3]); a.p[k].out = pview(W[k].T9, d[k].h, d[k].w, 16);
        }
        AFI_TRY(run_conv(ctx, prec, a, st));
        for (int k = 0; k < ncalls; k++)
            if (calls[k].logits) AFI_TRY(dhead_stencil16(W[k].T9, p->b[3], d[k].n, d[k].h, d[k].w, calls[k].logits, st));
    } else {
        for (int k = 0; k < ncalls; k++)
            if (calls[k].logits) AFI_TRY(dhead_forward(pview(W[k].A[3], d[k].h, d[k].w, DC[3]), dt, p->w[3], p->b[3], d[k].n, d[k].h, d[k].w, DC[3], W[k].T9, calls[k].logits, st));
    }
    return AFI_OK;
}

int afi_d_update_running(afi_ctx* ctx, int prec, const afi_d_params* p, const afi_d_call* calls, int ncalls, float momentum, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(ctx && p, "afi_d_update_running: null argument");
    DWs W[AFI_MAX_PROB]; Dim3 d[AFI_MAX_PROB];
    AFI_TRY(d_calls_check("afi_d_update_running", prec, calls, ncalls, 0, W, d));
    for (int i = 0; i < 3; i++) {
        const float* mean[AFI_MAX_PROB]; const float* var[AFI_MAX_PROB];
        for (int k = 0; k < ncalls; k++) { mean[k] = W[k].mean[i]; var[k] = W[k].var[i]; }
        AFI_TRY(bn_update_running(ncalls, mean, var, DC[i + 1], momentum, p->running_mean[i], p->running_var[i], p->num_batches_tracked[i], st));
    }
    return AFI_OK;
}

int afi_d_backward(afi_ctx* ctx, int prec, const afi_d_params* p, const void* packed, const afi_d_call* calls, int ncalls, int training,
                   float* gradacc, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(ctx && p && packed && gradacc, "afi_d_backward: null argument");
    DWs W[AFI_MAX_PROB]; Dim3 d[AFI_MAX_PROB];
    AFI_TRY(d_calls_check("afi_d_backward", prec, calls, ncalls, 1, W, d));
    g_ss.pairs = split_pairs_env(3);
    const int dt = prec_dt(prec); const size_t wes = prec_wes(prec);
    DPacked L = d_packed_layout();
    DGradAcc GL = d_gradacc_layout();
    PView X[AFI_MAX_PROB], DYv[AFI_MAX_PROB];
    const bool tc = prec_tc(prec);
    for (int k = 0; k < ncalls; k++) {
        AFI_REQUIRE(calls[k].dlogits, "afi_d_backward: call %d has no dlogits", k);
        AFI_REQUIRE((calls[k].dx != nullptr) == (calls[0].dx != nullptr), "afi_d_backward: either every call of a group or none asks for dx");
    }
    if (tc) {
        // head: dy3 = (sum_t g[q - tap_t] w4[:, t]) * lrelu'(a3) is 9 FMAs per element, so it is never stored: two grouped HBM-bound passes
        // over z3 recompute it, the first for the two BatchNorm-backward reductions of layer 3, the second for dz3 itself;
        // dW4 = g9^T a3 is a K = pixels GEMM on the tensor cores; db4 = sum g.
        WgradArgs g;
        memset(&g, 0, sizeof(g));
        g.cin = DC[3]; g.cout = 16; g.ntaps = 1; g.nprob = ncalls; g.dw = gradacc + GL.w[3];
        void* sums_p[AFI_MAX_PROB]; double* s0[AFI_MAX_PROB]; double* s1[AFI_MAX_PROB]; const float* mean_c[AFI_MAX_PROB]; const float* rstd_c[AFI_MAX_PROB];
        long long cnt[AFI_MAX_PROB]; PView Z3[AFI_MAX_PROB], DZ3[AFI_MAX_PROB];
        for (int k = 0; k < ncalls; k++) {
            sums_p[k] = W[k].sums; s0[k] = W[k].sums; s1[k] = W[k].sums + 1024; mean_c[k] = W[k].mean[2]; rstd_c[k] = W[k].rstd[2];
            cnt[k] = (long long)d[k].n * d[k].h * d[k].w;
            Z3[k] = pview(W[k].Z[2], d[k].h, d[k].w, DC[3]); DZ3[k] = pview(W[k].DY[2], d[k].h, d[k].w, DC[3]);
        }
        AFI_TRY(zero_group(ncalls, sums_p, 2 * 1024 * sizeof(double), st));
        for (int k = 0; k < ncalls; k++) {
            const int n = d[k].n, h = d[k].h, w = d[k].w;
            AFI_TRY(dhead_build_g9(calls[k].dlogits, n, h, w, W[k].G9, W[k].G9F, st));
            AFI_TRY(sum_f32(calls[k].dlogits, (long long)n * h * w, gradacc + GL.b[3], st));
            g.p[k].N = n; g.p[k].H = h; g.p[k].W = w; g.p[k].x = pview(W[k].A[3], h, w, DC[3]); g.p[k].dy = pview(W[k].G9, h, w, 16);
        }
        const void* g9b[AFI_MAX_PROB];
        for (int k = 0; k < ncalls; k++) g9b[k] = W[k].G9;
        // (an FMA formulation of these two passes was FMA-bound at 0.9 + 0.8 ms per step; the 9-tap product on warp-level MMAs is HBM-bound:
        //  profiles/r01_hbm_passes.md)
        for (int pass = 1; pass <= 2; pass++)
            AFI_TRY(dhead_backward_group_mma(pass, ncalls, g9b, Z3, DZ3, p->w[3], mean_c, rstd_c, p->gamma[2], p->beta[2], s0, s1,
                                             gradacc + GL.gamma[2], gradacc + GL.beta[2], cnt, DC[3], training ? 0 : 1, st));
        AFI_TRY(run_wgrad(ctx, prec, g, st));
    } else {
        for (int k = 0; k < ncalls; k++)   // head: dW4, db4 and dy3 = dA3 * lrelu'(a3) in one pass over a3
            AFI_TRY(dhead_backward(pview(W[k].A[3], d[k].h, d[k].w, DC[3]), dt, p->w[3], calls[k].dlogits, d[k].n, d[k].h, d[k].w, DC[3],
                                   gradacc + GL.w[3], gradacc + GL.b[3], pview(W[k].DY[2], d[k].h, d[k].w, DC[3]), st));
    }
    for (int i = 2; i >= 0; i--) {
        const int co = DC[i + 1], ci = DC[i];
        void* sums_p[AFI_MAX_PROB]; double* s0[AFI_MAX_PROB]; double* s1[AFI_MAX_PROB]; const float* mean_c[AFI_MAX_PROB]; const float* rstd_c[AFI_MAX_PROB];
        long long cnt[AFI_MAX_PROB]; PView Zv[AFI_MAX_PROB];
        for (int k = 0; k < ncalls; k++) {
            Zv[k] = pview(W[k].Z[i], d[k].h, d[k].w, co);
            DYv[k] = pview(W[k].DY[i], d[k].h, d[k].w, co);
            X[k] = pview(W[k].A[i], d[k].h, d[k].w, ci);
            sums_p[k] = W[k].sums; s0[k] = W[k].sums; s1[k] = W[k].sums + 1024; mean_c[k] = W[k].mean[i]; rstd_c[k] = W[k].rstd[i];
            cnt[k] = (long long)d[k].n * d[k].h * d[k].w;
        }
        // train-mode BatchNorm backward in closed form: two per-channel reductions, then one elementwise pass (in place: DY -> DZ);
        // each pass is one grouped launch over all calls
        if (!(tc && i == 2)) {      // (the tensor-core path did layer 3's BatchNorm backward in the fused head passes above)
            AFI_TRY(zero_group(ncalls, sums_p, 2 * 1024 * sizeof(double), st));
            AFI_TRY(col_reduce_group(1, ncalls, DYv, Zv, dt, mean_c, rstd_c, s0, s1, cnt, co, st));
            AFI_TRY(bn_bwd_apply_group(ncalls, DYv, Zv, dt, mean_c, rstd_c, p->gamma[i], s0, s1, gradacc + GL.gamma[i], gradacc + GL.beta[i], cnt, co,
                                       training ? 0 : 1, st));
        }
        AFI_TRY(wgrad_std(ctx, prec, ncalls, d, X, ci, DYv, co, gradacc + GL.w[i], st));
        // bias gradient: in training mode this bias feeds a batch-statistics BatchNorm, so dL/db = sum_p dz = 0 identically (the reference
        // gets ~1e-9 rounding noise there) and the accumulator slot stays at its zero-initialised value; in eval mode the
        // statistics are constants and db = sum_p dz.
        if (!training) for (int k = 0; k < ncalls; k++) AFI_TRY(col_sum_f32(DYv[k], dt, d[k].n, d[k].h, d[k].w, co, gradacc + GL.b[i], st));
        if (i == 0 && calls[0].dx) {   // input gradient (never needed by the stage-1/2 trainers, which detach D's input; kept for autograd completeness)
            ConvArgs a;
            conv_std(a, ncalls, d, co, ci, (const char*)packed + L.d[0] * wes);
            a.out_dt = DT_F32;
            for (int k = 0; k < ncalls; k++) { a.p[k].in[0] = DYv[k]; a.p[k].out = pview(W[k].DXb, d[k].h, d[k].w, ci); }
            AFI_TRY(run_conv(ctx, prec, a, st));
            afi_view4 none; memset(&none, 0, sizeof(none));
            for (int k = 0; k < ncalls; k++)
                AFI_TRY(nhwc_to_nchw<float>(pview(W[k].DXb, d[k].h, d[k].w, ci), pview_null(), none, 0, 0, 1.f, d[k].n, ci, d[k].h, d[k].w, calls[k].dx, st));
        }
        if (i > 0) {   // dA_i * lrelu'(a_i) -> DY[i-1]
            ConvArgs a;
            conv_std(a, ncalls, d, co, ci, (const char*)packed + L.d[i] * wes);
            a.out_dt = dt;
            for (int k = 0; k < ncalls; k++) { a.p[k].in[0] = DYv[k]; a.p[k].mask = X[k]; a.p[k].out = pview(W[k].DY[i - 1], d[k].h, d[k].w, ci); }
            // (emitting the layer below's two BatchNorm-backward reductions from this dgrad's epilogue measured 0.4 ms/step SLOWER than the
            //  separate grouped reduction pass -- 42.7 vs 42.3 ms: the extra operand stream and shuffles cost the GEMM more than the 0.37 ms
            //  pass they replace; round-1 experiment, removed)
            AFI_TRY(run_conv(ctx, prec, a, st));
        }
    }
    return AFI_OK;
}

int afi_d_unpack_grads(afi_ctx* ctx, int prec, const float* gradacc, const afi_d_grads* g, float scale, int accumulate, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(ctx && gradacc && g && prec_ok(prec), "afi_d_unpack_grads: bad argument");
    DGradAcc GL = d_gradacc_layout();
    const int nk = prec_nk(prec) ? 1 : 0, head_tc = prec_tc(prec) ? 1 : 0;
    UnpackJob jobs[AFI_MAX_UNPACK]; int nj = 0;
    auto cjob = [&](size_t off, float* dst, long long n) {
        if (!dst) return;
        UnpackJob& j = jobs[nj++]; memset(&j, 0, sizeof(j));
        j.src = gradacc + off; j.dst = dst; j.kind = UNPACK_COPY; j.total = n;
    };
    for (int i = 0; i < 3; i++) {
        if (g->w[i]) {
            UnpackJob& j = jobs[nj++]; memset(&j, 0, sizeof(j));
            j.src = gradacc + GL.w[i]; j.dst = g->w[i]; j.kind = UNPACK_STD; j.nk = nk; j.co = j.CO = DC[i + 1]; j.ci = j.CI = DC[i];
            j.total = (long long)9 * DC[i] * DC[i + 1];
        }
        cjob(GL.b[i], g->b[i], DC[i + 1]); cjob(GL.gamma[i], g->gamma[i], DC[i + 1]); cjob(GL.beta[i], g->beta[i], DC[i + 1]);
    }
    if (g->w[3]) {
        if (head_tc) {
            UnpackJob& j = jobs[nj++]; memset(&j, 0, sizeof(j));
            j.src = gradacc + GL.w[3]; j.dst = g->w[3]; j.kind = UNPACK_HEAD_TC; j.ci = DC[3]; j.total = (long long)9 * DC[3];
        } else cjob(GL.w[3], g->w[3], 9 * DC[3]);
    }
    cjob(GL.b[3], g->b[3], 1);
    return unpack_group(nj, jobs, scale, accumulate, st);
}

// =====================================================================================================
// single-layer entry points (unit tests / kernel benchmarks)
// =====================================================================================================
static inline int pad64(int c) { return c <= 64 ? c : (c + 63) / 64 * 64; }
static size_t conv_single_ws_bytes(int prec, int ks, int n, int cin, int h, int w, int cout) {
    size_t es = dt_size(prec_dt(prec)), wes = prec_wes(prec), P = (size_t)n * h * w, nt = (size_t)ks * ks;
    return align_up(P * pad64(cin) * es) + 2 * align_up(P * pad64(cout) * es) + 2 * align_up(nt * cin * cout * wes) +
           align_up(nt * cin * cout * 4) + align_up(P * cin * 4) + 4096 +
           (prec == AFI_PREC_SPLIT ? split_planes_bytes((long long)P, cin) + split_planes_bytes((long long)P, cout) + 256 : 0);
}
static void single_taps(ConvArgs& a, int ks) {
    if (ks == 1) { a.ntaps = 1; a.taps[0].dy = 0; a.taps[0].dx = 0; a.taps[0].view = 0; a.taps[0].slab = 0; }
}
// one stride-1 "same" convolution (ks = 1 or 3) through the GEMM engine of `prec`: y = [lrelu](conv(x, w) + b)
static int conv_single(afi_ctx* ctx, int prec, int ks, afi_view4 x, int n, int cin, int h, int w, const float* weight, const float* bias, int cout,
                       int lrelu, float* y, void* ws, size_t ws_bytes, cudaStream_t st) {
    AFI_REQUIRE(ctx && x.ptr && weight && y && ws && prec_ok(prec), "afi_conv: bad argument");
    AFI_REQUIRE(cin % 32 == 0 && cout % 32 == 0, "afi_conv: channels must be multiples of 32");
    if (ws_bytes < conv_single_ws_bytes(prec, ks, n, cin, h, w, cout)) { set_error("afi_conv: workspace too small"); return AFI_ERR_WORKSPACE; }
    int dt = prec_dt(prec); size_t es = dt_size(dt), wes = prec_wes(prec), P = (size_t)n * h * w, nt = (size_t)ks * ks;
    Carver cv(ws);
    void* X = cv.take(P * pad64(cin) * es); void* Y = cv.take(P * pad64(cout) * es); cv.take(P * pad64(cout) * es);
    void* Wp = cv.take(nt * cin * cout * wes); cv.take(nt * cin * cout * wes);
    cv.take(nt * cin * cout * 4); cv.take(P * cin * 4);
    g_ss.k0 = 0; g_ss.n[0] = split_planes_bytes((long long)P, cin) + split_planes_bytes((long long)P, cout); g_ss.p[0] = cv.take(g_ss.n[0]);
    g_ss.pairs = 6;
    AFI_TRY(to_nhwc(prec, x, n, cin, h, w, pview(X, h, w, cin), st));
    AFI_TRY(pack_weights(weight, cout, cin, pm(prec, ks == 1 ? 4 : 0), Wp, prec_wdt(prec), st));
    ConvArgs a;
    Dim3 d = {n, h, w};
    conv_std(a, 1, &d, cin, cout, Wp);
    single_taps(a, ks);
    a.bias = bias; a.act = lrelu; a.out_dt = dt;
    a.p[0].in[0] = pview(X, h, w, cin); a.p[0].out = pview(Y, h, w, cout);
    AFI_TRY(run_conv(ctx, prec, a, st));
    afi_view4 none; memset(&none, 0, sizeof(none));
    AFI_TRY(to_nchw(prec, pview(Y, h, w, cout), pview_null(), none, 0, 0, 1.f, n, cout, h, w, y, st));
    return AFI_OK;
}
static int conv_single_backward(afi_ctx* ctx, int prec, int ks, afi_view4 x, afi_view4 dy, int n, int cin, int h, int w, const float* weight, int cout,
                                float* dw, float* db, float* dxo, void* ws, size_t ws_bytes, cudaStream_t st) {
    AFI_REQUIRE(ctx && x.ptr && dy.ptr && weight && dw && ws && prec_ok(prec), "afi_conv_backward: bad argument");
    AFI_REQUIRE(cin % 32 == 0 && cout % 32 == 0, "afi_conv_backward: channels must be multiples of 32");
    if (ws_bytes < conv_single_ws_bytes(prec, ks, n, cin, h, w, cout)) { set_error("afi_conv_backward: workspace too small"); return AFI_ERR_WORKSPACE; }
    int dt = prec_dt(prec); size_t es = dt_size(dt), wes = prec_wes(prec), P = (size_t)n * h * w, nt = (size_t)ks * ks;
    Carver cv(ws);
    // pixel strides padded to whole 64-channel groups (the tensor-core wgrad reads operands through grouped TMA views)
    const int xs = pad64(cin), ys = pad64(cout);
    void* X = cv.take(P * xs * es); void* DYb = cv.take(P * ys * es); cv.take(P * ys * es);
    void* Wp = cv.take(nt * cin * cout * wes); cv.take(nt * cin * cout * wes);
    float* acc = (float*)cv.take(nt * cin * cout * 4);
    void* DX = cv.take(P * cin * 4);
    g_ss.k0 = 0; g_ss.n[0] = split_planes_bytes((long long)P, cin) + split_planes_bytes((long long)P, cout); g_ss.p[0] = cv.take(g_ss.n[0]);
    g_ss.pairs = split_pairs_env(3);
    if (xs != cin) AFI_CUDA(cudaMemsetAsync(X, 0, P * xs * es, st));
    if (ys != cout) AFI_CUDA(cudaMemsetAsync(DYb, 0, P * ys * es, st));
    PView Xv = pview(X, h, w, xs), DYv = pview(DYb, h, w, ys);
    AFI_TRY(to_nhwc(prec, x, n, cin, h, w, Xv, st));
    AFI_TRY(to_nhwc(prec, dy, n, cout, h, w, DYv, st));
    AFI_CUDA(cudaMemsetAsync(acc, 0, nt * cin * cout * 4, st));
    Dim3 d = {n, h, w};
    if (ks == 1) {
        WgradArgs g;
        memset(&g, 0, sizeof(g));
        g.cin = cin; g.cout = cout; g.ntaps = 1; g.nprob = 1; g.dw = acc;
        g.p[0].N = n; g.p[0].H = h; g.p[0].W = w; g.p[0].x = Xv; g.p[0].dy = DYv;
        AFI_TRY(run_wgrad(ctx, prec, g, st));
        AFI_TRY(unpack_1x1(acc, cout, cin, prec_nk(prec) ? 1 : 0, dw, 1.f, 0, st));
    } else {
        AFI_TRY(wgrad_std(ctx, prec, 1, &d, &Xv, cin, &DYv, cout, acc, st));
        AFI_TRY(unpack_wgrad(acc, cout, cin, prec_nk(prec) ? 1 : 0, 0, dw, 1.f, 0, st));
    }
    if (db) {
        AFI_CUDA(cudaMemsetAsync(db, 0, cout * sizeof(float), st));
        AFI_TRY(col_sum_f32(DYv, dt, n, h, w, cout, db, st));
    }
    if (dxo) {
        AFI_TRY(pack_weights(weight, cout, cin, pm(prec, ks == 1 ? 5 : 1), Wp, prec_wdt(prec), st));
        ConvArgs a;
        conv_std(a, 1, &d, cout, cin, Wp);
        single_taps(a, ks);
        a.out_dt = DT_F32;
        a.p[0].in[0] = DYv; a.p[0].out = pview(DX, h, w, cin);
        AFI_TRY(run_conv(ctx, prec, a, st));
        afi_view4 none; memset(&none, 0, sizeof(none));
        AFI_TRY(nhwc_to_nchw<float>(pview(DX, h, w, cin), pview_null(), none, 0, 0, 1.f, n, cin, h, w, dxo, st));
    }
    return AFI_OK;
}
// ---- 3x3 / stride 2 / pad 1 convolution (the bottom-up path of the PANet neck, pafpn_sr.py:103-117, 186-193) -----------------------------
// out[y, x] = sum_{dy, dx in -1..1} w[dy, dx] . in[2y + dy, 2x + dx].  With the four sub-pixel phase views of the input, in_ab[i, j] =
// in[2i + a, 2j + b] (strided views, no copy), every tap reads ONE phase at an offset of -1 or 0: tap dy -> phase |dy|, offset (dy == -1 ? -1 : 0).
// So the stride-2 conv is a 9-tap stride-1 implicit GEMM over four views on the OUTPUT grid: no wasted products, the engine's TMA
// out-of-bounds fill supplies the padding (each view with its own extent: the phases of an odd size differ by one row / column).
// Backward: the input gradient of phase (a, b) is a conv of dY with the taps |dy| = a, |dx| = b (1, 2, 2 and 4 taps) stored through a
// strided view; the weight gradient is one K = pixels GEMM per phase view with that phase's taps.
struct S2Phase { int rows, cols; };
static inline int s2_off(int d) { return d == -1 ? -1 : 0; }
static size_t conv_s2_ws_bytes(int prec, int n, int cin, int h, int w, int cout) {
    size_t es = dt_size(prec_dt(prec)), wes = prec_wes(prec), P = (size_t)n * h * w, Po = (size_t)n * ((h + 1) / 2) * ((w + 1) / 2);
    return align_up(P * pad64(cin) * es) + align_up(Po * pad64(cout) * es) + 2 * align_up((size_t)9 * cin * cout * wes) +
           align_up((size_t)9 * cin * cout * 4) + align_up(P * cin * 4) + 4096 +
           (prec == AFI_PREC_SPLIT ? split_planes_bytes((long long)P, cin) + 4 * 256 + split_planes_bytes((long long)Po, cout) + 256 : 0);
}
static PView s2_phase_view(void* base, int a, int b, int h, int w, int cs, size_t es) {
    PView v; v.ptr = (char*)base + ((size_t)a * w + b) * cs * es; v.sx = 2 * cs; v.sy = (long long)2 * w * cs; v.sn = (long long)h * w * cs; return v;
}
static int conv_s2_check(afi_ctx* ctx, int prec, int cin, int cout) {
    AFI_REQUIRE(ctx && prec_ok(prec), "afi_conv3x3s2: bad argument");
    AFI_REQUIRE(prec_nk(prec), "afi_conv3x3s2: the stride-2 convolution runs on the tensor-core engines only (precision bf16 or split)");
    AFI_REQUIRE(cin % 64 == 0 && cout % 32 == 0, "afi_conv3x3s2: cin must be a multiple of 64 and cout of 32");
    return AFI_OK;
}
size_t afi_conv3x3s2_workspace_bytes(int prec, int n, int cin, int h, int w, int cout) { return conv_s2_ws_bytes(prec, n, cin, h, w, cout); }
int afi_conv3x3s2(afi_ctx* ctx, int prec, afi_view4 x, int n, int cin, int h, int w, const float* weight, const float* bias, int cout, float* y,
                  void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_TRY(conv_s2_check(ctx, prec, cin, cout));
    AFI_REQUIRE(x.ptr && weight && y && ws && n >= 1 && h >= 1 && w >= 1, "afi_conv3x3s2: bad argument");
    if (ws_bytes < conv_s2_ws_bytes(prec, n, cin, h, w, cout)) { set_error("afi_conv3x3s2: workspace too small"); return AFI_ERR_WORKSPACE; }
    const int dt = prec_dt(prec), oh = (h + 1) / 2, ow = (w + 1) / 2;
    const size_t es = dt_size(dt), wes = prec_wes(prec), P = (size_t)n * h * w, Po = (size_t)n * oh * ow;
    Carver cv(ws);
    void* X = cv.take(P * pad64(cin) * es); void* Y = cv.take(Po * pad64(cout) * es);
    void* Wp = cv.take((size_t)9 * cin * cout * wes); cv.take((size_t)9 * cin * cout * wes);
    cv.take((size_t)9 * cin * cout * 4); cv.take(P * cin * 4);
    g_ss.k0 = 0; g_ss.n[0] = split_planes_bytes((long long)P, cin) + 4 * 256 + split_planes_bytes((long long)Po, cout); g_ss.p[0] = cv.take(g_ss.n[0]);
    g_ss.pairs = 6;
    AFI_TRY(to_nhwc(prec, x, n, cin, h, w, pview(X, h, w, cin), st));
    AFI_TRY(pack_weights(weight, cout, cin, pm(prec, 0), Wp, prec_wdt(prec), st));
    ConvArgs a;
    conv_args_init(a);
    a.cin = cin; a.cout = cout; a.ntaps = 9; a.nprob = 1; a.w = Wp; a.bias = bias; a.out_dt = dt;
    int t = 0;
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++, t++) {
            a.taps[t].dy = s2_off(dy); a.taps[t].dx = s2_off(dx); a.taps[t].view = (dy != 0) * 2 + (dx != 0); a.taps[t].slab = t;
        }
    a.p[0].N = n; a.p[0].H = oh; a.p[0].W = ow;
    for (int v = 0; v < 4; v++) {
        const int pa = v >> 1, pb = v & 1;
        a.p[0].in[v] = s2_phase_view(X, pa, pb, h, w, cin, es);
        a.p[0].vh[v] = (h - pa + 1) / 2; a.p[0].vw[v] = (w - pb + 1) / 2;
        if (a.p[0].vh[v] == 0 || a.p[0].vw[v] == 0) { a.p[0].in[v] = pview(X, 1, 1, cin); a.p[0].vh[v] = a.p[0].vw[v] = 1; }   // (h or w == 1: never in range)
    }
    a.p[0].out = pview(Y, oh, ow, cout);
    if (h == 1 || w == 1) AFI_REQUIRE(false, "afi_conv3x3s2: spatial size %dx%d not supported", h, w);
    AFI_TRY(run_conv(ctx, prec, a, st));
    afi_view4 none; memset(&none, 0, sizeof(none));
    AFI_TRY(to_nchw(prec, pview(Y, oh, ow, cout), pview_null(), none, 0, 0, 1.f, n, cout, oh, ow, y, st));
    return AFI_OK;
}
int afi_conv3x3s2_backward(afi_ctx* ctx, int prec, afi_view4 x, afi_view4 dy, int n, int cin, int h, int w, const float* weight, int cout,
                           float* dw, float* db, float* dxo, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_TRY(conv_s2_check(ctx, prec, cin, cout));
    AFI_REQUIRE(x.ptr && dy.ptr && weight && dw && ws && h >= 2 && w >= 2, "afi_conv3x3s2_backward: bad argument");
    AFI_REQUIRE(cout % 64 == 0, "afi_conv3x3s2_backward: cout must be a multiple of 64");
    if (ws_bytes < conv_s2_ws_bytes(prec, n, cin, h, w, cout)) { set_error("afi_conv3x3s2_backward: workspace too small"); return AFI_ERR_WORKSPACE; }
    const int dt = prec_dt(prec), oh = (h + 1) / 2, ow = (w + 1) / 2;
    const size_t es = dt_size(dt), wes = prec_wes(prec), P = (size_t)n * h * w, Po = (size_t)n * oh * ow;
    Carver cv(ws);
    void* X = cv.take(P * pad64(cin) * es); void* DYb = cv.take(Po * pad64(cout) * es);
    void* Wp = cv.take((size_t)9 * cin * cout * wes); cv.take((size_t)9 * cin * cout * wes);
    float* acc = (float*)cv.take((size_t)9 * cin * cout * 4);
    float* DX = (float*)cv.take(P * cin * 4);
    g_ss.k0 = 0; g_ss.n[0] = split_planes_bytes((long long)P, cin) + 4 * 256 + split_planes_bytes((long long)Po, cout); g_ss.p[0] = cv.take(g_ss.n[0]);
    g_ss.pairs = split_pairs_env(3);
    PView DYv = pview(DYb, oh, ow, cout);
    AFI_TRY(to_nhwc(prec, x, n, cin, h, w, pview(X, h, w, cin), st));
    AFI_TRY(to_nhwc(prec, dy, n, cout, oh, ow, DYv, st));
    AFI_CUDA(cudaMemsetAsync(acc, 0, (size_t)9 * cin * cout * 4, st));
    for (int v = 0; v < 4; v++) {       // weight gradient: one GEMM per phase view with that phase's taps
        const int pa = v >> 1, pb = v & 1;
        WgradArgs g;
        memset(&g, 0, sizeof(g));
        g.cin = cin; g.cout = cout; g.nprob = 1; g.dw = acc;
        for (int d0 = -1; d0 <= 1; d0++)
            for (int d1 = -1; d1 <= 1; d1++)
                if ((d0 != 0) == pa && (d1 != 0) == pb) { Tap& t = g.taps[g.ntaps++]; t.dy = s2_off(d0); t.dx = s2_off(d1); t.view = 0; t.slab = (d0 + 1) * 3 + d1 + 1; }
        g.p[0].N = n; g.p[0].H = oh; g.p[0].W = ow; g.p[0].dy = DYv;
        g.p[0].x = s2_phase_view(X, pa, pb, h, w, cin, es); g.p[0].xh = (h - pa + 1) / 2; g.p[0].xw = (w - pb + 1) / 2;
        AFI_TRY(run_wgrad(ctx, prec, g, st));
    }
    AFI_TRY(unpack_wgrad(acc, cout, cin, 1, 0, dw, 1.f, 0, st));
    if (db) {
        AFI_CUDA(cudaMemsetAsync(db, 0, cout * sizeof(float), st));
        AFI_TRY(col_sum_f32(DYv, dt, n, oh, ow, cout, db, st));
    }
    if (dxo) {
        // input gradient, phase by phase: dx[2i + a, 2j + b] = sum over the taps with |dy| = a, |dx| = b of W[dy, dx]^T . dY[i + (dy == -1), j + (dx == -1)]
        AFI_TRY(pack_weights(weight, cout, cin, pm(prec, 1), Wp, prec_wdt(prec), st));      // dgrad slabs: slab 8 - t holds W[tap t]^T
        for (int v = 0; v < 4; v++) {
            const int pa = v >> 1, pb = v & 1;
            const int rows = (h - pa + 1) / 2, cols = (w - pb + 1) / 2;
            ConvArgs a;
            conv_args_init(a);
            a.cin = cout; a.cout = cin; a.nprob = 1; a.w = Wp; a.out_dt = DT_F32; a.ntaps = 0; a.nslab = 9;
            for (int d0 = -1; d0 <= 1; d0++)
                for (int d1 = -1; d1 <= 1; d1++)
                    if ((d0 != 0) == pa && (d1 != 0) == pb) {
                        Tap& t = a.taps[a.ntaps++]; t.dy = d0 == -1 ? 1 : 0; t.dx = d1 == -1 ? 1 : 0; t.view = 0; t.slab = 8 - ((d0 + 1) * 3 + d1 + 1);
                    }
            a.p[0].N = n; a.p[0].H = rows; a.p[0].W = cols;
            a.p[0].in[0] = DYv; a.p[0].vh[0] = oh; a.p[0].vw[0] = ow;
            a.p[0].out = s2_phase_view(DX, pa, pb, h, w, cin, 4);
            AFI_TRY(run_conv(ctx, prec, a, st));
        }
        afi_view4 none; memset(&none, 0, sizeof(none));
        AFI_TRY(nhwc_to_nchw<float>(pview(DX, h, w, cin), pview_null(), none, 0, 0, 1.f, n, cin, h, w, dxo, st));
    }
    return AFI_OK;
}

// ---- BiFPN neck at inference: depthwise-separable conv and bottom-up fusion site -----------------------------------------------------------
size_t afi_sepconv_workspace_bytes(int prec, int n, int c, int h, int w, int cout) { return conv_single_ws_bytes(prec, 1, n, c, h, w, cout); }
int afi_sepconv(afi_ctx* ctx, int prec, afi_view4 x, int n, int c, int h, int w, const float* dw_w, const float* pw_w, const float* pw_b, int cout,
                int pre_swish, float* y, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(ctx && x.ptr && dw_w && pw_w && y && ws && prec_ok(prec), "afi_sepconv: bad argument");
    AFI_REQUIRE(c % 32 == 0 && cout % 32 == 0 && n >= 1 && h >= 1 && w >= 1, "afi_sepconv: channels must be multiples of 32");
    if (ws_bytes < conv_single_ws_bytes(prec, 1, n, c, h, w, cout)) { set_error("afi_sepconv: workspace too small"); return AFI_ERR_WORKSPACE; }
    const int dt = prec_dt(prec); const size_t es = dt_size(dt), wes = prec_wes(prec), P = (size_t)n * h * w;
    Carver cv(ws);
    void* X = cv.take(P * pad64(c) * es); void* Y = cv.take(P * pad64(cout) * es); cv.take(P * pad64(cout) * es);
    void* Wp = cv.take((size_t)c * cout * wes); cv.take((size_t)c * cout * wes);
    cv.take((size_t)c * cout * 4); cv.take(P * c * 4);
    g_ss.k0 = 0; g_ss.n[0] = split_planes_bytes((long long)P, c) + split_planes_bytes((long long)P, cout); g_ss.p[0] = cv.take(g_ss.n[0]);
    g_ss.pairs = split_pairs_env(3);                       // forward-only
    // depthwise 3x3 (+ swish in front) straight into the pointwise GEMM's operand layout
    if (dt == DT_F32) AFI_TRY(dw3x3_to_nhwc<float>(x, dw_w, n, c, h, w, pre_swish, pview(X, h, w, c), st));
    else AFI_TRY(dw3x3_to_nhwc<bf16>(x, dw_w, n, c, h, w, pre_swish, pview(X, h, w, c), st));
    AFI_TRY(pack_weights(pw_w, cout, c, pm(prec, 4), Wp, prec_wdt(prec), st));
    ConvArgs a;
    Dim3 d = {n, h, w};
    conv_std(a, 1, &d, c, cout, Wp);
    single_taps(a, 1);
    a.bias = pw_b; a.out_dt = dt;
    a.p[0].in[0] = pview(X, h, w, c); a.p[0].out = pview(Y, h, w, cout);
    AFI_TRY(run_conv(ctx, prec, a, st));
    afi_view4 none; memset(&none, 0, sizeof(none));
    AFI_TRY(to_nchw(prec, pview(Y, h, w, cout), pview_null(), none, 0, 0, 1.f, n, cout, h, w, y, st));
    return AFI_OK;
}
int afi_bifpn_fuse_down(afi_view4 a, afi_view4 b, afi_view4 down, const float* weights, int n, int c, int h, int w, int dh, int dw, float* out,
                        void* stream) {
    AFI_REQUIRE(a.ptr && down.ptr && out && n >= 1 && c >= 1 && h >= 1 && w >= 1, "afi_bifpn_fuse_down: bad argument");
    AFI_REQUIRE(dh >= 2 && dw >= 2 && (dh - 2) / 2 + 1 == h && (dw - 2) / 2 + 1 == w,
                "afi_bifpn_fuse_down: a %dx%d map is not the zero-padded 3x3 / stride-2 max-pool of a %dx%d one", h, w, dh, dw);
    return bifpn_fuse_down(a, b, down, weights, b.ptr ? 3 : 2, n, c, h, w, dh, dw, out, (cudaStream_t)stream);
}

size_t afi_conv3x3_workspace_bytes(int prec, int n, int cin, int h, int w, int cout) { return conv_single_ws_bytes(prec, 3, n, cin, h, w, cout); }
int afi_conv3x3(afi_ctx* ctx, int prec, afi_view4 x, int n, int cin, int h, int w, const float* weight, const float* bias, int cout,
                int lrelu, float* y, void* ws, size_t ws_bytes, void* stream) {
    return conv_single(ctx, prec, 3, x, n, cin, h, w, weight, bias, cout, lrelu, y, ws, ws_bytes, (cudaStream_t)stream);
}
int afi_conv3x3_backward(afi_ctx* ctx, int prec, afi_view4 x, afi_view4 dy, int n, int cin, int h, int w, const float* weight, int cout,
                         float* dw, float* db, float* dxo, void* ws, size_t ws_bytes, void* stream) {
    return conv_single_backward(ctx, prec, 3, x, dy, n, cin, h, w, weight, cout, dw, db, dxo, ws, ws_bytes, (cudaStream_t)stream);
}
size_t afi_conv1x1_workspace_bytes(int prec, int n, int cin, int h, int w, int cout) { return conv_single_ws_bytes(prec, 1, n, cin, h, w, cout); }
int afi_conv1x1(afi_ctx* ctx, int prec, afi_view4 x, int n, int cin, int h, int w, const float* weight, const float* bias, int cout,
                float* y, void* ws, size_t ws_bytes, void* stream) {
    return conv_single(ctx, prec, 1, x, n, cin, h, w, weight, bias, cout, 0, y, ws, ws_bytes, (cudaStream_t)stream);
}
int afi_conv1x1_backward(afi_ctx* ctx, int prec, afi_view4 x, afi_view4 dy, int n, int cin, int h, int w, const float* weight, int cout,
                         float* dw, float* db, float* dxo, void* ws, size_t ws_bytes, void* stream) {
    return conv_single_backward(ctx, prec, 1, x, dy, n, cin, h, w, weight, cout, dw, db, dxo, ws, ws_bytes, (cudaStream_t)stream);
}

}  // extern "C"
