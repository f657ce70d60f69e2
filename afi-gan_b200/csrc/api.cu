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

static inline bool prec_ok(int prec) { return prec == AFI_PREC_FP32 || prec == AFI_PREC_BF16 || prec == AFI_PREC_BF16_SIMT; }
static inline int prec_dt(int prec) { return prec == AFI_PREC_FP32 ? DT_F32 : DT_BF16; }
static inline size_t dt_size(int dt) { return dt == DT_F32 ? 4 : 2; }
static inline bool prec_tc(int prec) { return prec == AFI_PREC_BF16; }

static int run_conv(afi_ctx* ctx, int prec, const ConvArgs& a, cudaStream_t st) {
    if (prec == AFI_PREC_FP32) return conv_simt<float>(a, st);
    if (prec == AFI_PREC_BF16_SIMT) return conv_simt<bf16>(a, st);
    return conv_tc(ctx, a, st);
}
static int run_wgrad(afi_ctx* ctx, int prec, const WgradArgs& a, cudaStream_t st) {
    if (prec == AFI_PREC_FP32) return wgrad_simt<float>(a, st);
    if (prec == AFI_PREC_BF16_SIMT) return wgrad_simt<bf16>(a, st);
    return wgrad_tc(ctx, a, st);
}
static int to_nhwc(int prec, afi_view4 src, int n, int c, int h, int w, PView dst, cudaStream_t st) {
    return prec == AFI_PREC_FP32 ? nchw_to_nhwc<float>(src, n, c, h, w, dst, st) : nchw_to_nhwc<bf16>(src, n, c, h, w, dst, st);
}
static int to_nchw(int prec, PView a, PView lat, afi_view4 skip, int sh, int sw, float scale, int n, int c, int oh, int ow, float* dst,
                   cudaStream_t st) {
    return prec == AFI_PREC_FP32 ? nhwc_to_nchw<float>(a, lat, skip, sh, sw, scale, n, c, oh, ow, dst, st)
                                 : nhwc_to_nchw<bf16>(a, lat, skip, sh, sw, scale, n, c, oh, ow, dst, st);
}
// forward / dgrad pack modes per engine
static inline int pm(int prec, int kind) { return kind * 2 + (prec_tc(prec) ? 1 : 0); }

static int wgrad_std(afi_ctx* ctx, int prec, PView x, int cin, PView dy, int cout, int n, int h, int w, float* dw, cudaStream_t st) {
    WgradArgs g;
    memset(&g, 0, sizeof(g));
    g.N = n; g.H = h; g.W = w; g.cin = cin; g.cout = cout; g.ntaps = 9;
    set_std_taps(g.taps, 0, 0);
    g.x = x; g.dy = dy; g.dw = dw;
    return run_wgrad(ctx, prec, g, st);
}

// =====================================================================================================
// Generator
// =====================================================================================================
static const int C = AFI_CH, GR = 32, CB = AFI_CH + 4 * 32;   // 256, growth 32, dense buffer width 384

struct GPacked {   // element offsets into the packed weight buffer (dtype T)
    size_t head_f, head_d, rdb_f[AFI_MAX_RDB][5], rdb_d[AFI_MAX_RDB][5], post_f, post_d, up_f, up_d, out_f, out_d, total;
};
static GPacked g_packed_layout(int n_rdb) {
    GPacked L; size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 127) / 128 * 128; return r; };
    L.head_f = take(9 * C * C); L.head_d = take(9 * C * C);
    for (int r = 0; r < n_rdb; r++)
        for (int i = 0; i < 5; i++) {
            size_t n = (size_t)9 * (C + GR * i) * (i < 4 ? GR : C);
            L.rdb_f[r][i] = take(n); L.rdb_d[r][i] = take(n);
        }
    L.post_f = take(9 * C * C); L.post_d = take(9 * C * C);
    L.up_f = take(36 * C * C); L.up_d = take(36 * C * C);
    L.out_f = take(9 * C * C); L.out_d = take(9 * C * C);
    L.total = o;
    return L;
}
struct GGradAcc {  // float offsets
    size_t head_w, head_b, rdb_w[AFI_MAX_RDB][5], post_w, post_b, up_w, up_b, out_w, out_b, total;
};
static GGradAcc g_gradacc_layout(int n_rdb) {
    GGradAcc L; size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 63) / 64 * 64; return r; };
    L.head_w = take(9 * C * C); L.head_b = take(C);
    for (int r = 0; r < n_rdb; r++)
        for (int i = 0; i < 5; i++) L.rdb_w[r][i] = take((size_t)9 * (C + GR * i) * (i < 4 ? GR : C));
    L.post_w = take(9 * C * C); L.post_b = take(C);
    L.up_w = take(36 * C * C); L.up_b = take(C);
    L.out_w = take(9 * C * C); L.out_b = take(C);
    L.total = o;
    return L;
}
struct GWs {
    void *X0, *B[AFI_MAX_RDB], *H1, *H2, *H3, *Yb, *LX, *LAT;           // forward (X0..H3 saved for backward)
    void *G0, *G1, *G2, *dH1, *GA[2], *DC5, *GC, *GH, *DXb, *LW;        // backward scratch
    size_t total;
};
static GWs g_ws_layout(void* base, int prec, int n, int h, int w, int n_rdb, int lat_c, int backward) {
    GWs W; memset(&W, 0, sizeof(W));
    Carver cv(base);
    size_t es = dt_size(prec_dt(prec));
    size_t P = (size_t)n * h * w, P4 = 4 * P;
    W.X0 = cv.take(P * C * es);
    for (int r = 0; r < n_rdb; r++) W.B[r] = cv.take(P * CB * es);
    W.H1 = cv.take(P * C * es); W.H2 = cv.take(P * C * es); W.H3 = cv.take(P4 * C * es);
    W.Yb = cv.take(P4 * C * es);
    if (lat_c > 0) { W.LX = cv.take(P4 * lat_c * es); W.LAT = cv.take(P4 * C * es); W.LW = cv.take((size_t)C * lat_c * es); }
    if (backward) {
        W.G0 = cv.take(P4 * C * es); W.G1 = cv.take(P4 * C * es); W.G2 = cv.take(P * C * es);
        W.dH1 = cv.take(P * C * 4); W.GA[0] = cv.take(P * CB * 4); W.GA[1] = cv.take(P * CB * 4);
        W.DC5 = cv.take(P * C * es); W.GC = cv.take(P * GR * es); W.GH = cv.take(P * C * es);
        W.DXb = cv.take(P * C * 4);
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
    c->device = dev; c->sm_count = prop.multiProcessorCount; c->encode_tiled = nullptr; c->tile_counter = nullptr;
    int r = tc_init(c);
    if (r != AFI_OK) { delete c; return r; }
    *out = c;
    return AFI_OK;
}
void afi_destroy(afi_ctx* ctx) { delete ctx; }

size_t afi_g_packed_bytes(int prec, int n_rdb) { return g_packed_layout(n_rdb).total * dt_size(prec_dt(prec)); }
size_t afi_g_gradacc_bytes(int n_rdb) { return g_gradacc_layout(n_rdb).total * 4; }
size_t afi_g_workspace_bytes(int prec, int n, int h, int w, int n_rdb, int lat_c, int save_for_backward) {
    return g_ws_layout(nullptr, prec, n, h, w, n_rdb, lat_c, save_for_backward).total;
}

int afi_g_pack(afi_ctx* ctx, int prec, const afi_g_params* p, void* packed, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(ctx && p && packed, "afi_g_pack: null argument");
    AFI_TRY(g_check(prec, 1, 1, 1, p->n_rdb));
    GPacked L = g_packed_layout(p->n_rdb);
    int dt = prec_dt(prec); size_t es = dt_size(dt);
    char* b = (char*)packed;
    AFI_TRY(pack_weights(p->head_w, C, C, pm(prec, 0), b + L.head_f * es, dt, st));
    AFI_TRY(pack_weights(p->head_w, C, C, pm(prec, 1), b + L.head_d * es, dt, st));
    for (int r = 0; r < p->n_rdb; r++)
        for (int i = 0; i < 5; i++) {
            int co = i < 4 ? GR : C, ci = C + GR * i;
            AFI_TRY(pack_weights(p->rdb_w[r][i], co, ci, pm(prec, 0), b + L.rdb_f[r][i] * es, dt, st));
            AFI_TRY(pack_weights(p->rdb_w[r][i], co, ci, pm(prec, 1), b + L.rdb_d[r][i] * es, dt, st));
        }
    AFI_TRY(pack_weights(p->post_w, C, C, pm(prec, 0), b + L.post_f * es, dt, st));
    AFI_TRY(pack_weights(p->post_w, C, C, pm(prec, 1), b + L.post_d * es, dt, st));
    AFI_TRY(pack_weights(p->up_w, C, C, pm(prec, 2), b + L.up_f * es, dt, st));
    AFI_TRY(pack_weights(p->up_w, C, C, pm(prec, 3), b + L.up_d * es, dt, st));
    AFI_TRY(pack_weights(p->out_w, C, C, pm(prec, 0), b + L.out_f * es, dt, st));
    AFI_TRY(pack_weights(p->out_w, C, C, pm(prec, 1), b + L.out_d * es, dt, st));
    return AFI_OK;
}

int afi_g_forward(afi_ctx* ctx, int prec, const afi_g_params* p, const void* packed, afi_view4 x, int n, int h, int w, float* y,
                  int oh, int ow, const afi_lateral* lat, void* ws, size_t ws_bytes, int save, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(ctx && p && packed && x.ptr && y && ws, "afi_g_forward: null argument");
    AFI_TRY(g_check(prec, n, h, w, p->n_rdb));
    AFI_REQUIRE(oh >= 1 && ow >= 1 && oh <= 2 * h && ow <= 2 * w, "afi_g_forward: output crop %dx%d exceeds %dx%d", oh, ow, 2 * h, 2 * w);
    const int nr = p->n_rdb, dt = prec_dt(prec);
    const int es = (int)dt_size(dt);
    const int lat_c = lat ? lat->lat_c : 0;
    if (lat) AFI_REQUIRE(lat_c % 32 == 0 && lat_c <= 2048 && lat->lat_x.ptr && lat->lat_w, "afi_g_forward: bad lateral");
    GWs W = g_ws_layout(ws, prec, n, h, w, nr, lat_c, save);
    if (W.total > ws_bytes) { set_error("afi_g_forward: workspace %zu B < required %zu B", ws_bytes, W.total); return AFI_ERR_WORKSPACE; }
    if (n == 0) return AFI_OK;
    GPacked L = g_packed_layout(nr);
    const char* pk = (const char*)packed;
    const int H2x = 2 * h, W2x = 2 * w;

    AFI_TRY(to_nhwc(prec, x, n, C, h, w, pview(W.X0, h, w, C), st));
    ConvArgs a;
    // [0] head conv + bias + LeakyReLU -> B0[:, 0:256]                                   generator_rdb.py:91-93
    conv_args_init(a);
    a.N = n; a.H = h; a.W = w; a.cin = C; a.cout = C; a.ntaps = 9; set_std_taps(a.taps, 0, 0);
    a.in[0] = pview(W.X0, h, w, C); a.w = pk + L.head_f * es; a.bias = p->head_b; a.act = 1;
    a.out = pview(W.B[0], h, w, CB); a.out_dt = dt;
    AFI_TRY(run_conv(ctx, prec, a, st));
    // [1] residual-in-residual: dense blocks write their growth channels into slices of one 384-ch buffer   :39-71
    for (int r = 0; r < nr; r++) {
        PView Br = pview(W.B[r], h, w, CB);
        for (int i = 0; i < 4; i++) {
            conv_args_init(a);
            a.N = n; a.H = h; a.W = w; a.cin = C + GR * i; a.cout = GR; a.ntaps = 9; set_std_taps(a.taps, 0, 0);
            a.in[0] = Br; a.w = pk + L.rdb_f[r][i] * es; a.act = 1;
            a.out = pview_ch(Br, C + GR * i, es); a.out_dt = dt;
            AFI_TRY(run_conv(ctx, prec, a, st));
        }
        conv_args_init(a);
        a.N = n; a.H = h; a.W = w; a.cin = CB; a.cout = C; a.ntaps = 9; set_std_taps(a.taps, 0, 0);
        a.in[0] = Br; a.w = pk + L.rdb_f[r][4] * es;
        if (r + 1 < nr) {          // x_{r+1} = x_r + 0.2 * conv5
            a.alpha = 0.2f; a.r1 = Br; a.r1_dt = dt; a.beta1 = 1.f;
            a.out = pview(W.B[r + 1], h, w, CB);
        } else {                   // h1 = 0.2 * (x + 0.2 conv5) + h0                         :27-30
            a.alpha = 0.04f; a.r1 = Br; a.r1_dt = dt; a.beta1 = 0.2f;
            a.r2 = pview(W.B[0], h, w, CB); a.r2_dt = dt; a.beta2 = 1.f;
            a.out = pview(W.H1, h, w, C);
        }
        a.out_dt = dt;
        AFI_TRY(run_conv(ctx, prec, a, st));
    }
    // [2] post conv + bias + LeakyReLU                                                   :97-99
    conv_args_init(a);
    a.N = n; a.H = h; a.W = w; a.cin = C; a.cout = C; a.ntaps = 9; set_std_taps(a.taps, 0, 0);
    a.in[0] = pview(W.H1, h, w, C); a.w = pk + L.post_f * es; a.bias = p->post_b; a.act = 1;
    a.out = pview(W.H2, h, w, C); a.out_dt = dt;
    AFI_TRY(run_conv(ctx, prec, a, st));
    // [3] ConvTranspose2d k6 s2 p2 == four 3x3 sub-pixel convs with interleaved stores      :101-105, App. G
    for (int ph = 0; ph < 4; ph++) {
        int pa = ph >> 1, pb = ph & 1;
        conv_args_init(a);
        a.N = n; a.H = h; a.W = w; a.cin = C; a.cout = C; a.ntaps = 9; set_std_taps(a.taps, 0, 9 * ph);
        a.in[0] = pview(W.H2, h, w, C); a.w = pk + L.up_f * es; a.bias = p->up_b; a.act = 1;
        PView o; o.ptr = (char*)W.H3 + ((size_t)pa * W2x + pb) * C * es;
        o.sx = 2 * C; o.sy = (long long)2 * W2x * C; o.sn = (long long)H2x * W2x * C;
        a.out = o; a.out_dt = dt;
        AFI_TRY(run_conv(ctx, prec, a, st));
    }
    // [4] output conv + bias on the 2h x 2w grid                                           :107-108
    conv_args_init(a);
    a.N = n; a.H = H2x; a.W = W2x; a.cin = C; a.cout = C; a.ntaps = 9; set_std_taps(a.taps, 0, 0);
    a.in[0] = pview(W.H3, H2x, W2x, C); a.w = pk + L.out_f * es; a.bias = p->out_b;
    a.out = pview(W.Yb, H2x, W2x, C); a.out_dt = dt;
    AFI_TRY(run_conv(ctx, prec, a, st));
    // optional lateral 1x1 conv of the FPN merge (fpn_sr.py:152)
    PView latv = pview_null();
    float scale = 1.f;
    if (lat) {
        AFI_TRY(to_nhwc(prec, lat->lat_x, n, lat_c, oh, ow, pview(W.LX, oh, ow, lat_c), st));
        // 1x1 weight [256, lat_c, 1, 1] is already [co][ci]; KN engines want [ci][co]
        AFI_TRY(pack_weights(lat->lat_w, C, lat_c, pm(prec, 4), W.LW, dt, st));
        conv_args_init(a);
        a.N = n; a.H = oh; a.W = ow; a.cin = lat_c; a.cout = C; a.ntaps = 1;
        a.taps[0].dy = 0; a.taps[0].dx = 0; a.taps[0].view = 0; a.taps[0].slab = 0;
        a.in[0] = pview(W.LX, oh, ow, lat_c); a.w = W.LW; a.bias = lat->lat_b;
        a.out = pview(W.LAT, oh, ow, C); a.out_dt = dt;
        AFI_TRY(run_conv(ctx, prec, a, st));
        latv = pview(W.LAT, oh, ow, C);
        scale = lat->scale;
    }
    // y = (branch + bilinear2x(x) [+ lateral]) * scale, cropped to oh x ow                  :125,130; stage1_trainer.py:437-443
    AFI_TRY(to_nchw(prec, pview(W.Yb, H2x, W2x, C), latv, x, h, w, scale, n, C, oh, ow, y, st));
    return AFI_OK;
}

int afi_g_backward(afi_ctx* ctx, int prec, const afi_g_params* p, const void* packed, afi_view4 dy, int n, int h, int w, int oh,
                   int ow, void* ws, size_t ws_bytes, float* gradacc, float* dx, const afi_lateral* lat, float* lat_dx, float* lat_gw,
                   float* lat_gb, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(ctx && p && packed && dy.ptr && ws && gradacc, "afi_g_backward: null argument");
    AFI_TRY(g_check(prec, n, h, w, p->n_rdb));
    AFI_REQUIRE(!lat, "afi_g_backward: lateral backward is not implemented yet (use torch autograd around the lateral conv)");
    (void)lat_dx; (void)lat_gw; (void)lat_gb;
    const int nr = p->n_rdb, dt = prec_dt(prec);
    const int es = (int)dt_size(dt);
    GWs W = g_ws_layout(ws, prec, n, h, w, nr, 0, 1);
    if (W.total > ws_bytes) { set_error("afi_g_backward: workspace %zu B < required %zu B", ws_bytes, W.total); return AFI_ERR_WORKSPACE; }
    if (n == 0) return AFI_OK;
    GPacked L = g_packed_layout(nr);
    GGradAcc GL = g_gradacc_layout(nr);
    const char* pk = (const char*)packed;
    const int H2x = 2 * h, W2x = 2 * w;
    const size_t P = (size_t)n * h * w;
    PView X0 = pview(W.X0, h, w, C), H1 = pview(W.H1, h, w, C), H2 = pview(W.H2, h, w, C), H3 = pview(W.H3, H2x, W2x, C);
    PView G0 = pview(W.G0, H2x, W2x, C), G1 = pview(W.G1, H2x, W2x, C), G2 = pview(W.G2, h, w, C);
    PView dH1 = pview(W.dH1, h, w, C), DC5 = pview(W.DC5, h, w, C), GC = pview(W.GC, h, w, GR), GH = pview(W.GH, h, w, C);
    ConvArgs a;

    // dL/d(branch) on the full 2h x 2w grid: the crop's complement gets zero gradient
    AFI_CUDA(cudaMemsetAsync(W.G0, 0, 4 * P * C * es, st));
    AFI_TRY(to_nhwc(prec, dy, n, C, oh, ow, G0, st));
    // output conv
    AFI_TRY(wgrad_std(ctx, prec, H3, C, G0, C, n, H2x, W2x, gradacc + GL.out_w, st));
    AFI_TRY(col_sum_f32(G0, dt, n, H2x, W2x, C, gradacc + GL.out_b, st));
    conv_args_init(a);
    a.N = n; a.H = H2x; a.W = W2x; a.cin = C; a.cout = C; a.ntaps = 9; set_std_taps(a.taps, 0, 0);
    a.in[0] = G0; a.w = pk + L.out_d * es; a.mask = H3; a.out = G1; a.out_dt = dt;
    AFI_TRY(run_conv(ctx, prec, a, st));
    // transposed conv: wgrad per phase, bias, and dgrad as ONE 36-tap conv over the four phase views of G1
    PView ph[4];
    for (int i = 0; i < 4; i++) {
        int pa = i >> 1, pb = i & 1;
        ph[i].ptr = (char*)W.G1 + ((size_t)pa * W2x + pb) * C * es;
        ph[i].sx = 2 * C; ph[i].sy = (long long)2 * W2x * C; ph[i].sn = (long long)H2x * W2x * C;
        AFI_TRY(wgrad_std(ctx, prec, H2, C, ph[i], C, n, h, w, gradacc + GL.up_w + (size_t)i * 9 * C * C, st));
    }
    AFI_TRY(col_sum_f32(G1, dt, n, H2x, W2x, C, gradacc + GL.up_b, st));
    conv_args_init(a);
    a.N = n; a.H = h; a.W = w; a.cin = C; a.cout = C; a.ntaps = 36;
    for (int i = 0; i < 4; i++) { set_std_taps(a.taps + 9 * i, i, 9 * i); a.in[i] = ph[i]; }
    a.w = pk + L.up_d * es; a.mask = H2; a.out = G2; a.out_dt = dt;
    AFI_TRY(run_conv(ctx, prec, a, st));
    // post conv
    AFI_TRY(wgrad_std(ctx, prec, H1, C, G2, C, n, h, w, gradacc + GL.post_w, st));
    AFI_TRY(col_sum_f32(G2, dt, n, h, w, C, gradacc + GL.post_b, st));
    conv_args_init(a);
    a.N = n; a.H = h; a.W = w; a.cin = C; a.cout = C; a.ntaps = 9; set_std_taps(a.taps, 0, 0);
    a.in[0] = G2; a.w = pk + L.post_d * es; a.out = dH1; a.out_dt = DT_F32;
    AFI_TRY(run_conv(ctx, prec, a, st));
    // residual-in-residual: h1 = 0.2 * x_nr + h0.  d_out (fp32 view) = 0.2 * dH1 entering the last dense block.
    PView d_out = dH1; float d_scale = 0.2f; int cur = 0;
    for (int r = nr - 1; r >= 0; r--) {
        PView Br = pview(W.B[r], h, w, CB);
        PView GA = pview(W.GA[cur], h, w, CB);
        // dc5 = 0.2 * d_out (GEMM operand in storage dtype)
        AFI_TRY(ew_combine(DC5, dt, d_out, DT_F32, pview_null(), 0, pview_null(), 0, 0.2f, 0.2f * d_scale, n, h, w, C, st));
        AFI_TRY(wgrad_std(ctx, prec, Br, CB, DC5, C, n, h, w, gradacc + GL.rdb_w[r][4], st));
        // GA[:, 0:256] = d_out, GA[:, 256:384] = 0, then GA += dgrad(conv5)(dc5) over all 384 channels
        AFI_CUDA(cudaMemsetAsync(W.GA[cur], 0, P * CB * 4, st));
        AFI_TRY(ew_combine(GA, DT_F32, d_out, DT_F32, pview_null(), 0, pview_null(), 0, 0.2f, d_scale, n, h, w, C, st));
        conv_args_init(a);
        a.N = n; a.H = h; a.W = w; a.cin = C; a.cout = CB; a.ntaps = 9; set_std_taps(a.taps, 0, 0);
        a.in[0] = DC5; a.w = pk + L.rdb_d[r][4] * es; a.accin = GA; a.out = GA; a.out_dt = DT_F32;
        AFI_TRY(run_conv(ctx, prec, a, st));
        for (int i = 3; i >= 0; i--) {
            int cin_f = C + GR * i;
            // g = GA[:, slice_i] * lrelu'(c_{i+1})
            AFI_TRY(ew_combine(GC, dt, pview_ch(GA, cin_f, 4), DT_F32, pview_null(), 0, pview_ch(Br, cin_f, es), dt, 0.2f, 1.f, n, h, w, GR, st));
            AFI_TRY(wgrad_std(ctx, prec, Br, cin_f, GC, GR, n, h, w, gradacc + GL.rdb_w[r][i], st));
            conv_args_init(a);
            a.N = n; a.H = h; a.W = w; a.cin = GR; a.cout = cin_f; a.ntaps = 9; set_std_taps(a.taps, 0, 0);
            a.in[0] = GC; a.w = pk + L.rdb_d[r][i] * es; a.accin = GA; a.out = GA; a.out_dt = DT_F32;
            AFI_TRY(run_conv(ctx, prec, a, st));
        }
        d_out = GA; d_scale = 1.f; cur ^= 1;
    }
    // head conv: g_head = (d_out + dH1) * lrelu'(h0)
    AFI_TRY(ew_combine(GH, dt, d_out, DT_F32, dH1, DT_F32, pview(W.B[0], h, w, CB), dt, 0.2f, 1.f, n, h, w, C, st));
    AFI_TRY(wgrad_std(ctx, prec, X0, C, GH, C, n, h, w, gradacc + GL.head_w, st));
    AFI_TRY(col_sum_f32(GH, dt, n, h, w, C, gradacc + GL.head_b, st));
    if (dx) {
        // dx = dgrad(head)(g_head) + bilinear2x^T(dy): conv into an NHWC fp32 buffer, then to NCHW (+ skip adjoint, TODO)
        set_error("afi_g_backward: input gradient is not implemented yet");
        return AFI_ERR_INVALID;
    }
    return AFI_OK;
}

int afi_g_unpack_grads(afi_ctx* ctx, int prec, const float* gradacc, const afi_g_grads* g, float scale, int accumulate, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(ctx && gradacc && g, "afi_g_unpack_grads: null argument");
    AFI_TRY(g_check(prec, 1, 1, 1, g->n_rdb));
    GGradAcc GL = g_gradacc_layout(g->n_rdb);
    int nk = prec_tc(prec) ? 1 : 0;
    if (g->head_w) AFI_TRY(unpack_wgrad(gradacc + GL.head_w, C, C, nk, 0, g->head_w, scale, accumulate, st));
    if (g->head_b) AFI_TRY(axpby_f32(gradacc + GL.head_b, g->head_b, C, scale, accumulate, st));
    for (int r = 0; r < g->n_rdb; r++)
        for (int i = 0; i < 5; i++)
            if (g->rdb_w[r][i]) AFI_TRY(unpack_wgrad(gradacc + GL.rdb_w[r][i], i < 4 ? GR : C, C + GR * i, nk, 0, g->rdb_w[r][i], scale, accumulate, st));
    if (g->post_w) AFI_TRY(unpack_wgrad(gradacc + GL.post_w, C, C, nk, 0, g->post_w, scale, accumulate, st));
    if (g->post_b) AFI_TRY(axpby_f32(gradacc + GL.post_b, g->post_b, C, scale, accumulate, st));
    if (g->up_w) AFI_TRY(unpack_wgrad(gradacc + GL.up_w, C, C, nk, 1, g->up_w, scale, accumulate, st));
    if (g->up_b) AFI_TRY(axpby_f32(gradacc + GL.up_b, g->up_b, C, scale, accumulate, st));
    if (g->out_w) AFI_TRY(unpack_wgrad(gradacc + GL.out_w, C, C, nk, 0, g->out_w, scale, accumulate, st));
    if (g->out_b) AFI_TRY(axpby_f32(gradacc + GL.out_b, g->out_b, C, scale, accumulate, st));
    return AFI_OK;
}

}  // extern "C"

// =====================================================================================================
// Discriminator
// =====================================================================================================
namespace afi {
static const int DC[4] = {256, 512, 1024, 1024};

struct DPacked { size_t f[3], d[3], total; };
static DPacked d_packed_layout() {
    DPacked L; size_t o = 0;
    for (int i = 0; i < 3; i++) { L.f[i] = o; o += (size_t)9 * DC[i] * DC[i + 1]; L.d[i] = o; o += (size_t)9 * DC[i] * DC[i + 1]; }
    L.total = o;
    return L;
}
struct DGradAcc { size_t w[4], b[4], gamma[3], beta[3], total; };
static DGradAcc d_gradacc_layout() {
    DGradAcc L; size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 63) / 64 * 64; return r; };
    for (int i = 0; i < 3; i++) { L.w[i] = take((size_t)9 * DC[i] * DC[i + 1]); L.b[i] = take(DC[i + 1]); L.gamma[i] = take(DC[i + 1]); L.beta[i] = take(DC[i + 1]); }
    L.w[3] = take(9 * DC[3]); L.b[3] = take(1);
    L.total = o;
    return L;
}
struct DWs {
    void *A[4], *Z[3];
    float *mean[3], *rstd[3], *T9;
    double* sums;          // [2][1024]
    void *DY[3], *DXb;
    size_t total;
};
static DWs d_ws_layout(void* base, int prec, int n, int h, int w, int backward) {
    DWs W; memset(&W, 0, sizeof(W));
    Carver cv(base);
    size_t es = dt_size(prec_dt(prec)), P = (size_t)n * h * w;
    W.A[0] = cv.take(P * DC[0] * es);
    for (int i = 0; i < 3; i++) { W.Z[i] = cv.take(P * DC[i + 1] * es); W.A[i + 1] = cv.take(P * DC[i + 1] * es); }
    for (int i = 0; i < 3; i++) { W.mean[i] = (float*)cv.take(1024 * 4); W.rstd[i] = (float*)cv.take(1024 * 4); }
    W.T9 = (float*)cv.take(P * 9 * 4);
    W.sums = (double*)cv.take(2 * 1024 * 8);
    if (backward) {
        for (int i = 0; i < 3; i++) W.DY[i] = cv.take(P * DC[i + 1] * es);
        W.DXb = cv.take(P * DC[0] * 4);
    }
    W.total = cv.off;
    return W;
}
}  // namespace afi

extern "C" {

size_t afi_d_packed_bytes(int prec) { return d_packed_layout().total * dt_size(prec_dt(prec)); }
size_t afi_d_gradacc_bytes(void) { return d_gradacc_layout().total * 4; }
size_t afi_d_workspace_bytes(int prec, int n, int h, int w, int save_for_backward) { return d_ws_layout(nullptr, prec, n, h, w, save_for_backward).total; }

int afi_d_pack(afi_ctx* ctx, int prec, const afi_d_params* p, void* packed, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(ctx && p && packed && prec_ok(prec), "afi_d_pack: bad argument");
    DPacked L = d_packed_layout();
    int dt = prec_dt(prec); size_t es = dt_size(dt);
    for (int i = 0; i < 3; i++) {
        AFI_TRY(pack_weights(p->w[i], DC[i + 1], DC[i], pm(prec, 0), (char*)packed + L.f[i] * es, dt, st));
        AFI_TRY(pack_weights(p->w[i], DC[i + 1], DC[i], pm(prec, 1), (char*)packed + L.d[i] * es, dt, st));
    }
    return AFI_OK;
}

int afi_d_forward(afi_ctx* ctx, int prec, const afi_d_params* p, const void* packed, afi_view4 x, int n, int h, int w, float* logits,
                  int training, float momentum, float eps, void* ws, size_t ws_bytes, int save, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(ctx && p && packed && x.ptr && logits && ws && prec_ok(prec), "afi_d_forward: bad argument");
    AFI_REQUIRE(n >= 1 && h >= 1 && w >= 1, "afi_d_forward: bad shape n=%d h=%d w=%d", n, h, w);
    const int dt = prec_dt(prec); const size_t es = dt_size(dt);
    DWs W = d_ws_layout(ws, prec, n, h, w, save);
    if (W.total > ws_bytes) { set_error("afi_d_forward: workspace %zu B < required %zu B", ws_bytes, W.total); return AFI_ERR_WORKSPACE; }
    DPacked L = d_packed_layout();
    AFI_TRY(to_nhwc(prec, x, n, DC[0], h, w, pview(W.A[0], h, w, DC[0]), st));
    for (int i = 0; i < 3; i++) {
        // Conv2d 3x3 + bias -> BatchNorm (batch statistics of THIS call) -> LeakyReLU(0.2)   feature_patch_discriminator.py:36-38
        ConvArgs a;
        conv_args_init(a);
        a.N = n; a.H = h; a.W = w; a.cin = DC[i]; a.cout = DC[i + 1]; a.ntaps = 9; set_std_taps(a.taps, 0, 0);
        a.in[0] = pview(W.A[i], h, w, DC[i]); a.w = (const char*)packed + L.f[i] * es; a.bias = p->b[i];
        PView Z = pview(W.Z[i], h, w, DC[i + 1]);
        a.out = Z; a.out_dt = dt;
        AFI_TRY(run_conv(ctx, prec, a, st));
        if (training) {
            AFI_CUDA(cudaMemsetAsync(W.sums, 0, 2 * 1024 * sizeof(double), st));
            AFI_TRY(col_stats(Z, dt, n, h, w, DC[i + 1], W.sums, W.sums + 1024, st));
        }
        AFI_TRY(bn_finalize(W.sums, W.sums + 1024, (long long)n * h * w, DC[i + 1], eps, momentum, training, W.mean[i], W.rstd[i],
                            p->running_mean[i], p->running_var[i], p->num_batches_tracked[i], st));
        AFI_TRY(bn_apply_lrelu(Z, pview(W.A[i + 1], h, w, DC[i + 1]), dt, W.mean[i], W.rstd[i], p->gamma[i], p->beta[i], 0.2f, n, h, w,
                               DC[i + 1], st));
    }
    // Conv2d 1024 -> 1                                                                   feature_patch_discriminator.py:40-41
    AFI_TRY(dhead_forward(pview(W.A[3], h, w, DC[3]), dt, p->w[3], p->b[3], n, h, w, DC[3], W.T9, logits, st));
    return AFI_OK;
}

int afi_d_backward(afi_ctx* ctx, int prec, const afi_d_params* p, const void* packed, const float* dlogits, int n, int h, int w,
                   void* ws, size_t ws_bytes, float* gradacc, float* dx, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(ctx && p && packed && dlogits && ws && gradacc && prec_ok(prec), "afi_d_backward: bad argument");
    AFI_REQUIRE(dx == nullptr, "afi_d_backward: input gradient is not implemented (stage 1/2 detach the discriminator input)");
    const int dt = prec_dt(prec); const size_t es = dt_size(dt);
    DWs W = d_ws_layout(ws, prec, n, h, w, 1);
    if (W.total > ws_bytes) { set_error("afi_d_backward: workspace %zu B < required %zu B", ws_bytes, W.total); return AFI_ERR_WORKSPACE; }
    DPacked L = d_packed_layout();
    DGradAcc GL = d_gradacc_layout();
    // head: dW4, db4 and dy3 = dA3 * lrelu'(a3) in one pass over a3
    PView DY = pview(W.DY[2], h, w, DC[3]);
    AFI_TRY(dhead_backward(pview(W.A[3], h, w, DC[3]), dt, p->w[3], dlogits, n, h, w, DC[3], gradacc + GL.w[3], gradacc + GL.b[3], DY, st));
    for (int i = 2; i >= 0; i--) {
        const int co = DC[i + 1], ci = DC[i];
        PView Z = pview(W.Z[i], h, w, co);
        PView DYi = pview(W.DY[i], h, w, co);
        // train-mode BatchNorm backward in closed form: two per-channel reductions, then one elementwise pass (in place: DY -> DZ)
        AFI_CUDA(cudaMemsetAsync(W.sums, 0, 2 * 1024 * sizeof(double), st));
        AFI_TRY(bn_bwd_reduce(DYi, Z, dt, W.mean[i], W.rstd[i], n, h, w, co, W.sums, W.sums + 1024, st));
        AFI_TRY(bn_bwd_apply(DYi, Z, dt, W.mean[i], W.rstd[i], p->gamma[i], W.sums, W.sums + 1024, gradacc + GL.gamma[i], gradacc + GL.beta[i],
                             n, h, w, co, st));
        AFI_TRY(wgrad_std(ctx, prec, pview(W.A[i], h, w, ci), ci, DYi, co, n, h, w, gradacc + GL.w[i], st));
        // bias gradient: this bias feeds a train-mode BatchNorm, so dL/db = sum_p dz = 0 identically (the reference gets
        // ~1e-9 rounding noise there); the accumulator slot stays at its zero-initialised value.
        if (i > 0) {   // dA_{i} * lrelu'(a_i) -> DY[i-1]
            ConvArgs a;
            conv_args_init(a);
            a.N = n; a.H = h; a.W = w; a.cin = co; a.cout = ci; a.ntaps = 9; set_std_taps(a.taps, 0, 0);
            a.in[0] = DYi; a.w = (const char*)packed + L.d[i] * es; a.mask = pview(W.A[i], h, w, ci);
            a.out = pview(W.DY[i - 1], h, w, ci); a.out_dt = dt;
            AFI_TRY(run_conv(ctx, prec, a, st));
        }
    }
    return AFI_OK;
}

int afi_d_unpack_grads(afi_ctx* ctx, int prec, const float* gradacc, const afi_d_grads* g, float scale, int accumulate, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(ctx && gradacc && g && prec_ok(prec), "afi_d_unpack_grads: bad argument");
    DGradAcc GL = d_gradacc_layout();
    int nk = prec_tc(prec) ? 1 : 0;
    for (int i = 0; i < 3; i++) {
        if (g->w[i]) AFI_TRY(unpack_wgrad(gradacc + GL.w[i], DC[i + 1], DC[i], nk, 0, g->w[i], scale, accumulate, st));
        if (g->b[i]) AFI_TRY(axpby_f32(gradacc + GL.b[i], g->b[i], DC[i + 1], scale, accumulate, st));
        if (g->gamma[i]) AFI_TRY(axpby_f32(gradacc + GL.gamma[i], g->gamma[i], DC[i + 1], scale, accumulate, st));
        if (g->beta[i]) AFI_TRY(axpby_f32(gradacc + GL.beta[i], g->beta[i], DC[i + 1], scale, accumulate, st));
    }
    if (g->w[3]) AFI_TRY(axpby_f32(gradacc + GL.w[3], g->w[3], 9 * DC[3], scale, accumulate, st));
    if (g->b[3]) AFI_TRY(axpby_f32(gradacc + GL.b[3], g->b[3], 1, scale, accumulate, st));
    return AFI_OK;
}

// =====================================================================================================
// single-layer entry points (unit tests / kernel benchmarks)
// =====================================================================================================
static inline int pad64(int c) { return c <= 64 ? c : (c + 63) / 64 * 64; }
size_t afi_conv3x3_workspace_bytes(int prec, int n, int cin, int h, int w, int cout) {
    size_t es = dt_size(prec_dt(prec)), P = (size_t)n * h * w;
    return align_up(P * pad64(cin) * es) + 2 * align_up(P * pad64(cout) * es) + 2 * align_up((size_t)9 * cin * cout * es) +
           align_up((size_t)9 * cin * cout * 4) + align_up(P * cin * 4) + 4096;
}
int afi_conv3x3(afi_ctx* ctx, int prec, afi_view4 x, int n, int cin, int h, int w, const float* weight, const float* bias, int cout,
                int lrelu, float* y, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(ctx && x.ptr && weight && y && ws && prec_ok(prec), "afi_conv3x3: bad argument");
    AFI_REQUIRE(cin % 32 == 0 && cout % 32 == 0, "afi_conv3x3: channels must be multiples of 32");
    if (ws_bytes < afi_conv3x3_workspace_bytes(prec, n, cin, h, w, cout)) { set_error("afi_conv3x3: workspace too small"); return AFI_ERR_WORKSPACE; }
    int dt = prec_dt(prec); size_t es = dt_size(dt), P = (size_t)n * h * w;
    Carver cv(ws);
    void* X = cv.take(P * cin * es); void* Y = cv.take(P * cout * es); cv.take(P * cout * es);
    void* Wp = cv.take((size_t)9 * cin * cout * es);
    AFI_TRY(to_nhwc(prec, x, n, cin, h, w, pview(X, h, w, cin), st));
    AFI_TRY(pack_weights(weight, cout, cin, pm(prec, 0), Wp, dt, st));
    ConvArgs a;
    conv_args_init(a);
    a.N = n; a.H = h; a.W = w; a.cin = cin; a.cout = cout; a.ntaps = 9; set_std_taps(a.taps, 0, 0);
    a.in[0] = pview(X, h, w, cin); a.w = Wp; a.bias = bias; a.act = lrelu;
    a.out = pview(Y, h, w, cout); a.out_dt = dt;
    AFI_TRY(run_conv(ctx, prec, a, st));
    afi_view4 none; memset(&none, 0, sizeof(none));
    AFI_TRY(to_nchw(prec, pview(Y, h, w, cout), pview_null(), none, 0, 0, 1.f, n, cout, h, w, y, st));
    return AFI_OK;
}
int afi_conv3x3_backward(afi_ctx* ctx, int prec, afi_view4 x, afi_view4 dy, int n, int cin, int h, int w, const float* weight, int cout,
                         float* dw, float* dxo, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AFI_REQUIRE(ctx && x.ptr && dy.ptr && weight && dw && ws && prec_ok(prec), "afi_conv3x3_backward: bad argument");
    AFI_REQUIRE(cin % 32 == 0 && cout % 32 == 0, "afi_conv3x3_backward: channels must be multiples of 32");
    if (ws_bytes < afi_conv3x3_workspace_bytes(prec, n, cin, h, w, cout)) { set_error("afi_conv3x3_backward: workspace too small"); return AFI_ERR_WORKSPACE; }
    int dt = prec_dt(prec); size_t es = dt_size(dt), P = (size_t)n * h * w;
    Carver cv(ws);
    // pixel strides padded to whole 64-channel groups (the tensor-core wgrad reads operands through grouped TMA views)
    const int xs = pad64(cin), ys = pad64(cout);
    void* X = cv.take(P * xs * es); void* DYb = cv.take(P * ys * es); cv.take(P * ys * es);
    void* Wp = cv.take((size_t)9 * cin * cout * es); cv.take((size_t)9 * cin * cout * es);
    float* acc = (float*)cv.take((size_t)9 * cin * cout * 4);
    void* DX = cv.take(P * cin * 4);
    if (xs != cin) AFI_CUDA(cudaMemsetAsync(X, 0, P * xs * es, st));
    if (ys != cout) AFI_CUDA(cudaMemsetAsync(DYb, 0, P * ys * es, st));
    AFI_TRY(to_nhwc(prec, x, n, cin, h, w, pview(X, h, w, xs), st));
    AFI_TRY(to_nhwc(prec, dy, n, cout, h, w, pview(DYb, h, w, ys), st));
    AFI_CUDA(cudaMemsetAsync(acc, 0, (size_t)9 * cin * cout * 4, st));
    AFI_TRY(wgrad_std(ctx, prec, pview(X, h, w, xs), cin, pview(DYb, h, w, ys), cout, n, h, w, acc, st));
    AFI_TRY(unpack_wgrad(acc, cout, cin, prec_tc(prec) ? 1 : 0, 0, dw, 1.f, 0, st));
    if (dxo) {
        AFI_TRY(pack_weights(weight, cout, cin, pm(prec, 1), Wp, dt, st));
        ConvArgs a;
        conv_args_init(a);
        a.N = n; a.H = h; a.W = w; a.cin = cout; a.cout = cin; a.ntaps = 9; set_std_taps(a.taps, 0, 0);
        a.in[0] = pview(DYb, h, w, ys); a.w = Wp; a.out = pview(DX, h, w, cin); a.out_dt = DT_F32;
        AFI_TRY(run_conv(ctx, prec, a, st));
        afi_view4 none; memset(&none, 0, sizeof(none));
        AFI_TRY(nhwc_to_nchw<float>(pview(DX, h, w, cin), pview_null(), none, 0, 0, 1.f, n, cin, h, w, dxo, st));
    }
    return AFI_OK;
}

}  // extern "C"
