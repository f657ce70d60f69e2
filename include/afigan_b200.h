/*
 * afigan_b200.h -- C ABI of libafigan_b200.so: the B200 (sm_100a) implementation of AFI-GAN's hot path.
 *
 * The reference (inhavl-shlee/AFI-GAN) has no FFI of its own: its "operator API" for this path is the
 * nn.Module contract of Generator / Discriminator plus the loss block of the stage-1/2 trainers.  Each
 * entry point below names the reference interface it replaces (file:line under /root/reference).
 * The Python side (afi-gan_b200/afigan) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *  - plain C: device pointers, sizes, a cudaStream_t passed as void*.  No torch types.
 *  - every function returns 0 on success or a negative AFI_ERR_* code; afi_last_error() gives the text
 *    (thread-local).  Unsupported shapes / precision / architecture are hard errors: there is no fallback.
 *  - all work is enqueued on the caller's stream; nothing synchronises; nothing is allocated: the caller
 *    provides workspaces sized by the *_bytes() queries (torch owns all memory).
 *  - input tensors cross the boundary as strided fp32 or bf16 NCHW-logical views (afi_view4: element strides and a dtype
 *    tag), so contiguous and channels_last torch tensors, top-left crops and autocast (bf16) activations are accepted
 *    without copies or up-casts; outputs and gradients are contiguous fp32 NCHW.
 *  - parameters and gradients use the reference's state-dict layouts.
 */
#ifndef AFIGAN_B200_H
#define AFIGAN_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFI_ABI_VERSION 4   /* 2: afi_d_call.input_staged, afi_g_call.fuse_cur / fuse_w, afi_sgd_step_multi, afi_sizeof
                             * 3: AFI_PREC_SPLIT, afi_conv1x1*, afi_conv3x3s2* (struct layouts unchanged since 2)
                             * 4: afi_view4.dtype (fp32 or bf16 input views): afi_view4 and every struct that embeds one grew */

#define AFI_OK 0
#define AFI_ERR_INVALID (-1)   /* bad argument / unsupported shape */
#define AFI_ERR_CUDA (-2)      /* a CUDA runtime/driver call failed */
#define AFI_ERR_ARCH (-3)      /* device is not sm_100 */
#define AFI_ERR_WORKSPACE (-4) /* workspace too small */

/* Operand modes (bf16 operands cannot meet 1e-3 on gradients, fp32-accurate products can). */
#define AFI_PREC_FP32 0      /* fp32 storage + fp32 FFMA implicit GEMM: the parity mode              */
#define AFI_PREC_BF16 1      /* bf16 storage + tcgen05/TMEM implicit GEMM fed by TMA: throughput mode */
#define AFI_PREC_BF16_SIMT 2 /* bf16 storage + CUDA-core GEMM: on-device cross-check of mode 1        */
#define AFI_PREC_SPLIT 3     /* fp32 storage; every GEMM operand split on the fly into three bf16 planes (hi + mid + lo) and
                              * multiplied on the tcgen05 tensor cores as six plane-pair products with fp32 TMEM accumulation:
                              * fp32-accurate products at tensor-core rate -- the parity mode that meets north_star's tolerance */

#define AFI_MAX_RDB 4
#define AFI_MAX_CALLS 10 /* calls per grouped launch */
#define AFI_CH 256 /* feature channels of the AF interpolator and of the discriminator input */

typedef struct afi_ctx afi_ctx;

/* Strided view of a logical [N, C, H, W] tensor; strides in ELEMENTS of its dtype. */
#define AFI_DT_F32 0
#define AFI_DT_BF16 1
typedef struct {
    const void* ptr;
    long long sn, sc, sh, sw;
    int dtype;                  /* AFI_DT_F32 | AFI_DT_BF16 (a zeroed struct is an absent fp32 view) */
    int reserved;               /* must be 0 */
} afi_view4;

/* ---- library ------------------------------------------------------------------------------------ */
int afi_abi_version(void);
/* sizeof() of the boundary structs as this library was built (which: 0 afi_view4, 1 afi_g_params, 2 afi_lateral, 3 afi_g_call,
 * 4 afi_d_params, 5 afi_d_call, 6 afi_g_grads, 7 afi_d_grads; 0 for anything else): lets a binding check its own struct layouts */
size_t afi_sizeof(int which);
const char* afi_last_error(void);
/* Binds to the CURRENT device; fails with AFI_ERR_ARCH unless it is compute capability 10.x. */
int afi_create(afi_ctx** out);
void afi_destroy(afi_ctx* ctx);

/* ---- AF interpolator ("Generator")  generator_rdb.py:75-130 ---------------------------------------- */
typedef struct { /* device pointers, fp32, torch layouts:  Generators.0.* of App. B */
    int n_rdb;
    const float* head_w; const float* head_b;   /* Generators.0.0.0  [256,256,3,3] [256]   generator_rdb.py:91-93  */
    const float* rdb_w[AFI_MAX_RDB][5];          /* ...1.RDBs.r.conv{1..4}.0 / conv5        generator_rdb.py:39-55  */
    const float* post_w; const float* post_b;   /* Generators.0.2.0                        generator_rdb.py:97-99  */
    const float* up_w; const float* up_b;       /* Generators.0.3.0  ConvTranspose [256,256,6,6]   :101-105         */
    const float* out_w; const float* out_b;     /* Generators.0.4.0                        generator_rdb.py:107-108 */
} afi_g_params;

typedef struct { /* where gradients are written (same layouts); any pointer may be NULL to skip it */
    int n_rdb;
    float* head_w; float* head_b;
    float* rdb_w[AFI_MAX_RDB][5];
    float* post_w; float* post_b;
    float* up_w; float* up_b;
    float* out_w; float* out_b;
} afi_g_grads;

/* Optional fused lateral for the FPN/PAFPN top-down merge (fpn_sr.py:151-157, pafpn_sr.py:175-181):
 * y = (conv1x1(lat_x; lat_w, lat_b) + G(x)) * scale, scale = 0.5 for FUSE_TYPE "avg".  lat_x is [N, lat_c, 2h, 2w]. */
typedef struct {
    afi_view4 lat_x; int lat_c;
    const float* lat_w; const float* lat_b; /* [256, lat_c, 1, 1], [256] or NULL */
    float scale;
} afi_lateral;

size_t afi_g_packed_bytes(int prec, int n_rdb);                  /* packed forward+dgrad weights          */
size_t afi_g_gradacc_bytes(int n_rdb);                           /* packed fp32 gradient accumulator      */
/* lat_c: channels of the fused lateral input (0 = no lateral) */
size_t afi_g_workspace_bytes(int prec, int n, int h, int w, int n_rdb, int lat_c, int save_for_backward);

/* Re-layout the parameters for the GEMM kernels.  Call again whenever the parameters change. */
int afi_g_pack(afi_ctx*, int prec, const afi_g_params*, void* packed, void* stream);

/* One Generator call.  Calls of a group (e.g. the five pyramid levels of a stage-1 step) share the weights and are
 * evaluated by GROUPED kernel launches: one launch per layer covers every call of the group. */
typedef struct {
    afi_view4 x; int n, h, w;       /* input features [n,256,h,w]                                                     */
    float* y; int oh, ow;           /* forward: contiguous [n,256,oh,ow] output, oh <= 2h, ow <= 2w (top-left crop)     */
    afi_view4 dy; float* dx;        /* backward: dL/dy view of [n,256,oh,ow]; dx contiguous [n,256,h,w] or NULL         */
    const afi_lateral* lateral;     /* optional fused FPN lateral, may be NULL                                          */
    float* lat_dx; float* lat_gw; float* lat_gb; /* backward: d lat_x [n,lat_c,oh,ow], d lat_w [256,lat_c], d lat_b [256] (overwritten; may be NULL) */
    void* ws; size_t ws_bytes;      /* workspace of afi_g_workspace_bytes(); keeps the activations between fwd and bwd  */
    /* optional (forward only, fuse_w != NULL): the BiFPN fusion site (bifpn_sr.py:535-548) folded into the output pass:
     * y = fuse_w[0] * fuse_cur + fuse_w[1] * (Generators[0](x) + bilinear2x(x)); fuse_cur is a view of [n,256,oh,ow], fuse_w two
     * floats ON THE DEVICE (the raw attention weights `BiFPNLayer_*_w1`; {1, 1} for the attention-free sum) */
    afi_view4 fuse_cur; const float* fuse_w;
} afi_g_call;

/* Generator.forward (generator_rdb.py:123-130) for each call: y = (Generators[0](x) + bilinear2x(x))[:, :, :oh, :ow]
 * (the top-left crop of _reshape_stage1, stage1_trainer.py:437-443, folded in). */
int afi_g_forward(afi_ctx*, int prec, const afi_g_params*, const void* packed, const afi_g_call* calls, int ncalls,
                  int save_for_backward, void* stream);

/* Backward of the calls that filled their `ws`.  Weight/bias gradients of ALL calls are ADDED into `gradacc` (packed,
 * zero it with afi_zero first). */
int afi_g_backward(afi_ctx*, int prec, const afi_g_params*, const void* packed, const afi_g_call* calls, int ncalls,
                   float* gradacc, void* stream);

/* grads (torch layouts) = [grads +] scale * unpack(gradacc) */
int afi_g_unpack_grads(afi_ctx*, int prec, const float* gradacc, const afi_g_grads*, float scale, int accumulate, void* stream);

/* ---- feature-patch discriminator  feature_patch_discriminator.py:18-55 -------------------------------- */
typedef struct { /* Discriminators.0.* of App. B */
    const float* w[4]; const float* b[4];            /* [512,256,3,3] [1024,512,3,3] [1024,1024,3,3] [1,1024,3,3] */
    const float* gamma[3]; const float* beta[3];     /* ...norm.weight / norm.bias                                  */
    float* running_mean[3]; float* running_var[3];   /* updated in training mode (momentum, unbiased var); may be NULL */
    long long* num_batches_tracked[3];               /* +1 per training forward; may be NULL                         */
} afi_d_params;

typedef struct {
    float* w[4]; float* b[4]; float* gamma[3]; float* beta[3];
} afi_d_grads;

size_t afi_d_packed_bytes(int prec);
size_t afi_d_gradacc_bytes(void);
size_t afi_d_workspace_bytes(int prec, int n, int h, int w, int save_for_backward);
int afi_d_pack(afi_ctx*, int prec, const afi_d_params*, void* packed, void* stream);

/* One discriminator call; a group (e.g. level x {real, fake} of a stage-1 phase) runs as grouped launches per layer. */
typedef struct {
    afi_view4 x; int n, h, w;       /* input [n,256,h,w]                               */
    float* logits;                  /* forward: contiguous [n,1,h,w]; NULL (training, no backward) = the call only contributes
                                     * its BatchNorm batch statistics -- the dead D(hr) of stage1_trainer.py:400          */
    const float* dlogits;           /* backward: contiguous [n,1,h,w]                  */
    float* dx;                      /* backward: contiguous [n,256,h,w] or NULL           */
    void* ws; size_t ws_bytes;
    int input_staged;               /* forward: non-zero = ws still holds THIS input in the library's layout from an earlier
                                     * afi_d_forward on the same ws (the two phases of a stage-1 step feed the same tensors): the
                                     * layout conversion is skipped                                                          */
} afi_d_call;

/* Discriminators[0](x) (feature_patch_discriminator.py:32-41 as called at stage1_trainer.py:349-353) for each call, IN CALL
 * ORDER as far as the BatchNorm running buffers are concerned.  training != 0: each call is normalised with ITS OWN batch
 * statistics (biased var, eps) and updates the running buffers; training == 0: running statistics are used; training == 2: see
 * afi_d_update_running. */
int afi_d_forward(afi_ctx*, int prec, const afi_d_params*, const void* packed, const afi_d_call* calls, int ncalls,
                  int training, float momentum, float eps, int save_for_backward, void* stream);

/* training == 2 in afi_d_forward defers the running-buffer update: the calls (possibly issued as several groups on concurrent
 * streams) only record their batch statistics in ws; afi_d_update_running then applies the momentum update of `calls` IN THE GIVEN
 * ORDER (= the reference's call order) and bumps num_batches_tracked by ncalls. */
int afi_d_update_running(afi_ctx*, int prec, const afi_d_params*, const afi_d_call* calls, int ncalls, float momentum, void* stream);

/* Backward of the calls that filled their ws with save_for_backward; `training` must repeat the forward's mode (non-zero: batch
 * statistics; 0: running statistics, where the BatchNorm backward has no mean terms and the conv biases get a gradient).
 * Gradients of all calls are ADDED into gradacc (packed); dx (optional, all calls or none) is overwritten. */
int afi_d_backward(afi_ctx*, int prec, const afi_d_params*, const void* packed, const afi_d_call* calls, int ncalls, int training,
                   float* gradacc, void* stream);
int afi_d_unpack_grads(afi_ctx*, int prec, const float* gradacc, const afi_d_grads*, float scale, int accumulate, void* stream);

/* ---- losses of the stage-1/2 trainers ---------------------------------------------------------------- */
/* nn.BCEWithLogitsLoss() against a constant target map (stage1_trainer.py:154,355-359,408):
 * *loss_out = mean(max(x,0) - x t + log1p(exp(-|x|))); if loss_sum: *loss_sum += weight * mean;
 * if dlogits: dlogits = grad_scale * (sigmoid(x) - t) / count. */
int afi_bce_with_logits(const float* logits, long long count, float target, float* loss_out, float* loss_sum,
                        float weight, float* dlogits, float grad_scale, void* stream);
/* F.l1_loss(a, b) on [n,c,h,w] views (stage1_trainer.py:410): *loss_out = mean|a-b|; if loss_sum: += weight*mean;
 * if da: da (contiguous [n,c,h,w]) = grad_scale * sign(a-b) / numel. */
int afi_l1_loss(afi_view4 a, afi_view4 b, int n, int c, int h, int w, float* loss_out, float* loss_sum, float weight,
                float* da, float grad_scale, void* stream);

/* ---- optimiser step + helpers ------------------------------------------------------------------------- */
/* torch.optim.SGD as detectron2's build_optimizer configures it [upstream]: d = grad_scale*g + wd*p;
 * m = first ? d : momentum*m + d;  p -= lr*m.  grad_scale folds the 1/world_size of the gradient all-reduce. */
int afi_sgd_step(float* p, const float* g, float* m, long long count, float lr, float momentum, float wd,
                 float grad_scale, int first, void* stream);
/* The same update for n tensors in ONE launch (a detectron2 optimiser step over all parameters of a module, one param group each:
 * weight decay per tensor).  p, g, m, count, wd are HOST arrays of length n. */
int afi_sgd_step_multi(int n, float* const* p, const float* const* g, float* const* m, const long long* count, const float* wd,
                       float lr, float momentum, float grad_scale, int first, void* stream);
int afi_zero(void* ptr, size_t bytes, void* stream);

/* ---- single-layer entry points (unit tests and kernel benchmarks call the GEMM kernels through these) ---- */
/* y = [lrelu](conv3x3(x, w) + b) with stride 1, pad 1; x [n,cin,h,w] view, w [cout,cin,3,3], y contiguous [n,cout,h,w]. */
int afi_conv3x3(afi_ctx*, int prec, afi_view4 x, int n, int cin, int h, int w_, const float* weight, const float* bias,
                int cout, int lrelu, float* y, void* ws, size_t ws_bytes, void* stream);
/* dw [cout,cin,3,3] = wgrad of the same conv for upstream gradient dy [n,cout,h,w]; db [cout] = bias gradient (may be NULL);
 * dxo = dgrad, contiguous [n,cin,h,w] (may be NULL).  Also backs the necks' 3x3 output convs (fpn_sr.py:144-158). */
int afi_conv3x3_backward(afi_ctx*, int prec, afi_view4 x, afi_view4 dy, int n, int cin, int h, int w_, const float* weight,
                         int cout, float* dw, float* db, float* dxo, void* ws, size_t ws_bytes, void* stream);
size_t afi_conv3x3_workspace_bytes(int prec, int n, int cin, int h, int w_, int cout);
/* The same for 1x1 convolutions: the necks' lateral convs incl. the top-level one (fpn_sr.py:79-81,144-145), the pointwise half of the
 * BiFPN's depthwise-separable convs (bifpn_layers/wrappers.py:166-206) and its input laterals (bifpn_sr.py:160-183).  weight [cout,cin,1,1]. */
int afi_conv1x1(afi_ctx*, int prec, afi_view4 x, int n, int cin, int h, int w_, const float* weight, const float* bias, int cout,
                float* y, void* ws, size_t ws_bytes, void* stream);
int afi_conv1x1_backward(afi_ctx*, int prec, afi_view4 x, afi_view4 dy, int n, int cin, int h, int w_, const float* weight,
                         int cout, float* dw, float* db, float* dxo, void* ws, size_t ws_bytes, void* stream);
size_t afi_conv1x1_workspace_bytes(int prec, int n, int cin, int h, int w_, int cout);
/* 3x3 / stride 2 / pad 1 convolution (+bias): the bottom-up down-sampling convs of the PANet neck (pafpn_sr.py:103-117, 186-193).
 * x [n,cin,h,w] -> y contiguous [n,cout,ceil(h/2),ceil(w/2)]; backward: dw [cout,cin,3,3], db [cout] or NULL, dxo [n,cin,h,w] or NULL.
 * Tensor-core engines only (AFI_PREC_BF16, AFI_PREC_SPLIT); h, w >= 2. */
int afi_conv3x3s2(afi_ctx*, int prec, afi_view4 x, int n, int cin, int h, int w_, const float* weight, const float* bias, int cout,
                  float* y, void* ws, size_t ws_bytes, void* stream);
int afi_conv3x3s2_backward(afi_ctx*, int prec, afi_view4 x, afi_view4 dy, int n, int cin, int h, int w_, const float* weight,
                           int cout, float* dw, float* db, float* dxo, void* ws, size_t ws_bytes, void* stream);
size_t afi_conv3x3s2_workspace_bytes(int prec, int n, int cin, int h, int w_, int cout);

/* ---- BiFPN neck at inference (bifpn_sr.py:583-729) ------------------------------------------------------------------------------ */
/* SeparableConv2d of bifpn_layers/wrappers.py:166-206 with its (eval-mode) norm folded into pw_w / pw_b by the caller, fused with the swish the
 * neck applies in front of it: y = conv1x1(dw3x3(pre_swish ? swish(x) : x; dw_w); pw_w, pw_b).  x [n,c,h,w] view, dw_w [c,1,3,3],
 * pw_w [cout,c,1,1], pw_b [cout] or NULL, y contiguous [n,cout,h,w].  Forward only. */
int afi_sepconv(afi_ctx*, int prec, afi_view4 x, int n, int c, int h, int w_, const float* dw_w, const float* pw_w, const float* pw_b, int cout,
                int pre_swish, float* y, void* ws, size_t ws_bytes, void* stream);
size_t afi_sepconv_workspace_bytes(int prec, int n, int c, int h, int w_, int cout);
/* Bottom-up fusion site `_feature_funsion2` (bifpn_sr.py:550-564): out [n,c,h,w] = w[0]*a + w[1]*b + w[2]*pool(down), or w[0]*a + w[1]*pool(down)
 * when b.ptr == NULL; weights = device floats (the raw `BiFPNLayer_*_w2` parameter) or NULL for a plain sum; pool = the reference's 'static_same'
 * MaxPool2d(3, 2): zero padding on the right / bottom that takes part in the max; down is [n,c,dh,dw]. */
int afi_bifpn_fuse_down(afi_view4 a, afi_view4 b, afi_view4 down, const float* weights, int n, int c, int h, int w_, int dh, int dw, float* out,
                        void* stream);

/* Top-down fusion site `_feature_funsion` WITH autograd (bifpn_sr.py:542-548) and the swish the neck applies to it (:591-594) in one
 * elementwise pass each way:  s = w[0]*cur + w[1]*up (weights NULL: 1, 1);  out = act ? s*sigmoid(s) : s.  cur / up [n,c,h,w] views; s_out
 * (pre-activation, needed by the backward when act; may be NULL otherwise) and out contiguous fp32.
 * backward: ds = dout * d act(s);  d_cur = w[0]*ds, d_up = w[1]*ds (each may be NULL), d_w[2] = (sum ds*cur, sum ds*up) (may be NULL; overwritten). */
int afi_bifpn_fuse_act(afi_view4 cur, afi_view4 up, const float* weights, int act, int n, int c, int h, int w_, float* s_out, float* out,
                       void* stream);
int afi_bifpn_fuse_act_backward(afi_view4 dout, const float* s, afi_view4 cur, afi_view4 up, const float* weights, int act, int n, int c, int h,
                                int w_, float* d_cur, float* d_up, float* d_w, void* stream);

/* Per-launch CUDA-event timing of the implicit-GEMM kernels (bench.py's roofline leg).  begin: start recording up to
 * max_launches GEMM launches; end: device-synchronise and resolve the durations; get: record i = kind (0 conv tcgen05 per-tap
 * kernel, 1 wgrad tcgen05, 2 conv CUDA-core, 3 wgrad CUDA-core, 4 conv tcgen05 halo tiles on CTA pairs, 5 conv tcgen05 halo tiles on
 * single CTAs), algorithmic FLOPs (2*pixels*taps*cin*cout), milliseconds. */
int afi_profile_begin(int max_launches);
int afi_profile_end(int* n_launches);
int afi_profile_get(int i, int* kind, double* flops, float* ms, int* cin, int* cout, long long* pixels);

/* Kernel launches issued by this library since the last call with reset != 0 (bench.py's gpu_launches). */
long long afi_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* AFIGAN_B200_H */
