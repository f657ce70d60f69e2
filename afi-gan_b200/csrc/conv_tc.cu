// tcgen05 / TMEM / TMA implicit-GEMM engine (AFI_PREC_BF16): the throughput path of the library.  Three kernels:
//
//   k_conv_halo<EPI, PAIR> : every standard 3x3 conv / dgrad.  D[128 pixels x bn couts] per CTA; the M tile is a 16 x 8 (or 8 x 16) patch
//           of ONE image whose 18 x 10 HALO is fetched once per 64-channel chunk (one 4-D TMA box; out-of-image coordinates are
//           zero-filled = the conv padding, which also handles ragged 13x21 / 7x11 levels); the nine taps read it in place through
//           shifted UMMA descriptors.  PAIR: two CTAs of a cluster run ONE tcgen05.mma.cta_group::2 stream (M = 256) and split every
//           weight tile between them.  See the comment above the kernel.
//   k_conv_tc<EPI>         : per-tap variant (one shifted A box per tap, cta_group::1) for the narrow layers, 1x1 convs and any
//           tap set that is not a standard 3x3.
//   k_wgrad_tc<PAIR>       : D[128 couts x bn cins] += dY[64 px x 128 co]^T * X_shift[64 px x bn ci]: both operands MN-major (channels
//           contiguous), K = pixels; ONE 5-D grouped TMA box per operand per stage, split-K over the concatenated spatial patches of
//           all problems of the group, fp32 RED into dW.  PAIR: M = 256 couts across a CTA pair, the X tile split between them.
//   All are GROUPED: one launch covers up to AFI_MAX_PROB problems (pyramid levels / discriminator calls) that share the weights; the
//   persistent CTAs walk a largest-first work list spanning all of them.
//
// Warp roles (1 CTA/SM, persistent over a static work list):
//   warp 0   : TMA producer   -- mbarrier full/empty rings (4-10 stages depending on the N tile; separate halo / weight rings in
//                                k_conv_halo)
//   warp 1   : TMEM allocator + MMA issuer -- tcgen05.mma kind::f16 (bf16 x bf16 -> fp32), M = 128 (256 per pair), N = bn (runtime)
//              Both run their loops as CONVERGED warps and elect one lane only for the instruction issue (elect_one()).
//   warps 2+ : epilogue, 4 warps (long-K convs: 192 threads) or 8 warps (short-K convs, statistics below K = 8192, wgrad: 320 threads)
//                                            -- tcgen05.ld 32x32b from a double-buffered TMEM accumulator (2 x 256 columns)
//                                               so the epilogue of tile i overlaps the MMAs of tile i+1; optional fused per-channel
//                                               statistics (BatchNorm sum / sum of squares) via a 16-shuffle butterfly per chunk.
// Every mbarrier wait has a 4 s watchdog that traps instead of hanging the GPU.
#include <cuda.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include "common.cuh"

namespace afi {

namespace tc {
constexpr int STAGES = 4;             // stages of A_BYTES + B_BYTES_MAX that fit; narrower N tiles get more (Tiling.nstages)
constexpr int STAGES_MAX = 10;
constexpr int A_BYTES = 16384;          // 128 rows x 128 B
constexpr int B_BYTES_MAX = 32768;      // 256 rows x 128 B
constexpr int STAGE_BYTES = A_BYTES + B_BYTES_MAX;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
constexpr int NUM_THREADS = 320;        // warp 0: TMA, warp 1: MMA, warps 2-9: epilogue (two per TMEM lane quarter)
constexpr int EPI_THREADS = 256;
constexpr int ACC_COLS = 256;

struct alignas(64) Maps {
    CUtensorMap a[AFI_MAX_PROB][4];   // conv: input views per problem;  wgrad: a[p][0] = dY, a[p][1] = X
    CUtensorMap b;                    // conv: packed weights
};
struct TileP {                        // per problem
    int TH, TW, tiles_x, tiles_y;     // spatial patch and patch grid per image
    int begin;                        // conv: first work item (M tile x N tile) / wgrad: first K tile of this problem
    int prob;                         // index into ConvArgs.p / WgradArgs.p (problems are scheduled largest first)
    int orient;                       // halo kernels: 0 = 16 rows x 8 columns (8-pixel groups run along x), 1 = 8 rows x 16 columns (along y)
};
struct Tiling {
    int bn, n_tiles;                  // N tile (multiple of 16, <= 256)
    int kchunks;                      // ceil(cin / 64)
    int total;                        // work items
    int nprob;
    int m_tiles, ksplit, ktiles;      // wgrad only
    int nstages, stage_bytes;         // per-tap conv and wgrad: operand ring (a narrow N tile leaves room for more, smaller stages)
    // halo-tile convolution only: B ring of `sb` slots of b_slot bytes (one (tap, chunk) weight tile, or half of it in pair mode)
    int sb, b_slot, nviews;
    int tb;                           // taps per weight slot (1, 3 or 9): one TMA box / one full-empty barrier pair covers tb consecutive taps
    int view_slab0[4];                // first weight slab of view v (its nine taps use slab0 .. slab0 + 8 in standard order)
    // split-precision operands (ConvArgs.split): the K loop runs once per plane pair (pa[i], pb[i]); plane p of image n of an A view
    // is image p * N + n, plane p of weight slab s is slab p * nslab_total + s.  npairs = 1, pa = pb = {0} for plain bf16 operands.
    int npairs, nslab_total;
    int pa[6], pb[6];
    // split mode, long accumulation chains: the K loop alternates between the two TMEM accumulators and the epilogue adds them (DUAL_ACC
    // in k_conv_tc); the accumulator is then single-buffered (no epilogue / main-loop overlap)
    int dual;
    long long* dbg;                   // optional [grid][8] stall-cycle counters (AFIGAN_HALO_DBG)
    TileP p[AFI_MAX_PROB + 1];        // p[nprob].begin = end sentinel
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ unsigned int g_tc_timeout_flag = 0;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int who) {
    uint32_t done = 0;
    unsigned long long t0 = 0;
    uint32_t spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        if ((++spins & 1023u) == 0) {   // watchdog: a pipeline bug must trap, never hang the GPU
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 4000000000ull) {
                atomicExch(&g_tc_timeout_flag, 0x100u + (unsigned)who);
                __trap();
            }
        }
    }
}
// AFI_STALL_COUNTERS (build-time): per-thread stall-cycle counters of the producer / MMA threads, printed by conv_tc() when the
// environment has AFIGAN_HALO_DBG -- the instrument behind the pipeline notes in DESIGN.md.  Off in the shipped library: every
// instruction in those single-thread loops is on the critical path.
#ifdef AFI_STALL_COUNTERS
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define AFI_TSTAMP(slot) do { if (tl.dbg) tl.dbg[2048 + blockIdx.x * 8 + (slot)] = gtime(); } while (0)
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, int who, long long& acc, bool on) {
    if (on) { long long t0 = clock64(); mbar_wait(bar, parity, who); acc += clock64() - t0; }
    else mbar_wait(bar, parity, who);
}
#else
#define AFI_TSTAMP(slot) do { } while (0)
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, int who, long long&, bool) { mbar_wait(bar, parity, who); }
#endif
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
// ---- cta_group::2 (CTA pair) variants ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA 0 of the cluster (the pair's leader)
__device__ __forceinline__ void mbar_arrive_cta0(uint32_t bar) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(bar));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// TMA loads issued by either CTA of the pair into ITS OWN shared memory, completing bytes on the LEADER's barrier (peer bit cleared)
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {      // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Programmatic dependent launch (no-ops unless the kernel was launched with the programmatic-serialization attribute): pdl_launch() lets
// the NEXT kernel of the stream start its prologue (CTA launch, barrier / TMEM set-up) on idle SMs while this one still runs; pdl_wait()
// blocks until the PREVIOUS kernel has completed and its writes are visible -- it precedes every global-memory access.
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// one lane of a CONVERGED warp (the whole warp runs the role loop; only the instruction issue is elected -- the code the compiler emits
// for tcgen05 / TMA instructions inside a divergent `if (lane == 0)` region wraps each of them in an ELECT / branch loop)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
// UMMA shared-memory descriptor, SWIZZLE_128B, version 1 (sm_100).  Offsets in bytes.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;     // descriptor version
    d |= (uint64_t)2 << 61;     // SWIZZLE_128B
    return d;
}
// instruction descriptor kind::f16: D=f32, A=B=bf16, M=128
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn_major, int b_mn_major, int m = 128) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct Aux { uint4 a, b, c, d; };   // one 16-column chunk of auxiliary epilogue operands (64 B per thread)
struct Aux2 { Aux x, y; };          // fp32 auxiliary operands (split-precision mode): 16 floats of the r1 | mask slot, 16 of the r2 slot
template <bool AUX32> struct AuxSel { typedef Aux type; };
template <> struct AuxSel<true> { typedef Aux2 type; };

struct Smem {
    uint64_t full[STAGES_MAX];
    uint64_t empty[STAGES_MAX];
    uint64_t acc_full[2];
    uint64_t acc_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t setup(Smem& s, const Maps& maps, int epi_warps, int warp, int lane) {
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.b);
        prefetch_tmap(&maps.a[0][0]);
        for (int i = 0; i < STAGES_MAX; i++) { mbar_init(smem_u32(&s.full[i]), 1); mbar_init(smem_u32(&s.empty[i]), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(smem_u32(&s.acc_full[i]), 1); mbar_init(smem_u32(&s.acc_empty[i]), epi_warps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return s.tmem_base;
}
__device__ __forceinline__ void teardown(uint32_t tmem_base, int warp) {
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

__device__ __forceinline__ void ld16(const void* base, long long off, int dt, float* v) {
    if (dt == DT_F32) {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
#pragma unroll
        for (int i = 0; i < 4; i++) { float4 t = p[i]; v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w; }
    } else {
        const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(base) + off);
#pragma unroll
        for (int i = 0; i < 2; i++) {
            uint4 u = p[i];
            uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; j++) { v[8 * i + 2 * j] = __uint_as_float(w[j] << 16); v[8 * i + 2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u); }
        }
    }
}
__device__ __forceinline__ void st16(void* base, long long off, int dt, const float* v) {
    if (dt == DT_F32) {
        float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off);
#pragma unroll
        for (int i = 0; i < 4; i++) p[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
        uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(base) + off);
#pragma unroll
        for (int i = 0; i < 2; i++) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]);
                w[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            p[i] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

// Column sums of a [32 lanes x 16 columns] register tile in 16 shuffles (instead of 80): each step halves the number of
// columns a lane is responsible for while doubling the rows it has summed.  On return lane l holds in s[0] the sum over
// the whole warp of column butterfly_col(l) (both lanes of a pair hold the same value; even lanes publish it).
template <int C, int O>
__device__ __forceinline__ void butterfly_step(float (&s)[16], int lane) {
    const bool up = (lane & O) != 0;
#pragma unroll
    for (int i = 0; i < C; i++) {
        float send = up ? s[i] : s[i + C];
        float keep = up ? s[i + C] : s[i];
        s[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
    }
}
__device__ __forceinline__ void butterfly16(float (&s)[16], int lane) {
    butterfly_step<8, 16>(s, lane);     // fully static indexing: the arrays stay in registers
    butterfly_step<4, 8>(s, lane);
    butterfly_step<2, 4>(s, lane);
    butterfly_step<1, 2>(s, lane);
    s[0] += __shfl_xor_sync(0xffffffffu, s[0], 1);
}
__device__ __forceinline__ int butterfly_col(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }

// Epilogue warps of the convolution kernels (shared by the per-tap and the halo-tile main loops): TMEM -> registers -> bias / activation /
// residuals / mask / statistics -> global memory, for every work item of this CTA.
// PAIR: the CTA is one half of a cta_group::2 pair.  Work items are (pair of M tiles, N tile); this CTA owns M tile 2 * pair + rank
// (a duplicate of the last tile, with stores and statistics masked, when the problem has an odd tile count) and releases the
// accumulator stage on the LEADER's barrier, which counts the epilogue warps of both CTAs.
template <int EPI_WARPS, bool PAIR = false, bool AUX32 = false>
__device__ __forceinline__ void conv_epilogue(const ConvArgs& a, const Tiling& tl, uint64_t* acc_full, uint64_t* acc_empty,
                                              float (*sstat)[2][ACC_COLS / (EPI_WARPS / 4)], float (*sbias)[ACC_COLS], const uint32_t tmem_base, const int warp,
                                              const int lane, const int rank = 0) {
    constexpr int HALVES = EPI_WARPS / 4;
    const int item0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, item_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;       // 0 .. HALVES-1
    const int row = q * 32 + lane;
    int as = 0; uint32_t aphase = 0;
    int bias_buf = 0;
    // Column statistics stay in shared memory across the tiles of this CTA for as long as (problem, N tile) does not change
    // (with gridDim a multiple of n_tiles that is the whole run of a problem) and are published with ONE fp64 atomic per
    // column per run: same-address fp64 atomics from 148 CTAs on every tile were the bottleneck of the first version.
    int cur_prob = -1, cur_nt = -1;
    auto flush_stats = [&](int prob, int nt_) {
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
        const ConvProb& fp = a.p[prob];
        const int fn0 = nt_ * tl.bn;
        for (int c = row + half * 128; c < tl.bn; c += 32 * EPI_WARPS) {
            const int w0 = ((c >> 4) % HALVES) * 4;      // the four warps (one per lane quarter) that own this column's chunk
            const int lc = (((c >> 4) / HALVES) << 4) | (c & 15);      // ... and its slot in their compacted (every HALVES-th chunk) arrays
            float t0 = 0.f, t1 = 0.f;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                t0 += sstat[w0 + j][0][lc]; t1 += sstat[w0 + j][1][lc];
                sstat[w0 + j][0][lc] = 0.f; sstat[w0 + j][1][lc] = 0.f;
            }
            if (fn0 + c < a.cout) {
                atomicAdd(fp.stat0 + fn0 + c, (double)t0);
                atomicAdd(fp.stat1 + fn0 + c, (double)t1);
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
    };
    for (int tile = item0; tile < tl.total; tile += item_stride) {
        int ti = 0;
        while (tile >= tl.p[ti + 1].begin) ti++;
        const TileP& tp_ = tl.p[ti];
        const ConvProb& pr = a.p[tp_.prob];
        int local = tile - tp_.begin;
        int nt = local % tl.n_tiles, mt = local / tl.n_tiles;
        if (a.stat_mode && (tp_.prob != cur_prob || nt != cur_nt)) {
            if (cur_prob >= 0) flush_stats(cur_prob, cur_nt);
            cur_prob = tp_.prob; cur_nt = nt;
        }
        int tiles_per_img = tp_.tiles_x * tp_.tiles_y;
        bool tile_valid = true;
        if (PAIR) {
            mt = 2 * mt + rank;
            const int m_tiles = pr.N * tiles_per_img;
            if (mt >= m_tiles) { mt = m_tiles - 1; tile_valid = false; }
        }
        int img = mt / tiles_per_img, r = mt % tiles_per_img;
        const int ty = tp_.orient ? (row & 7) : row / tp_.TW, tx = tp_.orient ? (row >> 3) : row % tp_.TW;
        int y = (r / tp_.tiles_x) * tp_.TH + ty, x = (r % tp_.tiles_x) * tp_.TW + tx;
        const bool ok = tile_valid && (y < pr.H) && (x < pr.W);
        const int n0 = a.nphase ? 0 : nt * tl.bn;                                                   // phase mode: every N tile is a whole conv
        const long long out_off = a.nphase ? (nt >> 1) * (pr.out.sy >> 1) + (nt & 1) * (pr.out.sx >> 1) : 0;
        // Auxiliary epilogue operands (residuals / mask / BN input / fp32 accumulate-in) are prefetched one chunk ahead of their
        // use (a dependent global load per chunk made memory-bound epilogues ~8x slower than the MMA main loop).
        const long long offA = pr.mask.ptr ? img * pr.mask.sn + y * pr.mask.sy + x * pr.mask.sx
                                           : (pr.r1.ptr ? img * pr.r1.sn + y * pr.r1.sy + x * pr.r1.sx : 0);
        const long long offB = pr.r2.ptr ? img * pr.r2.sn + y * pr.r2.sy + x * pr.r2.sx : 0;
        typedef typename AuxSel<AUX32>::type AuxT;
        // element offsets are the same for bf16 and fp32 operands; the byte address depends on AUX32
        const char* pA = pr.mask.ptr ? (const char*)pr.mask.ptr + offA * (AUX32 ? 4 : 2) : (pr.r1.ptr ? (const char*)pr.r1.ptr + offA * (AUX32 ? 4 : 2) : nullptr);
        const char* pB = pr.r2.ptr ? (const char*)pr.r2.ptr + offB * (AUX32 ? 4 : 2) : nullptr;
        const float* pC = pr.accin.ptr ? (const float*)pr.accin.ptr + img * pr.accin.sn + y * pr.accin.sy + x * pr.accin.sx : nullptr;
        const int nchunks = tl.bn >> 4;
        auto aux_load = [&](int ch, AuxT& r) {
            const int col = n0 + (ch << 4);
            if (ch < nchunks && ok && col < a.cout) {
                if constexpr (AUX32) {
                    if (pC) {
                        const uint4* g = reinterpret_cast<const uint4*>(pC + col);
                        r.x.a = g[0]; r.x.b = g[1]; r.x.c = g[2]; r.x.d = g[3];
                    } else {
                        if (pA) { const uint4* g = reinterpret_cast<const uint4*>(pA + (size_t)col * 4); r.x.a = g[0]; r.x.b = g[1]; r.x.c = g[2]; r.x.d = g[3]; }
                        if (pB) { const uint4* g = reinterpret_cast<const uint4*>(pB + (size_t)col * 4); r.y.a = g[0]; r.y.b = g[1]; r.y.c = g[2]; r.y.d = g[3]; }
                    }
                } else {
                    if (pC) {
                        const uint4* g = reinterpret_cast<const uint4*>(pC + col);
                        r.a = g[0]; r.b = g[1]; r.c = g[2]; r.d = g[3];
                    } else {
                        if (pA) { const uint4* g = reinterpret_cast<const uint4*>(pA + (size_t)col * 2); r.a = g[0]; r.b = g[1]; }
                        if (pB) { const uint4* g = reinterpret_cast<const uint4*>(pB + (size_t)col * 2); r.c = g[0]; r.d = g[1]; }
                    }
                }
            }
        };
        auto unpack16 = [](const uint4& lo, const uint4& hi, float* t) {
            const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
            for (int j = 0; j < 8; j++) { t[2 * j] = __uint_as_float(w[j] << 16); t[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u); }
        };
        auto asfloat16 = [](const Aux& q, float* t) {
            const uint4 q4[4] = {q.a, q.b, q.c, q.d};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                t[4 * i] = __uint_as_float(q4[i].x); t[4 * i + 1] = __uint_as_float(q4[i].y);
                t[4 * i + 2] = __uint_as_float(q4[i].z); t[4 * i + 3] = __uint_as_float(q4[i].w);
            }
        };
        // the 16 values of the r1 | mask slot (which = 0), the r2 slot (1) or the fp32 accumulate-in operand (2) of this chunk
        auto aux_get = [&](const AuxT& ax, int which, float* t) {
            if constexpr (AUX32) {
                asfloat16(which == 1 ? ax.y : ax.x, t);
            } else {
                if (which == 2) asfloat16(ax, t);
                else if (which == 0) unpack16(ax.a, ax.b, t);
                else unpack16(ax.c, ax.d, t);
            }
        };
        // chunk ownership: the two warps of a TMEM lane quarter interleave the 16-column chunks (half = 0 / 1)
        const uint32_t taddr = tmem_base + as * ACC_COLS + ((uint32_t)(q * 32) << 16);
        auto process = [&](int ch, const AuxT& ax) {
            const int c0 = ch << 4;
            float v[16];
            tmem_ld16(taddr + c0, v);
            if (AUX32 && tl.dual) {     // split mode: the K loop alternated between the two accumulators (see DUAL_ACC below)
                float v2[16];
                tmem_ld16(taddr + ACC_COLS + c0, v2);
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] += v2[i];
            }
            const int col = n0 + c0;
            const bool valid = ok && col < a.cout;
            if (valid) {
                if (a.bias) {       // staged in shared memory while the MMAs of this tile ran (a dependent global load per chunk otherwise)
                    const float4* b4 = reinterpret_cast<const float4*>(&sbias[bias_buf][c0]);
#pragma unroll
                    for (int i = 0; i < 4; i++) { float4 b = b4[i]; v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w; }
                }
                if (a.act && !a.act_post) {
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = v[i] > 0.f ? v[i] : v[i] * a.slope;
                }
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] *= a.alpha;
                float t[16];
                if (pC) {
                    aux_get(ax, 2, t);
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] += t[i];
                } else {
                    if (pr.r1.ptr) {
                        aux_get(ax, 0, t);
#pragma unroll
                        for (int i = 0; i < 16; i++) v[i] += a.beta1 * t[i];
                    }
                    if (pr.r2.ptr) {
                        aux_get(ax, 1, t);
#pragma unroll
                        for (int i = 0; i < 16; i++) v[i] += a.beta2 * t[i];
                    }
                    if (pr.mask.ptr) {
                        aux_get(ax, 0, t);
#pragma unroll
                        for (int i = 0; i < 16; i++) v[i] *= t[i] > 0.f ? 1.f : a.mask_slope;
                    }
                }
                if (a.act_post || col < a.split_col) {
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = v[i] > 0.f ? v[i] : v[i] * a.slope;
                }
                if (col < a.split_col) st16(pr.out2.ptr, img * pr.out2.sn + y * pr.out2.sy + x * pr.out2.sx + col, a.out2_dt, v);
                else st16(pr.out.ptr, out_off + img * pr.out.sn + y * pr.out.sy + x * pr.out.sx + col, a.out_dt, v);
            }
            if (a.stat_mode) {      // warp-uniform: every lane takes part in the shuffles, invalid rows contribute zeros
                float s0[16], s1[16];
#pragma unroll
                for (int i = 0; i < 16; i++) { s0[i] = valid ? v[i] : 0.f; s1[i] = valid ? v[i] * v[i] : 0.f; }
                butterfly16(s0, lane);
                butterfly16(s1, lane);
                if ((lane & 1) == 0) {      // 16 lanes own 16 distinct columns of this warp's private accumulator
                    int cc = ((ch / HALVES) << 4) + butterfly_col(lane);      // this warp owns every HALVES-th chunk: compacted index
                    sstat[warp - 2][0][cc] += s0[0];
                    sstat[warp - 2][1][cc] += s1[0];
                }
            }
        };
        // two-deep ring of auxiliary-operand buffers per warp: chunk j of this warp lives in buffer j % 2 and is re-filled for
        // chunk j + 2 right after it has been consumed (no register rotation, so no load is waited on early).  With eight
        // epilogue warps that keeps 256 threads x 128 B = 32 KB of loads in flight per SM (~HBM latency x per-SM bandwidth).
        AuxT bA, bB;
        memset(&bA, 0, sizeof(bA));
        bB = bA;
        aux_load(half, bA);
        aux_load(half + HALVES, bB);
        if (a.bias) {           // this tile's bias columns -> shared memory (double-buffered over tiles: one barrier per tile is enough)
            bias_buf ^= 1;
            for (int c = row + half * 128; c < tl.bn; c += 32 * EPI_WARPS) sbias[bias_buf][c] = n0 + c < a.cout ? __ldg(a.bias + n0 + c) : 0.f;
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
        }
        mbar_wait(smem_u32(&acc_full[as]), aphase, 4);
        tc_fence_after();
        for (int ch = half; ch < nchunks; ch += 2 * HALVES) {
            process(ch, bA);
            aux_load(ch + 2 * HALVES, bA);
            if (ch + HALVES < nchunks) { process(ch + HALVES, bB); aux_load(ch + 3 * HALVES, bB); }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            if (PAIR) mbar_arrive_cta0(smem_u32(&acc_empty[as]));
            else mbar_arrive(smem_u32(&acc_empty[as]));
        }
        if (AUX32 && tl.dual) aphase ^= 1; // ONE accumulator stage (both TMEM halves belong to the tile in flight)
        else if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (a.stat_mode && cur_prob >= 0) flush_stats(cur_prob, cur_nt);
}

// ---------------------------------------------------------------------------------------------------
// convolution kernel
// ---------------------------------------------------------------------------------------------------
// EPI_WARPS = 4: one epilogue warp per TMEM lane quarter (long-K layers: the MMA main loop hides the epilogue; 192 threads leave
// 255 registers per thread).  EPI_WARPS = 8: two warps per quarter interleave the 16-column chunks (short-K layers whose
// epilogue -- residual / mask loads, statistics -- would otherwise be the critical path).
template <int EPI_WARPS, bool AUX32 = false>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 1)
k_conv_tc(const __grid_constant__ Maps maps, const __grid_constant__ ConvArgs a, const __grid_constant__ Tiling tl) {
    constexpr int NTHREADS = 64 + 32 * EPI_WARPS;
    extern __shared__ uint8_t smem_raw[];
    __shared__ Smem s;
    // per epilogue warp: no atomics (fp32 smem atomics are CAS loops); with eight warps each owns every second 16-column chunk, so the
    // arrays are half as long (8 KB either way: the halo kernel's weight ring gets the same three 48 KB slots in both instantiations)
    __shared__ float sstat[EPI_WARPS][2][ACC_COLS / (EPI_WARPS / 4)];
    __shared__ __align__(16) float sbias[2][ACC_COLS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4 * 2 * ACC_COLS; i += NTHREADS) (&sstat[0][0][0])[i] = 0.f;
    const uint32_t tiles0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t tmem_base = setup(s, maps, EPI_WARPS, warp, lane);
    const int iters = tl.npairs * a.ntaps * tl.kchunks;
    const int nstages = tl.nstages; const uint32_t stage_bytes = tl.stage_bytes;
    pdl_launch();
    if (warp != 1) pdl_wait();        // the producer and epilogue warps touch global memory; the MMA warp does not

    if (warp == 0) {
        {
            int stage = 0; uint32_t phase = 0;
            long long w_e = 0, w_l = 0; const bool dbg_on = tl.dbg != nullptr; const long long t_start = clock64();
            const uint32_t tx_bytes = A_BYTES + tl.bn * 128;
            for (int tile = blockIdx.x; tile < tl.total; tile += gridDim.x) {
                int ti = 0;
                while (tile >= tl.p[ti + 1].begin) ti++;
                const TileP& tp_ = tl.p[ti];
                int local = tile - tp_.begin;
                int nt = local % tl.n_tiles, mt = local / tl.n_tiles;
                int tiles_per_img = tp_.tiles_x * tp_.tiles_y;
                int img = mt / tiles_per_img, r = mt % tiles_per_img;
                int y0 = (r / tp_.tiles_x) * tp_.TH, x0 = (r % tp_.tiles_x) * tp_.TW;
                const int n0 = a.nphase ? 0 : nt * tl.bn, sl0 = a.nphase ? 9 * nt : 0;
                const int nimg = a.p[tp_.prob].N;
                for (int pr = 0; pr < tl.npairs; pr++) {      // plane pairs of split-precision operands (one pass for plain bf16)
                    const int img_p = img + tl.pa[pr] * nimg, sl_p = sl0 + tl.pb[pr] * tl.nslab_total;
                    for (int tp = 0; tp < a.ntaps; tp++) {
                        const Tap t = a.taps[tp];
                        const CUtensorMap* amap = &maps.a[tp_.prob][t.view];
                        for (int kc_ = 0; kc_ < tl.kchunks; kc_++) {
                            mbar_wait_t(smem_u32(&s.empty[stage]), phase ^ 1, 1, w_e, dbg_on);
                            uint32_t fb = smem_u32(&s.full[stage]);
                            uint32_t sa = tiles0 + stage * stage_bytes;
                            if (elect_one()) {
                                mbar_expect_tx(fb, tx_bytes);
                                tma_load_4d(amap, fb, sa, kc_ * 64, x0 + t.dx, y0 + t.dy, img_p);
                                tma_load_3d(&maps.b, fb, sa + A_BYTES, kc_ * 64, n0, sl_p + t.slab);
                            }
                            __syncwarp();
                            if (++stage == nstages) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
            if (dbg_on && lane == 0) { long long* d = tl.dbg + blockIdx.x * 8; d[0] = w_l; d[1] = w_e; d[2] = clock64() - t_start; }
        }
    } else if (warp == 1) {
        {
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            long long w_f = 0, w_acc = 0; const bool dbg_on = tl.dbg != nullptr; const long long t_start = clock64();
            const uint32_t idesc = make_idesc(tl.bn, 0, 0);
            const bool dual = AUX32 && tl.dual;
            for (int tile = blockIdx.x; tile < tl.total; tile += gridDim.x) {
                mbar_wait_t(smem_u32(&s.acc_empty[as]), aphase ^ 1, 2, w_acc, dbg_on);
                tc_fence_after();
                // DUAL_ACC (split mode): the tensor cores add into the fp32 accumulator with TRUNCATION, a bias of ~3e-8 of the
                // running sum per MMA that grows linearly with the length of the accumulation chain (measured: 1e-5 at K = 9216).
                // Alternating the K steps between the two 256-column accumulators halves every chain and decorrelates the bias; the
                // epilogue adds the halves in registers (round-to-nearest).  Costs the epilogue / main-loop overlap, which a six-pass
                // main loop does not need.
                for (int it = 0; it < iters; it++) {
                    const uint32_t d_tmem = tmem_base + (dual ? (it & 1) : as) * ACC_COLS;
                    mbar_wait_t(smem_u32(&s.full[stage]), phase, 3, w_f, dbg_on);
                    tc_fence_after();
                    uint32_t sa = tiles0 + stage * stage_bytes;
                    uint64_t ad = make_desc(sa, 16, 1024), bd = make_desc(sa + A_BYTES, 16, 1024);
                    const int first = dual ? (it >> 1) : it;        // 0 on the first K step into this accumulator
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; k++) umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (first | k) != 0);
                        umma_commit(smem_u32(&s.empty[stage]));
                    }
                    __syncwarp();
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma_commit(smem_u32(&s.acc_full[as]));
                __syncwarp();
                if (dual) aphase ^= 1;
                else if (++as == 2) { as = 0; aphase ^= 1; }
            }
            if (dbg_on && lane == 0) { long long* d = tl.dbg + blockIdx.x * 8; d[4] = w_f; d[5] = w_acc; d[6] = clock64() - t_start; }
        }
    } else {
        conv_epilogue<EPI_WARPS, false, AUX32>(a, tl, s.acc_full, s.acc_empty, sstat, sbias, tmem_base, warp, lane);
    }
    teardown(tmem_base, warp);
}

// ---------------------------------------------------------------------------------------------------
// halo-tile convolution kernel for standard 3x3 tap sets: the M tile is a 16 x 8 patch (8 pixels wide, so one 8-row swizzle atom of
// the A operand = one patch row).  Per (view, 64-channel chunk) the producer fetches the 18 x 10 HALO of the patch ONCE and all nine
// taps read it in place: tap (dy, dx) is the SAME shared-memory tile addressed through a UMMA descriptor whose start address is
// shifted by ((dy+1)*10 + (dx+1)) pixel rows of 128 B and whose 8-row-group stride (SBO) is 10 rows = 1280 B.  TMA and UMMA both
// apply the 128-byte swizzle to absolute shared-memory address bits, so the shifted views stay consistent (checked against the
// CUDA-core engine by tests/test_gpu_conv.py).  L2 -> SM traffic of the A operand drops ~6.4x (180 instead of 9 x 128 pixel rows per
// chunk); the weight tiles stream through their own, deeper ring.
//
// PAIR = true: the two CTAs of a cluster (the two SMs of a TPC) run ONE tcgen05.mma.cta_group::2 stream, M = 256 = two patches (one
// per CTA), N = bn.  Each CTA stages its own halo and only HALF of the weight tile (bn/2 rows); the tensor cores read the other half
// from the peer's shared memory.  Per SM that halves the weight bytes read from L2 and the shared-memory fill, cuts the operand reads
// per MMA from 12 KB to 8 KB, and gives the B ring three slots of a kernel ROW each (three taps: 48 KB at N = 256; see Tiling.tb).  Protocol (as CUTLASS' 2-SM pipelines): both producers wait on
// their OWN empty barriers and issue cta_group::2 TMA loads that complete bytes on the LEADER's full barrier (the leader's producer
// posts the expected bytes of both CTAs); the leader's MMA thread issues the MMAs and multicasts its commits to the empty /
// accumulator-full barriers of both CTAs; both CTAs' epilogue warps arrive on the leader's accumulator-empty barrier.
//
// The single producer / MMA threads are the serial resources of the design: their per-tap loops are fully unrolled over the nine taps
// with compile-time operand offsets and touch no dynamically indexed kernel parameter (a dependent LDC per tap cost ~300 cycles per
// tap in the first version, measured with per-thread stall counters).
// ---------------------------------------------------------------------------------------------------
// Ring depths (build-time overridable for A/B runs: AFI_EXTRA_NVCC_FLAGS="-DAFI_HALO_SA=4 -DAFI_HALO_SB_MAX=7").  Measured on the discriminator
// shapes, kernel alone (tools/ab_conv.py): two halo slots cost 11 % (1700 -> 1515 TFLOP/s on 1024->1024: one chunk of look-ahead does not cover
// the latency of the 180-row halo box), four or five are within noise of three, and 8 -> 11 (or 18, for 8 KB tiles) weight slots change nothing:
// what the weight ring was short of was not depth but the ~300 cycles the single producer / MMA threads spend per SLOT -- hence three taps per
// slot (conv_tc(): Tiling.tb).
#ifndef AFI_HALO_SA
#define AFI_HALO_SA 3
#endif
#ifndef AFI_HALO_SB_MAX
#define AFI_HALO_SB_MAX 8
#endif
constexpr int HALO_SA = AFI_HALO_SA, HALO_SB_MAX = AFI_HALO_SB_MAX;
constexpr int HALO_TH = 16, HALO_TW = 8;
constexpr int HALO_BYTES = (HALO_TH + 2) * (HALO_TW + 2) * 128;           // 23040
constexpr int HALO_SLOT = (HALO_BYTES + 1023) / 1024 * 1024;              // 23552
struct SmemH {
    uint64_t a_full[HALO_SA], a_empty[HALO_SA];
    uint64_t b_full[HALO_SB_MAX], b_empty[HALO_SB_MAX];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};
struct HaloTile { int prob, img, c1, c2, n0, sl0, nimg; };      // c1 / c2: TMA coordinates of the halo's first pixel along its fast / slow axis

template <int EPI_WARPS, bool PAIR, bool AUX32 = false>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 1)
k_conv_halo(const __grid_constant__ Maps maps, const __grid_constant__ ConvArgs a, const __grid_constant__ Tiling tl) {
    constexpr int NTHREADS = 64 + 32 * EPI_WARPS;
    extern __shared__ uint8_t smem_raw[];
    __shared__ SmemH s;
    __shared__ float sstat[EPI_WARPS][2][ACC_COLS / (EPI_WARPS / 4)];
    __shared__ __align__(16) float sbias[2][ACC_COLS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = PAIR ? (int)cluster_ctarank() : 0;
    const bool leader = rank == 0;
    if (threadIdx.x == 0) AFI_TSTAMP(0);
    for (int i = threadIdx.x; i < 4 * 2 * ACC_COLS; i += NTHREADS) (&sstat[0][0][0])[i] = 0.f;
    const uint32_t a0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b0 = a0 + HALO_SA * HALO_SLOT;
    const int sb = tl.sb;
    const uint32_t b_slot = tl.b_slot;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.b);
        prefetch_tmap(&maps.a[tl.p[0].prob][0]);
        for (int i = 0; i < HALO_SA; i++) { mbar_init(smem_u32(&s.a_full[i]), 1); mbar_init(smem_u32(&s.a_empty[i]), 1); }
        for (int i = 0; i < sb; i++) { mbar_init(smem_u32(&s.b_full[i]), 1); mbar_init(smem_u32(&s.b_empty[i]), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(smem_u32(&s.acc_full[i]), 1); mbar_init(smem_u32(&s.acc_empty[i]), (PAIR ? 2 : 1) * EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();      // the peer's barriers are initialised before anything signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s.tmem_base;
    if (threadIdx.x == 0) AFI_TSTAMP(1);
    pdl_launch();
    if (warp != 1) pdl_wait();        // the producer and epilogue warps touch global memory; the MMA warp does not
    if (threadIdx.x == 0) AFI_TSTAMP(2);
    const int kchunks = tl.kchunks;
    const int nchunks = tl.npairs * tl.nviews * kchunks;
    const int item0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, item_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int total = tl.total;
    const bool dbg_on = tl.dbg != nullptr;

    if (warp == 0) {
        {
            long long w_ae = 0, w_be = 0; const long long t_start = clock64();
            int ai = 0; uint32_t aph = 0; int bi = 0; uint32_t bph = 0;
            const int tb = tl.tb;
            const uint32_t b_tx = (uint32_t)tl.bn * 128u * (uint32_t)tb;   // bytes of a whole weight slot (both halves in pair mode)
            const int brow = PAIR ? rank * (tl.bn >> 1) : 0;
            const uint32_t bar_af = smem_u32(&s.a_full[0]), bar_ae = smem_u32(&s.a_empty[0]);
            const uint32_t bar_bf = smem_u32(&s.b_full[0]), bar_be = smem_u32(&s.b_empty[0]);
            auto decode = [&](int item, HaloTile& h) {
                int ti = 0;
                while (item >= tl.p[ti + 1].begin) ti++;
                const TileP& tp_ = tl.p[ti];
                const int local = item - tp_.begin;
                const int nt = local % tl.n_tiles;
                const int tiles_per_img = tp_.tiles_x * tp_.tiles_y;
                int mt = local / tl.n_tiles;
                if (PAIR) {
                    const int m_tiles = a.p[tp_.prob].N * tiles_per_img;
                    mt = 2 * mt + rank;
                    if (mt >= m_tiles) mt = m_tiles - 1;     // odd tile count: this CTA recomputes the last tile (its epilogue is masked)
                }
                const int r = mt % tiles_per_img;
                h.prob = tp_.prob; h.img = mt / tiles_per_img; h.nimg = a.p[tp_.prob].N;
                const int y0 = (r / tp_.tiles_x) * tp_.TH - 1, x0 = (r % tp_.tiles_x) * tp_.TW - 1;
                h.c1 = tp_.orient ? y0 : x0; h.c2 = tp_.orient ? x0 : y0;
                h.n0 = a.nphase ? 0 : nt * tl.bn; h.sl0 = a.nphase ? 9 * nt : 0;
            };
            auto issue_a = [&](const HaloTile& h, int view, int kc, int plane) {
                mbar_wait_t(bar_ae + 8 * ai, aph ^ 1, 21, w_ae, dbg_on);
                const uint32_t fb = bar_af + 8 * ai;
                const uint32_t dst = a0 + ai * HALO_SLOT;
                const CUtensorMap* amap = &maps.a[h.prob][view];
                if (elect_one()) {
                    if (PAIR) {
                        if (leader) mbar_expect_tx(fb, 2 * HALO_BYTES);
                        tma_load_4d_pair(amap, fb, dst, kc * 64, h.c1, h.c2, h.img + plane * h.nimg);
                    } else {
                        mbar_expect_tx(fb, HALO_BYTES);
                        tma_load_4d(amap, fb, dst, kc * 64, h.c1, h.c2, h.img + plane * h.nimg);
                    }
                }
                __syncwarp();
                if (++ai == HALO_SA) { ai = 0; aph ^= 1; }
            };
            HaloTile cur, nxt;
            int item = item0;
            if (item < total) { decode(item, cur); issue_a(cur, 0, 0, tl.pa[0]); }
            while (item < total) {
                const int nitem = item + item_stride;
                const bool has_next = nitem < total;
                if (has_next) decode(nitem, nxt);
                int pr = 0, view = 0, kc = 0;       // (plane pair of split-precision operands, input view, 64-channel chunk)
                for (int chunk = 0; chunk < nchunks; chunk++) {
                    const int slab0 = tl.view_slab0[view] + cur.sl0 + tl.pb[pr] * tl.nslab_total;
                    const int kcol = kc * 64, nrow = cur.n0 + brow;
                    int npr = pr, nview = view, nkc = kc + 1;
                    if (nkc == kchunks) { nkc = 0; if (++nview == tl.nviews) { nview = 0; if (++npr == tl.npairs) npr = 0; } }
                    // the halo of the NEXT chunk is requested before this chunk's weight tiles (three A slots: its slot was released long ago)
                    if (chunk + 1 < nchunks) issue_a(cur, nview, nkc, tl.pa[npr]);
                    else if (has_next) issue_a(nxt, 0, 0, tl.pa[0]);
                    for (int j = 0; j < 9; j += tb) {       // one box of tb taps per slot
                        mbar_wait_t(bar_be + 8 * bi, bph ^ 1, 22, w_be, dbg_on);
                        const uint32_t fb = bar_bf + 8 * bi;
                        const uint32_t dst = b0 + bi * b_slot;
                        if (elect_one()) {
                            if (PAIR) {
                                if (leader) mbar_expect_tx(fb, b_tx);
                                tma_load_3d_pair(&maps.b, fb, dst, kcol, nrow, slab0 + j);
                            } else {
                                mbar_expect_tx(fb, b_tx);
                                tma_load_3d(&maps.b, fb, dst, kcol, nrow, slab0 + j);
                            }
                        }
                        __syncwarp();
                        if (++bi == sb) { bi = 0; bph ^= 1; }
                    }
                    pr = npr; view = nview; kc = nkc;
                }
                cur = nxt; item = nitem;
            }
            if (dbg_on && lane == 0) { long long* d = tl.dbg + blockIdx.x * 8; d[0] = w_ae; d[1] = w_be; d[2] = clock64() - t_start; }
            if (PAIR) {
                // producer tail: every release the leader multicast to this CTA has landed before the CTA may exit
                for (int i = 0; i < HALO_SA; i++) { mbar_wait(bar_ae + 8 * ai, aph ^ 1, 33); if (++ai == HALO_SA) { ai = 0; aph ^= 1; } }
                for (int i = 0; i < sb; i++) { mbar_wait(bar_be + 8 * bi, bph ^ 1, 34); if (++bi == sb) { bi = 0; bph ^= 1; } }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            long long w_af = 0, w_bf = 0, w_acc = 0; const long long t_start = clock64();
            int ai = 0; uint32_t aph = 0; int bi = 0; uint32_t bph = 0;
            int as = 0; uint32_t aphase = 0;
            const uint32_t idesc = make_idesc(tl.bn, 0, 0, PAIR ? 256 : 128);
            const uint64_t adesc0 = make_desc(0, 16, 10 * 128), bdesc0 = make_desc(0, 16, 1024);
            const uint32_t bar_af = smem_u32(&s.a_full[0]), bar_ae = smem_u32(&s.a_empty[0]);
            const uint32_t bar_bf = smem_u32(&s.b_full[0]), bar_be = smem_u32(&s.b_empty[0]);
            const int cin = a.cin;
            const bool dual = AUX32 && tl.dual;
            // taps that open / close a weight slot (tb taps per slot): tb = 1 -> every tap, 3 -> taps 0, 3, 6 / 2, 5, 8, 9 -> tap 0 / 8
            const uint32_t slot_first = tl.tb == 1 ? 0x1FFu : (tl.tb == 3 ? 0x49u : 0x1u);
            const uint32_t slot_last = tl.tb == 1 ? 0x1FFu : (tl.tb == 3 ? 0x124u : 0x100u);
            const uint32_t b_tile = (uint32_t)tl.b_slot / (uint32_t)tl.tb;
            uint32_t bbase = b0;
            auto commit = [&](uint32_t bar) { if (PAIR) umma_commit_pair(bar); else umma_commit(bar); };
            int ti = 0;
            for (int item = item0; item < total; item += item_stride) {
                while (item >= tl.p[ti + 1].begin) ti++;
                const bool orient = tl.p[ti].orient != 0;
                mbar_wait_t(smem_u32(&s.acc_empty[as]), aphase ^ 1, 23, w_acc, dbg_on);
                tc_fence_after();
                uint32_t accum = 0;
                int kc = 0;
                for (int chunk = 0; chunk < nchunks; chunk++) {
                    // split mode: chunks alternate between the two accumulators (DUAL_ACC, see k_conv_tc); nchunks >= 6 there
                    const uint32_t d_tmem = tmem_base + (dual ? (chunk & 1) : as) * ACC_COLS;
                    if (dual) accum = chunk >= 2;
                    int nk = (cin - kc * 64 + 15) >> 4;              // 16-channel MMA steps with data in this chunk
                    if (++kc == kchunks) kc = 0;
                    mbar_wait_t(bar_af + 8 * ai, aph, 24, w_af, dbg_on);
                    if (item == item0 && chunk == 0 && lane == 0) AFI_TSTAMP(3);
                    const uint32_t abase = a0 + ai * HALO_SLOT;
#pragma unroll
                    for (int j = 0; j < 9; j++) {
                        // tap j = (dy, dx) = (j/3 - 1, j%3 - 1): shift of dy halo lines + dx pixels (orient 0) or dx lines + dy pixels (orient 1)
                        const uint32_t aaddr = abase + (orient ? (uint32_t)(((j % 3) * 10 + (j / 3)) * 128) : (uint32_t)(((j / 3) * 10 + (j % 3)) * 128));
                        if ((slot_first >> j) & 1u) {             // tap j opens a weight slot
                            mbar_wait_t(bar_bf + 8 * bi, bph, 25, w_bf, dbg_on);
                            tc_fence_after();
                            bbase = b0 + bi * b_slot;
                        }
                        const uint32_t baddr = bbase;
                        bbase += b_tile;
                        const uint64_t ad = adesc0 | (uint64_t)((aaddr >> 4) & 0x3FFFu), bd = bdesc0 | (uint64_t)((baddr >> 4) & 0x3FFFu);
                        if (elect_one()) {
                        if (nk >= 4) {
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                if (PAIR) umma_bf16_pair(d_tmem, ad + 2 * k, bd + 2 * k, idesc, accum | (uint32_t)(j | k));
                                else umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, accum | (uint32_t)(j | k));
                            }
                        } else {
                            for (int k = 0; k < nk; k++) {
                                if (PAIR) umma_bf16_pair(d_tmem, ad + 2 * k, bd + 2 * k, idesc, accum | (uint32_t)(j | k));
                                else umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, accum | (uint32_t)(j | k));
                            }
                        }
                        if ((slot_last >> j) & 1u) commit(bar_be + 8 * bi);
                        }
                        __syncwarp();
                        if ((slot_last >> j) & 1u) { if (++bi == sb) { bi = 0; bph ^= 1; } }
                    }
                    accum = 1;
                    if (elect_one()) commit(bar_ae + 8 * ai);
                    __syncwarp();
                    if (++ai == HALO_SA) { ai = 0; aph ^= 1; }
                }
                if (elect_one()) commit(smem_u32(&s.acc_full[as]));
                __syncwarp();
                if (lane == 0) AFI_TSTAMP(4);
                if (dual) aphase ^= 1;
                else if (++as == 2) { as = 0; aphase ^= 1; }
            }
            if (dbg_on && lane == 0) { long long* d = tl.dbg + blockIdx.x * 8; d[3] = w_af; d[4] = w_bf; d[5] = w_acc; d[6] = clock64() - t_start; }
        }
    } else {
        conv_epilogue<EPI_WARPS, PAIR, AUX32>(a, tl, s.acc_full, s.acc_empty, sstat, sbias, tmem_base, warp, lane, rank);
        if (warp == 2 && lane == 0) AFI_TSTAMP(6);
    }
    tc_fence_before();
    __syncwarp();
    if (threadIdx.x == 0) AFI_TSTAMP(5);
    if (PAIR) cluster_sync_all();      // neither CTA frees TMEM / exits while the other may still read its shared memory or signal it
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        if (lane == 0) AFI_TSTAMP(7);
    }
}

// ---------------------------------------------------------------------------------------------------
// weight-gradient kernel.  Work item = (tap, cout tile of 128, cin tile of bn, k split); K tile = TH x TW = 64 pixels.
// smem stage: A = two [64 px][64 co] SW128 boxes (8 KB each), B = bn/64 [64 px][64 ci] boxes.
// ---------------------------------------------------------------------------------------------------
// PAIR: the two CTAs of a cluster run one cta_group::2 MMA stream with M = 256 couts (128 per CTA) x N = bn cins: each CTA stages its own
// dY tile and HALF of the X tile (bn/2 channels), i.e. 32 KB instead of 48 KB per 64-pixel K tile, and the ring gets six stages.  Same
// barrier protocol as the pair convolution kernel.
template <bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_wgrad_tc(const __grid_constant__ Maps maps, const __grid_constant__ WgradArgs a, const __grid_constant__ Tiling tl) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Smem s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = PAIR ? (int)cluster_ctarank() : 0;
    const bool leader = rank == 0;
    const uint32_t tiles0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < STAGES_MAX; i++) { mbar_init(smem_u32(&s.full[i]), 1); mbar_init(smem_u32(&s.empty[i]), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(smem_u32(&s.acc_full[i]), 1); mbar_init(smem_u32(&s.acc_empty[i]), (PAIR ? 2 : 1) * (EPI_THREADS / 32)); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s.tmem_base;
    const int nb = tl.bn / 64;                      // 64-channel X boxes per K tile (both halves together in pair mode)
    const int nbl = PAIR ? nb / 2 : nb;             // ... staged by THIS CTA
    const int nstages = tl.nstages; const uint32_t stage_bytes = tl.stage_bytes;
    const int kper = (tl.ktiles + tl.ksplit - 1) / tl.ksplit;
    const int item0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, item_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const bool dual = tl.dual != 0;       // split mode: both TMEM halves accumulate the item in flight
    // work item = (k split, tap, cout tile, cin tile) with the K SPLIT SLOWEST: the CTAs that run concurrently stream the same range of
    // pixel tiles (different taps / channel tiles), so each operand tile is read from HBM once per wave and from L2 by everyone else
    const int kbase = tl.total / tl.ksplit;

    // the producer and MMA warps run their loops CONVERGED; only the TMA / tcgen05 instruction issue is elected (see elect_one())
    if (warp == 0) {
        int stage = 0; uint32_t phase = 0;
        const uint32_t tx_bytes = (PAIR ? 2u : 1u) * (A_BYTES + nbl * 8192);      // bytes of both CTAs land on the leader's barrier
        for (int item = item0; item < tl.total; item += item_stride) {
            int ks = item / kbase; int r = item % kbase;
            int nt = r % tl.n_tiles; r /= tl.n_tiles;
            int mt = r % tl.m_tiles; int tp = r / tl.m_tiles;
            if (PAIR) mt = 2 * mt + rank;               // m_tiles counts PAIRS of 128-cout tiles
            const Tap t = a.taps[tp];
            int k0 = ks * kper, k1 = min(k0 + kper, tl.ktiles);
            // position of K tile k0: (problem, image, patch row, patch column); advanced incrementally below -- integer divisions
            // per stage in this serial loop were eating the 512-cycle stage budget
            int ti0 = 0;
            while (k0 >= tl.p[ti0 + 1].begin) ti0++;
            const int lk0 = k0 - tl.p[ti0].begin;
            const int tpi = tl.p[ti0].tiles_x * tl.p[ti0].tiles_y;
            const int img0 = lk0 / tpi, rr0 = lk0 % tpi;
            for (int pr = 0; pr < tl.npairs; pr++) {     // plane pairs of split-precision operands walk the same K range (one pass for bf16)
                const int pa = tl.pa[pr], pb = tl.pb[pr];
                int ti = ti0, img = img0;
                int ty_ = rr0 / tl.p[ti].tiles_x, tx_ = rr0 % tl.p[ti].tiles_x;
                for (int kt = k0; kt < k1; kt++) {
                    const TileP& tp_ = tl.p[ti];
                    const int y0 = ty_ * tp_.TH, x0 = tx_ * tp_.TW;
                    const int nimg = a.p[tp_.prob].N;
                    mbar_wait(smem_u32(&s.empty[stage]), phase ^ 1, 11);
                    uint32_t fb = smem_u32(&s.full[stage]);
                    uint32_t sa = tiles0 + stage * stage_bytes;
                    if (elect_one()) {
                        // 5-D maps {64 ch, W, H, channel group, N}: ONE box lands as consecutive [64 px][128 B] SW128 blocks per group
                        if (PAIR) {
                            if (leader) mbar_expect_tx(fb, tx_bytes);
                            tma_load_5d_pair(&maps.a[tp_.prob][0], fb, sa, 0, x0, y0, mt * 2, img + pa * nimg);
                            tma_load_5d_pair(&maps.a[tp_.prob][1], fb, sa + A_BYTES, 0, x0 + t.dx, y0 + t.dy, nt * nb + rank * nbl, img + pb * nimg);
                        } else {
                            mbar_expect_tx(fb, tx_bytes);
                            tma_load_5d(&maps.a[tp_.prob][0], fb, sa, 0, x0, y0, mt * 2, img + pa * nimg);
                            tma_load_5d(&maps.a[tp_.prob][1], fb, sa + A_BYTES, 0, x0 + t.dx, y0 + t.dy, nt * nb, img + pb * nimg);
                        }
                    }
                    __syncwarp();
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                    if (++tx_ == tp_.tiles_x) {
                        tx_ = 0;
                        if (++ty_ == tp_.tiles_y) {
                            ty_ = 0;
                            if (kt + 1 >= tl.p[ti + 1].begin) { ti++; img = 0; } else img++;
                        }
                    }
                }
            }
        }
        if (PAIR) {     // producer tail: every release the leader multicast to this CTA has landed before the CTA may exit
            for (int i = 0; i < nstages; i++) { mbar_wait(smem_u32(&s.empty[stage]), phase ^ 1, 15); if (++stage == nstages) { stage = 0; phase ^= 1; } }
        }
    } else if (warp == 1) {
        if (leader) {
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            const uint32_t idesc = make_idesc(tl.bn, 1, 1, PAIR ? 256 : 128);
            for (int item = item0; item < tl.total; item += item_stride) {
                int ks = item / kbase;
                int k0 = ks * kper, k1 = min(k0 + kper, tl.ktiles);
                mbar_wait(smem_u32(&s.acc_empty[as]), aphase ^ 1, 12);
                tc_fence_after();
                const int ksteps = tl.npairs * (k1 - k0);        // split-precision operands: the K range once per plane pair
                for (int kt = 0; kt < ksteps; kt++) {
                    // split mode: K steps alternate between the two accumulators (DUAL_ACC, see k_conv_tc)
                    const uint32_t d_tmem = tmem_base + (dual ? (kt & 1) : as) * ACC_COLS;
                    const int kfirst = dual ? (kt >> 1) : kt;
                    mbar_wait(smem_u32(&s.full[stage]), phase, 13);
                    tc_fence_after();
                    uint32_t sa = tiles0 + stage * stage_bytes;
                    // MN-major SW128: LBO = stride between 64-channel groups (8192 B), SBO = stride between 8-pixel groups (1024 B)
                    uint64_t ad = make_desc(sa, 8192, 1024), bd = make_desc(sa + A_BYTES, 8192, 1024);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; k++) {     // 16 pixels = 2048 B per MMA
                            if (PAIR) umma_bf16_pair(d_tmem, ad + (uint64_t)(128 * k), bd + (uint64_t)(128 * k), idesc, (kfirst > 0) || (k != 0));
                            else umma_bf16(d_tmem, ad + (uint64_t)(128 * k), bd + (uint64_t)(128 * k), idesc, (kfirst > 0) || (k != 0));
                        }
                        if (PAIR) umma_commit_pair(smem_u32(&s.empty[stage])); else umma_commit(smem_u32(&s.empty[stage]));
                    }
                    __syncwarp();
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) { if (PAIR) umma_commit_pair(smem_u32(&s.acc_full[as])); else umma_commit(smem_u32(&s.acc_full[as])); }
                __syncwarp();
                if (dual) aphase ^= 1;
                else if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;       // the two warps of a lane quarter interleave the 16-column chunks
        const int row = q * 32 + lane;
        int as = 0; uint32_t aphase = 0;
        for (int item = item0; item < tl.total; item += item_stride) {
            int ks = item / kbase; int r = item % kbase;
            int nt = r % tl.n_tiles; r /= tl.n_tiles;
            int mt = r % tl.m_tiles; int tp = r / tl.m_tiles;
            if (PAIR) mt = 2 * mt + rank;
            int k0 = ks * kper, k1 = min(k0 + kper, tl.ktiles);
            const int co = mt * 128 + row;
            float* dst = a.dw + ((long long)a.taps[tp].slab * a.cout + co) * a.cin;
            mbar_wait(smem_u32(&s.acc_full[as]), aphase, 14);
            tc_fence_after();
            const uint32_t taddr = tmem_base + as * ACC_COLS + ((uint32_t)(q * 32) << 16);
            for (int c0 = half * 16; c0 < tl.bn; c0 += 32) {
                float v[16];
                tmem_ld16(taddr + c0, v);
                if (dual && k1 > k0) {
                    float v2[16];
                    tmem_ld16(taddr + ACC_COLS + c0, v2);
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] += v2[i];
                }
                const int ci = nt * tl.bn + c0;
                if (k1 > k0 && co < a.cout && ci < a.cin) {      // (cin is a multiple of 16: the whole chunk is inside the row, 16-byte aligned)
#pragma unroll
                    for (int i = 0; i < 16; i += 4)             // one vector reduction per four columns instead of four scalar ones
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + ci + i), "f"(v[i]), "f"(v[i + 1]), "f"(v[i + 2]),
                                     "f"(v[i + 3]) : "memory");
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_cta0(smem_u32(&s.acc_empty[as])); else mbar_arrive(smem_u32(&s.acc_empty[as])); }
            if (dual) aphase ^= 1;
            else if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }
    tc_fence_before();
    __syncwarp();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- host side -------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_map(afi_ctx* ctx, CUtensorMap* m, void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                      const cuuint32_t* box) {
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    AFI_REQUIRE(((uintptr_t)ptr & 15) == 0, "TMA: global address not 16-byte aligned");
    for (int i = 0; i < rank - 1; i++) AFI_REQUIRE(strides_bytes[i] % 16 == 0, "TMA: stride %d (%llu B) not a multiple of 16", i, (unsigned long long)strides_bytes[i]);
    CUresult r = ((EncodeTiledFn)ctx->encode_tiled)(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, ptr, dims, strides_bytes, box, es,
                                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return AFI_ERR_CUDA; }
    return AFI_OK;
}
// halo box {64 ch, 10, 18, 1}; orient 1 swaps the roles of x and y (the 10-pixel axis, along which the 8-pixel swizzle groups run, is y)
// (N counts images x operand planes: split-precision views stack their three bf16 planes along the image axis)
static int encode_view_halo(afi_ctx* ctx, CUtensorMap* m, const PView& v, int C, int W, int H, int N, int orient) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)(orient ? H : W), (cuuint64_t)(orient ? W : H), (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)(orient ? v.sy : v.sx) * 2, (cuuint64_t)(orient ? v.sx : v.sy) * 2, (cuuint64_t)v.sn * 2};
    cuuint32_t box[4] = {64, 10, 18, 1};
    return encode_map(ctx, m, v.ptr, 4, dims, strides, box);
}
static int encode_view(afi_ctx* ctx, CUtensorMap* m, const PView& v, int C, int W, int H, int N, int TW, int TH) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)v.sx * 2, (cuuint64_t)v.sy * 2, (cuuint64_t)v.sn * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, 1};
    return encode_map(ctx, m, v.ptr, 4, dims, strides, box);
}

// channel-grouped 5-D view for the MN-major operands of the weight-gradient GEMM
static int encode_view_grouped(afi_ctx* ctx, CUtensorMap* m, const PView& v, int C, int W, int H, int N, int TW, int TH, int gbox) {
    int G = (C + 63) / 64;
    AFI_REQUIRE(C <= 64 || (long long)G * 64 <= v.sx, "wgrad_tc: %d channels need %d groups of 64 but the pixel stride is %lld", C, G, v.sx);
    cuuint64_t dims[5] = {(cuuint64_t)(C < 64 ? C : 64), (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)G, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)v.sx * 2, (cuuint64_t)v.sy * 2, 128, (cuuint64_t)v.sn * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)gbox, 1};
    return encode_map(ctx, m, v.ptr, 5, dims, strides, box);
}

static void pick_patch(int H, int W, int pixels, int* TH, int* TW) {
    // spatial patch of `pixels` (= GEMM rows per tile) minimising padded area; TW multiple of 8 keeps swizzle atoms whole
    long long best = -1; int bh = pixels / 16, bw = 16;
    for (int tw = 8; tw <= pixels; tw *= 2) {
        int th = pixels / tw;
        if (th < 1 || tw > 256 || th > 256) continue;
        long long padded = (long long)((H + th - 1) / th) * th * (long long)((W + tw - 1) / tw) * tw;
        if (best < 0 || padded < best || (padded == best && tw == 16)) { best = padded; bh = th; bw = tw; }
    }
    *TH = bh; *TW = bw;
}

static int g_halo_dyn_max[2] = {0, 0};
// plane pairs of a split-precision product, smallest terms first (hi lo, lo hi, mid mid, hi mid, mid hi, hi hi): every pair whose
// weight is >= 2^-16 of the leading term; the dropped ones (mid lo, lo mid, lo lo) are below fp32's own rounding
// (split_pairs = 3: only hi mid, mid hi, hi hi -- products accurate to ~2^-17, for everything a result depends on linearly)
static void set_pairs(Tiling& tl, int split, int split_pairs, int nslab_total) {
    static const int PA[6] = {0, 2, 1, 0, 1, 0}, PB[6] = {2, 0, 1, 1, 0, 0};
    tl.nslab_total = nslab_total;
    if (split == 3) {
        const int first = split_pairs == 3 ? 3 : 0;
        tl.npairs = 6 - first;
        for (int i = first; i < 6; i++) { tl.pa[i - first] = PA[i]; tl.pb[i - first] = PB[i]; }
    } else { tl.npairs = 1; tl.pa[0] = tl.pb[0] = 0; }
}
// AFIGAN_CONV_HALO = 0: per-tap A tiles (k_conv_tc) everywhere; 1: halo tiles, one CTA per tile; 2: halo tiles on CTA pairs (cta_group::2)
static int halo_mode() {      // read on every call (tests switch variants in-process)
    const char* e = getenv("AFIGAN_CONV_HALO");
    int mode = e ? atoi(e) : 2;
    if (mode < 0 || mode > 2) mode = 2;
    return mode;
}
// the halo kernels take standard tap sets only: per view the nine taps (dy, dx) = (-1,-1) .. (1,1) in row-major order on consecutive slabs
static bool halo_eligible(const ConvArgs& a) {
    if (a.ntaps < 9 || a.ntaps % 9 != 0) return false;
    for (int i = 0; i < a.ntaps; i++) {
        const Tap& t = a.taps[i];
        const int j = i % 9;
        if (t.view != i / 9 || t.dy != j / 3 - 1 || t.dx != j % 3 - 1 || t.slab != a.taps[i - j].slab + j) return false;
    }
    return true;
}

}  // namespace tc

int tc_init(afi_ctx* ctx) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        set_error("cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
        return AFI_ERR_CUDA;
    }
    ctx->encode_tiled = fn;
    AFI_CUDA(cudaFuncSetAttribute(tc::k_conv_tc<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    AFI_CUDA(cudaFuncSetAttribute(tc::k_conv_tc<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    AFI_CUDA(cudaFuncSetAttribute(tc::k_conv_tc<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    AFI_CUDA(cudaFuncSetAttribute(tc::k_conv_tc<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    AFI_CUDA(cudaFuncSetAttribute(tc::k_wgrad_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    AFI_CUDA(cudaFuncSetAttribute(tc::k_wgrad_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    // halo-tile kernels: everything the SM has beyond their static shared memory (statistics scratch + barriers)
    cudaFuncAttributes fa;
    AFI_CUDA(cudaFuncGetAttributes(&fa, tc::k_conv_halo<4, true>));
    tc::g_halo_dyn_max[0] = 232448 - (int)fa.sharedSizeBytes;
    AFI_CUDA(cudaFuncGetAttributes(&fa, tc::k_conv_halo<8, true>));
    tc::g_halo_dyn_max[1] = 232448 - (int)fa.sharedSizeBytes;
    AFI_CUDA(cudaFuncSetAttribute(tc::k_conv_halo<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::g_halo_dyn_max[0]));
    AFI_CUDA(cudaFuncSetAttribute(tc::k_conv_halo<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::g_halo_dyn_max[1]));
    AFI_CUDA(cudaFuncSetAttribute(tc::k_conv_halo<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::g_halo_dyn_max[0]));
    AFI_CUDA(cudaFuncSetAttribute(tc::k_conv_halo<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::g_halo_dyn_max[1]));
    // fp32 auxiliary operands (split-precision mode): same shared-memory footprint as their bf16 twins
    AFI_CUDA(cudaFuncSetAttribute(tc::k_conv_halo<4, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::g_halo_dyn_max[0]));
    AFI_CUDA(cudaFuncSetAttribute(tc::k_conv_halo<8, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::g_halo_dyn_max[1]));
    return AFI_OK;
}

// problems are scheduled largest first (longest-processing-time order keeps the persistent CTAs' tail short)
static void order_by_size(int n, const long long* size, int* order) {
    for (int i = 0; i < n; i++) order[i] = i;
    for (int i = 1; i < n; i++) {
        int k = order[i], j = i - 1;
        while (j >= 0 && size[order[j]] < size[k]) { order[j + 1] = order[j]; j--; }
        order[j + 1] = k;
    }
}

int conv_tc(afi_ctx* ctx, const ConvArgs& a, cudaStream_t st) {
    using namespace tc;
    AFI_REQUIRE(ctx && ctx->encode_tiled, "conv_tc: context not initialised");
    AFI_REQUIRE(a.cin % 16 == 0 && a.cout % 16 == 0, "conv_tc: cin %d / cout %d must be multiples of 16", a.cin, a.cout);
    AFI_REQUIRE(a.nprob >= 1 && a.nprob <= AFI_MAX_PROB, "conv_tc: bad problem count %d", a.nprob);
    long long size[AFI_MAX_PROB], pixels = 0;
    for (int i = 0; i < a.nprob; i++) {
        const ConvProb& q = a.p[i];
        size[i] = (long long)q.N * q.H * q.W; pixels += size[i];
        // the epilogue prefetches its auxiliary operands through two slots (r1|mask, r2) or one fp32 slot (accin)
        AFI_REQUIRE(!(q.mask.ptr && q.r1.ptr), "conv_tc: unsupported epilogue operand combination");
        AFI_REQUIRE(!q.accin.ptr || !(q.mask.ptr || q.r1.ptr || q.r2.ptr), "conv_tc: accin excludes other epilogue operands");
        AFI_REQUIRE((!q.r1.ptr || a.r1_dt == (a.aux_f32 ? DT_F32 : DT_BF16)) && (!q.r2.ptr || a.r2_dt == (a.aux_f32 ? DT_F32 : DT_BF16)),
                    "conv_tc: residuals must be %s", a.aux_f32 ? "fp32" : "bf16");
        if (a.stat_mode) AFI_REQUIRE(a.stat_mode == 1 && q.stat0 && q.stat1, "conv_tc: missing statistics operands");
    }
    if (pixels == 0) return AFI_OK;
    AFI_REQUIRE(a.split == 0 || a.split == 3, "conv_tc: split %d", a.split);
    AFI_REQUIRE(a.split_col % 16 == 0 && (a.split_col == 0 || (!a.nphase && !a.stat_mode && a.split_col <= a.cout)), "conv_tc: bad split_col %d", a.split_col);
    AFI_REQUIRE(!a.split || a.split_pairs == 6 || a.split_pairs == 3, "conv_tc: split_pairs %d", a.split_pairs);
    const int planes = a.split ? (a.split_pairs == 3 ? 2 : 3) : 1;      // operand planes of the A views (the weights always carry three)
    int order[AFI_MAX_PROB];
    order_by_size(a.nprob, size, order);
    Tiling tl{};
    // N tile: 256 columns whenever the problem has at least one M tile per SM (a narrower tile was measured SLOWER on the training shapes:
    // 806 -> 605 TFLOP/s on 256->256).  SMALL problems -- inference on the coarse pyramid levels: a handful of M tiles -- are latency-
    // bound by the serial K loop of each tile (36 taps x 4 MMAs x 128 cycles = 10 us for a 256-column tile with one SM busy per tile):
    // there the N tile shrinks (128, 64) until the tiles cover the SMs, which shortens every tile's K loop proportionally.
    int bn_cap = 256;
    {
        const long long m_tiles_est = (pixels + 127) / 128;
        while (bn_cap > 64 && m_tiles_est * ((a.cout + bn_cap - 1) / bn_cap) < ctx->sm_count && a.cout > bn_cap / 2) bn_cap /= 2;
    }
    tl.n_tiles = (a.cout + bn_cap - 1) / bn_cap;
    tl.bn = ((a.cout + tl.n_tiles - 1) / tl.n_tiles + 15) / 16 * 16;
    if (a.nphase) {      // one N tile per phase
        AFI_REQUIRE(a.nphase == 4 && a.cout <= 256 && !a.stat_mode, "conv_tc: phase mode needs four phases of <= 256 couts without statistics");
        tl.n_tiles = a.nphase;
        tl.bn = (a.cout + 15) / 16 * 16;
    }
    tl.kchunks = (a.cin + 63) / 64;
    int nviews = 0;
    for (int i = 0; i < a.ntaps; i++) nviews = a.taps[i].view + 1 > nviews ? a.taps[i].view + 1 : nviews;
    AFI_REQUIRE(nviews >= 1 && nviews <= 4, "conv_tc: bad view count");
    Maps maps;
    // short-K layers (K = taps x cin < 4096) cannot hide their epilogue behind the MMAs: eight epilogue warps; so do the layers with
    // fused statistics below K = 8192 (measured: 512 -> 1024 with statistics on four warps is epilogue-bound in pair mode)
    const int K = a.ntaps * tl.kchunks * 64;
    const int Kp = K * (a.split ? a.split_pairs : 1);     // MMA work per tile: what the epilogue has to hide behind
    const int epi8 = (Kp < 4096 || (a.stat_mode && Kp < 8192)) ? 1 : 0;
    // halo tiles on CTA pairs: every 3x3 layer except the narrow ones (N tile < 128 or 32 input channels: the dense blocks' growth
    // convs and their dgrads), where the per-tap kernel measured 10-15 % faster (a pair halves the number of schedulable tiles and
    // an N = 32 MMA is bound by its A-operand read either way)
    int hmode = halo_eligible(a) ? halo_mode() : 0;
    if (a.aux_f32 && hmode == 1) hmode = 2;                // (the single-CTA halo kernel has no fp32-operand instantiation)
    if (hmode == 2 && K < 4096 && (tl.bn < 128 || a.cin < 64) && !getenv("AFIGAN_PAIR_ALL")) hmode = 0;
    // (Negative result, inference pyramid levels with fewer tiles than SMs: single-CTA halo tiles with 32/64-column N tiles and all nine
    // taps per weight slot instead of the per-tap kernel changed a 28-call image by < 2 %: those calls are bound by the ~7 us of
    // set-up, first TMA round trip, epilogue and exit of each of their 21 dependent kernels, not by the K loops.)
    const bool pair = hmode == 2;
#ifdef AFI_STALL_COUNTERS
    if (getenv("AFIGAN_HALO_DBG")) {
        static long long* dbg = nullptr;
        if (!dbg) cudaMalloc(&dbg, 4096 * sizeof(long long));
        cudaMemsetAsync(dbg, 0, 4096 * sizeof(long long), st);
        tl.dbg = dbg;
    }
#endif
    if (hmode) {
        tl.nviews = nviews;
        const int b_tile = pair ? tl.bn * 64 : tl.bn * 128;            // bytes of one tap's weight tile in this CTA (rows x 128 B: a multiple of 1024)
        // Taps per weight slot.  What the single producer / MMA threads spend per slot on barrier waits, expect_tx, TMA issue and commit
        // (~300 cycles) is more than the 4 MMAs of a 128-column tap (256 cycles) and not far below those of a 256-column tap (512): a slot
        // holds a whole kernel row (three taps, one TMA box, one barrier round trip).  Measured alone on 2 x 200 x 336: 256->128 766 -> 910,
        // 128->128 614 -> 677, 1024->1024 1728 -> 1769 TFLOP/s.
        tl.tb = 3;
        if (const char* e = getenv("AFIGAN_HALO_TB")) { int v = atoi(e); if (v == 1 || v == 3 || v == 9) tl.tb = v; }
        tl.b_slot = b_tile * tl.tb;
        const int room = g_halo_dyn_max[epi8] - 1024 - HALO_SA * HALO_SLOT;
        if (tl.tb == 9 && room / tl.b_slot < 2) { tl.tb = 3; tl.b_slot = b_tile * 3; }
        if (tl.tb == 3 && room / tl.b_slot < 3) { tl.tb = 1; tl.b_slot = b_tile; }
        tl.sb = room / tl.b_slot;
        if (tl.sb > HALO_SB_MAX) tl.sb = HALO_SB_MAX;
        AFI_REQUIRE(tl.sb >= 2, "conv_tc: shared memory budget leaves %d weight stages", tl.sb);
        for (int v = 0; v < nviews; v++) tl.view_slab0[v] = a.taps[9 * v].slab;
    }
    int begin = 0, np = 0;
    for (int oi = 0; oi < a.nprob; oi++) {
        const ConvProb& pr = a.p[order[oi]];
        if (size[order[oi]] == 0) continue;
        TileP& t = tl.p[np];
        t.orient = 0;
        if (hmode) {     // 16 x 8 or 8 x 16 patches, whichever pads the level less
            const long long n0 = (long long)((pr.H + 15) / 16) * ((pr.W + 7) / 8), n1 = (long long)((pr.H + 7) / 8) * ((pr.W + 15) / 16);
            t.orient = n1 < n0 ? 1 : 0;
            t.TH = t.orient ? 8 : 16; t.TW = t.orient ? 16 : 8;
        } else pick_patch(pr.H, pr.W, 128, &t.TH, &t.TW);
        t.tiles_x = (pr.W + t.TW - 1) / t.TW;
        t.tiles_y = (pr.H + t.TH - 1) / t.TH;
        t.begin = begin;
        t.prob = order[oi];
        const int m_tiles = pr.N * t.tiles_x * t.tiles_y;
        begin += (pair ? (m_tiles + 1) / 2 : m_tiles) * tl.n_tiles;
        for (int v = 0; v < nviews; v++) {
            const int vW = pr.vw[v] ? pr.vw[v] : pr.W, vH = pr.vh[v] ? pr.vh[v] : pr.H;      // the view's own extent (out-of-bounds reads are zero)
            if (hmode) AFI_TRY(encode_view_halo(ctx, &maps.a[t.prob][v], pr.in[v], a.cin, vW, vH, pr.N * planes, t.orient));
            else AFI_TRY(encode_view(ctx, &maps.a[t.prob][v], pr.in[v], a.cin, vW, vH, pr.N * planes, t.TW, t.TH));
        }
        np++;
    }
    tl.nprob = np;
    tl.p[np].begin = begin;
    tl.total = begin;
    int nslab = 0;
    for (int i = 0; i < a.ntaps; i++) nslab = a.taps[i].slab + 1 > nslab ? a.taps[i].slab + 1 : nslab;
    if (a.nphase) nslab += 9 * (a.nphase - 1);
    if (a.nslab > nslab) nslab = a.nslab;
    set_pairs(tl, a.split, a.split_pairs, nslab);
    // six-pair products (the sign-critical forward convs) alternate between the two accumulators; so do three-pair products with chains
    // of more than 160 MMAs per plane pair (K > 2560).  (Measured on the full-size step: with single accumulators on the generator's
    // K = 2304 layers the worst sampled generator gradient error rose from 3.7e-4 to 9.2e-4.)
    tl.dual = (a.split && (a.split_pairs == 6 || a.ntaps * tl.kchunks * 4 > 160)) ? 1 : 0;
    {
        cuuint64_t dims[3] = {(cuuint64_t)a.cin, (cuuint64_t)a.cout, (cuuint64_t)nslab * (a.split ? 3 : 1)};
        cuuint64_t strides[2] = {(cuuint64_t)a.cin * 2, (cuuint64_t)a.cin * a.cout * 2};
        cuuint32_t box[3] = {64, (cuuint32_t)(pair ? tl.bn / 2 : tl.bn), (cuuint32_t)(hmode ? tl.tb : 1)};
        AFI_TRY(encode_map(ctx, &maps.b, const_cast<void*>(a.w), 3, dims, strides, box));
    }
    tl.stage_bytes = A_BYTES + (tl.bn * 128 + 1023) / 1024 * 1024;
    tl.nstages = (SMEM_BYTES - 1024) / tl.stage_bytes;
    if (tl.nstages > STAGES_MAX) tl.nstages = STAGES_MAX;
    int grid = tl.total < ctx->sm_count ? tl.total : ctx->sm_count;
    if (pair) { int ncl = ctx->sm_count / 2; grid = 2 * (tl.total < ncl ? tl.total : ncl); }
    ProfScope prof(pair ? PROF_CONV_PAIR : (hmode ? PROF_CONV_HALO : PROF_CONV_TC), 2.0 * pixels * (double)a.ntaps * a.cin * a.cout * (a.nphase ? a.nphase : 1),
                   a.cin, a.cout * (a.nphase ? a.nphase : 1), pixels, st);
    // short-K layers (K = taps x cin < 4096) get eight epilogue warps
    // Programmatic dependent launch (AFIGAN_PDL=0 disables): the kernel may start its prologue (barrier init, TMEM allocation, tensor-map
    // prefetch) while its predecessor in the stream drains; every global-memory access waits for the predecessor (griddepcontrol.wait).
    // Measured: -18 % on the latency-bound inference sweep (BASELINE config 5: 9.8 -> 8.0 ms per image), +1.7 % on the generator alone.
    const char* pe = getenv("AFIGAN_PDL");
    const bool pdl = !pe || atoi(pe) != 0;
    {
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cudaLaunchAttribute attr[2];
        int na = 0;
        if (pair) {
            attr[na].id = cudaLaunchAttributeClusterDimension;
            attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
            na++;
        }
        if (pdl) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; na++; }
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64 + 32 * (epi8 ? 8 : 4)); cfg.stream = st;
        cfg.dynamicSmemBytes = hmode ? 1024 + HALO_SA * HALO_SLOT + tl.sb * tl.b_slot : SMEM_BYTES;
        cfg.attrs = attr; cfg.numAttrs = na;
        if (a.aux_f32) {
            if (pair) {
                if (epi8) AFI_CUDA(cudaLaunchKernelEx(&cfg, k_conv_halo<8, true, true>, maps, a, tl));
                else AFI_CUDA(cudaLaunchKernelEx(&cfg, k_conv_halo<4, true, true>, maps, a, tl));
            } else {
                if (epi8) AFI_CUDA(cudaLaunchKernelEx(&cfg, k_conv_tc<8, true>, maps, a, tl));
                else AFI_CUDA(cudaLaunchKernelEx(&cfg, k_conv_tc<4, true>, maps, a, tl));
            }
        } else if (pair) {
            if (epi8) AFI_CUDA(cudaLaunchKernelEx(&cfg, k_conv_halo<8, true>, maps, a, tl));
            else AFI_CUDA(cudaLaunchKernelEx(&cfg, k_conv_halo<4, true>, maps, a, tl));
        } else if (hmode) {
            if (epi8) AFI_CUDA(cudaLaunchKernelEx(&cfg, k_conv_halo<8, false>, maps, a, tl));
            else AFI_CUDA(cudaLaunchKernelEx(&cfg, k_conv_halo<4, false>, maps, a, tl));
        } else {
            if (epi8) AFI_CUDA(cudaLaunchKernelEx(&cfg, k_conv_tc<8>, maps, a, tl));
            else AFI_CUDA(cudaLaunchKernelEx(&cfg, k_conv_tc<4>, maps, a, tl));
        }
    }
#ifdef AFI_STALL_COUNTERS
    {
        if (tl.dbg) {      // experiment aid: per-CTA stall cycles of the producer / MMA threads
            long long h[256 * 8];
            cudaStreamSynchronize(st);
            cudaMemcpy(h, tl.dbg, sizeof(h), cudaMemcpyDeviceToHost);
            double sum[8] = {0};
            const int mg = pair ? grid / 2 : grid;
            for (int b = 0; b < grid; b++) for (int i = 0; i < 8; i++) sum[i] += (double)h[b * 8 + i];
            fprintf(stderr, "[halo dbg] cin %d cout %d sa %d sb %d tiles %d: producer wait a_empty %.0f b_empty %.0f of %.0f | mma wait a_full %.0f b_full %.0f acc_empty %.0f of %.0f (avg cycles per CTA)\n",
                    a.cin, a.cout, hmode ? HALO_SA : 0, tl.sb, tl.total, sum[0] / grid, sum[1] / grid, sum[2] / grid, sum[3] / mg, sum[4] / mg, sum[5] / mg, sum[6] / mg);
            if (hmode && grid <= 16) {      // timeline (ns since the first CTA's entry): entry, setup done, dependency wait done, first halo landed, MMAs issued, producer done, epilogue done, exit
                long long ts[16 * 8];
                cudaMemcpy(ts, tl.dbg + 2048, sizeof(ts), cudaMemcpyDeviceToHost);
                long long t0 = ts[0];
                for (int b = 0; b < grid; b++) if (ts[b * 8] && ts[b * 8] < t0) t0 = ts[b * 8];
                for (int b = 0; b < grid; b++) {
                    fprintf(stderr, "[halo timeline] cta %d:", b);
                    for (int i = 0; i < 8; i++) fprintf(stderr, " %lld", ts[b * 8 + i] ? ts[b * 8 + i] - t0 : -1);
                    fprintf(stderr, "\n");
                }
            }
        }
    }
#endif
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}

int wgrad_tc(afi_ctx* ctx, const WgradArgs& a, cudaStream_t st) {
    using namespace tc;
    AFI_REQUIRE(ctx && ctx->encode_tiled, "wgrad_tc: context not initialised");
    AFI_REQUIRE(a.cin % 16 == 0 && a.cout % 16 == 0, "wgrad_tc: cin %d / cout %d must be multiples of 16", a.cin, a.cout);
    AFI_REQUIRE(a.nprob >= 1 && a.nprob <= AFI_MAX_PROB, "wgrad_tc: bad problem count %d", a.nprob);
    long long size[AFI_MAX_PROB], pixels = 0;
    for (int i = 0; i < a.nprob; i++) { size[i] = (long long)a.p[i].N * a.p[i].H * a.p[i].W; pixels += size[i]; }
    if (pixels == 0) return AFI_OK;
    int order[AFI_MAX_PROB];
    order_by_size(a.nprob, size, order);
    AFI_REQUIRE(a.split == 0 || a.split == 3, "wgrad_tc: split %d", a.split);
    AFI_REQUIRE(!a.split || a.split_pairs == 6 || a.split_pairs == 3, "wgrad_tc: split_pairs %d", a.split_pairs);
    const int planes = a.split ? (a.split_pairs == 3 ? 2 : 3) : 1;
    Tiling tl{};
    set_pairs(tl, a.split, a.split_pairs, 0);
    tl.dual = a.split ? 1 : 0;
    tl.m_tiles = (a.cout + 127) / 128;
    tl.n_tiles = (a.cin + 255) / 256;
    tl.bn = ((a.cin + tl.n_tiles - 1) / tl.n_tiles + 63) / 64 * 64;
    // CTA pairs (M = 256 couts per MMA, X tile split across the pair) when the shape allows: whole 256-cout tiles and an even number of
    // 64-channel X boxes per N tile; AFIGAN_WGRAD_PAIR=0 keeps one CTA per tile
    const char* wp = getenv("AFIGAN_WGRAD_PAIR");
    const bool pair = (!wp || atoi(wp) != 0) && a.cout % 256 == 0 && (tl.bn / 64) % 2 == 0;
    Maps maps;
    int begin = 0, np = 0;
    for (int oi = 0; oi < a.nprob; oi++) {
        const WgradProb& pr = a.p[order[oi]];
        if (size[order[oi]] == 0) continue;
        TileP& t = tl.p[np];
        pick_patch(pr.H, pr.W, 64, &t.TH, &t.TW);
        t.tiles_x = (pr.W + t.TW - 1) / t.TW;
        t.tiles_y = (pr.H + t.TH - 1) / t.TH;
        t.begin = begin;
        t.prob = order[oi];
        begin += pr.N * t.tiles_x * t.tiles_y;
        AFI_TRY(encode_view_grouped(ctx, &maps.a[t.prob][0], pr.dy, a.cout, pr.W, pr.H, pr.N * planes, t.TW, t.TH, 2));
        AFI_TRY(encode_view_grouped(ctx, &maps.a[t.prob][1], pr.x, a.cin, pr.xw ? pr.xw : pr.W, pr.xh ? pr.xh : pr.H, pr.N * planes, t.TW, t.TH,
                                    pair ? tl.bn / 128 : tl.bn / 64));
        np++;
    }
    tl.nprob = np;
    tl.p[np].begin = begin;
    tl.ktiles = begin;
    maps.b = maps.a[tl.p[0].prob][0];
    if (pair) tl.m_tiles = a.cout / 256;
    const int workers = pair ? ctx->sm_count / 2 : ctx->sm_count;
    // split-K so that the work items fill (at most) two full waves of the persistent grid: base * ks <= 2 * workers
    int base = a.ntaps * tl.m_tiles * tl.n_tiles;
    int ks = (2 * workers) / base;
    if (ks > tl.ktiles) ks = tl.ktiles;
    if (ks < 1) ks = 1;
    int kper = (tl.ktiles + ks - 1) / ks;
    // split mode: the accumulation chain of one item is capped at 512 K tiles (1024 MMAs per accumulator and plane pair, a truncation
    // bias of ~2e-5, see DUAL_ACC in k_conv_tc); the partial sums meet in fp32 RED adds.  A tighter cap costs more than it buys: at
    // 64 K tiles the RED traffic of the discriminator's weight gradients (1.6 GB per launch) made them 1.75x slower (measured).
    if (a.split && kper > 512) kper = 512;
    tl.ksplit = (tl.ktiles + kper - 1) / kper;
    tl.total = base * tl.ksplit;
    tl.stage_bytes = A_BYTES + (tl.bn / 64) * (pair ? 4096 : 8192);
    tl.nstages = (SMEM_BYTES - 1024) / tl.stage_bytes;
    if (tl.nstages > STAGES_MAX) tl.nstages = STAGES_MAX;
    int grid = tl.total < workers ? tl.total : workers;
    ProfScope prof(PROF_WGRAD_TC, 2.0 * pixels * (double)a.ntaps * a.cin * a.cout, a.cin, a.cout, pixels, st);
    if (pair) {
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.gridDim = dim3(2 * grid); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = st;
        cfg.attrs = attr; cfg.numAttrs = 1;
        AFI_CUDA(cudaLaunchKernelEx(&cfg, k_wgrad_tc<true>, maps, a, tl));
    } else k_wgrad_tc<false><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(maps, a, tl);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}


// ---- split-precision front ends (AFI_PREC_SPLIT): fp32 views -> bf16 planes in the caller's scratch -> six-pair tcgen05 product --------
int conv_tc_split(afi_ctx* ctx, const ConvArgs& a, cudaStream_t st) {
    AFI_REQUIRE(a.nprob >= 1 && a.nprob <= AFI_MAX_PROB, "conv_tc_split: bad problem count %d", a.nprob);
    int nviews = 0;
    for (int i = 0; i < a.ntaps; i++) nviews = a.taps[i].view + 1 > nviews ? a.taps[i].view + 1 : nviews;
    AFI_REQUIRE(nviews >= 1 && nviews <= 4, "conv_tc_split: bad view count");
    AFI_REQUIRE(a.out_dt == DT_F32 && !a.stat_mode, "conv_tc_split: fp32 output without fused statistics");
    ConvArgs b = a;
    if (b.split_pairs != 3) b.split_pairs = 6;
    const int nplanes = b.split_pairs == 3 ? 2 : 3;
    SplitJob jobs[AFI_MAX_SPLIT]; int nj = 0;
    const int cpad = split_cpad(a.cin);
    for (int k = 0; k < a.nprob; k++) {
        const ConvProb& q = a.p[k];
        const long long P = (long long)q.N * q.H * q.W;
        if (P == 0) continue;
        size_t off = 0;
        for (int v = 0; v < nviews; v++) {
            const int vH = q.vh[v] ? q.vh[v] : q.H, vW = q.vw[v] ? q.vw[v] : q.W;
            const size_t each = split_planes_bytes((long long)q.N * vH * vW, a.cin);
            if (!q.sws || off + each > q.sws_bytes) {
                set_error("conv_tc_split: problem %d needs more than %zu B of split scratch", k, q.sws ? q.sws_bytes : (size_t)0);
                return AFI_ERR_WORKSPACE;
            }
            SplitJob& j = jobs[nj++];
            j.src = q.in[v]; j.n = q.N; j.h = vH; j.w = vW; j.c = a.cin; j.dst = (char*)q.sws + off; j.nplanes = nplanes;
            b.p[k].in[v] = pview(j.dst, vH, vW, cpad);
            off += each;
        }
    }
    AFI_TRY(split3_group(nj, jobs, st));
    b.split = 3; b.aux_f32 = 1;
    return conv_tc(ctx, b, st);
}

int wgrad_tc_split(afi_ctx* ctx, const WgradArgs& a, cudaStream_t st) {
    AFI_REQUIRE(a.nprob >= 1 && a.nprob <= AFI_MAX_PROB, "wgrad_tc_split: bad problem count %d", a.nprob);
    WgradArgs b = a;
    if (b.split_pairs != 6) b.split_pairs = 3;      // a weight gradient depends on its operands linearly
    const int nplanes = b.split_pairs == 3 ? 2 : 3;
    SplitJob jobs[AFI_MAX_SPLIT]; int nj = 0;
    for (int k = 0; k < a.nprob; k++) {
        const WgradProb& q = a.p[k];
        const long long P = (long long)q.N * q.H * q.W;
        if (P == 0) continue;
        const int xH = q.xh ? q.xh : q.H, xW = q.xw ? q.xw : q.W;
        const size_t bx = split_planes_bytes((long long)q.N * xH * xW, a.cin), by = split_planes_bytes(P, a.cout);
        if (!q.sws || bx + by > q.sws_bytes) {
            set_error("wgrad_tc_split: problem %d needs %zu B of split scratch, has %zu", k, bx + by, q.sws ? q.sws_bytes : (size_t)0);
            return AFI_ERR_WORKSPACE;
        }
        SplitJob& jx = jobs[nj++];
        jx.src = q.x; jx.n = q.N; jx.h = xH; jx.w = xW; jx.c = a.cin; jx.dst = q.sws; jx.nplanes = nplanes;
        SplitJob& jy = jobs[nj++];
        jy.src = q.dy; jy.n = q.N; jy.h = q.H; jy.w = q.W; jy.c = a.cout; jy.dst = (char*)q.sws + bx; jy.nplanes = nplanes;
        b.p[k].x = pview(jx.dst, xH, xW, split_cpad(a.cin));
        b.p[k].dy = pview(jy.dst, q.H, q.W, split_cpad(a.cout));
    }
    AFI_TRY(split3_group(nj, jobs, st));
    b.split = 3;
    return wgrad_tc(ctx, b, st);
}

}  // namespace afi
