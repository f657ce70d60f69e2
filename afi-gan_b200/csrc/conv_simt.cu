// CUDA-core (FFMA, fp32 accumulate) implicit-GEMM engine: the PARITY mode of the library (AFI_PREC_FP32: every
// product exact in fp32) and the on-device cross-check of the tcgen05 engine (AFI_PREC_BF16_SIMT).
//   conv : M = pixels (linear n,y,x index, no spatial tile waste), N = cout, K = taps x cin
//   wgrad: M = cin, N = cout, K = pixels (split-K over pixel ranges, fp32 atomics into a zeroed accumulator)
// 128 x {128|32} x 16 tiles, 256 threads, 8x8 / 4x4 register micro-tiles, double-buffered shared memory.
#include "common.cuh"

namespace afi {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int LDS_PAD = 4;

template <typename T, int E> struct VecLoad;
template <int E> struct VecLoad<float, E> {
    __device__ static __forceinline__ void load(const float* p, float* out) {
        if (E == 8) {
            float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
            out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w; out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
        } else if (E == 4) {
            float4 a = *reinterpret_cast<const float4*>(p);
            out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
        } else {
            float2 a = *reinterpret_cast<const float2*>(p);
            out[0] = a.x; out[1] = a.y;
        }
    }
};
template <int E> struct VecLoad<bf16, E> {
    __device__ static __forceinline__ void load(const bf16* p, float* out) {
        if (E == 8) {
            uint4 u = *reinterpret_cast<const uint4*>(p);
            uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int i = 0; i < 4; i++) { out[2 * i] = __uint_as_float(w[i] << 16); out[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
        } else if (E == 4) {
            uint2 u = *reinterpret_cast<const uint2*>(p);
            out[0] = __uint_as_float(u.x << 16); out[1] = __uint_as_float(u.x & 0xffff0000u);
            out[2] = __uint_as_float(u.y << 16); out[3] = __uint_as_float(u.y & 0xffff0000u);
        } else {
            uint32_t u = *reinterpret_cast<const uint32_t*>(p);
            out[0] = __uint_as_float(u << 16); out[1] = __uint_as_float(u & 0xffff0000u);
        }
    }
};

__device__ __forceinline__ float4 ld4g(const void* base, long long off, int dt) {
    if (dt == DT_F32) return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
    uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(base) + off);
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xffff0000u));
}
__device__ __forceinline__ void st4g(void* base, long long off, int dt, float4 v) {
    if (dt == DT_F32) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off) = v;
    } else {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 u; u.x = *reinterpret_cast<uint32_t*>(&lo); u.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(base) + off) = u;
    }
}

// acc[rs][i][cs][j] += As[k][row(rs,i)] * Bs[k][col(cs,j)]
template <int RS, int CS, int BN>
__device__ __forceinline__ void tile_fma(const float (*As)[BM + LDS_PAD], const float (*Bs)[BN + LDS_PAD], int ty, int tx,
                                         float (&acc)[RS][4][CS][4]) {
#pragma unroll
    for (int k = 0; k < BK; k++) {
        float4 av[RS], bv[CS];
#pragma unroll
        for (int rs = 0; rs < RS; rs++) av[rs] = *reinterpret_cast<const float4*>(&As[k][rs * 64 + ty * 4]);
#pragma unroll
        for (int cs = 0; cs < CS; cs++) bv[cs] = *reinterpret_cast<const float4*>(&Bs[k][cs * 64 + tx * 4]);
#pragma unroll
        for (int rs = 0; rs < RS; rs++) {
            float ar[4] = {av[rs].x, av[rs].y, av[rs].z, av[rs].w};
#pragma unroll
            for (int cs = 0; cs < CS; cs++) {
                float br[4] = {bv[cs].x, bv[cs].y, bv[cs].z, bv[cs].w};
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[rs][i][cs][j] = fmaf(ar[i], br[j], acc[rs][i][cs][j]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// convolution
// ---------------------------------------------------------------------------------------------------
struct SimtGroup {
    int mt_begin[AFI_MAX_PROB + 1];          // conv: first M tile of each problem
    long long px_begin[AFI_MAX_PROB + 1];    // wgrad: first global pixel of each problem
};

template <typename T, int BN>
__global__ void __launch_bounds__(256) k_conv_simt(const __grid_constant__ ConvArgs a, const __grid_constant__ SimtGroup grp) {
    constexpr int RS = BN == 128 ? 2 : 1, CS = RS;
    constexpr int TXN = BN / (4 * CS);
    constexpr int EB = BN / 16;
    __shared__ __align__(16) float As[2][BK][BM + LDS_PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + LDS_PAD];
    const int tid = threadIdx.x;
    int pi = 0;
    while (pi + 1 < a.nprob && (int)blockIdx.x >= grp.mt_begin[pi + 1]) pi++;
    const ConvProb& pr = a.p[pi];
    const int PH = pr.H, PW = pr.W;
    const long long M = (long long)pr.N * PH * PW;
    const long long m0 = (long long)((int)blockIdx.x - grp.mt_begin[pi]) * BM;
    const int n0 = blockIdx.y * BN;

    // A loader: one pixel row, 8 consecutive channels
    const int arow = tid >> 1, akoff = (tid & 1) * 8;
    long long ap = m0 + arow;
    const bool arow_ok = ap < M;
    int ax = 0, ay = 0, an = 0;
    if (arow_ok) { ax = (int)(ap % PW); long long t = ap / PW; ay = (int)(t % PH); an = (int)(t / PH); }
    // B loader
    const int bk = tid >> 4, bnoff = (tid & 15) * EB;
    const bool bcol_ok = (n0 + bnoff) < a.cout;

    const int kc_per_tap = a.cin / BK;
    const int iters = a.ntaps * kc_per_tap;
    const T* wbase = reinterpret_cast<const T*>(a.w);

    float ra[8], rb[EB];
    auto load_global = [&](int it) {
        int tp = it / kc_per_tap, c0 = (it - tp * kc_per_tap) * BK;
        Tap t = a.taps[tp];
        int iy = ay + t.dy, ix = ax + t.dx;
        if (arow_ok && iy >= 0 && iy < PH && ix >= 0 && ix < PW) {
            const PView& v = pr.in[t.view];
            const T* p = reinterpret_cast<const T*>(v.ptr) + an * v.sn + iy * v.sy + ix * v.sx + c0 + akoff;
            VecLoad<T, 8>::load(p, ra);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) ra[i] = 0.f;
        }
        if (bcol_ok) {
            const T* p = wbase + ((long long)t.slab * a.cin + c0 + bk) * a.cout + n0 + bnoff;
            VecLoad<T, EB>::load(p, rb);
        } else {
#pragma unroll
            for (int i = 0; i < EB; i++) rb[i] = 0.f;
        }
    };
    auto store_smem = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; i++) As[buf][akoff + i][arow] = ra[i];
#pragma unroll
        for (int i = 0; i < EB; i++) Bs[buf][bk][bnoff + i] = rb[i];
    };

    const int ty = tid / TXN, tx = tid % TXN;
    float acc[RS][4][CS][4];
#pragma unroll
    for (int rs = 0; rs < RS; rs++)
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int cs = 0; cs < CS; cs++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[rs][i][cs][j] = 0.f;

    load_global(0);
    store_smem(0);
    __syncthreads();
    for (int it = 0; it < iters; it++) {
        if (it + 1 < iters) load_global(it + 1);
        tile_fma<RS, CS, BN>(As[it & 1], Bs[it & 1], ty, tx, acc);
        if (it + 1 < iters) store_smem((it + 1) & 1);
        __syncthreads();
    }

    // epilogue
#pragma unroll
    for (int rs = 0; rs < RS; rs++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            long long p = m0 + rs * 64 + ty * 4 + i;
            if (p >= M) continue;
            int x = (int)(p % PW); long long t = p / PW; int y = (int)(t % PH); int n = (int)(t / PH);
#pragma unroll
            for (int cs = 0; cs < CS; cs++) {
                int col = n0 + cs * 64 + tx * 4;
                if (col >= a.cout) continue;
                float4 v = make_float4(acc[rs][i][cs][0], acc[rs][i][cs][1], acc[rs][i][cs][2], acc[rs][i][cs][3]);
                if (a.bias) { float4 b = *reinterpret_cast<const float4*>(a.bias + col); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
                if (a.act) {
                    v.x = v.x > 0.f ? v.x : v.x * a.slope; v.y = v.y > 0.f ? v.y : v.y * a.slope;
                    v.z = v.z > 0.f ? v.z : v.z * a.slope; v.w = v.w > 0.f ? v.w : v.w * a.slope;
                }
                v.x *= a.alpha; v.y *= a.alpha; v.z *= a.alpha; v.w *= a.alpha;
                if (pr.r1.ptr) {
                    float4 r = ld4g(pr.r1.ptr, n * pr.r1.sn + y * pr.r1.sy + x * pr.r1.sx + col, a.r1_dt);
                    v.x += a.beta1 * r.x; v.y += a.beta1 * r.y; v.z += a.beta1 * r.z; v.w += a.beta1 * r.w;
                }
                if (pr.r2.ptr) {
                    float4 r = ld4g(pr.r2.ptr, n * pr.r2.sn + y * pr.r2.sy + x * pr.r2.sx + col, a.r2_dt);
                    v.x += a.beta2 * r.x; v.y += a.beta2 * r.y; v.z += a.beta2 * r.z; v.w += a.beta2 * r.w;
                }
                if (pr.accin.ptr) {
                    float4 r = ld4g(pr.accin.ptr, n * pr.accin.sn + y * pr.accin.sy + x * pr.accin.sx + col, DT_F32);
                    v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
                }
                if (pr.mask.ptr) {
                    float4 m = ld4g(pr.mask.ptr, n * pr.mask.sn + y * pr.mask.sy + x * pr.mask.sx + col, dt_of<T>::v);
                    v.x *= m.x > 0.f ? 1.f : a.mask_slope; v.y *= m.y > 0.f ? 1.f : a.mask_slope;
                    v.z *= m.z > 0.f ? 1.f : a.mask_slope; v.w *= m.w > 0.f ? 1.f : a.mask_slope;
                }
                st4g(pr.out.ptr, n * pr.out.sn + y * pr.out.sy + x * pr.out.sx + col, a.out_dt, v);
            }
        }
    }
}

template <typename T>
int conv_simt(const ConvArgs& a, cudaStream_t st) {
    AFI_REQUIRE(a.cin % BK == 0 && a.cout % 4 == 0 && a.cout % 32 == 0, "conv_simt: cin %d / cout %d unsupported", a.cin, a.cout);
    AFI_REQUIRE(a.ntaps >= 1 && a.ntaps <= AFI_MAX_TAPS, "conv_simt: bad tap count");
    AFI_REQUIRE(a.nprob >= 1 && a.nprob <= AFI_MAX_PROB, "conv_simt: bad problem count %d", a.nprob);
    SimtGroup grp;
    long long M = 0;
    int mt = 0;
    for (int i = 0; i < a.nprob; i++) {
        long long Mi = (long long)a.p[i].N * a.p[i].H * a.p[i].W;
        grp.mt_begin[i] = mt;
        mt += (int)((Mi + BM - 1) / BM);
        M += Mi;
    }
    grp.mt_begin[a.nprob] = mt;
    if (M == 0) return AFI_OK;
    ProfScope prof(PROF_CONV_SIMT, 2.0 * M * (double)a.ntaps * a.cin * a.cout, a.cin, a.cout, M, st);
    if (a.cout % 128 == 0 || a.cout > 128) {
        dim3 grid((unsigned)mt, (a.cout + 127) / 128);
        k_conv_simt<T, 128><<<grid, 256, 0, st>>>(a, grp);
    } else {
        dim3 grid((unsigned)mt, (a.cout + 31) / 32);
        k_conv_simt<T, 32><<<grid, 256, 0, st>>>(a, grp);
    }
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
template int conv_simt<float>(const ConvArgs&, cudaStream_t);
template int conv_simt<bf16>(const ConvArgs&, cudaStream_t);

// ---------------------------------------------------------------------------------------------------
// weight gradient
// ---------------------------------------------------------------------------------------------------
template <typename T, int BN>
__global__ void __launch_bounds__(256) k_wgrad_simt(const __grid_constant__ WgradArgs a, const __grid_constant__ SimtGroup grp, int ksplit,
                                                    int chunk) {
    constexpr int RS = BN == 128 ? 2 : 1, CS = RS;
    constexpr int TXN = BN / (4 * CS);
    constexpr int EB = BN / 16;
    __shared__ __align__(16) float As[2][BK][BM + LDS_PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + LDS_PAD];
    const int tid = threadIdx.x;
    const long long P = grp.px_begin[a.nprob];
    const int ci0 = blockIdx.x * BM, co0 = blockIdx.y * BN;
    const int tp = blockIdx.z / ksplit, ks = blockIdx.z % ksplit;
    const Tap tap = a.taps[tp];
    long long pk0 = (long long)ks * chunk, pk1 = pk0 + chunk;
    if (pk1 > P) pk1 = P;
    if (pk0 >= pk1) return;
    const int iters = (int)((pk1 - pk0 + BK - 1) / BK);

    const int lk = tid >> 4;                      // pixel within the K chunk
    const int amoff = (tid & 15) * 8;             // 8 consecutive input channels
    const int bnoff = (tid & 15) * EB;
    const bool a_ok = (ci0 + amoff) < a.cin, b_ok = (co0 + bnoff) < a.cout;
    float ra[8], rb[EB];
    int pi = 0;      // the K range walks the concatenated pixels of all problems in order
    auto load_global = [&](int it) {
        long long p = pk0 + (long long)it * BK + lk;
        bool pok = p < pk1;
        int x = 0, y = 0, n = 0;
        while (pi + 1 < a.nprob && p >= grp.px_begin[pi + 1]) pi++;
        const WgradProb& pr = a.p[pi];
        if (pok) { long long q = p - grp.px_begin[pi]; x = (int)(q % pr.W); long long t = q / pr.W; y = (int)(t % pr.H); n = (int)(t / pr.H); }
        int iy = y + tap.dy, ix = x + tap.dx;
        if (pok && a_ok && iy >= 0 && iy < pr.H && ix >= 0 && ix < pr.W)
            VecLoad<T, 8>::load(reinterpret_cast<const T*>(pr.x.ptr) + n * pr.x.sn + iy * pr.x.sy + ix * pr.x.sx + ci0 + amoff, ra);
        else {
#pragma unroll
            for (int i = 0; i < 8; i++) ra[i] = 0.f;
        }
        if (pok && b_ok) VecLoad<T, EB>::load(reinterpret_cast<const T*>(pr.dy.ptr) + n * pr.dy.sn + y * pr.dy.sy + x * pr.dy.sx + co0 + bnoff, rb);
        else {
#pragma unroll
            for (int i = 0; i < EB; i++) rb[i] = 0.f;
        }
    };
    auto store_smem = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; i++) As[buf][lk][amoff + i] = ra[i];
#pragma unroll
        for (int i = 0; i < EB; i++) Bs[buf][lk][bnoff + i] = rb[i];
    };
    const int ty = tid / TXN, tx = tid % TXN;
    float acc[RS][4][CS][4];
#pragma unroll
    for (int rs = 0; rs < RS; rs++)
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int cs = 0; cs < CS; cs++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[rs][i][cs][j] = 0.f;
    load_global(0);
    store_smem(0);
    __syncthreads();
    for (int it = 0; it < iters; it++) {
        if (it + 1 < iters) load_global(it + 1);
        tile_fma<RS, CS, BN>(As[it & 1], Bs[it & 1], ty, tx, acc);
        if (it + 1 < iters) store_smem((it + 1) & 1);
        __syncthreads();
    }
    float* dw = a.dw + (long long)tap.slab * a.cin * a.cout;
#pragma unroll
    for (int rs = 0; rs < RS; rs++)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int ci = ci0 + rs * 64 + ty * 4 + i;
            if (ci >= a.cin) continue;
#pragma unroll
            for (int cs = 0; cs < CS; cs++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    int co = co0 + cs * 64 + tx * 4 + j;
                    if (co < a.cout) atomicAdd(dw + (long long)ci * a.cout + co, acc[rs][i][cs][j]);
                }
        }
}

template <typename T>
int wgrad_simt(const WgradArgs& a, cudaStream_t st) {
    AFI_REQUIRE(a.cin % 8 == 0 && a.cout % 32 == 0, "wgrad_simt: cin %d / cout %d unsupported", a.cin, a.cout);
    AFI_REQUIRE(a.nprob >= 1 && a.nprob <= AFI_MAX_PROB, "wgrad_simt: bad problem count %d", a.nprob);
    SimtGroup grp;
    long long P = 0;
    for (int i = 0; i < a.nprob; i++) { grp.px_begin[i] = P; P += (long long)a.p[i].N * a.p[i].H * a.p[i].W; }
    grp.px_begin[a.nprob] = P;
    if (P == 0) return AFI_OK;
    int bn = (a.cout >= 128) ? 128 : 32;
    int tiles = ((a.cin + BM - 1) / BM) * ((a.cout + bn - 1) / bn) * a.ntaps;
    long long ksplit = (148 * 4 + tiles - 1) / tiles;
    long long maxsplit = (P + 255) / 256;
    if (ksplit > maxsplit) ksplit = maxsplit;
    if (ksplit < 1) ksplit = 1;
    long long chunk = (P + ksplit - 1) / ksplit;
    chunk = (chunk + BK - 1) / BK * BK;
    ksplit = (P + chunk - 1) / chunk;
    dim3 grid((a.cin + BM - 1) / BM, (a.cout + bn - 1) / bn, (unsigned)(a.ntaps * ksplit));
    ProfScope prof(PROF_WGRAD_SIMT, 2.0 * P * (double)a.ntaps * a.cin * a.cout, a.cin, a.cout, P, st);
    if (bn == 128) k_wgrad_simt<T, 128><<<grid, 256, 0, st>>>(a, grp, (int)ksplit, (int)chunk);
    else k_wgrad_simt<T, 32><<<grid, 256, 0, st>>>(a, grp, (int)ksplit, (int)chunk);
    AFI_LAUNCH_CHECK();
    return AFI_OK;
}
template int wgrad_simt<float>(const WgradArgs&, cudaStream_t);
template int wgrad_simt<bf16>(const WgradArgs&, cudaStream_t);

}  // namespace afi
