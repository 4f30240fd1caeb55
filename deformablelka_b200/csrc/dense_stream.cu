// dense_stream.cu -- persistent streaming 1x1 contraction for the projections of the block
// (proj_1 + GELU, conv1 * u, proj_2 + x, conv8 + a, nn.Linear on tokens):  Y[M, N] = epi(X[M, K] . W^T + b),
// K = C in {32, 64, 96}, N <= 96, rows of X contiguous.  Replaces the nn.Conv3d(C, C, 1) / nn.Linear GEMMs of
// transformerblock.py:659-671 and MaxViT_deform_LKA.py:35-50 on the tcgen05 path.
//
// These layers are HBM-bound: read M*K floats, write M*N floats, ~100 FLOP per byte-pair in bf16x3 = far below the tensor
// roofline.  The generic kernel (mma_tc.cu) keeps at most ~32 KB of loads in flight per SM because the loads live in the
// registers of the producer warps (measured 2.5 TB/s = 38 % of the measured HBM peak).  Here the bytes in flight live in
// SHARED MEMORY: one elected thread streams whole 128-row tiles (contiguous K*512 bytes) with cp.async.bulk into a 2-deep raw
// ring (up to 96 KB in flight per SM), converter warps turn each 32-channel chunk of a raw tile into the bf16 hi / lo UMMA
// operand (conflict-free LDS.128 -> 2 x STS.64), one thread issues the N-stacked MMAs (one N = 2 NT over [W hi | W lo] + one
// N = NT with A lo: deform_ps.cu), and a dedicated epilogue warpgroup drains the double-buffered accumulator of tile t while
// tile t+1 is converted and tile t+2 is in flight.  Weights stay resident in shared memory for the whole kernel.
#include <cuda_bf16.h>

#include <cstring>

#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace dlka {
namespace {

using namespace ptx;

constexpr int DS_KC = 32;                 // channels per operand chunk
constexpr int DS_LBO = 2048 + 32;         // operand plane (8 channels x 128 rows) stride
constexpr int DS_APLANE = (DS_KC / 8) * DS_LBO, DS_ASLOT = 2 * DS_APLANE;   // hi + lo of one chunk
constexpr int DS_RS = 2, DS_OS = 3;       // raw-tile ring, operand-chunk ring
constexpr int DS_CONV_WARPS = 8, DS_EPI_WARPS = 16;   // epilogue: four warps per TMEM lane quadrant, a quarter of the columns each
constexpr int DS_EPI_SPLIT = DS_EPI_WARPS / 4;        // (the exact-erf GELU of proj_1 is a long dependent chain: ncu showed the epilogue pacing the kernel)
constexpr int DS_THREADS = (4 + DS_CONV_WARPS + DS_EPI_WARPS) * 32;   // 896

struct DenseStreamArgs {
    const float *X;      // [M][K] contiguous rows
    const uint8_t *Bp;   // deform3d_ps_pack layout, taps = 1: [chunk][4 planes][hi NT rows | lo NT rows][8 bf16]
    const float *bias;   // [N] or null
    const float *E;      // epilogue operand [M][ldE] (EPI_MUL / EPI_ADD)
    float *Y;            // [M][ldY]
    i64 M;
    int K, N, NT, ldE, ldY, epi, vec32, vec16e;   // vec32: 32-byte epilogue accesses; vec16e: E rows 16-byte aligned
    i64 tiles;
};

__global__ void __launch_bounds__(DS_THREADS, 1) dense_stream_kernel(const DenseStreamArgs a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int K = a.K, NT = a.NT, nch = K / DS_KC;
    const int RAW = 128 * K * 4;                                   // bytes of one raw tile
    const int B_LBO = 2 * NT * 16, B_SLOT = (DS_KC / 8) * B_LBO;   // weight chunk: 4 planes of (NT hi rows | NT lo rows)
    uint8_t *sRaw = smem;
    uint8_t *sA = sRaw + DS_RS * RAW;
    uint8_t *sB = sA + DS_OS * DS_ASLOT;
    float *sBias = reinterpret_cast<float *>(sB + nch * B_SLOT);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sBias + 128);
    constexpr int NBARS = 2 * DS_RS + 2 * DS_OS + 4 + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + NBARS);
    const uint32_t bar0 = smem_u32(bars);
    auto rawFull = [&](int s) { return bar0 + 8u * s; };
    auto rawEmpty = [&](int s) { return bar0 + 8u * (DS_RS + s); };
    auto opFull = [&](int s) { return bar0 + 8u * (2 * DS_RS + s); };
    auto opEmpty = [&](int s) { return bar0 + 8u * (2 * DS_RS + DS_OS + s); };
    auto accFull = [&](int s) { return bar0 + 8u * (2 * DS_RS + 2 * DS_OS + s); };
    auto accEmpty = [&](int s) { return bar0 + 8u * (2 * DS_RS + 2 * DS_OS + 2 + s); };
    const uint32_t bFull = bar0 + 8u * (2 * DS_RS + 2 * DS_OS + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t acc_stride = 2 * NT <= 128 ? 128u : 256u;      // two accumulator buffers of 2 NT columns
    const uint32_t tmem_cols = 2 * acc_stride;
    const i64 ntl = a.tiles > (i64)blockIdx.x ? (a.tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    if (tid == 0) {
        for (int s = 0; s < DS_RS; ++s) { mbar_init(rawFull(s), 1); mbar_init(rawEmpty(s), DS_CONV_WARPS); }
        for (int s = 0; s < DS_OS; ++s) { mbar_init(opFull(s), DS_CONV_WARPS); mbar_init(opEmpty(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(accFull(s), 1); mbar_init(accEmpty(s), DS_EPI_WARPS); }
        mbar_init(bFull, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(tmem_slot), tmem_cols);
        tmem_relinquish();
    }
    if (tid < 128) sBias[tid] = (a.bias && tid < a.N) ? __ldg(a.bias + tid) : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // =============================================== MMA issuer ===============================================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(128, NT), idesc2 = make_idesc_bf16(128, 2 * NT);
            mbar_wait(bFull, 0);
            uint32_t go = 0;   // running operand-chunk counter
            for (i64 i = 0; i < ntl; ++i) {
                const int buf = (int)(i & 1);
                if (i >= 2) mbar_wait(accEmpty(buf), (uint32_t)((i >> 1) - 1) & 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)buf * acc_stride;
                for (int c = 0; c < nch; ++c, ++go) {
                    const int os = go % DS_OS;
                    mbar_wait(opFull(os), (go / DS_OS) & 1);
                    tc_fence_after();
                    const uint32_t ahi = smem_u32(sA + os * DS_ASLOT), alo = ahi + DS_APLANE, bsl = smem_u32(sB + c * B_SLOT);
#pragma unroll
                    for (int kk = 0; kk < DS_KC / 16; ++kk) {
                        const uint64_t bd = make_smem_desc(bsl + kk * 2 * B_LBO, B_LBO, 128);
                        const uint64_t ad_hi = make_smem_desc(ahi + kk * 2 * DS_LBO, DS_LBO, 128), ad_lo = make_smem_desc(alo + kk * 2 * DS_LBO, DS_LBO, 128);
                        umma_bf16(d_tmem, ad_hi, bd, idesc2, (c | kk) != 0 ? 1u : 0u);   // [hi*hi | hi*lo]
                        umma_bf16(d_tmem, ad_lo, bd, idesc, 1u);                         // += lo*hi
                    }
                    umma_commit(opEmpty(os));
                }
                umma_commit(accFull(buf));
            }
        }
    } else if (warp == 1) {
        // =============================================== loader: weights once, then whole raw tiles ===============================================
        if (elect_one()) {
            mbar_arrive_expect_tx(bFull, (uint32_t)(nch * B_SLOT));
            bulk_g2s(smem_u32(sB), a.Bp, (uint32_t)(nch * B_SLOT), bFull);
            for (i64 i = 0; i < ntl; ++i) {
                const int rs = (int)(i % DS_RS);
                const i64 tile = blockIdx.x + i * gridDim.x, m0 = tile * 128;
                const i64 rows = a.M - m0 < 128 ? a.M - m0 : 128;
                mbar_wait(rawEmpty(rs), (uint32_t)((i / DS_RS) & 1) ^ 1u);
                const uint32_t bytes = (uint32_t)(rows * K * 4);
                mbar_arrive_expect_tx(rawFull(rs), bytes);
                bulk_g2s(smem_u32(sRaw + rs * RAW), a.X + m0 * K, bytes, rawFull(rs));
            }
        }
    } else if (warp >= 4 && warp < 4 + DS_CONV_WARPS) {
        // =============================================== converters: raw fp32 chunk -> bf16 hi / lo operand ===============================================
        // unit = (row, 4 channels): 8 consecutive lanes read one row's 128-byte chunk (conflict-free), 256 threads cover 32 rows
        // per pass, 4 passes per chunk.  Rows beyond M hold stale data: their accumulator rows are never stored.
        const int ct = tid - 128, cq = ct & 7, row0 = ct >> 3;
        const uint32_t boff0 = (uint32_t)((cq >> 1) * DS_LBO + (cq & 1) * 8);
        uint32_t go = 0;
        for (i64 i = 0; i < ntl; ++i) {
            const int rs = (int)(i % DS_RS);
            mbar_wait(rawFull(rs), (uint32_t)((i / DS_RS) & 1));
            const uint8_t *raw = sRaw + rs * RAW + cq * 16;
            for (int c = 0; c < nch; ++c, ++go) {
                const int os = go % DS_OS;
                mbar_wait(opEmpty(os), ((go / DS_OS) & 1) ^ 1);
                uint8_t *slot = sA + os * DS_ASLOT + boff0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int row = row0 + u * 32;
                    const float4 v = *reinterpret_cast<const float4 *>(raw + (size_t)row * K * 4 + c * 128);
                    uint2 hi, lo;
                    split_bf16x4(v, hi, lo);
                    *reinterpret_cast<uint2 *>(slot + row * 16) = hi;
                    *reinterpret_cast<uint2 *>(slot + DS_APLANE + row * 16) = lo;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(opFull(os));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(rawEmpty(rs));   // this warp has read its share of the raw tile
        }
    } else if (warp >= 4 + DS_CONV_WARPS) {
        // =============================================== epilogue: DS_EPI_SPLIT threads per accumulator row (column groups) ===============================================
        const int q = warp & 3, row = q * 32 + lane, grp = (warp - 4 - DS_CONV_WARPS) >> 2;
        const int per = ((NT / 8 + DS_EPI_SPLIT - 1) / DS_EPI_SPLIT) * 8;                  // columns per group (multiple of 8)
        const int c_begin = grp * per < NT ? grp * per : NT, c_end = c_begin + per < NT ? c_begin + per : NT;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool vec_y = (a.ldY & 3) == 0;
        for (i64 i = 0; i < ntl; ++i) {
            const int buf = (int)(i & 1);
            const i64 m = (blockIdx.x + i * gridDim.x) * 128 + row;
            const bool live = m < a.M;
            mbar_wait_sleep(accFull(buf), (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            const uint32_t tacc = tlane + (uint32_t)buf * acc_stride;
            float *yp = a.Y + (live ? m : 0) * a.ldY;
            const float *ep = a.E ? a.E + (live ? m : 0) * a.ldE : nullptr;
            for (int c0 = c_begin; c0 < c_end; c0 += 8) {
                float v[8], x[8], o[8];
                float4 e0 = f4zero(), e1 = f4zero();
                if (ep && live && c0 + 7 < a.N && a.vec16e) {
                    if (a.vec32) ldg8_stream(ep + c0, e0, e1);
                    else { e0 = ldg4_stream(ep + c0); e1 = ldg4_stream(ep + c0 + 4); }
                } else if (ep && live) {
                    float t[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) t[e] = c0 + e < a.N ? __ldg(ep + c0 + e) : 0.f;
                    e0 = make_float4(t[0], t[1], t[2], t[3]); e1 = make_float4(t[4], t[5], t[6], t[7]);
                }
                tmem_ld8(tacc + c0, v);
                tmem_ld8(tacc + NT + c0, x);
                const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float t = (v[e] + x[e]) + sBias[c0 + e];
                    if (a.epi == EPI_GELU) t = gelu_erf(t);
                    else if (a.epi == EPI_MUL) t *= ev[e];
                    else if (a.epi == EPI_ADD) t += ev[e];
                    o[e] = t;
                }
                if (!live) continue;
                if (a.vec32 && c0 + 7 < a.N) {
                    stg8(yp + c0, o);
                } else if (vec_y && c0 + 7 < a.N) {
                    *reinterpret_cast<float4 *>(yp + c0) = make_float4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<float4 *>(yp + c0 + 4) = make_float4(o[4], o[5], o[6], o[7]);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (c0 + e < a.N) yp[c0 + e] = o[e];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accEmpty(buf));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

size_t ds_smem_bytes(int K, int NT)
{
    return (size_t)DS_RS * 128 * K * 4 + (size_t)DS_OS * DS_ASLOT + (size_t)(K / DS_KC) * (DS_KC / 8) * 2 * NT * 16 + 128 * sizeof(float) +
           (2 * DS_RS + 2 * DS_OS + 5) * 8 + 16 + 128;
}

}  // namespace

// rows of X contiguous (ldX == K), K multiple of 32, one N tile of at most 128 columns, shared memory fits; M large enough
// that a persistent launch makes sense is the caller's call.
bool dense_stream_supported(const IgemmArgs &a)
{
    if (a.mode != IGEMM_DENSE || a.ksplit_steps) return false;
    const int K = a.geo.C, N = a.geo.Co;
    if (K % DS_KC != 0 || a.ldX != K || N > 128 || tc_nt(N) < N) return false;
    if (ds_smem_bytes(K, tc_nt(N)) > 225 * 1024) return false;
    if (((uintptr_t)a.X & 15) != 0) return false;
    if ((a.epi == EPI_MUL || a.epi == EPI_ADD) && !a.E) return false;
    return true;
}

// weights packed with deform3d_ps_pack(w, bp, Co, C, 1)
int dense_stream(const IgemmArgs &ga, const void *bp, cudaStream_t st)
{
    if (!dense_stream_supported(ga)) return DLKA_ERR_UNSUPPORTED;
    if (ga.M <= 0) return DLKA_OK;
    DenseStreamArgs a;
    memset(&a, 0, sizeof(a));
    a.X = ga.X; a.Bp = (const uint8_t *)bp; a.bias = ga.bias; a.E = ga.E; a.ldE = ga.ldE; a.Y = ga.Y; a.ldY = ga.ldY; a.M = ga.M;
    a.K = ga.geo.C; a.N = ga.geo.Co; a.NT = tc_nt(ga.geo.Co); a.epi = ga.epi;
    a.tiles = cdiv(ga.M, 128);
    auto al32 = [](const void *p, int ld) { return p == nullptr || (((uintptr_t)p & 31) == 0 && (ld & 7) == 0); };
    a.vec32 = al32(a.Y, a.ldY) && al32(a.E, a.ldE);
    a.vec16e = a.E == nullptr || ((((uintptr_t)a.E) & 15) == 0 && (a.ldE & 3) == 0);
    const size_t smem = ds_smem_bytes(a.K, a.NT);
    static SmemOptIn optin;
    DLKA_TRY(optin.ensure(dense_stream_kernel, smem));
    int dev = 0, sms = 148;
    DLKA_CUDA_TRY(cudaGetDevice(&dev));
    DLKA_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = a.tiles < sms ? (int)a.tiles : sms;
    DLKA_LAUNCH("tc_dense", st, (dense_stream_kernel<<<grid, DS_THREADS, smem, st>>>(a)));
    return DLKA_OK;
}

}  // namespace dlka
