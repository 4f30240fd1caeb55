// dense_persist.cu -- persistent, software-pipelined 1x1 contraction for the streaming projections of the block
// (proj_1 + GELU, conv1 * u, proj_2 + x, conv8 + a: Y[M, N] = epi(X[M, K] . W^T + b), K = C <= 96, N <= 128).
//
// These layers are HBM-bound (read M*K, write M*N floats).  The generic kernel (mma_tc.cu) runs one 128-row tile per CTA
// with its phases in sequence -- load, split, MMA, epilogue -- so an SM has loads in flight only a fraction of the time
// (measured 2.5 TB/s = 38 % of the HBM peak with 2 CTAs per SM).  Here one CTA per SM walks over its tiles with three roles
// running concurrently on different tiles:
//   producers (8 warps)  tile t+1: 12 x LDG.128 per thread in flight, bf16 hi/lo split, A slot (t+1) & 1
//   MMA issuer (1 thread) tile t  : 3 bf16 passes into TMEM accumulator t & 1      (weights stay resident in shared memory)
//   epilogue (8 warps)   tile t-1 : tcgen05.ld, bias / GELU / gate / residual, stores
// so global loads, tensor work and stores of neighbouring tiles overlap.  Same operand layout, packing (tc_pack_weight)
// and arithmetic as the generic kernel: results are bit-identical to it.
//
// STATUS: parity-green (tests/test_parity_gpu.py::test_linear_tokens_large_m_vs_torch with -DDLKA_DENSE_PERSIST) but measured
// SLOWER than the generic kernel at the headline shape (0.77 vs 0.64 ms for proj_1 + GELU): one producer group per SM keeps
// only 48 KB of loads in flight (the 2-CTA generic kernel has 96 KB) and still serialises load -> split -> store per tile.
// Compiled in, not dispatched (enable with -DDLKA_DENSE_PERSIST); a second producer group needs the register budget that
// setmaxnreg can move from the control warps -- round 2.
#include <atomic>
#include <mutex>

#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace dlka {
namespace {

using namespace ptx;

constexpr int DP_LBO = 2048 + 16;     // A plane stride (same bank spread as mma_tc.cu)
constexpr int DP_PROD = 8, DP_EPI = 8;
constexpr int DP_THREADS = (4 + DP_PROD + DP_EPI) * 32;

struct DpArgs {
    IgemmArgs g;
    const uint8_t *Bp;   // tc_pack_weight layout: [hi | lo][KC/8][NT][8 bf16]
    int NT, KC;
    i64 tiles;
};

__global__ void __launch_bounds__(DP_THREADS, 1) dense_persist_kernel(const DpArgs a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int NT = a.NT, KC = a.KC;
    const int A_PLANE = (KC / 8) * DP_LBO, A_SLOT = 2 * A_PLANE;
    const int B_PLANE = (KC / 8) * NT * 16;
    uint8_t *sA = smem;
    uint8_t *sB = sA + 2 * A_SLOT;
    float *sBias = reinterpret_cast<float *>(sB + 2 * B_PLANE);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sBias + 128);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 9);
    const uint32_t bar0 = smem_u32(bars);
    auto fullA = [&](int s) { return bar0 + 8u * s; };
    auto emptyA = [&](int s) { return bar0 + 8u * (2 + s); };
    auto accFull = [&](int s) { return bar0 + 8u * (4 + s); };
    auto accEmpty = [&](int s) { return bar0 + 8u * (6 + s); };
    const uint32_t fullB = bar0 + 64u;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tmem_cols = 2 * NT <= 32 ? 32u : 2 * NT <= 64 ? 64u : 2 * NT <= 128 ? 128u : 256u;
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(fullA(s), DP_PROD); mbar_init(emptyA(s), 1);
            mbar_init(accFull(s), 1); mbar_init(accEmpty(s), DP_EPI);
        }
        mbar_init(fullB, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(tmem_slot), tmem_cols);
        tmem_relinquish();
    }
    if (tid >= 128 && tid < 256) sBias[tid - 128] = (a.g.bias && tid - 128 < a.g.geo.Co) ? __ldg(a.g.bias + tid - 128) : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const i64 M = a.g.M;

    if (warp == 0) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(128, NT);
            mbar_wait(fullB, 0);
            const uint32_t bhi = smem_u32(sB), blo = bhi + B_PLANE;
            int it = 0;
            for (i64 tile = blockIdx.x; tile < a.tiles; tile += gridDim.x, ++it) {
                const int s = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                mbar_wait(fullA(s), ph);
                mbar_wait(accEmpty(s), ph ^ 1);
                tc_fence_after();
                const uint32_t ahi = smem_u32(sA + s * A_SLOT), alo = ahi + A_PLANE;
                const uint32_t d_tmem = tmem_base + (uint32_t)(s * NT);
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t ab = pass == 1 ? alo : ahi, bb = pass == 2 ? blo : bhi;
                    for (int kk = 0; kk < KC / 16; ++kk) {
                        const uint64_t ad = make_smem_desc(ab + kk * 2 * DP_LBO, DP_LBO, 128);
                        const uint64_t bd = make_smem_desc(bb + kk * 2 * NT * 16, NT * 16, 128);
                        umma_bf16(d_tmem, ad, bd, idesc, (pass | kk) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(emptyA(s));
                umma_commit(accFull(s));
            }
        }
    } else if (warp == 1) {
        // ===================== weights: loaded once, resident for every tile =====================
        if (elect_one()) {
            mbar_arrive_expect_tx(fullB, (uint32_t)(2 * B_PLANE));
            bulk_g2s(smem_u32(sB), a.Bp, (uint32_t)(2 * B_PLANE), fullB);
        }
    } else if (warp >= 4 && warp < 4 + DP_PROD) {
        // ===================== A producers =====================
        const int ptid = tid - 128;
        const int CG = KC / 4, UNITS = 128 * CG, PER = UNITS / (DP_PROD * 32);   // 12 / 8 / 4 float4 per thread
        int it = 0;
        for (i64 tile = blockIdx.x; tile < a.tiles; tile += gridDim.x, ++it) {
            const int s = it & 1;
            const i64 m0 = tile * 128;
            float4 v[12];
#pragma unroll
            for (int ui = 0; ui < 12; ++ui) {
                if (ui < PER) {
                    const int u = ptid + ui * (DP_PROD * 32);
                    const int row = u / CG, cg = u - row * CG;
                    const i64 m = m0 + row;
                    v[ui] = m < M ? ldg4(a.g.X + m * (i64)a.g.ldX + cg * 4) : f4zero();
                }
            }
            mbar_wait(emptyA(s), ((it >> 1) & 1) ^ 1);
            uint8_t *slot = sA + s * A_SLOT;
#pragma unroll
            for (int ui = 0; ui < 12; ++ui) {
                if (ui < PER) {
                    const int u = ptid + ui * (DP_PROD * 32);
                    const int row = u / CG, cg = u - row * CG;
                    uint2 hi, lo;
                    split_bf16x4(v[ui], hi, lo);
                    const int boff = (cg >> 1) * DP_LBO + row * 16 + (cg & 1) * 8;
                    *reinterpret_cast<uint2 *>(slot + boff) = hi;
                    *reinterpret_cast<uint2 *>(slot + A_PLANE + boff) = lo;
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(fullA(s));
        }
    } else if (warp >= 4 + DP_PROD) {
        // ===================== epilogue: 2 warps per TMEM lane quadrant, interleaved 16-column chunks =====================
        const int q = warp & 3, cc = (warp - 4 - DP_PROD) >> 2;   // cc = 0 / 1
        const int Nvalid = a.g.geo.Co;
        const bool vec_y = (a.g.ldY & 3) == 0, vec_e = (a.g.ldE & 3) == 0;
        const bool has_e = a.g.epi == EPI_MUL || a.g.epi == EPI_ADD;
        int it = 0;
        for (i64 tile = blockIdx.x; tile < a.tiles; tile += gridDim.x, ++it) {
            const int s = it & 1;
            const i64 m = tile * 128 + q * 32 + lane;
            const bool mv = m < M;
            float4 ev[4][4];
            if (has_e) {   // operand rows first: their latency hides behind the accumulator wait
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) {
                    const int c0 = (cc + ci * 2) * 16;
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const int n = c0 + j4 * 4;
                        float4 t = f4zero();
                        if (mv && c0 < NT && n < Nvalid) {
                            const float *ep = a.g.E + m * (i64)a.g.ldE + n;
                            if (vec_e && n + 3 < Nvalid) t = ldg4(ep);
                            else {
                                t.x = __ldg(ep);
                                if (n + 1 < Nvalid) t.y = __ldg(ep + 1);
                                if (n + 2 < Nvalid) t.z = __ldg(ep + 2);
                                if (n + 3 < Nvalid) t.w = __ldg(ep + 3);
                            }
                        }
                        ev[ci][j4] = t;
                    }
                }
            }
            mbar_wait(accFull(s), (it >> 1) & 1);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * NT);
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) {
                const int c0 = (cc + ci * 2) * 16;
                if (c0 >= NT) break;
                float v[16];
                tmem_ld16(trow + c0, v);
                if (!mv) continue;
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const int n = c0 + j4 * 4;
                    if (n >= Nvalid) break;
                    const float4 bv = *reinterpret_cast<const float4 *>(sBias + n);
                    float o[4] = {v[j4 * 4] + bv.x, v[j4 * 4 + 1] + bv.y, v[j4 * 4 + 2] + bv.z, v[j4 * 4 + 3] + bv.w};
                    if (a.g.epi == EPI_GELU) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) o[e] = gelu_erf(o[e]);
                    } else if (has_e) {
                        const float4 t = ev[ci][j4];
                        const float e4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) o[e] = a.g.epi == EPI_MUL ? o[e] * e4[e] : o[e] + e4[e];
                    }
                    float *yp = a.g.Y + m * (i64)a.g.ldY + n;
                    if (vec_y && n + 3 < Nvalid) {
                        *reinterpret_cast<float4 *>(yp) = make_float4(o[0], o[1], o[2], o[3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (n + e < Nvalid) yp[e] = o[e];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accEmpty(s));   // this warp has read its share of accumulator s
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

}  // namespace

bool dense_persist_supported(const IgemmArgs &a)
{
    if (a.mode != IGEMM_DENSE || a.ksplit_steps != 0 || !tc_supported(a)) return false;
    const int C = a.geo.C, Co = a.geo.Co;
    if (tc_kc(C) != C || C > 96) return false;          // single K step
    if (Co > 128) return false;                          // single N tile: the weights stay resident
    return a.M >= 128 * 148;                             // enough tiles for a persistent grid to pay off
}

int dense_persist(const IgemmArgs &g, const void *bp, cudaStream_t st)
{
    if (!dense_persist_supported(g)) return DLKA_ERR_UNSUPPORTED;
    DpArgs a;
    a.g = g;
    a.Bp = (const uint8_t *)bp;
    a.KC = g.geo.C;
    a.NT = tc_nt(g.geo.Co);
    a.tiles = cdiv(g.M, 128);
    const size_t smem = 2 * 2 * (size_t)(a.KC / 8) * DP_LBO + 2 * (size_t)(a.KC / 8) * a.NT * 16 + 128 * sizeof(float) + 9 * 8 + 16 + 128;
    static std::atomic<size_t> configured{0};
    static std::mutex configure_lock;
    if (smem > configured.load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> guard(configure_lock);
        if (smem > configured.load(std::memory_order_relaxed)) {
            DLKA_CUDA_TRY(cudaFuncSetAttribute(dense_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured.store(smem, std::memory_order_release);
        }
    }
    const int grid = (int)(a.tiles < 148 ? a.tiles : 148);
    DLKA_LAUNCH("tc_dense_persist", st, (dense_persist_kernel<<<grid, DP_THREADS, smem, st>>>(a)));
    return DLKA_OK;
}

}  // namespace dlka
