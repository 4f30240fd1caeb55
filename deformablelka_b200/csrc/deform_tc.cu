// deform_tc.cu -- 3D deformable convolution (groups 1, dg 1, stride 1) on tcgen05, L1-resident gather.
//
// Replaces the hot loop of D3D.deform_conv_forward (deformable_im2col_gpu_kernel + at::addmm,
// 3D/dcn/src/cuda/deform_im2col_cuda.cuh:192-265, deform_conv_cuda.cu:113-119) without any im2col buffer.
//
// Design (what the first ncu pass taught): the trilinear gather reads 8 corners x 27 taps x C channels per
// output voxel; served from L2 it saturates the L2 fabric (~12 TB/s measured).  So
//   * a CTA owns a compact 4 x 4 x 8 brick of output voxels (128 MMA rows), not 128 consecutive voxels,
//   * K is walked chunk-major: 32 channels (= one 128-byte line per voxel) over all taps, then the next 32,
//     so the working set of a pass (brick + halo, one line per voxel, ~100 KB) stays in L1,
//   * shared memory is kept under ~136 KB so the L1 carve-out stays large.
// Roles (default, L1 path): warp 0 MMA issuer, warp 1 weight loader (cp.async.bulk), warps 4-7 sample-parameter producers
// (positions, clamped corner offsets, masked trilinear weights: 64 B per (row, tap)), warps 8-23 gather / blend / convert
// producers that fill the UMMA A slots (two groups on alternate K steps); all 20 producer warps then run the fused
// epilogue chain (+bias -> conv1 -> * u -> proj_2 -> + x on the accumulator tile).
// (Round 1 also carried a TMA-staged shared-memory gather variant, measured slower, 13.8 vs 12.3 ms; removed in round 2 --
// the persistent kernel deform_ps.cu supersedes both for C <= 96; this kernel serves C > 96 and the K-split small volumes.)
#include <cuda_bf16.h>

#include <cstring>

#include <atomic>
#include <mutex>

#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace dlka {
namespace {

using namespace ptx;

constexpr int DF_KC = 32;                 // channels per K step (one 128-byte line per voxel)
#ifndef DLKA_DF_SA
#define DLKA_DF_SA 3
#define DLKA_DF_SB 3
#define DLKA_DF_SP 4
#endif
constexpr int DF_SA = DLKA_DF_SA, DF_SB = DLKA_DF_SB, DF_SP = DLKA_DF_SP;
constexpr int DF_LBO = 2048 + 32;         // A plane stride (bank-spread padding; see profiles/r01 notes)
constexpr int DF_APLANE = (DF_KC / 8) * DF_LBO;
constexpr int DF_ASLOT = 2 * DF_APLANE;
constexpr int DF_PARAM_WARPS = 4;
constexpr int DF_GATHER_WARPS = 16;
constexpr int DF_THREADS = (4 + DF_PARAM_WARPS + DF_GATHER_WARPS) * 32;
constexpr int DF_EPI_WARPS = DF_PARAM_WARPS + DF_GATHER_WARPS, DF_EWQ = DF_EPI_WARPS / 4;   // epilogue warps, per TMEM lane quadrant
#ifndef DLKA_DF_GROUPS
#define DLKA_DF_GROUPS 2   // gather warp groups taking alternate K steps (1: all 16 warps on every K step)
#endif
constexpr int DF_GROUPS = DLKA_DF_GROUPS;
constexpr int DF_GW = DF_GATHER_WARPS / DF_GROUPS;   // warps per group
constexpr int DF_GT = DF_GW * 32;                     // threads per group
constexpr int DF_UNITS = 128 * 8 / DF_GT;             // (row, float4) units per thread and K step
static_assert(DF_GROUPS == 1 || DF_GROUPS == 2, "gather groups");
constexpr int DF_BD = 4, DF_BH = 4, DF_BW = 8;  // brick
constexpr int DF_PSTRIDE = 5;                                      // int4 per parameter record (80 B: bank spread)

struct DeformTcArgs {
    ConvGeo g;
    const float *X;      // [B][D][H][W][C]
    const float *Off;    // [M][ldOff]
    int ldOff;
    const uint8_t *Bp;   // [n_tile][chunk][tap][hi|lo][4 planes][NT][8 bf16]
    const float *bias;
    float *Y;            // [M][ldY]
    int ldY;
    int NT;
    int tiles_d, tiles_h, tiles_w;
    i64 vol_c;
    // fused epilogue chain (C == Co <= 96): stage 1 = conv1 (1x1) then gate with U; stage 2 = proj_2 (1x1) + residual R
    int chain;            // 0 none, 1 conv1+gate, 2 conv1+gate+proj_2+residual
    const uint8_t *W1p;   // tc_pack_weight layout, KC = C
    const float *b1;
    const float *U;
    int ldU;
    const uint8_t *W2p;
    const float *b2;
    const float *R;
    int ldR;
    int chunk_split;      // > 0: grid.z slices of `chunk_split` channel chunks each; slice z writes its partial sum to Y + z * ysplit
    i64 ysplit;           // (bias only in slice 0); the host reduces the slices
};

struct DfRow {
    int b, d, h, w;
    int m;       // linear output row or -1
    int pad[3];
};

// 8 corner element offsets (clamped) + 8 masked trilinear weights of one (row, tap)
__device__ __forceinline__ void df_make_params(const DeformTcArgs &a, const DfRow &ri, int ii, int jj, int kk, float od, float oh,
                                               float ow, int4 *prm)
{
    const ConvGeo &g = a.g;
    int4 o0 = make_int4(0, 0, 0, 0), o1 = o0;
    float4 w0 = f4zero(), w1 = f4zero();
    if (ri.m >= 0) {
        const float pd = sample_pos(ri.d, g.sd, g.pd, ii, g.dd, od);
        const float ph = sample_pos(ri.h, g.sh, g.ph, jj, g.dh, oh);
        const float pw = sample_pos(ri.w, g.sw, g.pw, kk, g.dw, ow);
        const Sample3 s = make_sample3(pd, ph, pw, g.D, g.H, g.W);
        if (s.mask & 1) {
            const float ld = s.l[0], lh = s.l[1], lw = s.l[2], hd = 1.f - ld, hh = 1.f - lh, hw = 1.f - lw;
            const int d0 = max(s.lo[0], 0), d1 = min(s.lo[0] + 1, g.D - 1);
            const int h0 = max(s.lo[1], 0), h1 = min(s.lo[1] + 1, g.H - 1);
            const int x0 = max(s.lo[2], 0), x1 = min(s.lo[2] + 1, g.W - 1);
            const int sX = g.C, sH = g.W * sX, sD = g.H * sH;
            o0.x = d0 * sD + h0 * sH + x0 * sX; o0.y = d0 * sD + h0 * sH + x1 * sX;
            o0.z = d0 * sD + h1 * sH + x0 * sX; o0.w = d0 * sD + h1 * sH + x1 * sX;
            o1.x = d1 * sD + h0 * sH + x0 * sX; o1.y = d1 * sD + h0 * sH + x1 * sX;
            o1.z = d1 * sD + h1 * sH + x0 * sX; o1.w = d1 * sD + h1 * sH + x1 * sX;
            w0.x = (s.mask & (1 << 1)) ? hd * hh * hw : 0.f; w0.y = (s.mask & (1 << 2)) ? hd * hh * lw : 0.f;
            w0.z = (s.mask & (1 << 3)) ? hd * lh * hw : 0.f; w0.w = (s.mask & (1 << 4)) ? hd * lh * lw : 0.f;
            w1.x = (s.mask & (1 << 5)) ? ld * hh * hw : 0.f; w1.y = (s.mask & (1 << 6)) ? ld * hh * lw : 0.f;
            w1.z = (s.mask & (1 << 7)) ? ld * lh * hw : 0.f; w1.w = (s.mask & (1 << 8)) ? ld * lh * lw : 0.f;
        }
    }
    prm[0] = o0; prm[1] = o1;
    *reinterpret_cast<float4 *>(prm + 2) = w0;
    *reinterpret_cast<float4 *>(prm + 3) = w1;
}

__global__ void __launch_bounds__(DF_THREADS, 1) deform3d_tc_kernel(const DeformTcArgs a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const ConvGeo &g = a.g;
    const int NT = a.NT;
    const int B_PLANE = (DF_KC / 8) * NT * 16, B_SLOT = 2 * B_PLANE;
    uint8_t *sA = smem;
    uint8_t *sB = sA + DF_SA * DF_ASLOT;
    int4 *sPrm = reinterpret_cast<int4 *>(sB + DF_SB * B_SLOT);                 // [SP][128][PSTRIDE]
    DfRow *sRow = reinterpret_cast<DfRow *>(sPrm + DF_SP * 128 * DF_PSTRIDE);   // [128]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sRow + 128);
    constexpr int NBARS = 2 * DF_SA + 2 * DF_SB + 2 * DF_SP + 1 + 6;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + NBARS);
    float *sBias = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~(uintptr_t)15);                 // [3][128]: conv bias, conv1 bias, proj_2 bias (0 beyond Co)
    uint8_t *sChainW = sB;   // chain weights reuse the weight ring once the main loop is done
    const uint32_t bar0 = smem_u32(bars);
    auto fullA = [&](int s) { return bar0 + 8u * s; };
    auto emptyA = [&](int s) { return bar0 + 8u * (DF_SA + s); };
    auto fullB = [&](int s) { return bar0 + 8u * (2 * DF_SA + s); };
    auto emptyB = [&](int s) { return bar0 + 8u * (2 * DF_SA + DF_SB + s); };
    auto fullP = [&](int s) { return bar0 + 8u * (2 * DF_SA + 2 * DF_SB + s); };
    auto emptyP = [&](int s) { return bar0 + 8u * (2 * DF_SA + 2 * DF_SB + DF_SP + s); };
    const uint32_t accFull = bar0 + 8u * (NBARS - 7);
    const uint32_t barW1 = accFull + 8, barE1 = accFull + 16, barC1 = accFull + 24, barW2 = accFull + 32, barE2 = accFull + 40,
                   barC2 = accFull + 48;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tile = blockIdx.y;
    int bid = blockIdx.x;
    const int tw = bid % a.tiles_w; bid /= a.tiles_w;
    const int th = bid % a.tiles_h; bid /= a.tiles_h;
    const int td = bid % a.tiles_d;
    const int b = bid / a.tiles_d;
    const int nchunks = a.chunk_split > 0 ? a.chunk_split : g.C / DF_KC, K = g.K, KS = nchunks * K;
    const int chunk0 = a.chunk_split > 0 ? (int)blockIdx.z * a.chunk_split : 0;   // first channel chunk of this K slice
    const int tc_need = a.chain ? 2 * NT : NT;
    const uint32_t tmem_cols = tc_need <= 32 ? 32u : tc_need <= 64 ? 64u : tc_need <= 128 ? 128u : 256u;
    const int CP = g.C / 8;                               // chain: 16-byte K chunks per row
    const uint32_t chainA_lo = (uint32_t)CP * DF_LBO;      // byte offset of the lo plane set in sA (chain layout)
    const uint32_t chainB_lo = (uint32_t)CP * NT * 16;     // byte offset of the lo half in sB (chain layout)

    if (tid == 0) {
        for (int s = 0; s < DF_SA; ++s) { mbar_init(fullA(s), DF_GW); mbar_init(emptyA(s), 1); }
        for (int s = 0; s < DF_SB; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
        for (int s = 0; s < DF_SP; ++s) { mbar_init(fullP(s), DF_PARAM_WARPS); mbar_init(emptyP(s), DF_GW); }
        mbar_init(accFull, 1);
        mbar_init(barW1, 1); mbar_init(barE1, DF_EPI_WARPS); mbar_init(barC1, 1);
        mbar_init(barW2, 1); mbar_init(barE2, DF_EPI_WARPS); mbar_init(barC2, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(tmem_slot), tmem_cols);
        tmem_relinquish();
    }
    if (warp >= 8) {  // bias vectors of the three GEMM stages (the gather threads are idle during setup)
        for (int i = tid - 256; i < 3 * 128; i += DF_GATHER_WARPS * 32) {
            const int st = i >> 7, n = n_tile * NT + (i & 127);
            const float *src = st == 0 ? a.bias : st == 1 ? a.b1 : a.b2;
            sBias[i] = (src && (i & 127) < NT && n < g.Co && (st == 0 || st <= a.chain) && chunk0 == 0) ? __ldg(src + n) : 0.f;
        }
    }
    if (warp >= 4 && warp < 8) {  // brick row decode
        const int r = tid - 128;
        const int d = td * DF_BD + (r >> 5), h = th * DF_BH + ((r >> 3) & 3), w = tw * DF_BW + (r & 7);
        DfRow ri;
        ri.b = b; ri.d = d; ri.h = h; ri.w = w;
        ri.m = (d < g.Do && h < g.Ho && w < g.Wo) ? (int)((((i64)b * g.Do + d) * g.Ho + h) * g.Wo + w) : -1;
        ri.pad[0] = ri.pad[1] = ri.pad[2] = 0;
        sRow[r] = ri;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto epilogue = [&]() {
        // ===================== epilogue: all producer warps (parameter + gather) =====================
        // TMEM lane quadrant q = warp % 4 (hardware rule); the 5 warps of a quadrant split the columns in chunks of 8
        // (chunk it of this thread = columns (cc + it*DF_EWQ)*8 ...).  The gate (U) / residual (R) rows of the NEXT stage are
        // fetched before this thread waits for that stage's MMA, so their HBM latency hides behind the 1x1 GEMM, and the
        // bias vectors sit in shared memory: no dependent global load is left inside the per-chunk loop.
        constexpr int EP_MAXIT = (16 + DF_EWQ - 1) / DF_EWQ;   // NT <= 128 -> 16 chunks of 8 columns over DF_EWQ warps per quadrant
        const int q = warp & 3, cc = (warp - 4) >> 2;  // cc = 0..DF_EWQ-1
        const int row = q * 32 + lane;
        const DfRow ro = sRow[row];
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool vec_y = (a.ldY & 3) == 0;
        float4 e0[EP_MAXIT], e1[EP_MAXIT];
        auto fetch_operand = [&](const float *src, int ld) {
#pragma unroll
            for (int it = 0; it < EP_MAXIT; ++it) {
                const int c0 = cc * 8 + it * 8 * DF_EWQ, nb = n_tile * NT + c0;
                e0[it] = e1[it] = f4zero();
                if (c0 < NT && ro.m >= 0 && nb < g.Co) {
                    e0[it] = ldg4(src + (i64)ro.m * ld + nb);
                    e1[it] = ldg4(src + (i64)ro.m * ld + nb + 4);
                }
            }
        };
        mbar_wait(accFull, 0);
        tc_fence_after();
        for (int stage = 0; stage <= a.chain; ++stage) {
            if (stage == 1) mbar_wait(barC1, 0);
            if (stage == 2) mbar_wait(barC2, 0);
            if (stage) tc_fence_after();
            const bool last = stage == a.chain;
            const float *sb = sBias + stage * 128;
            const uint32_t tcol = trow + (stage == 1 ? (uint32_t)NT : 0u);
#pragma unroll
            for (int it = 0; it < EP_MAXIT; ++it) {
                const int c0 = cc * 8 + it * 8 * DF_EWQ;
                if (c0 >= NT) break;
                float v[8];
                tmem_ld8(tcol + c0, v);
                const int nb = n_tile * NT + c0;
                const bool live = ro.m >= 0 && nb < g.Co;
                const float4 b0 = *reinterpret_cast<const float4 *>(sb + c0), b1 = *reinterpret_cast<const float4 *>(sb + c0 + 4);
                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                const float ev[8] = {e0[it].x, e0[it].y, e0[it].z, e0[it].w, e1[it].x, e1[it].y, e1[it].z, e1[it].w};
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float t = live ? v[e] + bv[e] : 0.f;
                    if (stage == 1) t *= ev[e];
                    else if (stage == 2) t += ev[e];
                    o[e] = t;
                }
                if (last) {
                    if (live) {
                        float *yp = a.Y + (i64)blockIdx.z * a.ysplit + (i64)ro.m * a.ldY + nb;
                        if (vec_y && nb + 7 < g.Co) {
                            *reinterpret_cast<float4 *>(yp) = make_float4(o[0], o[1], o[2], o[3]);
                            *reinterpret_cast<float4 *>(yp + 4) = make_float4(o[4], o[5], o[6], o[7]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                if (nb + e < g.Co) yp[e] = o[e];
                        }
                    }
                } else {  // restage as the next GEMM's A operand: this thread's row, 16-byte chunk nb/8
                    uint2 hi0, lo0, hi1, lo1;
                    split_bf16x4(make_float4(o[0], o[1], o[2], o[3]), hi0, lo0);
                    split_bf16x4(make_float4(o[4], o[5], o[6], o[7]), hi1, lo1);
                    const int boff = (nb >> 3) * DF_LBO + row * 16;
                    *reinterpret_cast<uint4 *>(sA + boff) = make_uint4(hi0.x, hi0.y, hi1.x, hi1.y);
                    *reinterpret_cast<uint4 *>(sA + chainA_lo + boff) = make_uint4(lo0.x, lo0.y, lo1.x, lo1.y);
                }
            }
            if (!last) {
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(stage == 0 ? barE1 : barE2);
                // operand of the next stage: issued now, consumed after that stage's MMA has been waited for
                if (stage == 0) fetch_operand(a.U, a.ldU);
                else fetch_operand(a.R, a.ldR);
            }
        }
    };

    // Roles by warpgroup
    if (warp < 4) {
    if (warp == 0) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(128, NT);
            for (int ks = 0; ks < KS; ++ks) {
                const int bs = ks % DF_SB, as = ks % DF_SA;
                mbar_wait(fullB(bs), (ks / DF_SB) & 1);
                mbar_wait(fullA(as), (ks / DF_SA) & 1);
                tc_fence_after();
                const uint32_t ahi = smem_u32(sA + as * DF_ASLOT), alo = ahi + DF_APLANE;
                const uint32_t bhi = smem_u32(sB + bs * B_SLOT), blo = bhi + B_PLANE;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t ab = pass == 1 ? alo : ahi, bb = pass == 2 ? blo : bhi;
#pragma unroll
                    for (int kk = 0; kk < DF_KC / 16; ++kk) {
                        const uint64_t ad = make_smem_desc(ab + kk * 2 * DF_LBO, DF_LBO, 128);
                        const uint64_t bd = make_smem_desc(bb + kk * 2 * NT * 16, NT * 16, 128);
                        umma_bf16(tmem_base, ad, bd, idesc, (ks | pass | kk) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(emptyA(as));
                umma_commit(emptyB(bs));
            }
            umma_commit(accFull);
            // ---- fused 1x1 chain: A = epilogue-written rows in sA, B = whole weight matrix in sB ----
            for (int stage = 1; stage <= a.chain; ++stage) {
                mbar_wait(stage == 1 ? barW1 : barW2, 0);
                mbar_wait(stage == 1 ? barE1 : barE2, 0);
                tc_fence_after();
                const uint32_t ahi = smem_u32(sA), alo = ahi + chainA_lo, bhi = smem_u32(sChainW), blo = bhi + chainB_lo;
                const uint32_t d_tmem = tmem_base + (stage == 1 ? (uint32_t)NT : 0u);
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t ab = pass == 1 ? alo : ahi, bb = pass == 2 ? blo : bhi;
                    for (int kk = 0; kk < g.C / 16; ++kk) {
                        const uint64_t ad = make_smem_desc(ab + kk * 2 * DF_LBO, DF_LBO, 128);
                        const uint64_t bd = make_smem_desc(bb + kk * 2 * NT * 16, NT * 16, 128);
                        umma_bf16(d_tmem, ad, bd, idesc, (pass | kk) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(stage == 1 ? barC1 : barC2);
            }
        }
    } else if (warp == 1) {
        // ===================== weight loader =====================
        if (elect_one()) {
            const uint8_t *src = a.Bp + ((i64)n_tile * (g.C / DF_KC) + chunk0) * K * B_SLOT;
            for (int ks = 0; ks < KS; ++ks) {
                const int bs = ks % DF_SB;
                mbar_wait(emptyB(bs), ((ks / DF_SB) & 1) ^ 1);
                mbar_arrive_expect_tx(fullB(bs), (uint32_t)B_SLOT);
                bulk_g2s(smem_u32(sB + bs * B_SLOT), src + (i64)ks * B_SLOT, (uint32_t)B_SLOT, fullB(bs));
            }
            if (a.chain) {
                const uint32_t wbytes = 2u * chainB_lo;
                mbar_wait(accFull, 0);                      // every main-loop MMA has retired: sB is free
                mbar_arrive_expect_tx(barW1, wbytes);
                bulk_g2s(smem_u32(sChainW), a.W1p, wbytes, barW1);
                if (a.chain == 2) {
                    mbar_wait(barC1, 0);                    // conv1 MMAs done reading sB
                    mbar_arrive_expect_tx(barW2, wbytes);
                    bulk_g2s(smem_u32(sChainW), a.W2p, wbytes, barW2);
                }
            }
        }
    }
    } else if (warp < 8) {
        // ===================== sample-parameter producers (one thread per brick row) =====================
        const int r = tid - 128;
        const DfRow ri = sRow[r];
        const float *offrow = a.Off + (i64)(ri.m >= 0 ? ri.m : 0) * a.ldOff;
        // the 3 offsets of the NEXT (row, tap) are fetched one iteration ahead: their global latency overlaps this tap's math
        float od = __ldg(offrow), oh = __ldg(offrow + 1), ow = __ldg(offrow + 2);
        int ps = 0, tap = 0, ii = 0, jj = 0, kk = 0;
        uint32_t ph = 1;
        const int pf_ks = KS > 12 ? KS - 12 : 0;
        for (int ks = 0; ks < KS; ++ks) {
            const int ntap = tap + 1 == K ? 0 : tap + 1;
            const float nod = __ldg(offrow + ntap * 3), noh = __ldg(offrow + ntap * 3 + 1), now = __ldg(offrow + ntap * 3 + 2);
            if (ks == pf_ks && a.chain && ri.m >= 0) {
                // pull this row's gate / residual operand of the fused epilogue into L2 while the main loop still runs
                for (int c = 0; c < g.Co; c += 32) {
                    prefetch_l2(a.U + (i64)ri.m * a.ldU + c);
                    if (a.chain == 2) prefetch_l2(a.R + (i64)ri.m * a.ldR + c);
                }
            }
            mbar_wait(emptyP(ps), ph);
            df_make_params(a, ri, ii, jj, kk, od, oh, ow, sPrm + (ps * 128 + r) * DF_PSTRIDE);
            __syncwarp();
            if (lane == 0) mbar_arrive(fullP(ps));
            od = nod; oh = noh; ow = now;
            tap = ntap;
            if (++kk == g.kw) { kk = 0; if (++jj == g.kh) { jj = 0; if (++ii == g.kd) ii = 0; } }
            if (++ps == DF_SP) { ps = 0; ph ^= 1; }
        }
        epilogue();
    } else {
        // ===================== gather / blend / convert producers =====================
        // DF_GROUPS == 2: the 16 warps form two groups that take alternate K steps, so one group's load phase overlaps
        // the other's blend/convert/store phase instead of all 16 warps moving in lock step.
        const int gt = tid - 256;                      // 0..511
        const int grp = (warp - 8) / DF_GW;
        const int ggt = gt - grp * DF_GT;              // thread within its group
        // thread = (row, float4 cg of the 32-channel chunk); a warp instruction touches 4 lines ----
        const int cg = gt & 7;
        const float *Xb = a.X + (i64)b * a.vol_c + cg * 4;
        for (int ks = grp; ks < KS; ks += DF_GROUPS) {
            const int chunk = chunk0 + ks / K;
            const int as = ks % DF_SA, ps = ks % DF_SP;
            const uint32_t phA = ((ks / DF_SA) & 1) ^ 1, phP = (ks / DF_SP) & 1;   // emptyA starts "free", fullP "not ready"
            mbar_wait(fullP(ps), phP);
            mbar_wait(emptyA(as), phA);
            uint8_t *slot = sA + as * DF_ASLOT;
            const float *base = Xb + chunk * DF_KC;
#pragma unroll
            for (int u0 = 0; u0 < DF_UNITS; u0 += 2) {
                float4 acc[2];
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const int row = (ggt >> 3) + (u0 + s) * (DF_GT / 8);
                    const int4 *prm = sPrm + (ps * 128 + row) * DF_PSTRIDE;
                    const int4 o0 = prm[0], o1 = prm[1];
                    const float4 w0 = *reinterpret_cast<const float4 *>(prm + 2), w1 = *reinterpret_cast<const float4 *>(prm + 3);
                    const float4 v0 = ldg4(base + o0.x), v1 = ldg4(base + o0.y), v2 = ldg4(base + o0.z), v3 = ldg4(base + o0.w);
                    const float4 v4 = ldg4(base + o1.x), v5 = ldg4(base + o1.y), v6 = ldg4(base + o1.z), v7 = ldg4(base + o1.w);
                    float4 r = f4zero();
                    fma4(r, w0.x, v0); fma4(r, w0.y, v1); fma4(r, w0.z, v2); fma4(r, w0.w, v3);
                    fma4(r, w1.x, v4); fma4(r, w1.y, v5); fma4(r, w1.z, v6); fma4(r, w1.w, v7);
                    acc[s] = r;
                }
                if (u0 + 2 == DF_UNITS) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(emptyP(ps));    // parameters consumed
                }
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const int row = (ggt >> 3) + (u0 + s) * (DF_GT / 8);
                    uint2 hi, lo;
                    split_bf16x4(acc[s], hi, lo);
                    const int boff = (cg >> 1) * DF_LBO + row * 16 + (cg & 1) * 8;
                    *reinterpret_cast<uint2 *>(slot + boff) = hi;
                    *reinterpret_cast<uint2 *>(slot + DF_APLANE + boff) = lo;
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(fullA(as));
        }
        epilogue();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// weight [Co][C][taps] -> Bp[n_tile][chunk][tap][hi|lo][4 planes][NT][8]
__global__ void pack_weight_df_kernel(const float *__restrict__ w, __nv_bfloat16 *__restrict__ bp, int Co, int C, int taps, int NT,
                                      int n_tiles)
{
    const int nch = C / DF_KC;
    const i64 total = (i64)n_tiles * nch * taps * DF_KC * NT;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        const int e = (int)(i % 8);
        const int n = (int)((i / 8) % NT);
        const int p = (int)((i / (8 * NT)) % (DF_KC / 8));
        const int tap = (int)((i / ((i64)DF_KC * NT)) % taps);
        const int ch = (int)((i / ((i64)DF_KC * NT * taps)) % nch);
        const int nt = (int)(i / ((i64)DF_KC * NT * taps * nch));
        const int c = ch * DF_KC + p * 8 + e, co = nt * NT + n;
        const float v = co < Co ? w[((i64)co * C + c) * taps + tap] : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        const i64 slot = ((i64)(nt * nch + ch) * taps + tap) * (2 * DF_KC * NT);
        bp[slot + ((i64)p * NT + n) * 8 + e] = hi;
        bp[slot + (i64)DF_KC * NT + ((i64)p * NT + n) * 8 + e] = lo;
    }
}

size_t df_smem_bytes(int NT)
{
    return (size_t)DF_SA * DF_ASLOT + (size_t)DF_SB * 2 * (DF_KC / 8) * NT * 16 + (size_t)DF_SP * 128 * DF_PSTRIDE * 16 +
           128 * sizeof(DfRow) + (2 * DF_SA + 2 * DF_SB + 2 * DF_SP + 1 + 6) * 8 + 16 + 3 * 128 * sizeof(float) + 256;
}

}  // namespace

bool deform3d_tc_supported(const IgemmArgs &a)
{
    const ConvGeo &g = a.geo;
    if (a.mode != IGEMM_DEFORM || g.ndim != 3 || g.groups != 1 || g.dg != 1 || a.Mask) return false;
    if (g.C % DF_KC != 0 || a.epi != EPI_NONE) return false;
    if ((i64)g.D * g.H * g.W * g.C >= ((i64)1 << 31)) return false;
    if ((i64)g.B * g.Do * g.Ho * g.Wo >= ((i64)1 << 31)) return false;
    return true;
}

bool deform3d_chain_supported(const IgemmArgs &a)
{
    const ConvGeo &g = a.geo;
    return deform3d_tc_supported(a) && g.C == g.Co && g.C <= 96 && tc_kc(g.C) == g.C && tc_nt(g.Co) == g.Co;
}

int deform3d_tc(const IgemmArgs &ga, const float *w, void *bp, const DeformChain *chain, cudaStream_t st)
{
    if (!deform3d_tc_supported(ga)) return DLKA_ERR_UNSUPPORTED;
    if (chain && chain->stages && !deform3d_chain_supported(ga)) return DLKA_ERR_UNSUPPORTED;
    const ConvGeo &g = ga.geo;
    if (ga.M <= 0) return DLKA_OK;
    DeformTcArgs a;
    a.g = g; a.X = ga.X; a.Off = ga.Off; a.ldOff = ga.ldOff ? ga.ldOff : 3 * g.K; a.bias = ga.bias; a.Y = ga.Y; a.ldY = ga.ldY;
    a.NT = tc_nt(g.Co);
    const int n_tiles = (int)cdiv(g.Co, a.NT);
    a.tiles_d = (int)cdiv(g.Do, DF_BD); a.tiles_h = (int)cdiv(g.Ho, DF_BH); a.tiles_w = (int)cdiv(g.Wo, DF_BW);
    a.vol_c = (i64)g.D * g.H * g.W * g.C;
    a.chunk_split = 0; a.ysplit = 0;
    a.chain = 0; a.W1p = a.W2p = nullptr; a.b1 = a.b2 = a.U = a.R = nullptr; a.ldU = a.ldR = 0;
    if (chain && chain->stages) {
        a.chain = chain->stages;
        a.W1p = (const uint8_t *)chain->W1p; a.b1 = chain->b1; a.U = chain->U; a.ldU = chain->ldU;
        a.W2p = (const uint8_t *)chain->W2p; a.b2 = chain->b2; a.R = chain->R; a.ldR = chain->ldR;
    }
    if (!pack_skipped()) {
        const i64 total = (i64)n_tiles * g.K * g.C * a.NT;
        const int blocks = (int)(cdiv(total, 256) < 148 * 8 ? cdiv(total, 256) : 148 * 8);
        DLKA_LAUNCH("pack_weight_df", st,
                    pack_weight_df_kernel<<<blocks, 256, 0, st>>>(w, (__nv_bfloat16 *)bp, g.Co, g.C, g.K, a.NT, n_tiles));
    }
    a.Bp = (const uint8_t *)bp;
    const size_t smem = df_smem_bytes(a.NT);
    // the attribute is per function and process-wide: keep a process-wide monotonic maximum (a thread_local cache let a second
    // host thread -- e.g. the autograd engine's -- lower the limit under a launch that needs more)
    static SmemOptIn optin;   // per launch site (= per kernel instantiation), per device
    DLKA_TRY(optin.ensure(deform3d_tc_kernel, smem));
    dim3 grid((unsigned)((i64)g.B * a.tiles_d * a.tiles_h * a.tiles_w), (unsigned)n_tiles);
    if (ga.ksplit_steps > 0 && !a.chain) {   // K split over channel chunks (small volumes: a few tiles, a long K loop)
        a.chunk_split = ga.ksplit_steps;
        a.ysplit = ga.ysplit_stride;
        grid.z = (unsigned)((g.C / DF_KC) / a.chunk_split);
    }
    DLKA_LAUNCH(a.chain ? "tc_deform3d_chain" : "tc_deform3d", st, (deform3d_tc_kernel<<<grid, DF_THREADS, smem, st>>>(a)));
    return DLKA_OK;
}

}  // namespace dlka
