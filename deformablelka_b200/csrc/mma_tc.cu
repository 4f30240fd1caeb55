// mma_tc.cu -- tcgen05 tensor-core implicit GEMM over channels-last volumes (DLKA_MATH_BF16X3).
//
//   Y[M x N] = epi( A[M x K] * W^T + bias ),   K = taps * C,  A produced on the fly per tap:
//     dense      : rows of X                       (proj_1 / conv1 / proj_2)
//     conv       : zero-padded shifted rows        (offset nets, conv_offset)
//     deformable : trilinear samples               (3D deformable conv; no im2col buffer in HBM)
//
// Precision: every fp32 operand is split a = hi + lo (two bf16), and the product is accumulated as
// hi*hi + lo*hi + hi*lo in fp32 TMEM accumulators (3 tcgen05.mma per K step): error ~2^-16 relative.
//
// CTA layout ((4 + NPW) warps, 1 CTA / SM):
//   warp 0 (one elected lane)  MMA issuer: waits full barriers, issues tcgen05.mma, commits to empty barriers
//   warp 1 (one elected lane)  weight loader: cp.async.bulk (TMA engine) of pre-packed B tiles -> smem
//   warps 4..                  A producers: gather / convert / st.shared into the UMMA canonical layout,
//                              then (first 4*MT of them) the epilogue: tcgen05.ld -> bias/GELU/gate/residual -> global
// Shared-memory operand layout (K-major, SWIZZLE_NONE): 16-byte K chunk j of row r at
//   j * LBO + (r / 8) * 128 + (r % 8) * 16   -> a 128-row slot is contiguous per chunk (LBO = 2048 + pad).
#include <cuda_bf16.h>

#include <atomic>
#include <mutex>

#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace dlka {
namespace {

using namespace ptx;

#ifndef DLKA_A_PAD
#define DLKA_A_PAD 16   // LBO = 2064: the 8 K-chunk planes a half-warp writes land on 8 different 4-bank groups (2048 gave an 8-way conflict)
#endif
constexpr int A_PAD = DLKA_A_PAD;       // extra bytes on LBO_A (bank spread; descriptors only need 16-byte granularity)
constexpr int LBO_A = 2048 + A_PAD;     // bytes between 16-byte K chunks of an A slot (128 rows)
constexpr int CTRL_WARPS = 4;

struct RowInfo {
    int b, d, h, w;
};

struct TcArgs {
    IgemmArgs g;
    const uint8_t *Bp;  // packed bf16 weights [n_tiles][KS][2 (hi,lo)][KC/8][NT][8]
    int NT;             // N tile (multiple of 16, <= 128)
    int KS;             // K steps = taps * (C / KC)
    i64 vol_c;          // D*H*W*C of one sample (elements)
};

template <int KC>
__host__ __device__ constexpr int a_plane_bytes() { return (KC / 8) * LBO_A; }

// ---- per-mode production of one float4 (4 consecutive channels of one A row) ------------------------
template <int MODE>
__device__ __forceinline__ float4 produce4(const TcArgs &a, const RowInfo &ri, bool row_valid, i64 m, int tap, int c,
                                           const int4 *prm /* deform: 4 x int4 = 8 offsets + 8 weights */)
{
    const ConvGeo &g = a.g.geo;
    if (MODE == IGEMM_DENSE) {
        return row_valid ? ldg4(a.g.X + m * (i64)a.g.ldX + c) : f4zero();
    } else if (MODE == IGEMM_CONV) {
        const int kk = tap % g.kw, jj = (tap / g.kw) % g.kh, ii = tap / (g.kw * g.kh);
        const int d = ri.d * g.sd - g.pd + ii * g.dd, h = ri.h * g.sh - g.ph + jj * g.dh, w = ri.w * g.sw - g.pw + kk * g.dw;
        if (!row_valid || (unsigned)d >= (unsigned)g.D || (unsigned)h >= (unsigned)g.H || (unsigned)w >= (unsigned)g.W) return f4zero();
        return ldg4(a.g.X + ((((i64)ri.b * g.D + d) * g.H + h) * g.W + w) * (i64)g.C + c);
    } else {
        const int4 o0 = prm[0], o1 = prm[1];
        const float4 w0 = *reinterpret_cast<const float4 *>(prm + 2), w1 = *reinterpret_cast<const float4 *>(prm + 3);
        const float *base = a.g.X + (i64)ri.b * a.vol_c + c;
        float4 acc = f4zero();
        fma4(acc, w0.x, ldg4(base + o0.x));
        fma4(acc, w0.y, ldg4(base + o0.y));
        fma4(acc, w0.z, ldg4(base + o0.z));
        fma4(acc, w0.w, ldg4(base + o0.w));
        fma4(acc, w1.x, ldg4(base + o1.x));
        fma4(acc, w1.y, ldg4(base + o1.y));
        fma4(acc, w1.z, ldg4(base + o1.z));
        fma4(acc, w1.w, ldg4(base + o1.w));
        return acc;
    }
}

// deform: sample parameters of one (row, tap): 8 clamped element offsets + 8 weights (invalid corners -> 0)
__device__ __forceinline__ void make_params(const TcArgs &a, const RowInfo &ri, bool row_valid, i64 m, int tap, int4 *prm)
{
    const ConvGeo &g = a.g.geo;
    int4 o0 = make_int4(0, 0, 0, 0), o1 = o0;
    float4 w0 = f4zero(), w1 = f4zero();
    if (row_valid) {
        const int kk = tap % g.kw, jj = (tap / g.kw) % g.kh, ii = tap / (g.kw * g.kh);
        const float *off = a.g.Off + m * (i64)(a.g.ldOff ? a.g.ldOff : 3 * g.K) + tap * 3;
        const float pd = sample_pos(ri.d, g.sd, g.pd, ii, g.dd, __ldg(off));
        const float ph = sample_pos(ri.h, g.sh, g.ph, jj, g.dh, __ldg(off + 1));
        const float pw = sample_pos(ri.w, g.sw, g.pw, kk, g.dw, __ldg(off + 2));
        const Sample3 s = make_sample3(pd, ph, pw, g.D, g.H, g.W);
        if (s.mask & 1) {
            const float ld = s.l[0], lh = s.l[1], lw = s.l[2], hd = 1.f - ld, hh = 1.f - lh, hw = 1.f - lw;
            // clamped corner indices: always inside the volume; the weight decides whether they count
            const int d0 = max(s.lo[0], 0), d1 = min(s.lo[0] + 1, g.D - 1);
            const int h0 = max(s.lo[1], 0), h1 = min(s.lo[1] + 1, g.H - 1);
            const int x0 = max(s.lo[2], 0), x1 = min(s.lo[2] + 1, g.W - 1);
            const int sH = g.W * g.C, sD = g.H * sH;
            o0.x = d0 * sD + h0 * sH + x0 * g.C; o0.y = d0 * sD + h0 * sH + x1 * g.C;
            o0.z = d0 * sD + h1 * sH + x0 * g.C; o0.w = d0 * sD + h1 * sH + x1 * g.C;
            o1.x = d1 * sD + h0 * sH + x0 * g.C; o1.y = d1 * sD + h0 * sH + x1 * g.C;
            o1.z = d1 * sD + h1 * sH + x0 * g.C; o1.w = d1 * sD + h1 * sH + x1 * g.C;
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

// SA / SB: A / B ring depth.  Dense 1x1 launches use SA = MT, SB = 1 (a single K step) so two CTAs fit per SM.
template <int MODE, int KC, int MT, int NPW, int SA, int SB>
__global__ void __launch_bounds__((CTRL_WARPS + NPW) * 32, (MODE == IGEMM_DENSE && MT == 1) ? 2 : 1) tc_igemm_kernel(const TcArgs a)
{
    constexpr int NPT = NPW * 32;
    constexpr int A_PLANE = a_plane_bytes<KC>();
    constexpr int A_SLOT = 2 * A_PLANE;
    constexpr int CG = KC / 4;            // float4 groups per row per K step
    constexpr int UNITS = 128 * CG;       // (row, group) units per A slot
    constexpr int NKC_MAXROWS = MT * 128;

    extern __shared__ __align__(128) uint8_t smem[];
    const int NT = a.NT;
    const int B_PLANE = (KC / 8) * NT * 16, B_SLOT = 2 * B_PLANE;
    uint8_t *sA = smem;
    uint8_t *sB = sA + SA * A_SLOT;
    int4 *sPrm = reinterpret_cast<int4 *>(sB + SB * B_SLOT);                       // [SA][128][4] int4 (deform)
    RowInfo *sRow = reinterpret_cast<RowInfo *>(reinterpret_cast<uint8_t *>(sPrm) + (MODE == IGEMM_DEFORM ? SA * 128 * 64 : 0));
    uint64_t *bars = reinterpret_cast<uint64_t *>(sRow + NKC_MAXROWS);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * SA + 2 * SB + 1);
    float *sBias = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~(uintptr_t)15);   // [128], 0 beyond Co

    const uint32_t bar0 = smem_u32(bars);
    auto fullA = [&](int s) { return bar0 + 8u * s; };
    auto emptyA = [&](int s) { return bar0 + 8u * (SA + s); };
    auto fullB = [&](int s) { return bar0 + 8u * (2 * SA + s); };
    auto emptyB = [&](int s) { return bar0 + 8u * (2 * SA + SB + s); };
    const uint32_t accFull = bar0 + 8u * (2 * SA + 2 * SB);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const i64 m0 = (i64)blockIdx.x * (MT * 128);
    const int n_tile = blockIdx.y;
    // split-K: slice blockIdx.z covers K steps [ks0, ks0 + KS)
    const int ks0 = a.g.ksplit_steps ? (int)blockIdx.z * a.g.ksplit_steps : 0;
    const int KS = a.g.ksplit_steps ? (a.KS - ks0 < a.g.ksplit_steps ? a.KS - ks0 : a.g.ksplit_steps) : a.KS;
    const int nkc = a.g.geo.C / KC;
    const uint32_t tmem_cols = (MT * NT <= 32) ? 32u : (MT * NT <= 64) ? 64u : (MT * NT <= 128) ? 128u : (MT * NT <= 256) ? 256u : 512u;

    if (tid == 0) {
        for (int s = 0; s < SA; ++s) { mbar_init(fullA(s), NPW); mbar_init(emptyA(s), 1); }
        for (int s = 0; s < SB; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
        mbar_init(accFull, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(tmem_slot), tmem_cols);
        tmem_relinquish();
    }
    if (warp >= CTRL_WARPS) {  // bias vector + row decode, once per CTA
        for (int i = tid - CTRL_WARPS * 32; i < 128; i += NPT) {
            const int n = n_tile * a.NT + i;
            sBias[i] = (a.g.bias && blockIdx.z == 0 && i < a.NT && n < a.g.geo.Co) ? __ldg(a.g.bias + n) : 0.f;
        }
        for (int r = tid - CTRL_WARPS * 32; r < MT * 128; r += NPT) {
            i64 m = m0 + r;
            RowInfo ri = {0, 0, 0, 0};
            if (MODE != IGEMM_DENSE && m < a.g.M) {   // dense rows need no coordinates
                const ConvGeo &g = a.g.geo;
                ri.w = (int)(m % g.Wo);
                i64 t = m / g.Wo;
                ri.h = (int)(t % g.Ho);
                t /= g.Ho;
                ri.d = (int)(t % g.Do);
                ri.b = (int)(t / g.Do);
            }
            sRow[r] = ri;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(128, NT);
            for (int ks = 0; ks < KS; ++ks) {
                const int bs = ks % SB;
                mbar_wait(fullB(bs), (ks / SB) & 1);
                const uint32_t bhi = smem_u32(sB + bs * B_SLOT), blo = bhi + B_PLANE;
                for (int h = 0; h < MT; ++h) {
                    const int it = ks * MT + h, as = it % SA;
                    mbar_wait(fullA(as), (it / SA) & 1);
                    tc_fence_after();
                    const uint32_t ahi = smem_u32(sA + as * A_SLOT), alo = ahi + A_PLANE;
                    const uint32_t d_tmem = tmem_base + (uint32_t)(h * NT);
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint32_t ab = pass == 1 ? alo : ahi, bb = pass == 2 ? blo : bhi;
#pragma unroll
                        for (int kk = 0; kk < KC / 16; ++kk) {
                            const uint64_t ad = make_smem_desc(ab + kk * 2 * LBO_A, LBO_A, 128);
                            const uint64_t bd = make_smem_desc(bb + kk * 2 * NT * 16, NT * 16, 128);
                            umma_bf16(d_tmem, ad, bd, idesc, (ks | pass | kk) != 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(emptyA(as));
                }
                umma_commit(emptyB(bs));
            }
            umma_commit(accFull);
        }
    } else if (warp == 1) {
        // ===================== weight loader (bulk async copy) =====================
        if (elect_one()) {
            const uint8_t *src = a.Bp + (i64)n_tile * a.KS * B_SLOT;
            for (int ks = 0; ks < KS; ++ks) {
                const int bs = ks % SB;
                mbar_wait(emptyB(bs), ((ks / SB) & 1) ^ 1);
                mbar_arrive_expect_tx(fullB(bs), (uint32_t)B_SLOT);
                bulk_g2s(smem_u32(sB + bs * B_SLOT), src + (i64)(ks0 + ks) * B_SLOT, (uint32_t)B_SLOT, fullB(bs));
            }
        }
    } else if (warp >= CTRL_WARPS) {
        // ===================== A producers =====================
        const int ptid = tid - CTRL_WARPS * 32;
        for (int ks = 0; ks < KS; ++ks) {
            const int tap = (ks0 + ks) / nkc, kc = (ks0 + ks) - tap * nkc;
            for (int h = 0; h < MT; ++h) {
                const int it = ks * MT + h, as = it % SA;
                mbar_wait(emptyA(as), ((it / SA) & 1) ^ 1);
                if (MODE == IGEMM_DEFORM) {
                    if (ptid < 128) {
                        const i64 m = m0 + h * 128 + ptid;
                        make_params(a, sRow[h * 128 + ptid], m < a.g.M, m, tap, sPrm + (as * 128 + ptid) * 4);
                    }
                    asm volatile("bar.sync 1, %0;" ::"r"(NPT) : "memory");
                }
                uint8_t *slot = sA + as * A_SLOT;
                // dense / conv rows are plain loads: keep every load of the slot in flight (HBM latency is paid once per slot)
                static_assert(UNITS % NPT == 0, "units per producer thread");
#pragma unroll(MODE == IGEMM_DEFORM ? 2 : UNITS / NPT)
                for (int ui = 0; ui < UNITS / NPT; ++ui) {
                    const int u = ptid + ui * NPT;
                    const int row = u / CG, cg = u - row * CG;
                    const i64 m = m0 + h * 128 + row;
                    const float4 v = produce4<MODE>(a, sRow[h * 128 + row], m < a.g.M, m, tap, kc * KC + cg * 4,
                                                    sPrm + (as * 128 + row) * 4);
                    uint2 hi, lo;
                    split_bf16x4(v, hi, lo);
                    const int boff = (cg >> 1) * LBO_A + row * 16 + (cg & 1) * 8;
                    *reinterpret_cast<uint2 *>(slot + boff) = hi;
                    *reinterpret_cast<uint2 *>(slot + A_PLANE + boff) = lo;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(fullA(as));
            }
        }
        // ===================== epilogue: every producer warp =====================
        // TMEM lane quadrant q = warp % 4 (hardware rule); the WPP warps that share a (quadrant, 128-row half) interleave
        // the 16-column chunks.  Bias comes from shared memory and the multiply / add operand rows are fetched BEFORE the
        // accumulator wait, so no dependent global load sits between tcgen05.ld and the store.
        constexpr int WPP = NPW / (4 * MT);             // warps per (quadrant, half)
        constexpr int EP_MAXIT = (8 + WPP - 1) / WPP;    // NT <= 128 -> 8 chunks of 16 columns
        const int pw = warp - CTRL_WARPS;
        const int q = warp & 3, h = (pw >> 2) % MT, cc = (pw >> 2) / MT;
        const i64 m = m0 + h * 128 + q * 32 + lane;
        const bool mv = m < a.g.M;
        const int Nvalid = a.g.geo.Co;
        const bool vec_y = (a.g.ldY & 3) == 0, vec_e = (a.g.ldE & 3) == 0;
        const bool has_e = (a.g.epi == EPI_MUL || a.g.epi == EPI_ADD) && blockIdx.z == 0;
        float *Yz = a.g.Y + (i64)blockIdx.z * a.g.ysplit_stride;
        float4 ev[EP_MAXIT][4];
        if (has_e) {
#pragma unroll
            for (int it = 0; it < EP_MAXIT; ++it) {
                const int nb = n_tile * NT + (cc + it * WPP) * 16;
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const int n = nb + j4 * 4;
                    float4 t = f4zero();
                    if (mv && (cc + it * WPP) * 16 < NT && n < Nvalid) {
                        const float *ep = a.g.E + m * (i64)a.g.ldE + n;
                        if (vec_e && n + 3 < Nvalid) t = ldg4(ep);
                        else {
                            t.x = __ldg(ep);
                            if (n + 1 < Nvalid) t.y = __ldg(ep + 1);
                            if (n + 2 < Nvalid) t.z = __ldg(ep + 2);
                            if (n + 3 < Nvalid) t.w = __ldg(ep + 3);
                        }
                    }
                    ev[it][j4] = t;
                }
            }
        }
        mbar_wait(accFull, 0);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * NT);
#pragma unroll
        for (int it = 0; it < EP_MAXIT; ++it) {
            const int c0 = (cc + it * WPP) * 16;
            if (c0 >= NT) break;
            float v[16];
            tmem_ld16(trow + c0, v);
            if (!mv) continue;
            const int nb = n_tile * NT + c0;
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const int n = nb + j4 * 4;
                if (n >= Nvalid) break;
                const float4 bv = *reinterpret_cast<const float4 *>(sBias + c0 + j4 * 4);
                float o[4] = {v[j4 * 4] + bv.x, v[j4 * 4 + 1] + bv.y, v[j4 * 4 + 2] + bv.z, v[j4 * 4 + 3] + bv.w};
                if (a.g.epi == EPI_GELU) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[e] = gelu_erf(o[e]);
                } else if (has_e) {
                    const float4 t = ev[it][j4];
                    const float e4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[e] = a.g.epi == EPI_MUL ? o[e] * e4[e] : o[e] + e4[e];
                }
                float *yp = Yz + m * (i64)a.g.ldY + n;
                if (vec_y && n + 3 < Nvalid) {
                    *reinterpret_cast<float4 *>(yp) = make_float4(o[0], o[1], o[2], o[3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (n + e < Nvalid) yp[e] = o[e];
                }
            }
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

// weight [Co][C][taps] fp32 -> Bp[n_tile][ks = tap*nkc + kc][hi|lo][j = KC/8][n = NT][8 bf16]
__global__ void pack_weight_tc_kernel(const float *__restrict__ w, __nv_bfloat16 *__restrict__ bp, int Co, int C, int taps,
                                      int KC, int NT, int n_tiles)
{
    const int nkc = C / KC, KS = taps * nkc;
    const i64 per_slot = (i64)2 * KC * NT;  // elements (hi + lo)
    const i64 total = (i64)n_tiles * KS * KC * NT;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        const int e = (int)(i % 8);
        const int n = (int)((i / 8) % NT);
        const int j = (int)((i / (8 * NT)) % (KC / 8));
        const int ks = (int)((i / ((i64)KC * NT)) % KS);
        const int nt = (int)(i / ((i64)KC * NT * KS));
        const int tap = ks / nkc, kc = ks % nkc;
        const int c = kc * KC + j * 8 + e, co = nt * NT + n;
        const float v = co < Co ? w[((i64)co * C + c) * taps + tap] : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        const i64 base = ((i64)nt * KS + ks) * per_slot + ((i64)j * NT + n) * 8 + e;
        bp[base] = hi;
        bp[base + (i64)KC * NT] = lo;
    }
}

template <int MODE, int KC, int MT, int NPW, int SA, int SB>
int launch_tc(const TcArgs &a, int n_tiles, cudaStream_t st)
{
    const int NT = a.NT;
    const size_t smem = (size_t)SA * 2 * a_plane_bytes<KC>() + (size_t)SB * 2 * (KC / 8) * NT * 16 +
                        (MODE == IGEMM_DEFORM ? (size_t)SA * 128 * 64 : 0) + (size_t)MT * 128 * sizeof(RowInfo) +
                        (2 * SA + 2 * SB + 1) * 8 + 16 + 16 + 128 * sizeof(float) + 128;
    auto kern = tc_igemm_kernel<MODE, KC, MT, NPW, SA, SB>;
    // the attribute is per function and process-wide: keep a process-wide monotonic maximum (a thread_local cache let a second
    // host thread -- e.g. the autograd engine's -- lower the limit under a launch that needs more)
    static SmemOptIn optin;   // per launch site (= per kernel instantiation), per device
    DLKA_TRY(optin.ensure(kern, smem));
    dim3 grid((unsigned)cdiv(a.g.M, MT * 128), (unsigned)n_tiles, (unsigned)(a.g.ksplit_steps ? cdiv(a.KS, a.g.ksplit_steps) : 1));
    const char *name = MODE == IGEMM_DENSE ? "tc_dense" : MODE == IGEMM_CONV ? "tc_conv" : "tc_deform";
    DLKA_LAUNCH(name, st, (kern<<<grid, (CTRL_WARPS + NPW) * 32, smem, st>>>(a)));
    return DLKA_OK;
}

template <int MODE, int KC>
int launch_tc_mt(const TcArgs &a, int n_tiles, cudaStream_t st)
{
    constexpr int NPW = MODE == IGEMM_DEFORM ? 16 : 8;
    if (MODE == IGEMM_DENSE && a.KS == 1) return launch_tc<MODE, KC, 1, NPW, 1, 1>(a, n_tiles, st);  // streaming 1x1: 2 CTAs / SM
    const bool big = a.g.M >= (i64)2 * 128 * 148 * 2;
    // two 128-row halves per CTA share every weight tile (halves the L2 -> smem weight traffic)
    if (big) return launch_tc<MODE, KC, 2, NPW, 2, 2>(a, n_tiles, st);
    return launch_tc<MODE, KC, 1, NPW, 2, 2>(a, n_tiles, st);
}

template <int MODE>
int launch_tc_kc(const TcArgs &a, int KC, int n_tiles, cudaStream_t st)
{
    switch (KC) {
    case 32: return launch_tc_mt<MODE, 32>(a, n_tiles, st);
    case 64: return launch_tc_mt<MODE, 64>(a, n_tiles, st);
    case 96: return launch_tc_mt<MODE, 96>(a, n_tiles, st);
    default: return DLKA_ERR_UNSUPPORTED;
    }
}

}  // namespace

bool tc_supported(const IgemmArgs &a)
{
    const ConvGeo &g = a.geo;
    if (g.groups != 1 || g.dg != 1) return false;
    if (g.C % 16 != 0) return false;
    if (a.mode == IGEMM_DEFORM && (g.ndim != 3 || a.Mask != nullptr)) return false;
    if (a.mode == IGEMM_DENSE && (a.ldX % 4 != 0)) return false;
    if ((i64)g.D * g.H * g.W * g.C >= ((i64)1 << 31)) return false;
    const int kc = tc_kc(g.C);
    return kc != 0;
}

int tc_kc(int C)
{
    if (C <= 96) return (C == 32 || C == 64 || C == 96) ? C : (C % 32 == 0 ? 32 : 0);
    if (C % 64 == 0) return 64;
    if (C % 96 == 0) return 96;
    if (C % 32 == 0) return 32;
    return 0;
}

int tc_nt(int Co) { return Co >= 128 ? 128 : (int)cdiv(Co, 16) * 16; }

size_t tc_packed_weight_bytes(int Co, int C, int taps)
{
    const int NT = tc_nt(Co), n_tiles = (int)cdiv(Co, NT);
    return (size_t)n_tiles * taps * C * NT * 2 * sizeof(__nv_bfloat16);
}

int tc_pack_weight(const float *w, void *bp, int Co, int C, int taps, cudaStream_t st)
{
    if (pack_skipped()) return DLKA_OK;   // prepacked weights: see PackSkipScope
    const int KC = tc_kc(C), NT = tc_nt(Co), n_tiles = (int)cdiv(Co, NT);
    if (KC == 0) return DLKA_ERR_UNSUPPORTED;
    const i64 total = (i64)n_tiles * taps * C * NT;
    const int blocks = (int)(cdiv(total, 256) < 148 * 8 ? cdiv(total, 256) : 148 * 8);
    DLKA_LAUNCH("pack_weight_tc", st,
                pack_weight_tc_kernel<<<blocks, 256, 0, st>>>(w, (__nv_bfloat16 *)bp, Co, C, taps, KC, NT, n_tiles));
    return DLKA_OK;
}

int igemm_tc(const IgemmArgs &g, const void *bp, cudaStream_t st)
{
    if (!tc_supported(g)) return DLKA_ERR_UNSUPPORTED;
    if (g.M <= 0) return DLKA_OK;
    TcArgs a;
    a.g = g;
    a.Bp = (const uint8_t *)bp;
    const int KC = tc_kc(g.geo.C);
    a.NT = tc_nt(g.geo.Co);
    const int n_tiles = (int)cdiv(g.geo.Co, a.NT);
    a.KS = g.geo.K * (g.geo.C / KC);
    a.vol_c = (i64)g.geo.D * g.geo.H * g.geo.W * g.geo.C;
    switch (g.mode) {
    case IGEMM_DENSE: return launch_tc_kc<IGEMM_DENSE>(a, KC, n_tiles, st);
    case IGEMM_CONV: return launch_tc_kc<IGEMM_CONV>(a, KC, n_tiles, st);
    default: return launch_tc_kc<IGEMM_DEFORM>(a, KC, n_tiles, st);
    }
}

}  // namespace dlka
