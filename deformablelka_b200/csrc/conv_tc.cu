// conv_tc.cu -- regular (zero-padded, stride-1) convolution as a ZERO-COPY implicit GEMM on tcgen05.
//
// Used for the offset-predicting convolutions of the D-LKA blocks (3D conv_offset 3x3x3 C->81,
// synapse/deform_conv.py:80-85; 2D offset_net 5x5 / 7x7-dil-3, deformable_LKA.py:10-16).
//
// Idea: a CTA owns a block of MT "M tiles"; an M tile is 16 x 8 output positions of one (d) slice
// (3D: tile t = slice d0+t; 2D: tile t = rows h0+16t..).  For a 16-channel K chunk the input region
// (tile + halo) is loaded ONCE, split into bf16 hi/lo and written to shared memory as
//     [hi|lo][plane p = 8 channels][region voxel v = (z*RH + y)*RW + x][16 bytes].
// Because 8 consecutive x positions are 8 consecutive 16-byte rows and consecutive tile rows (y) are RW*16
// bytes apart, the A operand of EVERY tap is the same buffer addressed through a K-major SWIZZLE_NONE
// UMMA descriptor with  start = base + v0(tap, t)*16,  SBO = RW*16,  LBO = plane stride.
// No per-tap im2col copy exists anywhere -- not in HBM, not in shared memory.
//
// Pipeline per CTA: region producers (warps 4..) double-buffer K chunks; warp 1 streams the per-(chunk,tap)
// weight tiles with cp.async.bulk; warp 0 issues 3 (bf16 hi/lo split) x MT tcgen05.mma per tap; the
// producers then run the epilogue (tcgen05.ld -> +bias -> global).
#include <cuda_bf16.h>

#include <atomic>
#include <mutex>

#include <cstdlib>

#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace dlka {
namespace {

using namespace ptx;

constexpr int CT_CTRL_WARPS = 4;
constexpr int CT_NPW = 8;          // producer / epilogue warps
#ifndef DLKA_CT_MTMAX
#define DLKA_CT_MTMAX 2
#define DLKA_CT_MINB 2
#endif
// weight ring depth (ConvTileArgs::sb): 3 when two CTAs share an SM and it still fits half the shared memory (headline conv_offset:
// 2 -> 3 slots took it from 2.39 to 2.19 ms -- a 6 KB tile per tap from L2 against ~290 cycles of MMAs), else 2; up to
// CT_SB_MAX when the region buffers leave room for one CTA only (2D 7x7 dil 3: a 7 KB tile per tap from L2 for ~340 cycles of
// MMAs -- with 2 slots in flight the loop ran at the copy latency, 3x under the tensor rate)
constexpr int CT_SB_MAX = 6;
constexpr int CT_KCH = 16;         // channels per K chunk (= one UMMA K step)
constexpr int CT_LPAD = 64;        // bytes added to the plane stride (bank spread between the 2 planes)

struct ConvTileArgs {
    ConvGeo g;
    const float *X;      // [B][D][H][W][C], or chunk-major [C/32][B][D][H][W][32] when xch != 0
    i64 xch;             // floats between 32-channel chunks (0: channels-last)
    const uint8_t *Bp;   // packed weights [n_tiles][chunk][tap][hi|lo][plane 2][NT][8 bf16]
    const float *bias;   // [Co] or null
    int act;             // 0 none, 1 LeakyReLU(slope), 2 (+E) then LeakyReLU(slope)   (UnetResBlock, row N3)
    float slope;
    const float *E;      // residual operand [M][ldE] for act == 2
    int ldE;
    float *Y;            // [M][ldY], or brick-major [brick][Co][128] when ybrick (bricks of 4 x 4 x 8 output voxels)
    int ldY;
    int ybrick, bt_d, bt_h, bt_w;   // brick grid extents
    int NT;              // N tile
    int MT;              // M tiles per CTA
    int RD, RH, RW;      // region extent (voxels)
    int tiles_d, tiles_h, tiles_w;  // CTA grid decomposition
    int lbo;             // plane stride in bytes = RV*16 + pad
    int sb;              // weight ring depth, 2 .. CT_SB_MAX
    int csplit;          // > 0: grid.z slices of `csplit` K chunks; slice z writes to Y + z * ysplit (bias in slice 0 only)
    i64 ysplit;
};

template <int MT>
__global__ void __launch_bounds__((CT_CTRL_WARPS + CT_NPW) * 32, DLKA_CT_MINB) conv_tiled_kernel(const ConvTileArgs a)
{
    constexpr int NPT = CT_NPW * 32;
    extern __shared__ __align__(128) uint8_t smem[];
    const ConvGeo &g = a.g;
    const int NT = a.NT, RV = a.RD * a.RH * a.RW;
    const int LBO = a.lbo;                   // bytes between the two 8-channel planes
    const int R_HALF = 2 * LBO;              // hi (or lo) part of one region buffer: 2 planes
    const int R_BUF = 2 * R_HALF;            // hi + lo
    const int B_HALF = 2 * NT * 16, B_SLOT = 2 * B_HALF;
    uint8_t *sR = smem;
    uint8_t *sB = sR + 2 * R_BUF;
    const int CT_SB = a.sb;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + CT_SB * B_SLOT);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4 + 2 * CT_SB_MAX + 1);
    const uint32_t bar0 = smem_u32(bars);
    auto fullR = [&](int s) { return bar0 + 8u * s; };
    auto emptyR = [&](int s) { return bar0 + 8u * (2 + s); };
    auto fullB = [&](int s) { return bar0 + 8u * (4 + s); };
    auto emptyB = [&](int s) { return bar0 + 8u * (4 + CT_SB_MAX + s); };
    const uint32_t accFull = bar0 + 8u * (4 + 2 * CT_SB_MAX);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tile = blockIdx.y;
    // CTA -> (b, block origin)
    int bid = blockIdx.x;
    const int tw = bid % a.tiles_w; bid /= a.tiles_w;
    const int th = bid % a.tiles_h; bid /= a.tiles_h;
    const int td = bid % a.tiles_d;
    const int b = bid / a.tiles_d;
    const bool is3d = g.ndim == 3;
    const int d0 = is3d ? td * MT : 0, h0 = is3d ? th * 16 : th * 16 * MT, w0 = tw * 8;
    const int nchunks_all = g.C / CT_KCH, K = g.K;
    const int nchunks = a.csplit > 0 ? a.csplit : nchunks_all, chunk0 = a.csplit > 0 ? (int)blockIdx.z * a.csplit : 0;
    const uint32_t tmem_cols = (MT * NT <= 32) ? 32u : (MT * NT <= 64) ? 64u : (MT * NT <= 128) ? 128u : (MT * NT <= 256) ? 256u : 512u;

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(fullR(s), CT_NPW); mbar_init(emptyR(s), 1); }
        for (int s = 0; s < CT_SB; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
        mbar_init(accFull, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(tmem_slot), tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(128, NT);
            const uint32_t sbo = (uint32_t)a.RW * 16u;
            int bs = 0;           // weight ring slot and its phase bit (running counters: the ring depth is a runtime value)
            uint32_t bph = 0;
            for (int c = 0; c < nchunks; ++c) {
                const int rb = c & 1;
                mbar_wait(fullR(rb), (c >> 1) & 1);
                tc_fence_after();
                const uint32_t rhi = smem_u32(sR + rb * R_BUF), rlo = rhi + R_HALF;
                int kk = 0, jj = 0, ii = 0;   // tap = (ii * kh + jj) * kw + kk, kept as counters (no divisions on the issue path)
                for (int tap = 0; tap < K; ++tap) {
                    mbar_wait(fullB(bs), bph);
                    tc_fence_after();
                    const uint32_t bhi = smem_u32(sB + bs * B_SLOT), blo = bhi + B_HALF;
                    const uint64_t bd_hi = make_smem_desc(bhi, NT * 16, 128), bd_lo = make_smem_desc(blo, NT * 16, 128);
#pragma unroll
                    for (int t = 0; t < MT; ++t) {
                        // first region voxel read by row 0 of tile t for this tap
                        const int v0 = is3d ? ((t + ii * g.dd) * a.RH + jj * g.dh) * a.RW + kk * g.dw
                                            : ((t * 16 + jj * g.dh) * a.RW + kk * g.dw);
                        const uint64_t ad_hi = make_smem_desc(rhi + v0 * 16, LBO, sbo), ad_lo = make_smem_desc(rlo + v0 * 16, LBO, sbo);
                        const uint32_t d_tmem = tmem_base + (uint32_t)(t * NT);
                        const uint32_t acc = (c | tap) != 0 ? 1u : 0u;
                        umma_bf16(d_tmem, ad_hi, bd_hi, idesc, acc);
                        umma_bf16(d_tmem, ad_lo, bd_hi, idesc, 1u);
                        umma_bf16(d_tmem, ad_hi, bd_lo, idesc, 1u);
                    }
                    umma_commit(emptyB(bs));
                    if (++bs == CT_SB) { bs = 0; bph ^= 1u; }
                    if (++kk == g.kw) { kk = 0; if (++jj == g.kh) { jj = 0; ++ii; } }
                }
                umma_commit(emptyR(rb));
            }
            umma_commit(accFull);
        }
    } else if (warp == 1) {
        // ===================== weight loader =====================
        if (elect_one()) {
            const int total = nchunks * K;
            const uint8_t *src = a.Bp + ((i64)n_tile * nchunks_all + chunk0) * K * B_SLOT;
            int bs = 0;
            uint32_t bph = 1;   // emptyB starts "free"
            for (int bi = 0; bi < total; ++bi) {
                mbar_wait(emptyB(bs), bph);
                mbar_arrive_expect_tx(fullB(bs), (uint32_t)B_SLOT);
                bulk_g2s(smem_u32(sB + bs * B_SLOT), src + (i64)bi * B_SLOT, (uint32_t)B_SLOT, fullB(bs));
                if (++bs == CT_SB) { bs = 0; bph ^= 1u; }
            }
        }
    } else if (warp >= CT_CTRL_WARPS) {
        // ===================== region producers =====================
        const int ptid = tid - CT_CTRL_WARPS * 32;
        const int rd0 = d0 - g.pd, rh0 = h0 - g.ph, rw0 = w0 - g.pw;  // region origin in input coordinates
        // voxel stride / per-K-chunk base: channels-last (ldv = C, chunk c at + 16 c) or chunk-major (ldv = 32, 32-channel
        // chunk c/2 at + (c/2) * xch, its second half at + 16)
        const int ldv = a.xch ? 32 : g.C;
        const float *Xb = a.X + (i64)b * g.D * g.H * g.W * ldv;
        const int units = RV * 4;  // (voxel, float4 of the 16-channel chunk)
        for (int c = 0; c < nchunks; ++c) {
            const int rb = c & 1;
            mbar_wait(emptyR(rb), ((c >> 1) & 1) ^ 1);
            uint8_t *buf = sR + rb * R_BUF;
            const int cg = chunk0 + c;   // global K chunk
            const float *Xc = a.xch ? Xb + (i64)(cg >> 1) * a.xch + (cg & 1) * CT_KCH : Xb + cg * CT_KCH;
            for (int u0 = ptid; u0 < units; u0 += 4 * NPT) {
                float4 v[4];
                int vox[4];
#pragma unroll
                for (int s = 0; s < 4; ++s) {  // 4 independent loads in flight per thread
                    const int u = u0 + s * NPT;
                    v[s] = f4zero();
                    vox[s] = -1;
                    if (u < units) {
                        const int q = u & 3, vv = u >> 2;
                        vox[s] = vv;
                        const int x = vv % a.RW, y = (vv / a.RW) % a.RH, z = vv / (a.RW * a.RH);
                        const int di = rd0 + z, hi_ = rh0 + y, wi = rw0 + x;
                        if ((unsigned)di < (unsigned)g.D && (unsigned)hi_ < (unsigned)g.H && (unsigned)wi < (unsigned)g.W)
                            v[s] = ldg4(Xc + (((i64)di * g.H + hi_) * g.W + wi) * ldv + q * 4);
                    }
                }
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    if (vox[s] < 0) continue;
                    const int q = (u0 + s * NPT) & 3;
                    uint2 hi, lo;
                    split_bf16x4(v[s], hi, lo);
                    const int boff = (q >> 1) * LBO + vox[s] * 16 + (q & 1) * 8;
                    *reinterpret_cast<uint2 *>(buf + boff) = hi;
                    *reinterpret_cast<uint2 *>(buf + R_HALF + boff) = lo;
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(fullR(rb));
        }
        // ===================== epilogue =====================
        mbar_wait(accFull, 0);
        tc_fence_after();
        const int pw = warp - CT_CTRL_WARPS, q = warp & 3;
        const bool vec_y = (a.ldY & 3) == 0;
        for (int t = pw >> 2; t < MT; t += CT_NPW / 4) {
            const int r = q * 32 + lane, oh = r >> 3, ow = r & 7;
            const int od = is3d ? d0 + t : 0, ohh = is3d ? h0 + oh : h0 + t * 16 + oh, oww = w0 + ow;
            const bool mv = od < g.Do && ohh < g.Ho && oww < g.Wo;
            const i64 m = (((i64)b * g.Do + od) * g.Ho + ohh) * g.Wo + oww;
            // brick-major output: 32 consecutive lanes are 4 h-rows x 8 w of one brick slice = 32 consecutive floats per column
            const i64 ybase = a.ybrick ? ((((i64)b * a.bt_d + (od >> 2)) * a.bt_h + (ohh >> 2)) * a.bt_w + (oww >> 3)) * (i64)g.Co * 128 +
                                             ((od & 3) * 32 + (ohh & 3) * 8 + (oww & 7))
                                       : 0;
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * NT);
            for (int c0 = 0; c0 < NT; c0 += 16) {
                float v[16];
                tmem_ld16(trow + c0, v);
                if (!mv) continue;
                const int nb = n_tile * NT + c0;
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const int n = nb + j4 * 4;
                    if (n >= g.Co) break;
                    float o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int ne = n + e < g.Co ? n + e : g.Co - 1;
                        o[e] = v[j4 * 4 + e] + ((a.bias && chunk0 == 0) ? __ldg(a.bias + ne) : 0.f);
                    }
                    if (a.act) {
                        if (a.act == 2) {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (n + e < g.Co) o[e] += __ldg(a.E + m * (i64)a.ldE + n + e);
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) o[e] = o[e] > 0.f ? o[e] : a.slope * o[e];
                    }
                    float *Ys = a.Y + (i64)blockIdx.z * a.ysplit;
                    if (a.ybrick) {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (n + e < g.Co) Ys[ybase + (i64)(n + e) * 128] = o[e];
                        continue;
                    }
                    float *yp = Ys + m * (i64)a.ldY + n;
                    if (vec_y && n + 3 < a.ldY) {
                        *reinterpret_cast<float4 *>(yp) = make_float4(o[0], o[1], o[2], o[3]);  // pad columns hold junk-free bias values
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (n + e < g.Co) yp[e] = o[e];
                    }
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

// weight [Co][C][taps] -> Bp[n_tile][chunk][tap][hi|lo][plane 2][NT][8]
__global__ void pack_weight_ct_kernel(const float *__restrict__ w, const float *__restrict__ wscale, __nv_bfloat16 *__restrict__ bp,
                                      int Co, int C, int taps, int NT, int n_tiles)
{
    const int nch = C / CT_KCH;
    const i64 total = (i64)n_tiles * nch * taps * 16 * NT;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        const int e = (int)(i % 8);
        const int n = (int)((i / 8) % NT);
        const int p = (int)((i / (8 * NT)) % 2);
        const int tap = (int)((i / (16 * NT)) % taps);
        const int ch = (int)((i / ((i64)16 * NT * taps)) % nch);
        const int nt = (int)(i / ((i64)16 * NT * taps * nch));
        const int c = ch * CT_KCH + p * 8 + e, co = nt * NT + n;
        float v = co < Co ? w[((i64)co * C + c) * taps + tap] : 0.f;
        if (wscale && co < Co) v *= wscale[co];   // folded eval-mode BatchNorm scale (per output channel)
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        const i64 slot = ((i64)(nt * nch + ch) * taps + tap) * (32 * NT);  // elements per slot: hi+lo = 2*16*NT
        bp[slot + ((i64)p * NT + n) * 8 + e] = hi;
        bp[slot + 16 * NT + ((i64)p * NT + n) * 8 + e] = lo;
    }
}

size_t ct_smem_bytes(const ConvTileArgs &a)
{
    const size_t rbuf = (size_t)4 * a.lbo;  // hi+lo, 2 planes each
    return 2 * rbuf + (size_t)a.sb * 4 * a.NT * 16 + (4 + 2 * CT_SB_MAX + 1) * 8 + 16 + 128;
}

bool ct_plan(const IgemmArgs &g, ConvTileArgs &a)
{
    const ConvGeo &geo = g.geo;
    if (g.mode != IGEMM_CONV || geo.groups != 1 || geo.C % CT_KCH != 0) return false;
    if (geo.sd != 1 || geo.sh != 1 || geo.sw != 1) return false;
    if (g.epi != EPI_NONE) return false;
    if (g.xch && geo.C % 32 != 0) return false;
    if (g.ybrick && geo.ndim != 3) return false;
    a.ybrick = g.ybrick; a.bt_d = (int)cdiv(geo.Do, 4); a.bt_h = (int)cdiv(geo.Ho, 4); a.bt_w = (int)cdiv(geo.Wo, 8);
    a.g = geo; a.X = g.X; a.xch = g.xch; a.bias = g.bias; a.Y = g.Y; a.ldY = g.ldY;
    a.act = 0; a.slope = 0.f; a.E = nullptr; a.ldE = 0; a.csplit = 0; a.ysplit = 0; a.sb = 2;
    a.NT = tc_nt(geo.Co);
    const bool is3d = geo.ndim == 3;
    // tuning knobs (environment, read once): DLKA_CT_MT = largest MT tried, DLKA_CT_SB = ring depth when two CTAs share an SM
    static const int env_mt = [] { const char *e = getenv("DLKA_CT_MT"); return e ? atoi(e) : DLKA_CT_MTMAX; }();
    static const int env_sb = [] { const char *e = getenv("DLKA_CT_SB"); return e ? atoi(e) : 3; }();
    for (int mt = env_mt; mt >= 1; mt >>= 1) {
        a.MT = mt;
        a.RD = is3d ? mt + (geo.kd - 1) * geo.dd : 1;
        a.RH = (is3d ? 16 : 16 * mt) + (geo.kh - 1) * geo.dh;
        a.RW = 8 + (geo.kw - 1) * geo.dw;
        const i64 rv = (i64)a.RD * a.RH * a.RW;
        a.lbo = (int)(rv * 16 + CT_LPAD);
        if (a.lbo / 16 >= (1 << 14) || a.RW * 16 / 16 >= (1 << 14)) continue;
        if (mt * a.NT > 512) continue;
        if (ct_smem_bytes(a) > 220 * 1024) continue;
        // small problems: prefer fewer tiles per CTA so that the grid still fills the SMs.  2D nets with large kernels (7x7 dil 3,
        // C up to 384) are bound by streaming the weights from L2 -- every CTA walks the whole packed tensor, 7 KB per tap for 3
        // MMAs -- so there two tiles per CTA (half the weight bytes per output) win as long as half the SMs get a CTA
        const i64 blocks = (i64)geo.B * (is3d ? cdiv(geo.Do, mt) * cdiv(geo.Ho, 16) : cdiv(geo.Ho, 16 * mt)) * cdiv(geo.Wo, 8);
        if (mt > 1 && blocks < (is3d ? 148 : 74)) continue;
        a.tiles_d = is3d ? (int)cdiv(geo.Do, mt) : 1;
        a.tiles_h = is3d ? (int)cdiv(geo.Ho, 16) : (int)cdiv(geo.Ho, 16 * mt);
        a.tiles_w = (int)cdiv(geo.Wo, 8);
        // one CTA per SM anyway (> half the shared memory): deepen the weight ring as far as it fits
        if (ct_smem_bytes(a) > 112 * 1024) {
            while (a.sb < CT_SB_MAX) {
                ++a.sb;
                if (ct_smem_bytes(a) > 220 * 1024) { --a.sb; break; }
            }
        } else {
            while (a.sb < env_sb) {
                ++a.sb;
                if (ct_smem_bytes(a) > 112 * 1024) { --a.sb; break; }
            }
        }
        return true;
    }
    return false;
}

template <int MT>
int launch_ct(const ConvTileArgs &a, int n_tiles, cudaStream_t st, int zsplit = 1)
{
    const size_t smem = ct_smem_bytes(a);
    auto kern = conv_tiled_kernel<MT>;
    // the attribute is per function and process-wide: keep a process-wide monotonic maximum (a thread_local cache let a second
    // host thread -- e.g. the autograd engine's -- lower the limit under a launch that needs more)
    static SmemOptIn optin;   // per launch site (= per kernel instantiation), per device
    DLKA_TRY(optin.ensure(kern, smem));
    dim3 grid((unsigned)((i64)a.g.B * a.tiles_d * a.tiles_h * a.tiles_w), (unsigned)n_tiles, (unsigned)zsplit);
    DLKA_LAUNCH("tc_conv_tiled", st, (kern<<<grid, (CT_CTRL_WARPS + CT_NPW) * 32, smem, st>>>(a)));
    return DLKA_OK;
}

}  // namespace

bool conv_tiled_supported(const IgemmArgs &g)
{
    ConvTileArgs a;
    return ct_plan(g, a);
}

size_t conv_tiled_packed_bytes(int Co, int C, int taps)
{
    const int NT = tc_nt(Co), n_tiles = (int)cdiv(Co, NT);
    return (size_t)n_tiles * (C / CT_KCH) * taps * 32 * NT * sizeof(__nv_bfloat16);
}

int conv_tiled(const IgemmArgs &g, const float *w, void *bp, cudaStream_t st)
{
    return conv_tiled_ex(g, w, nullptr, 0, 0.f, nullptr, 0, bp, st);
}

// wscale: per-output-channel factor folded into the packed weights; act / slope / E: see ConvTileArgs
int conv_tiled_ex(const IgemmArgs &g, const float *w, const float *wscale, int act, float slope, const float *E, int ldE, void *bp,
                  cudaStream_t st)
{
    ConvTileArgs a;
    if (!ct_plan(g, a)) return DLKA_ERR_UNSUPPORTED;
    a.act = act; a.slope = slope; a.E = E; a.ldE = ldE;
    const ConvGeo &geo = g.geo;
    const int n_tiles = (int)cdiv(geo.Co, a.NT);
    if (!pack_skipped()) {
        const i64 total = (i64)n_tiles * (geo.C / CT_KCH) * geo.K * 16 * a.NT;
        const int blocks = (int)(cdiv(total, 256) < 148 * 8 ? cdiv(total, 256) : 148 * 8);
        DLKA_LAUNCH("pack_weight_ct", st,
                    pack_weight_ct_kernel<<<blocks, 256, 0, st>>>(w, wscale, (__nv_bfloat16 *)bp, geo.Co, geo.C, geo.K, a.NT, n_tiles));
    }
    a.Bp = (const uint8_t *)bp;
    // K split over channel chunks when the grid would leave most SMs idle behind a long K loop (the offset nets of the deep network
    // stages: 24 x 14 x 14 pixels, 4^3 volumes): S slices write partial outputs, reduce_partials sums them into Y
    int S = 1;
    const i64 blocks = (i64)geo.B * a.tiles_d * a.tiles_h * a.tiles_w * n_tiles;
    const int nch = geo.C / CT_KCH;
    const i64 ybuf = a.ybrick ? (i64)geo.B * a.bt_d * a.bt_h * a.bt_w * geo.Co * 128 : g.M * (i64)a.ldY;
    if (act == 0 && g.split_scratch && blocks < 100 && nch >= 4) {
        for (int s = 2; s <= nch; ++s)
            if (nch % s == 0 && blocks * s <= 296 && (i64)s * ybuf <= g.split_scratch_floats) S = s;
    }
    float *Yfinal = a.Y;
    if (S > 1) { a.csplit = nch / S; a.ysplit = ybuf; a.Y = g.split_scratch; }
    int rc;
    switch (a.MT) {
    case 4: rc = launch_ct<4>(a, n_tiles, st, S); break;
    case 2: rc = launch_ct<2>(a, n_tiles, st, S); break;
    default: rc = launch_ct<1>(a, n_tiles, st, S); break;
    }
    if (rc != DLKA_OK || S == 1) return rc;
    return reduce_partials(g.split_scratch, Yfinal, ybuf, S, st);
}

}  // namespace dlka
