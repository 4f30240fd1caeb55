// deform_ps.cu -- 3D deformable convolution (groups 1, dg 1, stride 1) + fused 1x1 chain, PERSISTENT kernel (round 2).
//
// Replaces the hot loop of D3D.deform_conv_forward (deformable_im2col_gpu_kernel + at::addmm,
// 3D/dcn/src/cuda/deform_im2col_cuda.cuh:192-265, deform_conv_cuda.cu:113-119) and, in the D-LKA block, the two 1x1x1
// convolutions + gate + residual behind it (transformerblock.py:648-651,669-671).  No im2col buffer exists.
//
// What changed against deform_tc.cu (round 1: 12.3 ms at the headline shape) and why -- each item is a measurement:
//  * The gather source is CHUNK-MAJOR, X[chunk][b][d][h][w][32 ch]: the two w-corners of a trilinear sample are then two
//    ADJACENT 128-byte lines.  tools/l1_probe.cu: an L1-hit LDG.128 whose 16 lanes cover such a pair costs 1.49 cycles per
//    line, against 1.84 for four unrelated lines per instruction (the round-1 mapping).  ncu on the first version of this
//    kernel then showed the binding unit is the L1TEX DATA PIPE, one wavefront per cycle for every global line, shared-memory
//    wavefront and shuffle alike (90 % busy: 1295 global + 1127 shared wavefronts per K step).  So the mapping minimises
//    wavefronts: 8 lanes serve one (row, tap) with 32-byte loads (LDG.E.256: lanes 0-3 the low-w line, lanes 4-7 the high-w
//    line, 4 loads each for the (d, h) corner pairs), which halves the per-lane parameter reads; the two sides swap HALF of
//    their partial sums (4 shuffles), each lane then owns 4 complete channels and writes their bf16 hi / lo with two
//    conflict-free 8-byte stores; the offsets are read coalesced (brick-major from the offset conv, or straight from the
//    NCDHW tensor of the operator entry) instead of one line per row.
//  * Persistent CTAs (one per SM) walk a static list of 4x4x8 output bricks; the accumulator is DOUBLE-BUFFERED in tensor
//    memory and a dedicated epilogue warpgroup drains tile t while the gather / MMA pipeline already runs tile t+1.  Round 1
//    spent ~10 % of every tile filling the pipeline and running the 3-stage epilogue with nothing overlapped.
//  * The 1x1 chain (conv1 -> * u -> proj_2 -> + x) takes its A operand FROM TENSOR MEMORY: the epilogue threads own one
//    accumulator row each, add the bias, split to bf16 hi/lo and write the row back with tcgen05.st; the MMAs are the
//    tcgen05 "TS" form.  No shared-memory restaging (48 KB in round 1), and the chain weights stream through the same
//    weight ring as the main loop (6 extra slots per tile), so shared memory stays at ~115 KB and L1 keeps ~110 KB.
//
// Roles (28 warps): warp 0 MMA issuer, warp 1 weight loader (cp.async.bulk), warps 4-7 sample-parameter producers (one
// thread per brick row: position, validity, 4 pair offsets, 2 x 4 masked weights = 48 B per (row, tap)), warps 8-23 gather /
// blend / convert producers in two groups on alternate K steps, warps 24-27 epilogue (one thread per accumulator row).
#include <cuda_bf16.h>

#include <cstring>

#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace dlka {
namespace {

using namespace ptx;

constexpr int PS_KC = 32;                  // channels per K step = one 128-byte line per voxel
constexpr int PS_SA = 3, PS_SB = 3, PS_SP = 4;
constexpr int PS_LBO = 2048 + 32;          // A plane (8 channels x 128 rows) stride, padded to spread banks
constexpr int PS_APLANE = (PS_KC / 8) * PS_LBO;
constexpr int PS_ASLOT = 2 * PS_APLANE;    // hi + lo
constexpr int PS_PSIDE = 128 * 16 + 64;    // parameter stage: [128 rows x 16 B, low-w side] [+64 B bank shift] [128 x 16 B, high-w side]
constexpr int PS_PSTAGE = 2 * PS_PSIDE;    // record = one 16-byte {offset|flags, wx, ld, lh} per row and w-side (ps_make_params)
constexpr int PS_GROUPS = 2;               // gather warp groups on alternate K steps
constexpr int PS_GW = 8;                   // warps per gather group
constexpr int PS_PARAM_WARPS = 4, PS_EPI_WARPS = 4;
constexpr int PS_THREADS = (4 + PS_PARAM_WARPS + PS_GROUPS * PS_GW + PS_EPI_WARPS) * 32;   // 896
constexpr int PS_BD = 4, PS_BH = 4, PS_BW = 8;   // output brick = 128 MMA rows
// Tensor-memory map (512 columns).  Every GEMM stage is "N-stacked": D[:, 0:NT) accumulates hi*hi + lo*hi and D[:, NT:2NT) the
// cross term hi*lo, produced by ONE MMA with N = 2 NT over the weight rows [hi | lo] plus one MMA with N = NT (A lo x W hi):
// 4 MMAs per K step instead of 6 and 34 KB instead of 42 KB of operand reads.  The epilogue adds the two halves.
//   accumulator buffers at columns 0 and S (S = 192 for NT <= 96, 256 for NT = 128), each 2 NT wide; with the chain (NT <= 96)
//   the A operand (bf16 hi | lo, NT columns) sits at 384, and the chain's GEMM output X reuses the accumulator buffer of the
//   tile being drained (free once stage 0 has read it).
constexpr int PS_TM_A = 384, PS_TM_COLS = 512;

struct DeformPsArgs {
    ConvGeo g;
    const float *X;      // chunk-major [C/32][B][D][H][W][32]
    i64 xch;             // floats between channel chunks (= B*D*H*W*32)
    const float *Off;    // offsets, element (row, column c) at Off[row_term + c * off_cs]:
    i64 off_cs;          //   brick-major (conv_tiled, off_mode 1): row_term = tile*3K*128 + r, off_cs = 128
    int off_mode;        //   NCDHW operator input (off_mode 2): row_term = b*3K*S + voxel, off_cs = S;  [M][ld] rows (0): m*ld, 1
    int ldOff;
    const uint8_t *Bp;   // [chunk][tap][4 planes][hi NT rows | lo NT rows][8 bf16]
    const uint8_t *W1p;  // [chunk][4 planes][hi | lo][8 bf16]   (chain)
    const uint8_t *W2p;
    const float *bias, *b1, *b2;
    const float *U;      // gate operand [M][ldU]      (chain >= 1)
    const float *R;      // residual operand [M][ldR]  (chain == 2)
    int ldU, ldR;
    float *Y;            // [M][ldY]
    int ldY;
    int NT;              // N tile (multiple of 16, <= 128; chain: == C == Co <= 96)
    int chain;           // 0: Y = conv + bias; 1: Y = (conv1(conv + bias) + b1) * U; 2: Y = proj_2(that) + b2 + R
    int tiles_d, tiles_h, tiles_w, ntiles;
    int vec32;           // Y / U / R rows are 32-byte aligned: 32-byte epilogue accesses
    int KS, S1, S2;      // K steps per tile; main-loop positions at which the previous tile's chain stages are issued
};

struct PsRow {
    int m;        // linear output row, -1 outside the volume
    int d, h, w;
};

__device__ __forceinline__ void ps_tile_coords(const DeformPsArgs &a, int tile, int &b, int &td, int &th, int &tw)
{
    tw = tile % a.tiles_w; tile /= a.tiles_w;
    th = tile % a.tiles_h; tile /= a.tiles_h;
    td = tile % a.tiles_d;
    b = tile / a.tiles_d;
}

// brick row r (0..127 = 4 d x 4 h x 8 w) of a tile: pure arithmetic, so the parameter warps and the epilogue warps each derive
// it themselves and no row table has to be handed between them through shared memory
__device__ __forceinline__ PsRow ps_row(const DeformPsArgs &a, int tile, int r, int &b)
{
    int td, th, tw;
    ps_tile_coords(a, tile, b, td, th, tw);
    const ConvGeo &g = a.g;
    PsRow ri;
    ri.d = td * PS_BD + (r >> 5); ri.h = th * PS_BH + ((r >> 3) & 3); ri.w = tw * PS_BW + (r & 7);
    ri.m = (ri.d < g.Do && ri.h < g.Ho && ri.w < g.Wo) ? (int)((((i64)b * g.Do + ri.d) * g.Ho + ri.h) * g.Wo + ri.w) : -1;
    return ri;
}

// Sampling parameters of one (row, tap): the position / validity rules are the reference's (cuh:245-248, 30-65) through
// make_sample3; corners are addressed as 4 (d, h) PAIRS of w-adjacent lines starting at xb = clamp(floor(pw), 0, W-2), and
// the w-interpolation weights are attached to whichever side of the pair holds that corner's voxel.
__device__ __forceinline__ void ps_make_params(const ConvGeo &g, const PsRow &ri, int ii, int jj, int kk, float od, float oh, float ow,
                                               uint8_t *rec)
{
    // record (32 B per row, one 16-byte half per w-side): {word0, wx, ld, lh} with
    //   word0 = byte offset of the (d0, h0) line pair (multiple of 128) | bit 0: the h1 row differs | bit 1: the d1 plane differs |
    //           bits 2..5: the reference keeps the d-low / d-high / h-low / h-high corner (index inside the volume, cuh:43-65),
    //   wx = w-interpolation weight of this side of the pair, ld / lh = fractional parts along d / h.
    // A gather lane reads ONE 16-byte word and forms its 4 corner weights (hd|ld)*(hh|lh)*wx exactly as the reference orders the
    // products (cuh:67-68).  An invalid sample has wx = 0 on both sides.
    uint32_t word0 = 0u;
    float wx0 = 0.f, wx1 = 0.f, ld = 0.f, lh = 0.f;
    if (ri.m >= 0) {
        const float pd = sample_pos(ri.d, g.sd, g.pd, ii, g.dd, od);
        const float ph = sample_pos(ri.h, g.sh, g.ph, jj, g.dh, oh);
        const float pw = sample_pos(ri.w, g.sw, g.pw, kk, g.dw, ow);
        const Sample3 s = make_sample3(pd, ph, pw, g.D, g.H, g.W);
        if (s.mask & 1) {
            ld = s.l[0]; lh = s.l[1];
            const float lw = s.l[2], hw = 1.f - lw;
            const int d0 = max(s.lo[0], 0), d1 = min(s.lo[0] + 1, g.D - 1);
            const int h0 = max(s.lo[1], 0), h1 = min(s.lo[1] + 1, g.H - 1);
            const int xb = min(max(s.lo[2], 0), g.W - 2);
            const float fxl = s.lo[2] >= 0 ? hw : 0.f, fxh = s.lo[2] + 1 <= g.W - 1 ? lw : 0.f;   // corner at lo / at lo + 1
            // side 0 reads voxel xb, side 1 voxel xb + 1
            wx0 = (xb == s.lo[2] ? fxl : 0.f) + (xb == s.lo[2] + 1 ? fxh : 0.f);
            wx1 = (xb + 1 == s.lo[2] ? fxl : 0.f) + (xb + 1 == s.lo[2] + 1 ? fxh : 0.f);
            const uint32_t sH = (uint32_t)g.W * 128u, sD = (uint32_t)g.H * sH;
            word0 = ((uint32_t)d0 * sD + (uint32_t)h0 * sH + (uint32_t)xb * 128u) | (h1 != h0 ? 1u : 0u) | (d1 != d0 ? 2u : 0u) |
                    (s.lo[0] >= 0 ? 4u : 0u) | (s.lo[0] + 1 <= g.D - 1 ? 8u : 0u) | (s.lo[1] >= 0 ? 16u : 0u) |
                    (s.lo[1] + 1 <= g.H - 1 ? 32u : 0u);
        }
    }
    *reinterpret_cast<uint4 *>(rec) = make_uint4(word0, __float_as_uint(wx0), __float_as_uint(ld), __float_as_uint(lh));
    *reinterpret_cast<uint4 *>(rec + PS_PSIDE) = make_uint4(word0, __float_as_uint(wx1), __float_as_uint(ld), __float_as_uint(lh));
}

__global__ void __launch_bounds__(PS_THREADS, 1) deform3d_ps_kernel(const DeformPsArgs a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const ConvGeo &g = a.g;
    const int NT = a.NT;
    const int B_LBO = 2 * NT * 16, B_SLOT = (PS_KC / 8) * B_LBO;   // weight plane (8 channels) = NT hi rows then NT lo rows
    const uint32_t acc_stride = NT > 96 ? 256u : 192u;
    uint8_t *sA = smem;
    uint8_t *sB = sA + PS_SA * PS_ASLOT;
    uint8_t *sPrm = sB + PS_SB * B_SLOT;                                    // [SP][PS_PSTAGE]
    float *sBias = reinterpret_cast<float *>(sPrm + PS_SP * PS_PSTAGE);     // [3][128]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sBias + 3 * 128);
    constexpr int NBARS = 2 * PS_SA + 2 * PS_SB + 2 * PS_SP + 4 + 4;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + NBARS);
    const uint32_t bar0 = smem_u32(bars);
    auto fullA = [&](int s) { return bar0 + 8u * s; };
    auto emptyA = [&](int s) { return bar0 + 8u * (PS_SA + s); };
    auto fullB = [&](int s) { return bar0 + 8u * (2 * PS_SA + s); };
    auto emptyB = [&](int s) { return bar0 + 8u * (2 * PS_SA + PS_SB + s); };
    auto fullP = [&](int s) { return bar0 + 8u * (2 * PS_SA + 2 * PS_SB + s); };
    auto emptyP = [&](int s) { return bar0 + 8u * (2 * PS_SA + 2 * PS_SB + PS_SP + s); };
    const uint32_t barX = bar0 + 8u * (2 * PS_SA + 2 * PS_SB + 2 * PS_SP);
    auto accFull = [&](int s) { return barX + 8u * s; };
    auto accEmpty = [&](int s) { return barX + 8u * (2 + s); };
    const uint32_t barE1 = barX + 32, barC1 = barX + 40, barE2 = barX + 48, barC2 = barX + 56;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = g.K, KS = a.KS, nchunks = g.C / PS_KC;
    const int ntl = a.ntiles > (int)blockIdx.x ? (a.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;   // tiles of this CTA
    const i64 vol32 = (i64)g.D * g.H * g.W * PS_KC;

    if (tid == 0) {
        for (int s = 0; s < PS_SA; ++s) { mbar_init(fullA(s), PS_GW); mbar_init(emptyA(s), 1); }
        for (int s = 0; s < PS_SB; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
        for (int s = 0; s < PS_SP; ++s) { mbar_init(fullP(s), PS_PARAM_WARPS); mbar_init(emptyP(s), PS_GW); }
        for (int s = 0; s < 2; ++s) { mbar_init(accFull(s), 1); mbar_init(accEmpty(s), PS_EPI_WARPS); }
        mbar_init(barE1, PS_EPI_WARPS); mbar_init(barC1, 1); mbar_init(barE2, PS_EPI_WARPS); mbar_init(barC2, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(tmem_slot), PS_TM_COLS);
        tmem_relinquish();
    }
    for (int i = tid; i < 3 * 128; i += PS_THREADS) {   // bias vectors of the three GEMM stages
        const int st = i >> 7, n = i & 127;
        const float *src = st == 0 ? a.bias : st == 1 ? a.b1 : a.b2;
        sBias[i] = (src && n < NT && n < g.Co && (st == 0 || st <= a.chain)) ? __ldg(src + n) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // =============================================== MMA issuer ===============================================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(128, NT), idesc2 = make_idesc_bf16(128, 2 * NT);
            uint32_t ga = 0, gb = 0;   // running A-slot / B-slot counters
            auto chain_stage = [&](int stage, int j) {   // stage of local tile j: X = A(tmem) * W(stage), into tile j's drained buffer
                mbar_wait(stage == 1 ? barE1 : barE2, (uint32_t)j & 1u);
                tc_fence_after();
                const uint32_t xacc = tmem_base + (uint32_t)(j & 1) * acc_stride;
                for (int c = 0; c < nchunks; ++c, ++gb) {
                    const int bs = gb % PS_SB;
                    mbar_wait(fullB(bs), (gb / PS_SB) & 1);
                    tc_fence_after();
                    const uint32_t bsl = smem_u32(sB + bs * B_SLOT);
                    const uint32_t a_hi = tmem_base + PS_TM_A + (uint32_t)(c * (PS_KC / 2)), a_lo = a_hi + (uint32_t)(g.C / 2);
#pragma unroll
                    for (int kk = 0; kk < PS_KC / 16; ++kk) {
                        const uint64_t bd = make_smem_desc(bsl + kk * 2 * B_LBO, B_LBO, 128);
                        umma_bf16_ts(xacc, a_hi + kk * 8, bd, idesc2, (c | kk) != 0 ? 1u : 0u);   // [hi*hi | hi*lo]
                        umma_bf16_ts(xacc, a_lo + kk * 8, bd, idesc, 1u);                          // += lo*hi
                    }
                    umma_commit(emptyB(bs));
                }
                umma_commit(stage == 1 ? barC1 : barC2);
            };
            for (int i = 0; i < ntl; ++i) {
                const int buf = i & 1;
                if (i >= 2) mbar_wait(accEmpty(buf), (uint32_t)((i >> 1) - 1) & 1u);   // epilogue of tile i-2 has drained acc[buf]
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)buf * acc_stride;
                for (int ks = 0; ks < KS; ++ks, ++ga, ++gb) {
                    if (a.chain && i > 0 && ks == a.S1) chain_stage(1, i - 1);
                    if (a.chain == 2 && i > 0 && ks == a.S2) chain_stage(2, i - 1);
                    const int bs = gb % PS_SB, as = ga % PS_SA;
                    mbar_wait(fullB(bs), (gb / PS_SB) & 1);
                    mbar_wait(fullA(as), (ga / PS_SA) & 1);
                    tc_fence_after();
                    const uint32_t ahi = smem_u32(sA + as * PS_ASLOT), alo = ahi + PS_APLANE;
                    const uint32_t bsl = smem_u32(sB + bs * B_SLOT);
#pragma unroll
                    for (int kk = 0; kk < PS_KC / 16; ++kk) {
                        const uint64_t bd = make_smem_desc(bsl + kk * 2 * B_LBO, B_LBO, 128);
                        const uint64_t ad_hi = make_smem_desc(ahi + kk * 2 * PS_LBO, PS_LBO, 128), ad_lo = make_smem_desc(alo + kk * 2 * PS_LBO, PS_LBO, 128);
                        umma_bf16(d_tmem, ad_hi, bd, idesc2, (ks | kk) != 0 ? 1u : 0u);   // N = 2 NT: [hi*hi | hi*lo]
                        umma_bf16(d_tmem, ad_lo, bd, idesc, 1u);                          // N = NT:   += lo*hi
                    }
                    umma_commit(emptyA(as));
                    umma_commit(emptyB(bs));
                }
                umma_commit(accFull(buf));
            }
            if (a.chain && ntl > 0) {
                chain_stage(1, ntl - 1);
                if (a.chain == 2) chain_stage(2, ntl - 1);
            }
        }
    } else if (warp == 1) {
        // =============================================== weight loader (same slot order as the issuer) ===============================================
        if (elect_one()) {
            uint32_t gb = 0;
            auto push = [&](const uint8_t *src) {
                const int bs = gb % PS_SB;
                mbar_wait(emptyB(bs), ((gb / PS_SB) & 1) ^ 1);
                mbar_arrive_expect_tx(fullB(bs), (uint32_t)B_SLOT);
                bulk_g2s(smem_u32(sB + bs * B_SLOT), src, (uint32_t)B_SLOT, fullB(bs));
                ++gb;
            };
            for (int i = 0; i < ntl; ++i) {
                for (int ks = 0; ks < KS; ++ks) {
                    if (a.chain && i > 0 && ks == a.S1)
                        for (int c = 0; c < nchunks; ++c) push(a.W1p + (i64)c * B_SLOT);
                    if (a.chain == 2 && i > 0 && ks == a.S2)
                        for (int c = 0; c < nchunks; ++c) push(a.W2p + (i64)c * B_SLOT);
                    push(a.Bp + (i64)ks * B_SLOT);
                }
            }
            if (a.chain && ntl > 0) {
                for (int c = 0; c < nchunks; ++c) push(a.W1p + (i64)c * B_SLOT);
                if (a.chain == 2)
                    for (int c = 0; c < nchunks; ++c) push(a.W2p + (i64)c * B_SLOT);
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // =============================================== sample-parameter producers: one thread per brick row ===============================================
        const int r = tid - 128;
        uint32_t gp = 0;
        for (int i = 0; i < ntl; ++i) {
            int b;
            const PsRow ri = ps_row(a, (int)blockIdx.x + i * (int)gridDim.x, r, b);
            // offsets of this row: column c at offrow[c * cs] (coalesced across the warp for the brick-major and NCDHW layouts)
            const i64 cs = a.off_cs;
            const float *offrow = a.Off;
            if (a.off_mode == 1) offrow += (i64)((int)blockIdx.x + i * (int)gridDim.x) * (3 * K) * 128 + r;
            else if (a.off_mode == 2) offrow += (i64)b * (3 * K) * cs + (ri.m >= 0 ? (i64)ri.m - (i64)b * cs : 0);
            else offrow += (i64)(ri.m >= 0 ? ri.m : 0) * a.ldOff;
            // the 3 offsets of the NEXT tap are fetched one iteration ahead: their latency overlaps this tap's arithmetic
            float od = __ldg(offrow), oh = __ldg(offrow + cs), ow = __ldg(offrow + 2 * cs);
            int tap = 0, ii = 0, jj = 0, kk = 0;
            const int pf_ks = KS > 12 ? KS - 12 : 0;
            for (int ks = 0; ks < KS; ++ks, ++gp) {
                const int ntap = tap + 1 == K ? 0 : tap + 1;
                const float *on = offrow + (i64)(ntap * 3) * cs;
                const float nod = __ldg(on), noh = __ldg(on + cs), now = __ldg(on + 2 * cs);
                if (ks == pf_ks && a.chain && ri.m >= 0) {   // pull this row's gate / residual operands into L2 ahead of the epilogue
                    for (int c = 0; c < g.Co; c += 32) {
                        prefetch_l2(a.U + (i64)ri.m * a.ldU + c);
                        if (a.chain == 2) prefetch_l2(a.R + (i64)ri.m * a.ldR + c);
                    }
                }
                const int ps = gp % PS_SP;
                mbar_wait(emptyP(ps), ((gp / PS_SP) & 1) ^ 1);
                ps_make_params(g, ri, ii, jj, kk, od, oh, ow, sPrm + ps * PS_PSTAGE + r * 16);
                __syncwarp();
                if (lane == 0) mbar_arrive(fullP(ps));
                od = nod; oh = noh; ow = now;
                tap = ntap;
                if (++kk == g.kw) { kk = 0; if (++jj == g.kh) { jj = 0; if (++ii == g.kd) ii = 0; } }
            }
        }
    } else if (warp >= 8 && warp < 8 + PS_GROUPS * PS_GW) {
        // =============================================== gather / blend / convert producers ===============================================
        // 8 lanes per (row, tap): lane l < 4 loads 32 B (8 channels) of the low-w line of each of the 4 (d, h) corner pairs,
        // lane l >= 4 the same 32 B of the adjacent high-w line.  The two sides swap half of their partial sums, so every lane
        // ends with 4 complete channels.  A group of 8 warps covers 32 rows per pass, 4 passes per K step; the two groups take
        // alternate K steps so one's loads overlap the other's arithmetic.
        const int gt = tid - 256, grp = gt >> 8, ggt = gt & 255;
        const int l8 = ggt & 7, side = l8 >> 2, q8 = l8 & 3;   // q8: which 8-channel slice of the 32-channel line
        const int row0 = ggt >> 3;                              // 0..31
        const int tiles_per_sample = a.tiles_d * a.tiles_h * a.tiles_w;
        const uint32_t sHb = (uint32_t)g.W * 128u, sDb = (uint32_t)g.H * sHb;   // byte strides of the h / d axes in the gather source
        const char *Xl = reinterpret_cast<const char *>(a.X) + l8 * 32;   // lanes 4..7 land on the next line (+128 B)
        const uint32_t boff0 = (uint32_t)(q8 * PS_LBO + side * 8);         // A operand: plane q8, channels 4*side .. +3 of its 8
        int i = 0, ks = grp;                                   // local tile, K step inside it (this group's first step)
        while (ks >= KS) { ks -= KS; ++i; }
        int b = ((int)blockIdx.x + i * (int)gridDim.x) / tiles_per_sample;
        for (uint32_t gk = grp; i < ntl; gk += PS_GROUPS) {
            int chunk = 0;
            for (int t = ks; t >= K; t -= K) ++chunk;
            const char *base = Xl + ((i64)chunk * a.xch + (i64)b * vol32) * 4;
            const int as = gk % PS_SA, ps = gk % PS_SP;
            mbar_wait(fullP(ps), (gk / PS_SP) & 1);
            mbar_wait(emptyA(as), ((gk / PS_SA) & 1) ^ 1);
            uint8_t *slot = sA + as * PS_ASLOT + boff0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int row = row0 + u * 32;
                const uint4 rc = *reinterpret_cast<const uint4 *>(sPrm + ps * PS_PSTAGE + side * PS_PSIDE + row * 16);   // {word0, wx, ld, lh}
                const uint32_t o00 = rc.x & ~127u, dh = (rc.x & 1u) ? sHb : 0u, dd = (rc.x & 2u) ? sDb : 0u;
                const float wx = __uint_as_float(rc.y), fl_d = __uint_as_float(rc.z), fl_h = __uint_as_float(rc.w);
                const float fd0 = (rc.x & 4u) ? 1.f - fl_d : 0.f, fd1 = (rc.x & 8u) ? fl_d : 0.f;
                const float fh0 = (rc.x & 16u) ? 1.f - fl_h : 0.f, fh1 = (rc.x & 32u) ? fl_h : 0.f;
                const float4 w = make_float4((fd0 * fh0) * wx, (fd0 * fh1) * wx, (fd1 * fh0) * wx, (fd1 * fh1) * wx);
                float4 a0, b0, a1, b1, a2, b2, a3, b3;
                ldg8(base + o00, a0, b0); ldg8(base + (o00 + dh), a1, b1); ldg8(base + (o00 + dd), a2, b2); ldg8(base + (o00 + dd + dh), a3, b3);
                if (u == 3) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(emptyP(ps));    // parameters consumed
                }
                float4 lo4 = f4zero(), hi4 = f4zero();         // channels 8*q8 .. +3 and +4 .. +7 of this side
                fma4(lo4, w.x, a0); fma4(hi4, w.x, b0); fma4(lo4, w.y, a1); fma4(hi4, w.y, b1);
                fma4(lo4, w.z, a2); fma4(hi4, w.z, b2); fma4(lo4, w.w, a3); fma4(hi4, w.w, b3);
                // side 0 keeps the low 4 channels and hands its high 4 to side 1, and vice versa
                const float4 keep = side ? hi4 : lo4, give = side ? lo4 : hi4;
                float4 r;
                r.x = keep.x + __shfl_xor_sync(0xffffffffu, give.x, 4); r.y = keep.y + __shfl_xor_sync(0xffffffffu, give.y, 4);
                r.z = keep.z + __shfl_xor_sync(0xffffffffu, give.z, 4); r.w = keep.w + __shfl_xor_sync(0xffffffffu, give.w, 4);
                uint2 hi, lo;
                split_bf16x4(r, hi, lo);
                *reinterpret_cast<uint2 *>(slot + row * 16) = hi;
                *reinterpret_cast<uint2 *>(slot + PS_APLANE + row * 16) = lo;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(fullA(as));
            ks += PS_GROUPS;
            if (ks >= KS) {
                while (ks >= KS) { ks -= KS; ++i; }
                b = ((int)blockIdx.x + i * (int)gridDim.x) / tiles_per_sample;
            }
        }
    } else if (warp >= 8 + PS_GROUPS * PS_GW) {
        // =============================================== epilogue: one thread per accumulator row ===============================================
        const int q = warp & 3, row = q * 32 + lane;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t tA_hi = tlane + PS_TM_A, tA_lo = tA_hi + (uint32_t)(g.C / 2);
        const bool vec_y = (a.ldY & 3) == 0;
        for (int i = 0; i < ntl; ++i) {
            const int buf = i & 1;
            mbar_wait_sleep(accFull(buf), (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            int b_unused;
            const PsRow ro = ps_row(a, (int)blockIdx.x + i * (int)gridDim.x, row, b_unused);
            const bool live = ro.m >= 0;
            const uint32_t tacc = tlane + (uint32_t)buf * acc_stride, tX = tacc;   // the chain's X lives in this tile's drained buffer
            float *yp = a.Y + (i64)(live ? ro.m : 0) * a.ldY;
            const float *up = a.U ? a.U + (i64)(live ? ro.m : 0) * a.ldU : nullptr;
            const float *rp = a.R ? a.R + (i64)(live ? ro.m : 0) * a.ldR : nullptr;
            auto store_y = [&](int c0, const float(&o)[8]) {
                if (!live) return;
                if (a.vec32 && c0 + 7 < g.Co) {
                    stg8(yp + c0, o);
                } else if (vec_y && c0 + 7 < g.Co) {
                    *reinterpret_cast<float4 *>(yp + c0) = make_float4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<float4 *>(yp + c0 + 4) = make_float4(o[4], o[5], o[6], o[7]);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (c0 + e < g.Co) yp[c0 + e] = o[e];
                }
            };
            auto restage = [&](int c0, const float(&o)[8]) {   // this row's 8 values -> bf16 hi / lo A operand in tensor memory
                uint2 h0, l0, h1, l1;
                split_bf16x4(make_float4(o[0], o[1], o[2], o[3]), h0, l0);
                split_bf16x4(make_float4(o[4], o[5], o[6], o[7]), h1, l1);
                tmem_st4(tA_hi + (uint32_t)(c0 >> 1), h0.x, h0.y, h1.x, h1.y);
                tmem_st4(tA_lo + (uint32_t)(c0 >> 1), l0.x, l0.y, l1.x, l1.y);
            };
            // ---- stage 0: deformable-conv accumulator + bias ----
            for (int c0 = 0; c0 < NT; c0 += 8) {
                float v[8], x[8], o[8];
                tmem_ld8(tacc + c0, v);
                tmem_ld8(tacc + NT + c0, x);
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = (v[e] + x[e]) + sBias[c0 + e];
                if (a.chain) restage(c0, o);
                else store_y(c0, o);
            }
            if (a.chain) tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a.chain ? barE1 : accEmpty(buf));   // no chain: the accumulator buffer is free again
            if (!a.chain) continue;
            // ---- stage 1: conv1 + bias, gate with U ----
            mbar_wait_sleep(barC1, (uint32_t)i & 1u);
            tc_fence_after();
            for (int c0 = 0; c0 < NT; c0 += 8) {
                float v[8], o[8];
                float4 u0 = f4zero(), u1 = f4zero();
                if (live) {
                    if (a.vec32) ldg8_stream(up + c0, u0, u1);
                    else { u0 = ldg4_stream(up + c0); u1 = ldg4_stream(up + c0 + 4); }
                }
                float x[8];
                tmem_ld8(tX + c0, v);
                tmem_ld8(tX + NT + c0, x);
                const float uv[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = ((v[e] + x[e]) + sBias[128 + c0 + e]) * uv[e];
                if (a.chain == 2) restage(c0, o);
                else store_y(c0, o);
            }
            if (a.chain == 2) {
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(barE2);
                // ---- stage 2: proj_2 + bias + residual ----
                mbar_wait_sleep(barC2, (uint32_t)i & 1u);
                tc_fence_after();
                for (int c0 = 0; c0 < NT; c0 += 8) {
                    float v[8], o[8];
                    float4 r0 = f4zero(), r1 = f4zero();
                    if (live) {
                        if (a.vec32) ldg8_stream(rp + c0, r0, r1);
                        else { r0 = ldg4_stream(rp + c0); r1 = ldg4_stream(rp + c0 + 4); }
                    }
                    float x[8];
                    tmem_ld8(tX + c0, v);
                    tmem_ld8(tX + NT + c0, x);
                    const float rv[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = (v[e] + x[e]) + sBias[256 + c0 + e] + rv[e];
                    store_y(c0, o);
                }
            }
            // the chain's X (= this buffer) and the A operand are fully consumed: the buffer may take tile i+2's accumulator
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
        tmem_dealloc(tmem_base, PS_TM_COLS);
    }
}

// weight [Co][C][taps] -> Bp[chunk][tap][4 planes][hi NT rows | lo NT rows][8]   (one N tile)
__global__ void ps_pack_weight_kernel(const float *__restrict__ w, __nv_bfloat16 *__restrict__ bp, int Co, int C, int taps, int NT)
{
    const i64 total = (i64)(C / PS_KC) * taps * PS_KC * NT;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        const int e = (int)(i % 8);
        const int n = (int)((i / 8) % NT);
        const int p = (int)((i / (8 * NT)) % (PS_KC / 8));
        const int tap = (int)((i / ((i64)PS_KC * NT)) % taps);
        const int ch = (int)(i / ((i64)PS_KC * NT * taps));
        const int c = ch * PS_KC + p * 8 + e;
        const float v = n < Co ? w[((i64)n * C + c) * taps + tap] : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        const i64 slot = ((i64)ch * taps + tap) * (2 * PS_KC * NT);   // plane p = 2 NT rows of 8 channels: hi rows, then lo rows
        bp[slot + ((i64)p * 2 * NT + n) * 8 + e] = hi;
        bp[slot + ((i64)p * 2 * NT + NT + n) * 8 + e] = lo;
    }
}

size_t ps_smem_bytes(int NT)
{
    return (size_t)PS_SA * PS_ASLOT + (size_t)PS_SB * 2 * (PS_KC / 8) * NT * 16 + (size_t)PS_SP * PS_PSTAGE +
           3 * 128 * sizeof(float) + (2 * PS_SA + 2 * PS_SB + 2 * PS_SP + 8) * 8 + 16 + 128;
}

}  // namespace

// groups 1, dg 1, stride 1, one N tile; chain: C == Co <= 96.  X must be chunk-major (see deform_ps.cu header).
bool deform3d_ps_supported(const IgemmArgs &a, int chain_stages)
{
    const ConvGeo &g = a.geo;
    if (a.mode != IGEMM_DEFORM || g.ndim != 3 || g.groups != 1 || g.dg != 1 || a.Mask || a.epi != EPI_NONE) return false;
    if (g.sd != 1 || g.sh != 1 || g.sw != 1) return false;
    if (g.C % PS_KC != 0 || g.W < 2) return false;
    if (tc_nt(g.Co) > 128 || g.Co > tc_nt(g.Co)) return false;   // a single N tile
    if (chain_stages && (g.C != g.Co || g.C > 96)) return false;
    if ((i64)g.D * g.H * g.W * 128 >= ((i64)1 << 32)) return false;   // 32-bit byte offsets inside one (chunk, sample) volume
    if ((i64)g.B * g.Do * g.Ho * g.Wo >= ((i64)1 << 31)) return false;
    return true;
}

size_t deform3d_ps_offset_floats(int B, int D, int H, int W, int cols)
{
    return (size_t)B * cdiv(D, PS_BD) * cdiv(H, PS_BH) * cdiv(W, PS_BW) * cols * 128;
}

size_t deform3d_ps_packed_bytes(int Co, int C, int taps) { return (size_t)2 * C * taps * tc_nt(Co) * sizeof(__nv_bfloat16); }

int deform3d_ps_pack(const float *w, void *bp, int Co, int C, int taps, cudaStream_t st)
{
    if (pack_skipped()) return DLKA_OK;   // prepacked weights: see PackSkipScope
    const int NT = tc_nt(Co);
    const i64 total = (i64)taps * C * NT;
    const int blocks = (int)(cdiv(total, 256) < 148 * 8 ? cdiv(total, 256) : 148 * 8);
    DLKA_LAUNCH("pack_weight_ps", st, ps_pack_weight_kernel<<<blocks, 256, 0, st>>>(w, (__nv_bfloat16 *)bp, Co, C, taps, NT));
    return DLKA_OK;
}

// ga.X: chunk-major gather source; xch: floats between 32-channel chunks.  Weights are expected PACKED (deform3d_ps_pack).
// off_mode: 0 = ga.Off is [M][ldOff] rows, 1 = brick-major [tile][3K][128] (conv_tiled_ex ... brick output), 2 = NCDHW [B][3K][S]
int deform3d_ps(const IgemmArgs &ga, i64 xch, const void *bp, const DeformChain *chain, int off_mode, cudaStream_t st)
{
    const int stages = chain ? chain->stages : 0;
    if (!deform3d_ps_supported(ga, stages)) return DLKA_ERR_UNSUPPORTED;
    const ConvGeo &g = ga.geo;
    if (ga.M <= 0) return DLKA_OK;
    DeformPsArgs a;
    memset(&a, 0, sizeof(a));
    a.g = g; a.X = ga.X; a.xch = xch; a.Off = ga.Off; a.ldOff = ga.ldOff ? ga.ldOff : 3 * g.K;
    a.off_mode = off_mode;
    a.off_cs = off_mode == 1 ? 128 : off_mode == 2 ? (i64)g.Do * g.Ho * g.Wo : 1; a.bias = ga.bias; a.Y = ga.Y; a.ldY = ga.ldY;
    a.NT = tc_nt(g.Co);
    a.Bp = (const uint8_t *)bp;
    a.chain = stages;
    if (stages) {
        a.W1p = (const uint8_t *)chain->W1p; a.b1 = chain->b1; a.U = chain->U; a.ldU = chain->ldU;
        a.W2p = (const uint8_t *)chain->W2p; a.b2 = chain->b2; a.R = chain->R; a.ldR = chain->ldR;
    }
    a.tiles_d = (int)cdiv(g.Do, PS_BD); a.tiles_h = (int)cdiv(g.Ho, PS_BH); a.tiles_w = (int)cdiv(g.Wo, PS_BW);
    a.ntiles = g.B * a.tiles_d * a.tiles_h * a.tiles_w;
    a.KS = (g.C / PS_KC) * g.K;
    a.S1 = a.KS / 8 + 2 < a.KS - 1 ? a.KS / 8 + 2 : a.KS - 1;
    a.S2 = a.KS / 2 > a.S1 ? a.KS / 2 : a.S1;
    if (a.S2 > a.KS - 1) a.S2 = a.KS - 1;
    {
        auto al32 = [](const void *p, int ld) { return p == nullptr || (((uintptr_t)p & 31) == 0 && (ld & 7) == 0); };
        a.vec32 = al32(a.Y, a.ldY) && al32(a.U, a.ldU) && al32(a.R, a.ldR) && (g.Co & 7) == 0;
    }
    const size_t smem = ps_smem_bytes(a.NT);
    static SmemOptIn optin;
    DLKA_TRY(optin.ensure(deform3d_ps_kernel, smem));
    int dev = 0, sms = 148;
    DLKA_CUDA_TRY(cudaGetDevice(&dev));
    DLKA_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = a.ntiles < sms ? a.ntiles : sms;
    DLKA_LAUNCH(stages ? "tc_deform3d_chain" : "tc_deform3d", st, (deform3d_ps_kernel<<<grid, PS_THREADS, smem, st>>>(a)));
    return DLKA_OK;
}

}  // namespace dlka
