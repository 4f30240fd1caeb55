// dwconv_smem.cu -- depthwise K^3 stencil with dilation L (5^3/L=1 and 7^3/L=3 of LKA3d_deform,
// transformerblock.py:637-638) on CUDA cores, fp32, shared-memory plane streaming.
//
// A dilated conv with dilation L only couples voxels of the same residue mod L on every axis, so the volume is
// processed as L^3 independent sub-lattices ("phases"), each a dense K^3 conv with halo (K-1)/2.
// One CTA = (batch b, 32-channel chunk, phase, lattice tile TD x TH x TW).  Input planes of the tile
// (TH+K-1) x (TW+K-1) voxels x 32 channels (128 B per voxel) stream through a double buffer, ONE TMA tile copy
// (cp.async.bulk.tensor.5d, traversal stride L, hardware zero fill outside the volume) per plane; every
// thread owns 4 channels (float4) x R outputs along w x TD outputs along d in registers, so each input row loaded
// from shared memory feeds up to TD*K*R FMAs and each weight vector (broadcast across the warp) R FMAs.
#include "kernels.cuh"
#include "tc_ptx.cuh"
#include "tma_host.cuh"

namespace dlka {
namespace {

// tile shapes per stencil (lattice voxels): TD x TH x TW outputs per CTA, R outputs along w per thread.
// 5^3:    D,H,W = 64,128,128 divide evenly (TH = 8 variants with 6 / 8 output planes measured 1.52-1.55 ms against 1.22).
// 7^3 dil 3: the sub-lattices of the headline volume are 22 x 43 x 43: 11 x 22 tiles waste 5 % of the lanes (16 x 16: 20 %).
#ifndef DLKA_DS5_TD
#define DLKA_DS5_TD 4
#define DLKA_DS5_TH 16
#define DLKA_DS5_TW 16
#define DLKA_DS5_R 8
#endif
#ifndef DLKA_DS5_VW
#define DLKA_DS5_VW 2   // channels per thread (4: float4, 8 lanes per voxel; 2: float2, 16 lanes per voxel)
#endif
#ifndef DLKA_DS7_TD
#define DLKA_DS7_TD 4    // 4 output planes per CTA (10 input planes / 4 outputs instead of 8 / 2: fewer plane hand-offs per output);
#define DLKA_DS7_TH 11   // 11 x 22 tiles = 4 x 2 per 43 x 43 sub-lattice plane, 352 threads.  Measured at the headline shape:
#define DLKA_DS7_TW 22   // (2,15,22) 3.23 ms, (3,15,22) 3.19, (4,11,22) 3.17, (5,9,22) 3.83; the ragged last d-tile costs nothing
#define DLKA_DS7_R 11    // (td_here in the kernel)
#endif
#ifndef DLKA_DS7_VW
#define DLKA_DS7_VW 2
#endif
constexpr int DS_CCH = 32;   // channels per CTA = one 128-byte line per voxel (16 per CTA, two CTAs per SM, measured slower: 8 lanes per
                             // voxel break the conflict-free LDS.64 pattern, 4.05 / 1.76 ms)

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool valid)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// per-thread channel vector: float4 (8 lanes per voxel) or float2 (16 lanes per voxel: half the registers per thread, twice
// the warps per scheduler for the same LSU / FMA totals)
template <int VW> struct DsVec;
template <> struct DsVec<4> {
    typedef float4 T;
    static __device__ __forceinline__ T zero() { return f4zero(); }
    static __device__ __forceinline__ T ldg(const float *p) { return ldg4(p); }
    static __device__ __forceinline__ void fma(T &a, const T &w, const T &x) { fma4v(a, w, x); }
};
template <> struct DsVec<2> {
    typedef float2 T;
    static __device__ __forceinline__ T zero() { return make_float2(0.f, 0.f); }
    static __device__ __forceinline__ T ldg(const float *p) { return __ldg(reinterpret_cast<const float2 *>(p)); }
    static __device__ __forceinline__ void fma(T &a, const T &w, const T &x) { fma2v(a, w, x); }
};

// plane buffers per CTA (a third buffer fits for the 11 x 22 tile and was measured: no better, 3.14 vs 3.09 ms)
__host__ __device__ constexpr int ds_nbuf(int, int, int, int) { return 2; }

template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int KD, int K, int LD, int L, int DS_TD, int DS_TH, int DS_TW, int DS_R, int VW>
__global__ void __launch_bounds__((DS_CCH / VW) * (DS_TW / DS_R) * DS_TH, 1)
    dwconv_smem_kernel(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ wp, const float *__restrict__ bias,
                       float *__restrict__ y, int C, int D, int H, int W, int tiles_d, int tiles_h, int tiles_w, i64 ych, int yldv)
{
    typedef DsVec<VW> V;
    typedef typename V::T vec;
    constexpr int LPV = DS_CCH / VW;                          // lanes (vectors) per voxel
    constexpr int DS_THREADS = LPV * (DS_TW / DS_R) * DS_TH;
    static_assert(DS_TW % DS_R == 0 && DS_THREADS <= 1024, "tile shape");
    constexpr int P = (K - 1) / 2, PD = (KD - 1) / 2;   // halo along h / w and along d
    constexpr int PH = DS_TH + K - 1, PW = DS_TW + K - 1;   // plane extent (lattice voxels)
    constexpr int PLANE_F4 = PH * PW * (DS_CCH / 4);         // float4 elements per plane (128 B per voxel)
    constexpr int NPLANES = DS_TD + KD - 1;
    constexpr int NBUF = ds_nbuf(KD, K, DS_TH, DS_TW);       // plane buffers
    extern __shared__ __align__(128) float4 smem4[];
    float4 *sP = smem4;                                      // [2][PH][PW][8] float4 (TMA destination, 128-byte aligned)
    float4 *sW = sP + NBUF * PLANE_F4;                       // [KD*K*K][8] float4 : weights of this channel chunk
    uint64_t *bars = reinterpret_cast<uint64_t *>(sW + KD * K * K * (DS_CCH / 4));   // one per plane buffer

    const int tid = threadIdx.x;
    constexpr int WRUNS = DS_TW / DS_R;
    const int q = tid % LPV, wr = (tid / LPV) % WRUNS, hl = (tid / LPV) / WRUNS;
    // CTA decomposition: x = tile, y = phase * nchunks + chunk, z = batch
    int bid = blockIdx.x;
    const int tw = bid % tiles_w; bid /= tiles_w;
    const int th = bid % tiles_h; bid /= tiles_h;
    const int td = bid;
    const int nchunks = C / DS_CCH;
    const int chunk = blockIdx.y % nchunks, phase = blockIdx.y / nchunks;
    const int pw_ = phase % L, ph_ = (phase / L) % L, pd_ = phase / (L * L);   // pd_ < LD
    const int b = blockIdx.z;
    const int c0 = chunk * DS_CCH;
    const int zd0 = td * DS_TD, zh0 = th * DS_TH, zw0 = tw * DS_TW;   // lattice tile origin
    // ragged last d-tile: only td_here of the DS_TD output planes exist on this phase's sub-lattice, so the plane loop stops
    // early and the FMA blocks of the missing outputs are skipped (CTA-uniform branches): no wasted work when DS_TD does not
    // divide the lattice depth
    const int ld_phase = (D - pd_ + LD - 1) / LD;
    const int td_here = ld_phase - zd0 < DS_TD ? ld_phase - zd0 : DS_TD;
    const int nplanes = td_here + KD - 1;
    if (td_here <= 0) return;   // nothing issued yet: the whole CTA leaves

    const uint32_t bar0 = ptx::smem_u32(bars);
    if (tid == 0) {
        for (int i = 0; i < NBUF; ++i) ptx::mbar_init(bar0 + 8u * i, 1);
        ptx::fence_barrier_init();
    }
    // weights of the chunk -> smem ([tap][C] packed layout in global: 128 contiguous bytes per tap)
    for (int i = tid; i < KD * K * K * (DS_CCH / 4); i += DS_THREADS) {
        const int tap = i / (DS_CCH / 4), qq = i % (DS_CCH / 4);
        cp_async16(sW + i, wp + (i64)tap * C + c0 + qq * 4, true);
    }
    cp_async_commit();
    __syncthreads();   // barrier init visible before anyone waits

    // one TMA tile copy per plane: out-of-volume voxels (halo, ragged lattice edge, planes above / below) are zero-filled
    auto load_plane = [&](int s, int buf) {
        const uint32_t bar = bar0 + 8u * buf;
        ptx::mbar_arrive_expect_tx(bar, (uint32_t)(PLANE_F4 * sizeof(float4)));
        ptx::tma_load_5d(ptx::smem_u32(sP + buf * PLANE_F4), &tmap, bar, c0, pw_ + L * (zw0 - P), ph_ + L * (zh0 - P),
                         pd_ + LD * (zd0 - PD + s), b);
    };

    vec acc[DS_TD][DS_R];
    {
        const vec bv = bias ? V::ldg(bias + c0 + q * VW) : V::zero();
#pragma unroll
        for (int t = 0; t < DS_TD; ++t)
#pragma unroll
            for (int r = 0; r < DS_R; ++r) acc[t][r] = bv;
    }

    if (tid == 0)
        for (int i = 0; i < NBUF && i < nplanes; ++i) load_plane(i, i);
    cp_async_wait<0>();
    __syncthreads();   // weights landed
#pragma unroll 1
    for (int s = 0; s < nplanes; ++s) {
        const int buf = s % NBUF;
        const uint32_t bph = (uint32_t)(s / NBUF) & 1u;
        ptx::mbar_wait(bar0 + 8u * buf, bph);
        const vec *pl = reinterpret_cast<const vec *>(sP + buf * PLANE_F4);
        // plane s contributes to output t with depth tap i = s - t
#pragma unroll
        for (int j = 0; j < K; ++j) {
            vec in[DS_R + K - 1];
            const vec *row = pl + ((hl + j) * PW + wr * DS_R) * LPV + q;
#pragma unroll
            for (int e = 0; e < DS_R + K - 1; ++e) in[e] = row[e * LPV];
#pragma unroll
            for (int t = 0; t < DS_TD; ++t) {
                const int i = s - t;
                if (i < 0 || i >= KD || t >= td_here) continue;  // uniform across the CTA
                const vec *wrow = reinterpret_cast<const vec *>(sW) + ((i * K + j) * K) * LPV + q;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const vec wv = wrow[k * LPV];
#pragma unroll
                    for (int r = 0; r < DS_R; ++r) V::fma(acc[t][r], wv, in[r + k]);
                }
            }
        }
        // split hand-off: only the warp that refills buffer `buf` with plane s + NBUF waits until every warp has finished reading
        // plane s (named barrier 1 + buf: bar.sync by warp 0, bar.arrive by the others, who run on into the planes already
        // resident).  A warp can arrive for plane s + NBUF on the same barrier only after that plane's TMA, which follows this
        // bar.sync.  (3.16 -> 3.09 ms and 1.22 -> 1.16 ms with two buffers against the CTA-wide barrier.)
        if (s + NBUF < nplanes) {
            const int id = 1 + buf;
            if (tid < 32) {
                asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(DS_THREADS) : "memory");
                if (tid == 0) load_plane(s + NBUF, buf);
            } else {
                asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(DS_THREADS) : "memory");
            }
        }
    }

    // store: real coordinates of this thread's outputs.  One 64-bit address per (thread, t); the R outputs along w are L voxels apart
    const int hr = ph_ + L * (zh0 + hl);
    if (hr < H) {
        const int w0r = pw_ + L * (zw0 + wr * DS_R);           // real w of output r = 0
        const i64 wstep = (i64)L * yldv;
#pragma unroll
        for (int t = 0; t < DS_TD; ++t) {
            const int dr = pd_ + LD * (zd0 + t);
            if (dr >= D) continue;
            float *yp = y + (i64)(c0 >> 5) * ych + ((((i64)b * D + dr) * H + hr) * W + w0r) * yldv + (c0 & 31) + q * VW;
#pragma unroll
            for (int r = 0; r < DS_R; ++r)
                if (w0r + L * r < W) *reinterpret_cast<vec *>(yp + r * wstep) = acc[t][r];
        }
    }
}

template <int KD, int K, int LD, int L, int DS_TD, int DS_TH, int DS_TW, int DS_R, int VW>
int launch_ds(const float *x, const float *wp, const float *bias, float *y, int B, int C, int D, int H, int W, cudaStream_t st, bool cm)
{
    // output addressing: y + chunk * ych + voxel * yldv + channel-in-chunk.  channels-last: ych = 32, yldv = C;
    // chunk-major [C/32][B][D][H][W][32] (the gather layout of deform_ps.cu): ych = B*D*H*W*32, yldv = 32
    const i64 ych = cm ? (i64)B * D * H * W * 32 : 32;   // the chunk-major layout is defined on 32-channel chunks whatever DS_CCH is
    const int yldv = cm ? 32 : C;
    constexpr int PH = DS_TH + K - 1, PW = DS_TW + K - 1;
    constexpr int DS_THREADS = (DS_CCH / VW) * (DS_TW / DS_R) * DS_TH;
    static_assert(PW * L <= 256 && PH * L <= 256, "TMA box extent");
    const size_t smem = ((size_t)KD * K * K * (DS_CCH / 4) + ds_nbuf(KD, K, DS_TH, DS_TW) * (size_t)PH * PW * (DS_CCH / 4)) * sizeof(float4) + 64;
    auto kern = dwconv_smem_kernel<KD, K, LD, L, DS_TD, DS_TH, DS_TW, DS_R, VW>;
    static SmemOptIn optin;   // per template instance, per device
    DLKA_TRY(optin.ensure(kern, smem));
    // tensor map of the channels-last activation; box = one plane of the lattice tile (32 channels x PW x PH voxels),
    // traversal stride L along w and h picks the sub-lattice of the CTA's phase
    CUtensorMap tmap;
    if (!make_tmap_cl5(&tmap, x, B, C, D, H, W, DS_CCH, PW, PH, 1, L)) return DLKA_ERR_CUDA;
    // lattice extents of the largest phase
    const int ld = (int)cdiv(D, LD), lh = (int)cdiv(H, L), lw = (int)cdiv(W, L);
    const int tiles_d = (int)cdiv(ld, DS_TD), tiles_h = (int)cdiv(lh, DS_TH), tiles_w = (int)cdiv(lw, DS_TW);
    dim3 grid((unsigned)(tiles_d * tiles_h * tiles_w), (unsigned)(LD * L * L * (C / DS_CCH)), (unsigned)B);
    if (grid.y > 65535u || grid.z > 65535u) return DLKA_ERR_UNSUPPORTED;
    DLKA_LAUNCH(KD == 5 && K == 5 ? "dwconv3d_smem_k5" : KD == 7 ? "dwconv3d_smem_k7d3" : "dwconv3d_smem_aniso", st,
                (kern<<<grid, DS_THREADS, smem, st>>>(tmap, wp, bias, y, C, D, H, W, tiles_d, tiles_h, tiles_w, ych, yldv)));
    return DLKA_OK;
}

}  // namespace

bool dwconv_smem_supported(int C, int kd, int kh, int kw, int dd, int dh, int dw)
{
    if (C % DS_CCH != 0 || kh != kw || dh != dw) return false;
    return (kd == 5 && kh == 5 && dd == 1 && dh == 1) || (kd == 7 && kh == 7 && dd == 3 && dh == 3) ||   // synapse (transformerblock.py:637-638)
           (kd == 5 && kh == 7 && dd == 3 && dh == 3) || (kd == 3 && kh == 5 && dd == 1 && dh == 3) ||   // acdc dims 32/64, 128
           (kd == 3 && kh == 3 && dd == 1 && dh == 1);                                                    // acdc dim 256 (acdc/transformerblock.py:214-236)
}

// wp: packed [taps][C] weights (pack_dw layout)
int dwconv_smem(const float *x, const float *wp, const float *bias, float *y, int B, int C, int D, int H, int W, int kd, int k,
                int dd, int dil, cudaStream_t st, bool cm)
{
    if (kd == 5 && k == 5 && dd == 1 && dil == 1) {
        // mid-size volumes (32^3 stage of the 3D nets): with 4 output planes per CTA the grid is under half a wave
        const i64 ctas = cdiv(D, DLKA_DS5_TD) * cdiv(H, DLKA_DS5_TH) * cdiv(W, DLKA_DS5_TW) * (C / DS_CCH) * B;
        if (ctas < 148) return launch_ds<5, 5, 1, 1, 2, DLKA_DS5_TH, DLKA_DS5_TW, DLKA_DS5_R, DLKA_DS5_VW>(x, wp, bias, y, B, C, D, H, W, st, cm);
        return launch_ds<5, 5, 1, 1, DLKA_DS5_TD, DLKA_DS5_TH, DLKA_DS5_TW, DLKA_DS5_R, DLKA_DS5_VW>(x, wp, bias, y, B, C, D, H, W, st, cm);
    }
    if (kd == 7 && k == 7 && dd == 3 && dil == 3) {
        // sub-lattices of at most 12 x 12 (volumes up to 36 wide): an 11 x 12 tile instead of 15 x 22 (2.5x fewer lanes per CTA)
        if (cdiv(H, 3) <= 11 && cdiv(W, 3) <= 12) return launch_ds<7, 7, 3, 3, 2, 11, 12, 6, 2>(x, wp, bias, y, B, C, D, H, W, st, cm);
        return launch_ds<7, 7, 3, 3, DLKA_DS7_TD, DLKA_DS7_TH, DLKA_DS7_TW, DLKA_DS7_R, DLKA_DS7_VW>(x, wp, bias, y, B, C, D, H, W, st, cm);
    }
    if (kd == 5 && k == 7 && dd == 3 && dil == 3) return launch_ds<5, 7, 3, 3, 2, 15, 22, 11, 2>(x, wp, bias, y, B, C, D, H, W, st, cm);
    if (kd == 3 && k == 5 && dd == 1 && dil == 3) return launch_ds<3, 5, 1, 3, 4, 16, 16, 8, 2>(x, wp, bias, y, B, C, D, H, W, st, cm);
    if (kd == 3 && k == 3 && dd == 1 && dil == 1) return launch_ds<3, 3, 1, 1, 4, 16, 16, 8, 2>(x, wp, bias, y, B, C, D, H, W, st, cm);
    return DLKA_ERR_UNSUPPORTED;
}

}  // namespace dlka
