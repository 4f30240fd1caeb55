// igemm_simt.cu -- fp32 CUDA-core implicit GEMM over channels-last volumes (DLKA_MATH_FP32_SIMT).
//
// One kernel, three A-operand loaders:
//   * dense      : plain 1x1(x1) projection rows            (proj_1, conv1, proj_2)
//   * conv       : zero-padded regular conv taps            (offset nets; conv_offset 3x3x3)
//   * deformable : bi/trilinear sample per (row, tap)       (deformable conv, no im2col buffer)
// and a fused epilogue (bias, exact-erf GELU, gate multiply, residual add).
// This is the exact-fp32 validation path; the tensor-core path lives in mma_tc.cu.
#include "kernels.cuh"

namespace dlka {

namespace {

constexpr int BM = 128, BK = 16, NTHREADS = 256;

struct RowPos {
    int b, d, h, w;
    bool valid;
};

__device__ __forceinline__ RowPos decode_row(i64 m, i64 M, int Do, int Ho, int Wo)
{
    RowPos r;
    r.valid = m < M;
    if (!r.valid) m = 0;
    r.w = (int)(m % Wo);
    i64 t = m / Wo;
    r.h = (int)(t % Ho);
    t /= Ho;
    r.d = (int)(t % Do);
    r.b = (int)(t / Do);
    return r;
}

// A[m][k] for k = tap*Cg + c ; loads 4 consecutive channels.
template <int MODE>
__device__ __forceinline__ float4 load_a4(const IgemmArgs &a, const RowPos &r, i64 m, int k, int group)
{
    const ConvGeo &g = a.geo;
    const int Cg = g.C / g.groups;
    if (!r.valid || k >= a.Ktot) return f4zero();
    const int tap = k / Cg, c = k - tap * Cg + group * Cg;
    if (MODE == IGEMM_DENSE) {
        return ldg4(a.X + m * (i64)a.ldX + c);
    } else {
        const int kk = tap % g.kw, jj = (tap / g.kw) % g.kh, ii = tap / (g.kw * g.kh);
        if (MODE == IGEMM_CONV) {
            const int d = r.d * g.sd - g.pd + ii * g.dd, h = r.h * g.sh - g.ph + jj * g.dh, w = r.w * g.sw - g.pw + kk * g.dw;
            if ((unsigned)d >= (unsigned)g.D || (unsigned)h >= (unsigned)g.H || (unsigned)w >= (unsigned)g.W) return f4zero();
            return ldg4(a.X + ((((i64)r.b * g.D + d) * g.H + h) * g.W + w) * (i64)g.C + c);
        } else {  // IGEMM_DEFORM
            const int dgi = c / (g.C / g.dg);
            const float *vol = a.X + (i64)r.b * g.D * g.H * g.W * g.C + c;
            if (g.ndim == 3) {
                const float *off = a.Off + m * (i64)(a.ldOff ? a.ldOff : g.dg * 3 * g.K) + (dgi * g.K + tap) * 3;
                const float pd = sample_pos(r.d, g.sd, g.pd, ii, g.dd, __ldg(off));
                const float ph = sample_pos(r.h, g.sh, g.ph, jj, g.dh, __ldg(off + 1));
                const float pw = sample_pos(r.w, g.sw, g.pw, kk, g.dw, __ldg(off + 2));
                const Sample3 s = make_sample3(pd, ph, pw, g.D, g.H, g.W);
                return trilinear4(vol, s, g.H, g.W, g.C);
            } else {
                const float *off = a.Off + m * (i64)(a.ldOff ? a.ldOff : g.dg * 2 * g.K) + (dgi * g.K + tap) * 2;
                const float ph = sample_pos(r.h, g.sh, g.ph, jj, g.dh, __ldg(off));
                const float pw = sample_pos(r.w, g.sw, g.pw, kk, g.dw, __ldg(off + 1));
                const Sample2 s = make_sample2(ph, pw, g.H, g.W);
                float4 v = bilinear4(vol, s, g.W, g.C);
                if (a.Mask) {
                    const float mk = __ldg(a.Mask + m * (i64)(g.dg * g.K) + dgi * g.K + tap);
                    v.x *= mk; v.y *= mk; v.z *= mk; v.w *= mk;
                }
                return v;
            }
        }
    }
}

template <int MODE, int TN>
__global__ void __launch_bounds__(NTHREADS) igemm_simt_kernel(const IgemmArgs a)
{
    constexpr int BN = 16 * TN;
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int group = blockIdx.z;
    const i64 m0 = (i64)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int Ng = a.geo.Co / a.geo.groups;  // real output channels per group
    const float *Wp = a.Wp + (i64)group * a.Ktot * a.Npad;

    // A-fill role: row = tid % 128, float4 slots q and q+2 of the BK=16 chunk
    const int frow = tid & (BM - 1), fq = tid >> 7;
    const RowPos rp = decode_row(m0 + frow, a.M, a.geo.Do, a.geo.Ho, a.geo.Wo);

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < a.Ktot; k0 += BK) {
        const float4 a0 = load_a4<MODE>(a, rp, m0 + frow, k0 + 4 * fq, group);
        const float4 a1 = load_a4<MODE>(a, rp, m0 + frow, k0 + 4 * (fq + 2), group);
        __syncthreads();  // previous tile fully consumed
        As[4 * fq + 0][frow] = a0.x; As[4 * fq + 1][frow] = a0.y; As[4 * fq + 2][frow] = a0.z; As[4 * fq + 3][frow] = a0.w;
        As[4 * fq + 8][frow] = a1.x; As[4 * fq + 9][frow] = a1.y; As[4 * fq + 10][frow] = a1.z; As[4 * fq + 11][frow] = a1.w;
        for (int i = tid; i < BK * BN / 4; i += NTHREADS) {
            const int kk = i / (BN / 4), nn = (i % (BN / 4)) * 4;
            float4 v = f4zero();
            if (k0 + kk < a.Ktot) v = ldg4(Wp + (i64)(k0 + kk) * a.Npad + n0 + nn);
            *reinterpret_cast<float4 *>(&Bs[kk][nn]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 x0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8]);
            const float4 x1 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8 + 4]);
            const float av[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            float bv[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }

    // epilogue: bias -> (GELU | *U | +S) -> store channels-last
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const i64 m = m0 + ty * 8 + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx + 16 * j;
            if (n >= Ng) continue;
            const int nc = group * Ng + n;
            float v = acc[i][j] + (a.bias ? __ldg(a.bias + nc) : 0.f);
            if (a.epi == EPI_GELU) v = gelu_erf_libm(v);
            else if (a.epi == EPI_MUL) v *= __ldg(a.E + m * (i64)a.ldE + nc);
            else if (a.epi == EPI_ADD) v += __ldg(a.E + m * (i64)a.ldE + nc);
            a.Y[m * (i64)a.ldY + nc] = v;
        }
    }
}

// weight [Co][Cg][taps] (PyTorch conv layout) -> Wp[group][tap*Cg + c][Npad], zero padded columns
__global__ void pack_weight_kernel(const float *__restrict__ w, float *__restrict__ wp, int Co, int Cg, int taps, int groups,
                                   int Npad)
{
    const int Ng = Co / groups, Ktot = taps * Cg;
    const i64 total = (i64)groups * Ktot * Npad;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        const int n = (int)(i % Npad);
        const int k = (int)((i / Npad) % Ktot);
        const int g = (int)(i / ((i64)Npad * Ktot));
        const int tap = k / Cg, c = k % Cg;
        wp[i] = n < Ng ? w[((i64)(g * Ng + n) * Cg + c) * taps + tap] : 0.f;
    }
}

template <int MODE>
int launch_mode(const IgemmArgs &a, dim3 grid, int tn, cudaStream_t st)
{
    const char *name = MODE == IGEMM_DENSE ? "igemm_simt_dense" : MODE == IGEMM_CONV ? "igemm_simt_conv" : "igemm_simt_deform";
    switch (tn) {
    case 2: DLKA_LAUNCH(name, st, (igemm_simt_kernel<MODE, 2><<<grid, NTHREADS, 0, st>>>(a))); break;
    case 4: DLKA_LAUNCH(name, st, (igemm_simt_kernel<MODE, 4><<<grid, NTHREADS, 0, st>>>(a))); break;
    case 6: DLKA_LAUNCH(name, st, (igemm_simt_kernel<MODE, 6><<<grid, NTHREADS, 0, st>>>(a))); break;
    default: DLKA_LAUNCH(name, st, (igemm_simt_kernel<MODE, 8><<<grid, NTHREADS, 0, st>>>(a))); break;
    }
    return DLKA_OK;
}

}  // namespace

int igemm_simt_npad(int n_per_group)
{
    if (n_per_group <= 32) return 32;
    if (n_per_group <= 64) return 64;
    if (n_per_group <= 96) return 96;
    return (int)cdiv(n_per_group, 128) * 128;
}

int pack_weight(const float *w, float *wp, int Co, int Cg, int taps, int groups, int Npad, cudaStream_t st)
{
    if (pack_skipped()) return DLKA_OK;   // prepacked weights: see PackSkipScope
    const i64 total = (i64)groups * taps * Cg * Npad;
    const int blocks = (int)(cdiv(total, 256) < 1184 ? cdiv(total, 256) : 1184);
    DLKA_LAUNCH("pack_weight", st, pack_weight_kernel<<<blocks, 256, 0, st>>>(w, wp, Co, Cg, taps, groups, Npad));
    return DLKA_OK;
}

int igemm_simt(const IgemmArgs &a, cudaStream_t st)
{
    const int Cg = a.geo.C / a.geo.groups;
    if (Cg % 4 != 0 || a.ldX % 4 != 0) return DLKA_ERR_UNSUPPORTED;
    if (a.M <= 0) return DLKA_OK;
    const int Ng = a.geo.Co / a.geo.groups;
    const int bn = a.Npad < 128 ? a.Npad : 128;
    dim3 grid((unsigned)cdiv(a.M, BM), (unsigned)cdiv(Ng, bn), (unsigned)a.geo.groups);
    const int tn = bn / 16;
    switch (a.mode) {
    case IGEMM_DENSE: return launch_mode<IGEMM_DENSE>(a, grid, tn, st);
    case IGEMM_CONV: return launch_mode<IGEMM_CONV>(a, grid, tn, st);
    default: return launch_mode<IGEMM_DEFORM>(a, grid, tn, st);
    }
}

}  // namespace dlka
