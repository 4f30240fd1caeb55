// deform_bwd.cu -- backward of the 3D deformable convolution, any group / deformable_group.
// (Weight groups run through the dense kernels as a block-diagonal [K*C][Co] matrix with explicit zeros: G times the flops of a
// grouped GEMM, exact results; the D-LKA block itself only uses group = deformable_group = 1.)
//
// Replaces D3D.deform_conv_backward (3D/dcn/src/cuda/deform_conv_cuda.cu:128-285: at::mm -> columns, deformable_col2im_coord
// -> grad_offset, deformable_col2im -> grad_input, deformable_im2col + at::addmm -> grad_weight, at::addmv -> grad_bias) with
// the same mathematics on channels-last data, streamed over row chunks so the "columns" buffer is bounded (the reference
// allocates C*K x B*Do*Ho*Wo floats: 21.7 GB at the headline shape):
//   per chunk of Mc output voxels
//     gcol [Mc][K*C]      = gout [Mc][Co] . Wt^T            tcgen05 dense kernel (Wt[(tap, c)][co] = W[co][c][tap])
//     grad_input, grad_offset <- gcol                        one thread per (voxel, tap): trilinear scatter with vector
//                                                            reductions (red.global.add.v4.f32) + coordinate gradients
//     col  [Mc][K*C]      = trilinear samples of the input   (same sampler as the forward pass)
//     gWt  [K*C][Co]     += col^T . gout                     tcgen05 dense kernel, K dimension = the chunk's rows
// The gradients are those of the forward definition (validity and corner masks of dmcn_im2col_bilinear, cuh:30-65; floor has
// zero derivative), i.e. what dmcn_get_gradient_weight / dmcn_get_coordinate_weight (cuh:74-190) compute.  The reference's
// col2im_coord kernel indexes the offset tensor with pad_h/pad_w mixed up (cuh:448): that defect is NOT
// reproduced; parity is against autograd through a pure-torch restatement of the forward pass (oracle/oracle.py).
#include "kernels.cuh"

namespace dlka {
namespace {

// Wt[(tap*C + c)][co] = W[co][c - c0(co)][tap] when c lies in co's weight group, else 0
// (W: [Co][C/G][K] as in the reference's state_dict; group of co = co / (Co/G), its input channels c0 .. c0 + C/G)
__global__ void bwd_pack_wt_kernel(const float *__restrict__ w, float *__restrict__ wt, int Co, int C, int K, int G)
{
    const i64 total = (i64)Co * C * K;
    const int cpg = C / G, copg = Co / G;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        const int co = (int)(i % Co);
        const int c = (int)((i / Co) % C);
        const int tap = (int)(i / ((i64)Co * C));
        wt[i] = (c / cpg == co / copg) ? w[((i64)co * cpg + c % cpg) * K + tap] : 0.f;
    }
}

// gW[co][cl][tap] = gWt[(tap*C + c0(co) + cl)][co]   (the out-of-group entries of gWt are not gradients of anything)
__global__ void bwd_unpack_gw_kernel(const float *__restrict__ gwt, float *__restrict__ gw, int Co, int C, int K, int G)
{
    const int cpg = C / G, copg = Co / G;
    const i64 total = (i64)Co * cpg * K;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        const int tap = (int)(i % K);
        const int cl = (int)((i / K) % cpg);
        const int co = (int)(i / ((i64)K * cpg));
        gw[i] = gwt[((i64)tap * C + (co / copg) * cpg + cl) * Co + co];
    }
}

// grad_bias[co] = sum_m gout[m][co]      (at::addmv with a ones vector, deform_conv_cuda.cu:274)
__global__ void __launch_bounds__(256) bwd_bias_kernel(const float *__restrict__ gout, float *__restrict__ gb, i64 M, int Co)
{
    const int co = blockIdx.x;
    float s = 0.f;
    for (i64 m = threadIdx.x; m < M; m += blockDim.x) s += gout[m * Co + co];
    __shared__ float red[256];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) gb[co] = red[0];
}

__device__ __forceinline__ void red_add4(float *p, const float4 &v)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct BwdRow {
    int b, d, h, w;
};

__device__ __forceinline__ BwdRow decode_row(const ConvGeo &g, i64 m)
{
    BwdRow r;
    r.w = (int)(m % g.Wo);
    i64 t = m / g.Wo;
    r.h = (int)(t % g.Ho);
    t /= g.Ho;
    r.d = (int)(t % g.Do);
    r.b = (int)(t / g.Do);
    return r;
}

// LPG lanes per (row of the chunk, tap): lane l owns channels 4l, 4l + 4*LPG, ... so gcol reads, corner reads and the vector
// reductions into grad_input are contiguous across the group; the three coordinate gradients are reduced with shuffles.
//   grad_input[corner] += w_corner * gcol           (deformable_col2im, cuh:300-350)
//   grad_offset[m][3*tap + a] = sum_c gcol[c] * d val_c / d p_a   (deformable_col2im_coord, cuh:352-405)
template <int LPG>
__global__ void __launch_bounds__(256) bwd_scatter_kernel(const float *__restrict__ gcol, const float *__restrict__ x,
                                                          const float *__restrict__ off, float *__restrict__ gin,
                                                          float *__restrict__ goff, ConvGeo g, i64 m0, int Mc, i64 M)
{
    const int K = g.K, C = g.C, dg = g.dg, cpd = C / dg;   // cpd: channels per deformable group (share one sampling position)
    const i64 total = (i64)Mc * K * dg;                  // Mc: valid rows of this chunk
    const int lg = threadIdx.x % LPG;
    const i64 ngroups = (i64)gridDim.x * (blockDim.x / LPG);
    // the loop condition is warp-uniform (index of the warp's first group), so the full-mask shuffles below are executed by
    // all 32 lanes even in the last, partially filled round
    const int gw_ = (threadIdx.x % 32) / LPG;            // group within the warp
    for (i64 i = (i64)blockIdx.x * (blockDim.x / LPG) + threadIdx.x / LPG; i - gw_ < total; i += ngroups) {
        const bool act = i < total;
        const i64 ic = act ? i : total - 1;
        const int dgi = (int)(ic % dg);
        const int tap = (int)((ic / dg) % K);
        const int r = (int)(ic / ((i64)dg * K));
        const i64 m = m0 + r;
        const BwdRow ro = decode_row(g, m);
        const int kk = tap % g.kw, jj = (tap / g.kw) % g.kh, ii = tap / (g.kw * g.kh);
        const i64 oidx = m * (3 * (i64)K * dg) + 3 * ((i64)dgi * K + tap);   // offset channel = ((dgi * K + tap) * 3 + axis)
        const float *o = off + oidx;
        const float pd = sample_pos(ro.d, g.sd, g.pd, ii, g.dd, __ldg(o));
        const float ph = sample_pos(ro.h, g.sh, g.ph, jj, g.dh, __ldg(o + 1));
        const float pw = sample_pos(ro.w, g.sw, g.pw, kk, g.dw, __ldg(o + 2));
        const Sample3 s = make_sample3(pd, ph, pw, g.D, g.H, g.W);
        float gd = 0.f, gh = 0.f, gw = 0.f;
        if (act && (s.mask & 1)) {
            const float ld = s.l[0], lh = s.l[1], lw = s.l[2], hd = 1.f - ld, hh = 1.f - lh, hw = 1.f - lw;
            const i64 sW = C, sH = (i64)g.W * C, sD = (i64)g.H * g.W * C;
            const i64 base = (i64)ro.b * g.D * sD + (i64)s.lo[0] * sD + (i64)s.lo[1] * sH + (i64)s.lo[2] * sW;
            // corner order and validity bits as make_sample3: (d0|d1, h0|h1, w0|w1)
            const float wt[8] = {hd * hh * hw, hd * hh * lw, hd * lh * hw, hd * lh * lw, ld * hh * hw, ld * hh * lw, ld * lh * hw, ld * lh * lw};
            const float cd[8] = {-hh * hw, -hh * lw, -lh * hw, -lh * lw, hh * hw, hh * lw, lh * hw, lh * lw};   // d weight / d p_d
            const float ch[8] = {-hd * hw, -hd * lw, hd * hw, hd * lw, -ld * hw, -ld * lw, ld * hw, ld * lw};   // d weight / d p_h
            const float cw[8] = {-hd * hh, hd * hh, -hd * lh, hd * lh, -ld * hh, ld * hh, -ld * lh, ld * lh};   // d weight / d p_w
            const float *gc = gcol + (i64)r * K * C + (i64)tap * C;
            for (int c = dgi * cpd + lg * 4; c < (dgi + 1) * cpd; c += LPG * 4) {
                const float4 gv = *reinterpret_cast<const float4 *>(gc + c);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (!(s.mask & (2 << q))) continue;
                    const i64 a = base + ((q >> 2) & 1) * sD + ((q >> 1) & 1) * sH + (q & 1) * sW + c;
                    red_add4(gin + a, make_float4(wt[q] * gv.x, wt[q] * gv.y, wt[q] * gv.z, wt[q] * gv.w));
                    const float4 xv = ldg4(x + a);
                    const float dot = (gv.x * xv.x + gv.y * xv.y) + (gv.z * xv.z + gv.w * xv.w);
                    gd = fmaf(cd[q], dot, gd);
                    gh = fmaf(ch[q], dot, gh);
                    gw = fmaf(cw[q], dot, gw);
                }
            }
        }
#pragma unroll
        for (int o2 = LPG / 2; o2 > 0; o2 >>= 1) {   // every lane of the group runs the same trip count: full-mask shuffles are safe
            gd += __shfl_xor_sync(0xffffffffu, gd, o2);
            gh += __shfl_xor_sync(0xffffffffu, gh, o2);
            gw += __shfl_xor_sync(0xffffffffu, gw, o2);
        }
        if (act && lg == 0) {
            float *go = goff + oidx;
            go[0] = gd; go[1] = gh; go[2] = gw;
        }
    }
}

// gwt (+)= sum_z partial[z]
__global__ void bwd_reduce_partials_kernel(const float *__restrict__ partial, float *__restrict__ gwt, i64 n, int nsplit, int accumulate)
{
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        float s = accumulate ? gwt[i] : 0.f;
        for (int z = 0; z < nsplit; ++z) s += partial[(i64)z * n + i];
        gwt[i] = s;
    }
}

// col[r][tap*C + c] = trilinear sample (the forward pass's im2col row, cuh:192-265), zero rows past M
__global__ void __launch_bounds__(256) bwd_im2col_kernel(const float *__restrict__ x, const float *__restrict__ off,
                                                         float *__restrict__ col, ConvGeo g, i64 m0, int Mc, i64 M)
{
    const int K = g.K, C = g.C, C4 = C / 4;
    const i64 total = (i64)Mc * K * C4;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        const int tap = (int)((i / C4) % K);
        const int r = (int)(i / ((i64)C4 * K));
        const i64 m = m0 + r;
        float4 v = f4zero();
        if (m < M) {
            const BwdRow ro = decode_row(g, m);
            const int kk = tap % g.kw, jj = (tap / g.kw) % g.kh, ii = tap / (g.kw * g.kh);
            const float *o = off + m * (3 * (i64)K * g.dg) + 3 * ((i64)(c / (C / g.dg)) * K + tap);
            const float pd = sample_pos(ro.d, g.sd, g.pd, ii, g.dd, o[0]);
            const float ph = sample_pos(ro.h, g.sh, g.ph, jj, g.dh, o[1]);
            const float pw = sample_pos(ro.w, g.sw, g.pw, kk, g.dw, o[2]);
            const Sample3 s = make_sample3(pd, ph, pw, g.D, g.H, g.W);
            v = trilinear4(x + (i64)ro.b * g.D * g.H * g.W * C + c, s, g.H, g.W, C);
        }
        *reinterpret_cast<float4 *>(col + i * 4) = v;
    }
}

// out[r][c] = in[m0 + r][c] for r < valid rows, zero after (chunk of gout with zero padding for the K-dimension GEMM)
__global__ void bwd_copy_rows_kernel(const float *__restrict__ in, float *__restrict__ out, i64 m0, int Mc, i64 M, int Co)
{
    const i64 total = (i64)Mc * Co;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        const i64 m = m0 + i / Co;
        out[i] = m < M ? in[m * Co + i % Co] : 0.f;
    }
}

int grid_for(i64 total, int threads) { return (int)(cdiv(total, threads) < 148 * 32 ? cdiv(total, threads) : 148 * 32); }

}  // namespace

// out = sum_z partial[z] (n elements per slice): reduction of K-split partial products
int reduce_partials(const float *partial, float *out, i64 n, int nsplit, cudaStream_t st)
{
    if (n <= 0) return DLKA_OK;
    DLKA_LAUNCH("reduce_partials", st, bwd_reduce_partials_kernel<<<grid_for(n, 256), 256, 0, st>>>(partial, out, n, nsplit, 0));
    return DLKA_OK;
}

int deform3d_bwd_chunk_rows(i64 M) { return (int)(M < 8192 ? cdiv(M, 64) * 64 : 8192); }

// all pointers channels-last; gin / gwt zero-initialised by the caller; workspace pieces supplied by the caller (api.cu)
int deform3d_backward_cl(const ConvGeo &g, const float *x, const float *off, const float *w, const float *gout, float *gin, float *goff,
                         float *gw, float *gb, float *wt, float *gwt, float *colbuf, float *colT, float *gchunk, float *gchunkT,
                         float *partial, float *wscratch, int math, cudaStream_t st)
{
    const int K = g.K, C = g.C, Co = g.Co, KC = K * C;
    const i64 M = (i64)g.B * g.Do * g.Ho * g.Wo;
    if (M <= 0) return DLKA_OK;
    const int Mc = deform3d_bwd_chunk_rows(M);
    DLKA_LAUNCH("bwd_pack_wt", st, bwd_pack_wt_kernel<<<grid_for((i64)Co * KC, 256), 256, 0, st>>>(w, wt, Co, C, K, g.groups));
    DLKA_LAUNCH("bwd_bias", st, bwd_bias_kernel<<<Co, 256, 0, st>>>(gout, gb, M, Co));
    bool first = true;
    for (i64 m0 = 0; m0 < M; m0 += Mc) {
        const i64 rows = M - m0 < Mc ? M - m0 : Mc;
        // columns = W^T . grad_output (deform_conv_cuda.cu:226-230), for this chunk only
        DLKA_TRY(dense_cl(gout + m0 * Co, Co, rows, Co, KC, wt, nullptr, EPI_NONE, nullptr, 0, colbuf, KC, math, wscratch, st));
        {
            const int v4 = C / g.dg / 4;   // float4 per (row, tap, deformable group)
            const int lpg = v4 >= 32 ? 32 : v4 > 8 ? 16 : v4 > 4 ? 8 : 4;   // (24 -> 16 lanes, two passes)
            const int blocks = grid_for(rows * K * g.dg * lpg, 256);
            if (lpg == 32) DLKA_LAUNCH("bwd_scatter", st, bwd_scatter_kernel<32><<<blocks, 256, 0, st>>>(colbuf, x, off, gin, goff, g, m0, (int)rows, M));
            else if (lpg == 16) DLKA_LAUNCH("bwd_scatter", st, bwd_scatter_kernel<16><<<blocks, 256, 0, st>>>(colbuf, x, off, gin, goff, g, m0, (int)rows, M));
            else if (lpg == 8) DLKA_LAUNCH("bwd_scatter", st, bwd_scatter_kernel<8><<<blocks, 256, 0, st>>>(colbuf, x, off, gin, goff, g, m0, (int)rows, M));
            else DLKA_LAUNCH("bwd_scatter", st, bwd_scatter_kernel<4><<<blocks, 256, 0, st>>>(colbuf, x, off, gin, goff, g, m0, (int)rows, M));
        }
        // grad_weight += grad_output . columns^T with columns = im2col(input) (deform_conv_cuda.cu:251-273)
        DLKA_LAUNCH("bwd_im2col", st,
                    bwd_im2col_kernel<<<grid_for((i64)Mc * K * (C / 4), 256), 256, 0, st>>>(x, off, colbuf, g, m0, Mc, M));
        DLKA_TRY(transpose_sc_to_cs(colbuf, colT, 1, KC, Mc, st));                       // [Mc][KC] -> [KC][Mc]
        DLKA_LAUNCH("bwd_copy_rows", st, bwd_copy_rows_kernel<<<grid_for((i64)Mc * Co, 256), 256, 0, st>>>(gout, gchunk, m0, Mc, M, Co));
        DLKA_TRY(transpose_sc_to_cs(gchunk, gchunkT, 1, Co, Mc, st));                    // [Mc][Co] -> [Co][Mc]
        // K dimension = the chunk's rows: split-K over grid.z (a [KC x Co] output alone would occupy only KC/128 CTAs)
        int nsplit = 0;
        const int rc = (math == DLKA_MATH_BF16X3) ? dense_splitk_cl(colT, Mc, KC, Mc, Co, gchunkT, partial, 8, &nsplit, wscratch, st)
                                                  : DLKA_ERR_UNSUPPORTED;
        if (rc == DLKA_OK) {
            DLKA_LAUNCH("bwd_reduce_partials", st,
                        bwd_reduce_partials_kernel<<<grid_for((i64)KC * Co, 256), 256, 0, st>>>(partial, gwt, (i64)KC * Co, nsplit, first ? 0 : 1));
        } else if (rc == DLKA_ERR_UNSUPPORTED) {
            DLKA_TRY(dense_cl(colT, Mc, KC, Mc, Co, gchunkT, nullptr, first ? EPI_NONE : EPI_ADD, first ? nullptr : gwt, Co, gwt, Co, math,
                              wscratch, st));
        } else {
            return rc;
        }
        first = false;
    }
    DLKA_LAUNCH("bwd_unpack_gw", st, bwd_unpack_gw_kernel<<<grid_for((i64)Co * KC, 256), 256, 0, st>>>(gwt, gw, Co, C, K, g.groups));
    return DLKA_OK;
}

}  // namespace dlka
