// edgeconv.cu -- SURVEY section 8f rank 1: the EdgeConv layer WITHOUT the (B,2C,N,k) edge tensor.
//
// The reference builds every DGCNN layer as
//     get_graph_feature(x) -> 1x1 Conv2d [-> BatchNorm2d] -> LeakyReLU -> max over the k neighbours
// (PointDA/Models.py:114-128 with conv_2d of PointDA/model_utils.py:45-63; PointSegDA/Models.py:171-184, where the
// convolutions are stacked without BatchNorm or activation).  The convolution is linear in the edge feature
// [x_j - x_i | x_i], so with W = [Wa | Wb]
//     h[b,o,i,j] = Y[b,idx[b,i,j],o] + Z[b,i,o],   Y = Wa x,  Z = (Wb - Wa) x  (+ bias)
// and the two point-wise products are one plain GEMM on (B*N, C) x (C, 2O) (library call on the host side).
// BatchNorm (training statistics over all B*N*k edges) is a per-channel affine map a*h + c with the sign of gamma,
// LeakyReLU is increasing, so   max_j lrelu(a h_j + c) = lrelu(a * max_j h_j + c)  for a >= 0; a caller with a negative
// scale folds its sign into the rows of W (h' = -h, a' = -a), so these kernels always take a max.
// What remains for this file is HBM/L2-bound index work:
//   edgeconv_reduce_kernel : per point, gather k rows of Y (L2 resident), track the extreme value and its slot, the
//                            row sum and -- for the batch statistics -- sum h and sum h^2 per channel (fp64 partials)
//   edgeconv_coeffs_kernel : mean / biased variance -> a, c, mean, invstd
//   edgeconv_apply_kernel  : out (B,O,N) = lrelu(a * hsel + c), transposed through shared memory
//   backward               : dy = g * lrelu'(.)  (+ the two BatchNorm sums), scatter of dy to the selected neighbour,
//                            and for training-mode BatchNorm the dense mean/variance terms, which need the in-degree of
//                            every point and T[j] = sum_{i : j in nbr(i)} Z[i] (128-bit reductions into L2, like edge.cu)
// Layouts are point-major (B,N,O): a thread owns one float4 of channels of one point, so every access is a coalesced
// 128-bit load and the k neighbour-row gathers of a point are whole 4*O-byte rows.
#include "common.cuh"

namespace mlsp {

static constexpr int EC_THREADS = 256;

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.0f ? v : slope * v; }

// ------------------------------------------------------------------------------------------------ forward
// grid-stride over (point, float4-of-channels) items with a stride that is a multiple of G = O/4, so a thread keeps
// its channels and the statistics stay in registers until the end.
// COOP (G a power of two): the min(G,32) lanes that share a point load its neighbour indices once, coalesced, and hand
// them round with shuffles -- the k row gathers then depend on no load and are issued five at a time.
template <bool STATS, bool COOP>
__global__ void __launch_bounds__(EC_THREADS, 4)
edgeconv_reduce_kernel(const float4 *__restrict__ yz, const int64_t *__restrict__ idx, int N, int G, int k, long long items,
                       float4 *__restrict__ hsel, uchar4 *__restrict__ slot, float4 *__restrict__ rowsum,
                       double *__restrict__ stats)
{
    __shared__ double red[EC_THREADS * 2];
    const int c4 = threadIdx.x % G;                           // blockDim.x % G == 0
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int gw = G < 32 ? G : 32;                           // lanes of this warp that work on the same point
    const int lig = lane_id() & (gw - 1), gbase = lane_id() & ~(gw - 1);
    double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += stride) {
        const long long p = it / G;                           // global point b*N + i
        const long long b = p / N;
        const float4 z = yz[p * 2 * G + G + c4];
        const float4 *yb = yz + b * N * 2 * G + c4;
        const int64_t *ip = idx + p * k;
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        uchar4 sl = make_uchar4(0, 0, 0, 0);
        float4 rs = make_float4(0.f, 0.f, 0.f, 0.f), rq = make_float4(0.f, 0.f, 0.f, 0.f);
        auto visit = [&](int j, int nb) {
            const float4 y = __ldg(yb + (long long)nb * 2 * G);
            const float4 h = make_float4(y.x + z.x, y.y + z.y, y.z + z.z, y.w + z.w);
            if (h.x > m.x) { m.x = h.x; sl.x = (unsigned char)j; }      // strict: the first maximum wins
            if (h.y > m.y) { m.y = h.y; sl.y = (unsigned char)j; }
            if (h.z > m.z) { m.z = h.z; sl.z = (unsigned char)j; }
            if (h.w > m.w) { m.w = h.w; sl.w = (unsigned char)j; }
            if (STATS) {
                rs.x += h.x; rs.y += h.y; rs.z += h.z; rs.w += h.w;
                rq.x = fmaf(h.x, h.x, rq.x); rq.y = fmaf(h.y, h.y, rq.y); rq.z = fmaf(h.z, h.z, rq.z); rq.w = fmaf(h.w, h.w, rq.w);
            }
        };
        if (COOP) {
            const unsigned mask = __activemask();             // whole groups are active or not (items % G == 0)
            for (int j0 = 0; j0 < k; j0 += gw) {
                const int cnt = min(gw, k - j0);
                const int mine = lig < cnt ? (int)ip[j0 + lig] : 0;
#pragma unroll 5
                for (int jj = 0; jj < cnt; ++jj) visit(j0 + jj, __shfl_sync(mask, mine, gbase + jj));
            }
        } else {
#pragma unroll 5
            for (int j = 0; j < k; ++j) visit(j, (int)ip[j]);
        }
        hsel[p * G + c4] = m;
        slot[p * G + c4] = sl;
        if (STATS) {
            rowsum[p * G + c4] = rs;
            s1[0] += rs.x; s1[1] += rs.y; s1[2] += rs.z; s1[3] += rs.w;
            s2[0] += rq.x; s2[1] += rq.y; s2[2] += rq.z; s2[3] += rq.w;
        }
    }
    if (STATS) {
        // threads tid, tid+G, tid+2G, ... share channels: one shared-memory pass per component, then 8 atomics per channel group
        const int O = 4 * G;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            red[threadIdx.x] = s1[w];
            red[EC_THREADS + threadIdx.x] = s2[w];
            __syncthreads();
            if ((int)threadIdx.x < G) {
                double a = 0.0, q = 0.0;
                for (int t = threadIdx.x; t < (int)blockDim.x; t += G) { a += red[t]; q += red[EC_THREADS + t]; }
                atomicAdd(stats + 4 * c4 + w, a);
                atomicAdd(stats + O + 4 * c4 + w, q);
            }
            __syncthreads();
        }
    }
}

// stats (2,O) double: sum h, sum h^2 over `count` edges -> coef (4,O): a = gamma*invstd, c = beta - a*mean, mean, invstd.
// With running_mean / running_var the kernel also makes BatchNorm's running-statistics update (torch semantics:
// r = (1-m) r + m * batch value, the variance unbiased by count/(count-1)); `sign` (O, +-1 or NULL) maps the mean back
// when the caller has folded the sign of gamma into the weight rows (h' = sign*h has mean' = sign*mean, same variance).
__global__ void edgeconv_coeffs_kernel(const double *__restrict__ stats, const float *__restrict__ gamma,
                                       const float *__restrict__ beta, int O, double count, float eps, float *__restrict__ coef,
                                       float *__restrict__ running_mean, float *__restrict__ running_var,
                                       const float *__restrict__ sign, float momentum)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= O) return;
    const double mean = stats[o] / count;
    double var = stats[O + o] / count - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const double invstd = 1.0 / sqrt(var + (double)eps);
    const double g = gamma ? (double)gamma[o] : 1.0, bt = beta ? (double)beta[o] : 0.0;
    coef[o] = (float)(g * invstd);
    coef[O + o] = (float)(bt - g * invstd * mean);
    coef[2 * O + o] = (float)mean;
    coef[3 * O + o] = (float)invstd;
    if (running_mean) {
        const float m = momentum, sg = sign ? sign[o] : 1.0f;
        const float unbiased = (float)(var * (count / (count > 1.0 ? count - 1.0 : 1.0)));
        running_mean[o] = (1.0f - m) * running_mean[o] + m * (sg * (float)mean);
        running_var[o] = (1.0f - m) * running_var[o] + m * unbiased;
    }
}

// out[b,o,i] = lrelu(a_o * hsel[b,i,o] + c_o): 32 points x 32 channels per block through a padded shared tile
__global__ void __launch_bounds__(256)
edgeconv_apply_kernel(const float *__restrict__ hsel, const float *__restrict__ coef, int N, int O, float slope,
                      float *__restrict__ out)
{
    __shared__ float tile[32][33];
    const int b = blockIdx.z, i0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int o = o0 + tx;
    const float a = o < O ? coef[o] : 0.f, c = o < O ? coef[O + o] : 0.f;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r;
        if (i < N && o < O) tile[r][tx] = lrelu(fmaf(a, hsel[((long long)b * N + i) * O + o], c), slope);
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int oo = o0 + r, i = i0 + tx;
        if (oo < O && i < N) out[((long long)b * O + oo) * N + i] = tile[tx][r];
    }
}

// ------------------------------------------------------------------------------------------------ backward
// dy[b,i,o] = g[b,o,i] * lrelu'(a*hsel + c); with SUMS also dsum[o] += dy, dsum[O+o] += dy * (hsel - mean) * invstd.
// A block owns a 32-channel x 32-point tile (both global loads of a thread are issued before the barrier) and ends
// with 64 fp64 atomics.
static constexpr int EC_PTS = 32;
template <bool SUMS>
__global__ void __launch_bounds__(256)
edgeconv_bwd_prepare_kernel(const float *__restrict__ g, long long g_bstride, const float *__restrict__ hsel,
                            const float *__restrict__ coef, int N, int O, float slope, float *__restrict__ dy,
                            double *__restrict__ dsum)
{
    __shared__ float tile[32][33];
    __shared__ double red[2][8][32];
    const int b = blockIdx.z, o0 = blockIdx.y * 32, i0 = blockIdx.x * EC_PTS;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int o = o0 + tx;
    const float a = o < O ? coef[o] : 0.f, c = o < O ? coef[O + o] : 0.f;
    const float mean = o < O ? coef[2 * O + o] : 0.f, invstd = o < O ? coef[3 * O + o] : 0.f;
    float hs[4], gv[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty + 8 * r;                       // point-major side, coalesced along o
        hs[r] = (i < N && o < O) ? hsel[((long long)b * N + i) * O + o] : 0.f;
        const int oo = o0 + ty + 8 * r, ig = i0 + tx;        // g tile, coalesced along i
        gv[r] = (oo < O && ig < N) ? g[b * g_bstride + (long long)oo * N + ig] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) tile[ty + 8 * r][tx] = gv[r];
    __syncthreads();
    double sb = 0.0, sgm = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty + 8 * r;
        if (i < N && o < O) {
            const float d = tile[tx][ty + 8 * r] * (fmaf(a, hs[r], c) > 0.0f ? 1.0f : slope);
            dy[((long long)b * N + i) * O + o] = d;
            if (SUMS) { sb += d; sgm += (double)d * (double)((hs[r] - mean) * invstd); }
        }
    }
    if (SUMS) {
        red[0][ty][tx] = sb;
        red[1][ty][tx] = sgm;
        __syncthreads();
        if (ty == 0 && o < O) {
            double x = 0.0, y = 0.0;
#pragma unroll
            for (int r = 0; r < 8; ++r) { x += red[0][r][tx]; y += red[1][r][tx]; }
            atomicAdd(dsum + o, x);
            atomicAdd(dsum + O + o, y);
        }
    }
}

// scatter: dYacc[b, idx[b,i,slot], o] += dy[b,i,o]; with DENSE also T[b, idx[b,i,j], :] += Z[b,i,:] for every j and the
// in-degree histogram deg[b, idx[b,i,j]] += 1
template <bool DENSE>
__global__ void __launch_bounds__(EC_THREADS)
edgeconv_bwd_scatter_kernel(const float4 *__restrict__ dy, const uchar4 *__restrict__ slot, const int64_t *__restrict__ idx,
                            const float4 *__restrict__ yz, int N, int G, int k, long long items, float *__restrict__ dyacc,
                            float4 *__restrict__ T, int *__restrict__ deg)
{
    const long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= items) return;
    const long long p = it / G;
    const int c4 = (int)(it - p * G);
    const long long b = p / N;
    const int64_t *ip = idx + p * k;
    const float4 d = dy[it];
    const uchar4 sl = slot[it];
    float *acc = dyacc + b * N * 4 * G + 4 * c4;
    atomicAdd(acc + ip[sl.x] * 4 * G + 0, d.x);
    atomicAdd(acc + ip[sl.y] * 4 * G + 1, d.y);
    atomicAdd(acc + ip[sl.z] * 4 * G + 2, d.z);
    atomicAdd(acc + ip[sl.w] * 4 * G + 3, d.w);
    if (DENSE) {
        const float4 z = yz[p * 2 * G + G + c4];
        float4 *tb = T + b * N * G + c4;
        int *db = deg + b * N;
#pragma unroll 4
        for (int j = 0; j < k; ++j) {
            const long long nb = ip[j];
            atomicAdd(tb + nb * G, z);                       // red.global.add.v4.f32
            if (c4 == 0) atomicAdd(db + nb, 1);
        }
    }
}

// dyz (B,N,2O) = [dY | dZ]:
//   plain : dY = a * dYacc,                                              dZ = a * dy
//   DENSE : dY = a * (dYacc - (deg*dbeta + dgamma*(deg*(Y-mean) + T)*invstd) / M)
//           dZ = a * (dy    - (k*dbeta   + dgamma*(rowsum - k*mean)*invstd) / M)
template <bool DENSE>
__global__ void __launch_bounds__(EC_THREADS)
edgeconv_bwd_finish_kernel(const float4 *__restrict__ dy, const float4 *__restrict__ dyacc, const float4 *__restrict__ T,
                           const int *__restrict__ deg, const float4 *__restrict__ rowsum, const float4 *__restrict__ yz,
                           const float *__restrict__ coef, const double *__restrict__ dsum, int G, int k, long long items,
                           double M, float4 *__restrict__ dyz)
{
    const long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= items) return;
    const long long p = it / G;
    const int c4 = (int)(it - p * G);
    const int O = 4 * G;
    const float4 a = reinterpret_cast<const float4 *>(coef)[c4];
    const float4 d = dy[it], da = dyacc[it];
    float4 oy, oz;
    if (!DENSE) {
        oy = make_float4(a.x * da.x, a.y * da.y, a.z * da.z, a.w * da.w);
        oz = make_float4(a.x * d.x, a.y * d.y, a.z * d.z, a.w * d.w);
    } else {
        const float4 mean = reinterpret_cast<const float4 *>(coef + 2 * O)[c4];
        const float4 istd = reinterpret_cast<const float4 *>(coef + 3 * O)[c4];
        const float4 y = yz[p * 2 * G + c4], t = T[it], rs = rowsum[it];
        const float dg = (float)deg[p], kf = (float)k, invM = (float)(1.0 / M);
        const float db[4] = {(float)dsum[4 * c4], (float)dsum[4 * c4 + 1], (float)dsum[4 * c4 + 2], (float)dsum[4 * c4 + 3]};
        const float dgm[4] = {(float)dsum[O + 4 * c4], (float)dsum[O + 4 * c4 + 1], (float)dsum[O + 4 * c4 + 2],
                              (float)dsum[O + 4 * c4 + 3]};
        const float av[4] = {a.x, a.y, a.z, a.w}, mv[4] = {mean.x, mean.y, mean.z, mean.w};
        const float iv[4] = {istd.x, istd.y, istd.z, istd.w}, yv[4] = {y.x, y.y, y.z, y.w}, tv[4] = {t.x, t.y, t.z, t.w};
        const float rv[4] = {rs.x, rs.y, rs.z, rs.w}, dv[4] = {d.x, d.y, d.z, d.w}, dav[4] = {da.x, da.y, da.z, da.w};
        float ry[4], rz[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            ry[w] = av[w] * (dav[w] - (dg * db[w] + dgm[w] * (dg * (yv[w] - mv[w]) + tv[w]) * iv[w]) * invM);
            rz[w] = av[w] * (dv[w] - (kf * db[w] + dgm[w] * (rv[w] - kf * mv[w]) * iv[w]) * invM);
        }
        oy = make_float4(ry[0], ry[1], ry[2], ry[3]);
        oz = make_float4(rz[0], rz[1], rz[2], rz[3]);
    }
    dyz[p * 2 * G + c4] = oy;
    dyz[p * 2 * G + G + c4] = oz;
}

// dgamma_dbeta (2,O) fp32 <- dsum (2,O) fp64: [0] = d gamma = sum dy * hhat, [1] = d beta = sum dy
__global__ void edgeconv_param_grads_kernel(const double *__restrict__ dsum, int O, float *__restrict__ out)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o < O) {
        out[o] = (float)dsum[O + o];
        out[O + o] = (float)dsum[o];
    }
}

// ---- weight side of the layer's GEMM (one launch each instead of a dozen elementwise torch kernels per layer and step)
// W (O,2C) = [Wa | Wb] over [x_j - x_i | x_i], scale (O) or NULL = the BatchNorm weight / fixed affine scale, bias (O) or NULL
//   -> Wcat (2O,C) = [s Wa ; s (Wb - Wa)],  sgn (O) = s = sign(scale) (+1 where scale >= 0 or absent),  zb (2O) = [0 ; s bias]
__global__ void edgeconv_weight_prep_kernel(const float *__restrict__ W, const float *__restrict__ scale, const float *__restrict__ bias,
                                            int O, int C, float *__restrict__ Wcat, float *__restrict__ sgn, float *__restrict__ zb)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= O * C) return;
    const int o = t / C, c = t - o * C;
    const float s = (scale && scale[o] < 0.0f) ? -1.0f : 1.0f;
    const float wa = W[(size_t)o * 2 * C + c], wb = W[(size_t)o * 2 * C + C + c];
    Wcat[(size_t)o * C + c] = s * wa;
    Wcat[(size_t)(O + o) * C + c] = s * (wb - wa);
    if (c == 0) {
        sgn[o] = s;
        if (zb) {
            zb[o] = 0.0f;
            zb[O + o] = bias ? s * bias[o] : 0.0f;
        }
    }
}

// part (Z,2O,C): partial products dyz^T x of the batch slices -> gW (O,2C) = [s (gY - gZ) | s gZ]  (d/dWa = gY - gZ, d/dWb = gZ),
// summed over Z in a fixed order (deterministic)
__global__ void edgeconv_weight_grad_kernel(const float *__restrict__ part, int Z, const float *__restrict__ sgn, int O, int C,
                                            float *__restrict__ gW)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= O * C) return;
    const int o = t / C, c = t - o * C;
    const size_t plane = (size_t)2 * O * C;
    float gy = 0.0f, gz = 0.0f;
    for (int z = 0; z < Z; ++z) {
        gy += part[z * plane + (size_t)o * C + c];
        gz += part[z * plane + (size_t)(O + o) * C + c];
    }
    const float s = sgn ? sgn[o] : 1.0f;
    gW[(size_t)o * 2 * C + c] = s * (gy - gz);
    gW[(size_t)o * 2 * C + C + c] = s * gz;
}

static int ec_block(int G) { return EC_THREADS / G * G; }      // largest multiple of G that fits a block

size_t edgeconv_workspace_bytes(int B, int O, int N)
{
    // dsum (2,O) f64 | dYacc (B,N,O) | T (B,N,O) | deg (B,N) i32   [zeroed]   | dy (B,N,O)
    const size_t pts = (size_t)B * N;
    return align_up((size_t)2 * O * 8, 256) + 3 * align_up(pts * O * 4, 256) + align_up(pts * 4, 256);
}

}  // namespace mlsp

using namespace mlsp;

extern "C" {

int mlsp_edgeconv_reduce_fwd(const float *yz, const int64_t *idx, int B, int N, int O, int k, float *hsel, uint8_t *slot,
                             float *rowsum, double *stats, void *stream)
{
    MLSP_REQUIRE(yz && idx && hsel && slot, MLSP_EINVAL, "edgeconv_reduce_fwd: null pointer");
    MLSP_REQUIRE(B > 0 && N > 0 && O > 0 && k >= 1 && k <= N, MLSP_EINVAL, "edgeconv_reduce_fwd: bad shape B=%d N=%d O=%d k=%d", B, N, O, k);
    MLSP_REQUIRE(O % 4 == 0 && O <= 4 * EC_THREADS, MLSP_EUNSUPPORTED, "edgeconv_reduce_fwd: O=%d must be a multiple of 4, <= %d", O, 4 * EC_THREADS);
    MLSP_REQUIRE(k <= 255, MLSP_EUNSUPPORTED, "edgeconv_reduce_fwd: k=%d > 255 (slots are bytes)", k);
    MLSP_REQUIRE((stats == nullptr) == (rowsum == nullptr), MLSP_EINVAL, "edgeconv_reduce_fwd: stats and rowsum go together");
    cudaStream_t s = as_stream(stream);
    const int G = O / 4, threads = ec_block(G);
    const long long items = (long long)B * N * G;
    // persistent grid: exactly the resident blocks (one wave), so that the grid-stride loop is balanced
    using Kern = void (*)(const float4 *, const int64_t *, int, int, int, long long, float4 *, uchar4 *, float4 *, double *);
    const bool coop = (G & (G - 1)) == 0;
    const Kern kern = coop ? (stats ? (Kern)edgeconv_reduce_kernel<true, true> : (Kern)edgeconv_reduce_kernel<false, true>)
                           : (stats ? (Kern)edgeconv_reduce_kernel<true, false> : (Kern)edgeconv_reduce_kernel<false, false>);
    int per_sm = 0;
    MLSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0));
    const long long want = (items + threads - 1) / threads, resident = (long long)(per_sm > 0 ? per_sm : 1) * sm_count();
    const int grid = (int)(want < resident ? want : resident);
    if (stats) MLSP_CUDA(cudaMemsetAsync(stats, 0, (size_t)2 * O * sizeof(double), s));
    kern<<<grid, threads, 0, s>>>(reinterpret_cast<const float4 *>(yz), idx, N, G, k, items, reinterpret_cast<float4 *>(hsel),
                                  reinterpret_cast<uchar4 *>(slot), reinterpret_cast<float4 *>(rowsum), stats);
    MLSP_LAUNCH_CHECK("edgeconv_reduce_kernel");
    return MLSP_OK;
}

int mlsp_edgeconv_bn_coeffs(const double *stats, const float *gamma, const float *beta, int O, double count, float eps,
                            float *coef, float *running_mean, float *running_var, const float *sign, float momentum,
                            void *stream)
{
    MLSP_REQUIRE(stats && coef, MLSP_EINVAL, "edgeconv_bn_coeffs: null pointer");
    MLSP_REQUIRE(O > 0 && count > 0, MLSP_EINVAL, "edgeconv_bn_coeffs: bad shape O=%d count=%g", O, count);
    MLSP_REQUIRE((running_mean == nullptr) == (running_var == nullptr), MLSP_EINVAL, "edgeconv_bn_coeffs: running_mean and running_var go together");
    edgeconv_coeffs_kernel<<<(O + 127) / 128, 128, 0, as_stream(stream)>>>(stats, gamma, beta, O, count, eps, coef, running_mean,
                                                                           running_var, sign, momentum);
    MLSP_LAUNCH_CHECK("edgeconv_coeffs_kernel");
    return MLSP_OK;
}

int mlsp_edgeconv_apply_fwd(const float *hsel, const float *coef, int B, int N, int O, float slope, float *out, void *stream)
{
    MLSP_REQUIRE(hsel && coef && out, MLSP_EINVAL, "edgeconv_apply_fwd: null pointer");
    MLSP_REQUIRE(B > 0 && N > 0 && O > 0 && B <= 65535, MLSP_EINVAL, "edgeconv_apply_fwd: bad shape B=%d N=%d O=%d", B, N, O);
    dim3 grid((N + 31) / 32, (O + 31) / 32, B);
    edgeconv_apply_kernel<<<grid, 256, 0, as_stream(stream)>>>(hsel, coef, N, O, slope, out);
    MLSP_LAUNCH_CHECK("edgeconv_apply_kernel");
    return MLSP_OK;
}

int mlsp_edgeconv_bwd(const float *g, int64_t g_bstride, const float *yz, const int64_t *idx, const float *hsel, const uint8_t *slot,
                      const float *rowsum, const float *coef, int B, int N, int O, int k, float slope, int bn_train,
                      float *dyz, float *dgamma_dbeta, void *ws, size_t ws_bytes, void *stream)
{
    MLSP_REQUIRE(g && yz && idx && hsel && slot && coef && dyz && ws, MLSP_EINVAL, "edgeconv_bwd: null pointer");
    MLSP_REQUIRE(g_bstride >= (int64_t)O * N, MLSP_EINVAL, "edgeconv_bwd: g batch stride %lld < O*N", (long long)g_bstride);
    MLSP_REQUIRE(B > 0 && N > 0 && O > 0 && k >= 1 && k <= N && B <= 65535, MLSP_EINVAL, "edgeconv_bwd: bad shape B=%d N=%d O=%d k=%d", B, N, O, k);
    MLSP_REQUIRE(O % 4 == 0 && O <= 4 * EC_THREADS && k <= 255, MLSP_EUNSUPPORTED, "edgeconv_bwd: O=%d k=%d unsupported", O, k);
    MLSP_REQUIRE(!bn_train || (rowsum && dgamma_dbeta), MLSP_EINVAL, "edgeconv_bwd: training-mode BatchNorm needs rowsum and dgamma_dbeta");
    MLSP_REQUIRE(ws_bytes >= edgeconv_workspace_bytes(B, O, N), MLSP_EWORKSPACE, "edgeconv_bwd: workspace too small");
    cudaStream_t s = as_stream(stream);
    const size_t pts = (size_t)B * N;
    uint8_t *w = static_cast<uint8_t *>(ws);
    double *dsum = reinterpret_cast<double *>(w);
    w += align_up((size_t)2 * O * 8, 256);
    float *dyacc = reinterpret_cast<float *>(w);
    w += align_up(pts * O * 4, 256);
    float *T = reinterpret_cast<float *>(w);
    w += align_up(pts * O * 4, 256);
    int *deg = reinterpret_cast<int *>(w);
    w += align_up(pts * 4, 256);
    float *dy = reinterpret_cast<float *>(w);
    const size_t zero_bytes = bn_train ? (size_t)(reinterpret_cast<uint8_t *>(dy) - static_cast<uint8_t *>(ws))
                                       : (size_t)(reinterpret_cast<uint8_t *>(T) - static_cast<uint8_t *>(ws));
    MLSP_CUDA(cudaMemsetAsync(ws, 0, zero_bytes, s));
    dim3 pgrid((N + EC_PTS - 1) / EC_PTS, (O + 31) / 32, B);
    if (dgamma_dbeta)
        edgeconv_bwd_prepare_kernel<true><<<pgrid, 256, 0, s>>>(g, (long long)g_bstride, hsel, coef, N, O, slope, dy, dsum);
    else
        edgeconv_bwd_prepare_kernel<false><<<pgrid, 256, 0, s>>>(g, (long long)g_bstride, hsel, coef, N, O, slope, dy, dsum);
    MLSP_LAUNCH_CHECK("edgeconv_bwd_prepare_kernel");
    const int G = O / 4;
    const long long items = (long long)pts * G;
    const int grid = (int)((items + EC_THREADS - 1) / EC_THREADS);
    const float4 *dy4 = reinterpret_cast<const float4 *>(dy);
    const float4 *yz4 = reinterpret_cast<const float4 *>(yz);
    if (bn_train) {
        edgeconv_bwd_scatter_kernel<true><<<grid, EC_THREADS, 0, s>>>(dy4, reinterpret_cast<const uchar4 *>(slot), idx, yz4, N, G, k,
                                                                      items, dyacc, reinterpret_cast<float4 *>(T), deg);
        MLSP_LAUNCH_CHECK("edgeconv_bwd_scatter_kernel");
        edgeconv_bwd_finish_kernel<true><<<grid, EC_THREADS, 0, s>>>(dy4, reinterpret_cast<const float4 *>(dyacc),
                                                                     reinterpret_cast<const float4 *>(T), deg,
                                                                     reinterpret_cast<const float4 *>(rowsum), yz4, coef, dsum, G, k,
                                                                     items, (double)pts * k, reinterpret_cast<float4 *>(dyz));
        MLSP_LAUNCH_CHECK("edgeconv_bwd_finish_kernel");
    } else {
        edgeconv_bwd_scatter_kernel<false><<<grid, EC_THREADS, 0, s>>>(dy4, reinterpret_cast<const uchar4 *>(slot), idx, yz4, N, G, k,
                                                                       items, dyacc, nullptr, nullptr);
        MLSP_LAUNCH_CHECK("edgeconv_bwd_scatter_kernel");
        edgeconv_bwd_finish_kernel<false><<<grid, EC_THREADS, 0, s>>>(dy4, reinterpret_cast<const float4 *>(dyacc), nullptr, nullptr,
                                                                      nullptr, yz4, coef, dsum, G, k, items, (double)pts * k,
                                                                      reinterpret_cast<float4 *>(dyz));
        MLSP_LAUNCH_CHECK("edgeconv_bwd_finish_kernel");
    }
    if (dgamma_dbeta) {
        edgeconv_param_grads_kernel<<<(O + 127) / 128, 128, 0, s>>>(dsum, O, dgamma_dbeta);
        MLSP_LAUNCH_CHECK("edgeconv_param_grads_kernel");
    }
    return MLSP_OK;
}
int mlsp_edgeconv_weight_prep(const float *W, const float *scale, const float *bias, int O, int C, float *Wcat, float *sgn,
                              float *zb, void *stream)
{
    MLSP_REQUIRE(W && Wcat && sgn, MLSP_EINVAL, "edgeconv_weight_prep: null pointer");
    MLSP_REQUIRE(O > 0 && C > 0 && (long long)O * C < (1ll << 30), MLSP_EINVAL, "edgeconv_weight_prep: bad shape O=%d C=%d", O, C);
    MLSP_REQUIRE(zb || !bias, MLSP_EINVAL, "edgeconv_weight_prep: bias without zb");
    edgeconv_weight_prep_kernel<<<(O * C + 255) / 256, 256, 0, as_stream(stream)>>>(W, scale, bias, O, C, Wcat, sgn, zb);
    MLSP_LAUNCH_CHECK("edgeconv_weight_prep_kernel");
    return MLSP_OK;
}

int mlsp_edgeconv_weight_grad(const float *part, int Z, const float *sgn, int O, int C, float *gW, void *stream)
{
    MLSP_REQUIRE(part && gW, MLSP_EINVAL, "edgeconv_weight_grad: null pointer");
    MLSP_REQUIRE(Z > 0 && O > 0 && C > 0 && (long long)O * C < (1ll << 30), MLSP_EINVAL, "edgeconv_weight_grad: bad shape");
    edgeconv_weight_grad_kernel<<<(O * C + 255) / 256, 256, 0, as_stream(stream)>>>(part, Z, sgn, O, C, gW);
    MLSP_LAUNCH_CHECK("edgeconv_weight_grad_kernel");
    return MLSP_OK;
}
}
