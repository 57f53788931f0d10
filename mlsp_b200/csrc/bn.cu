// bn.cu -- training-mode BatchNorm fused with the activation that follows it, forward and backward, for the point-wise layers
// of the DGCNN callers (conv_2d / fc_layer of PointDA/model_utils.py:45-89: Conv -> BatchNorm -> LeakyReLU(0.2); the heads of
// PointDA/Models.py:165-285 and PointSegDA/Models.py:245-392: Conv1d -> BatchNorm1d -> ReLU).
//
// torch runs this as cuDNN BatchNorm (3 passes forward, 5 backward over the activation) plus an elementwise activation kernel
// (2 passes forward, 3 backward): 13 passes over tensors of up to 335 MB (the transform net's edge features).  Here:
//   forward  : one statistics pass (per-channel sum and sum of squares, fp64 accumulation) + one apply pass
//              y = act(x * scale + shift)                                                              -> 3 passes
//   backward : one reduction pass (sum dz, sum dz * xhat with dz = dy * act'(z) recomputed from x) + one apply pass
//              dx = scale * (dz - mean(dz) - xhat * mean(dz * xhat))                                   -> 5 passes
// Both HBM-bound streaming kernels; 128-bit accesses; two memory layouts:
//   ROWS : x[r * C + c], r < R            (channels-last 4-D tensors (B,C,N,k), 2-D (B,C) inputs of the fc layers)
//   NCL  : x[(b * C + c) * L + l]         (contiguous (B,C,N) maps of the Conv1d heads)
// Semantics = torch.nn.functional.batch_norm(training=True) + leaky_relu(slope) (slope 0 = ReLU, slope 1 = no activation):
// biased variance for the normalisation, running_var updated with the unbiased one, momentum as in torch.
#include "common.cuh"

namespace mlsp {

constexpr int BN_THREADS = 256;

__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MLSP_FULL, v, o);
    return v;
}

__device__ __forceinline__ float act_fwd(float z, float slope) { return z > 0.0f ? z : z * slope; }
__device__ __forceinline__ float act_grad(float z, float slope) { return z > 0.0f ? 1.0f : slope; }

// ---------------------------------------------------------------------------------------------------------- ROWS layout
// thread = (row residue, float4 channel group); rows_per_iter = BN_THREADS / C4 rows in flight per CTA
// MODE 0: acc[c] += x, acc[C + c] += x^2          MODE 1: acc[c] += dz, acc[C + c] += dz * xhat
template <int MODE>
__global__ void __launch_bounds__(BN_THREADS)
bn_rows_reduce_kernel(const float4 *__restrict__ x, const float4 *__restrict__ dy, long long R, int C4, const float *__restrict__ mean,
                      const float *__restrict__ invstd, const float *__restrict__ gamma, const float *__restrict__ beta, float slope,
                      double *__restrict__ acc)
{
    __shared__ double red[BN_THREADS][8];
    const int rpi = BN_THREADS / C4;
    const int q = threadIdx.x % C4, rr = threadIdx.x / C4;
    double s[4] = {0, 0, 0, 0}, t[4] = {0, 0, 0, 0};
    if (rr < rpi) {
        float m[4] = {0, 0, 0, 0}, is[4] = {1, 1, 1, 1}, g[4] = {1, 1, 1, 1}, b[4] = {0, 0, 0, 0};
        if (MODE == 1) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                m[e] = mean[4 * q + e];
                is[e] = invstd[4 * q + e];
                g[e] = gamma ? gamma[4 * q + e] : 1.0f;
                b[e] = beta ? beta[4 * q + e] : 0.0f;
            }
        }
        for (long long r = (long long)blockIdx.x * rpi + rr; r < R; r += (long long)gridDim.x * rpi) {
            const float4 v = __ldg(x + r * C4 + q);
            const float xv[4] = {v.x, v.y, v.z, v.w};
            if (MODE == 0) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    s[e] += (double)xv[e];
                    t[e] += (double)xv[e] * (double)xv[e];
                }
            } else {
                const float4 d = __ldg(dy + r * C4 + q);
                const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float xh = (xv[e] - m[e]) * is[e];
                    const float dz = dv[e] * act_grad(xh * g[e] + b[e], slope);
                    s[e] += (double)dz;
                    t[e] += (double)dz * (double)xh;
                }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        red[threadIdx.x][e] = s[e];
        red[threadIdx.x][4 + e] = t[e];
    }
    __syncthreads();
    if (rr == 0) {
        for (int o = 1; o < rpi; ++o)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                s[e] += red[o * C4 + q][e];
                t[e] += red[o * C4 + q][4 + e];
            }
        const int C = 4 * C4;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            atomicAdd(acc + 4 * q + e, s[e]);
            atomicAdd(acc + C + 4 * q + e, t[e]);
        }
    }
}

// per-channel coefficients from the accumulated sums; the CTA with blockIdx 0 also publishes the saved statistics and
// updates the running ones.  sc / sh: shared arrays of C floats.
__device__ __forceinline__ void bn_coeffs(const double *__restrict__ acc, int C, double M, const float *__restrict__ gamma,
                                          const float *__restrict__ beta, float eps, float momentum, float *running_mean,
                                          float *running_var, float *save_mean, float *save_invstd, bool publish, float *sc,
                                          float *sh)
{
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double mean = acc[c] / M;
        double var = acc[C + c] / M - mean * mean;
        var = var > 0.0 ? var : 0.0;
        const float is = (float)(1.0 / sqrt(var + (double)eps));
        const float g = gamma ? gamma[c] : 1.0f, b = beta ? beta[c] : 0.0f;
        sc[c] = g * is;
        sh[c] = b - (float)mean * g * is;
        if (publish) {
            save_mean[c] = (float)mean;
            save_invstd[c] = is;
            if (running_mean) running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
            if (running_var) running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)(var * (M > 1.0 ? M / (M - 1.0) : 1.0));
        }
    }
}

__global__ void __launch_bounds__(BN_THREADS)
bn_rows_apply_kernel(const float4 *__restrict__ x, float4 *__restrict__ y, long long R, int C4, const double *__restrict__ acc,
                     const float *__restrict__ gamma, const float *__restrict__ beta, float eps, float momentum, float slope,
                     float *running_mean, float *running_var, float *save_mean, float *save_invstd)
{
    extern __shared__ float coef[];                       // scale [C] | shift [C]
    const int C = 4 * C4;
    float *sc = coef, *sh = coef + C;
    bn_coeffs(acc, C, (double)R, gamma, beta, eps, momentum, running_mean, running_var, save_mean, save_invstd, blockIdx.x == 0, sc, sh);
    __syncthreads();
    const long long total = R * C4;
    for (long long i = (long long)blockIdx.x * BN_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * BN_THREADS) {
        const int q = (int)(i % C4);
        const float4 v = __ldcs(x + i);
        const float4 a = *reinterpret_cast<const float4 *>(sc + 4 * q), b = *reinterpret_cast<const float4 *>(sh + 4 * q);
        float4 o;
        o.x = act_fwd(fmaf(v.x, a.x, b.x), slope);
        o.y = act_fwd(fmaf(v.y, a.y, b.y), slope);
        o.z = act_fwd(fmaf(v.z, a.z, b.z), slope);
        o.w = act_fwd(fmaf(v.w, a.w, b.w), slope);
        y[i] = o;
    }
}

// dx = gamma * invstd * (dz - sum_dz / M - xhat * sum_dz_xhat / M); blockIdx 0 writes dgamma = sum dz xhat, dbeta = sum dz
__global__ void __launch_bounds__(BN_THREADS)
bn_rows_bwd_apply_kernel(const float4 *__restrict__ x, const float4 *__restrict__ dy, float4 *__restrict__ dx, long long R, int C4,
                         const double *__restrict__ acc, const float *__restrict__ mean, const float *__restrict__ invstd,
                         const float *__restrict__ gamma, const float *__restrict__ beta, float slope, float *dgamma, float *dbeta)
{
    extern __shared__ float coef[];                       // mean | invstd | gamma | beta | k1 = sum_dz / M | k2 = sum_dz_xhat / M
    const int C = 4 * C4;
    float *cm = coef, *ci = coef + C, *cg = coef + 2 * C, *cb = coef + 3 * C, *k1 = coef + 4 * C, *k2 = coef + 5 * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        cm[c] = mean[c];
        ci[c] = invstd[c];
        cg[c] = gamma ? gamma[c] : 1.0f;
        cb[c] = beta ? beta[c] : 0.0f;
        k1[c] = (float)(acc[c] / (double)R);
        k2[c] = (float)(acc[C + c] / (double)R);
        if (blockIdx.x == 0) {
            if (dgamma) dgamma[c] = (float)acc[C + c];
            if (dbeta) dbeta[c] = (float)acc[c];
        }
    }
    __syncthreads();
    const long long total = R * C4;
    for (long long i = (long long)blockIdx.x * BN_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * BN_THREADS) {
        const int q = (int)(i % C4);
        const float4 v = __ldcs(x + i), d = __ldcs(dy + i);
        const float xv[4] = {v.x, v.y, v.z, v.w}, dv[4] = {d.x, d.y, d.z, d.w};
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = 4 * q + e;
            const float xh = (xv[e] - cm[c]) * ci[c];
            const float dz = dv[e] * act_grad(xh * cg[c] + cb[c], slope);
            o[e] = cg[c] * ci[c] * (dz - k1[c] - xh * k2[c]);
        }
        dx[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// ----------------------------------------------------------------------------------------------------------- NCL layout
// one warp per (b, c) row of L contiguous values
template <int MODE>
__global__ void __launch_bounds__(BN_THREADS)
bn_ncl_reduce_kernel(const float *__restrict__ x, const float *__restrict__ dy, int B, int C, int L, long long xbs, long long obs,
                     const float *__restrict__ mean,
                     const float *__restrict__ invstd, const float *__restrict__ gamma, const float *__restrict__ beta, float slope,
                     double *__restrict__ acc)
{
    const int lane = threadIdx.x & 31;
    const long long rows = (long long)B * C;
    for (long long row = (long long)blockIdx.x * (BN_THREADS / 32) + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * (BN_THREADS / 32)) {
        const int c = (int)(row % C);
        const long long bb = row / C;
        const float *xr = x + bb * xbs + (long long)c * L;
        const float *dr = MODE == 1 ? dy + bb * obs + (long long)c * L : nullptr;
        float m = 0, is = 1, g = 1, b = 0;
        if (MODE == 1) {
            m = mean[c];
            is = invstd[c];
            g = gamma ? gamma[c] : 1.0f;
            b = beta ? beta[c] : 0.0f;
        }
        double s = 0, t = 0;
        const bool vec = (L & 3) == 0 && ((reinterpret_cast<uintptr_t>(xr) & 15) == 0) && (MODE == 0 || (reinterpret_cast<uintptr_t>(dr) & 15) == 0);
        if (vec) {
            for (int l = lane * 4; l < L; l += 128) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(xr + l));
                const float xv[4] = {v.x, v.y, v.z, v.w};
                float dv[4] = {0, 0, 0, 0};
                if (MODE == 1) {
                    const float4 d = __ldg(reinterpret_cast<const float4 *>(dr + l));
                    dv[0] = d.x; dv[1] = d.y; dv[2] = d.z; dv[3] = d.w;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (MODE == 0) {
                        s += (double)xv[e];
                        t += (double)xv[e] * (double)xv[e];
                    } else {
                        const float xh = (xv[e] - m) * is;
                        const float dz = dv[e] * act_grad(xh * g + b, slope);
                        s += (double)dz;
                        t += (double)dz * (double)xh;
                    }
                }
            }
        } else {
            for (int l = lane; l < L; l += 32) {
                const float xv = xr[l];
                if (MODE == 0) {
                    s += (double)xv;
                    t += (double)xv * (double)xv;
                } else {
                    const float xh = (xv - m) * is;
                    const float dz = dr[l] * act_grad(xh * g + b, slope);
                    s += (double)dz;
                    t += (double)dz * (double)xh;
                }
            }
        }
        s = warp_sum_d(s);
        t = warp_sum_d(t);
        if (lane == 0) {
            atomicAdd(acc + c, s);
            atomicAdd(acc + C + c, t);
        }
    }
}

// one CTA per (b, c) row chunk; MODE 0 = forward apply, MODE 1 = backward apply
template <int MODE>
__global__ void __launch_bounds__(BN_THREADS)
bn_ncl_apply_kernel(const float *__restrict__ x, const float *__restrict__ dy, float *__restrict__ out, int B, int C, int L,
                    long long xbs, long long obs, const double *__restrict__ acc, const float *__restrict__ gamma, const float *__restrict__ beta, float eps,
                    float momentum, float slope, float *running_mean, float *running_var, float *save_mean, float *save_invstd,
                    float *dgamma, float *dbeta)
{
    const long long rows = (long long)B * C;
    const double M = (double)B * (double)L;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int c = (int)(row % C);
        const float g = gamma ? gamma[c] : 1.0f, b = beta ? beta[c] : 0.0f;
        float mean, is, k1 = 0, k2 = 0;
        if (MODE == 0) {
            const double mu = acc[c] / M;
            double var = acc[C + c] / M - mu * mu;
            var = var > 0.0 ? var : 0.0;
            mean = (float)mu;
            is = (float)(1.0 / sqrt(var + (double)eps));
            if (row < C && threadIdx.x == 0) {               // rows 0..C-1 are cloud 0: each channel once
                save_mean[c] = mean;
                save_invstd[c] = is;
                if (running_mean) running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mean;
                if (running_var) running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)(var * (M > 1.0 ? M / (M - 1.0) : 1.0));
            }
        } else {
            mean = save_mean[c];
            is = save_invstd[c];
            k1 = (float)(acc[c] / M);
            k2 = (float)(acc[C + c] / M);
            if (row < C && threadIdx.x == 0) {
                if (dgamma) dgamma[c] = (float)acc[C + c];
                if (dbeta) dbeta[c] = (float)acc[c];
            }
        }
        const float sc = g * is, sh = b - mean * g * is;
        const long long bb = row / C;
        const float *xr = x + bb * xbs + (long long)c * L;
        const float *dr = MODE == 1 ? dy + bb * obs + (long long)c * L : nullptr;
        float *orow = out + bb * obs + (long long)c * L;
        const bool vec = (L & 3) == 0 && (((reinterpret_cast<uintptr_t>(xr) | reinterpret_cast<uintptr_t>(orow)) & 15) == 0) &&
                         (MODE == 0 || (reinterpret_cast<uintptr_t>(dr) & 15) == 0);
        if (vec) {
            for (int l = threadIdx.x * 4; l < L; l += BN_THREADS * 4) {
                const float4 v = __ldcs(reinterpret_cast<const float4 *>(xr + l));
                const float xv[4] = {v.x, v.y, v.z, v.w};
                float o[4];
                if (MODE == 0) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[e] = act_fwd(fmaf(xv[e], sc, sh), slope);
                } else {
                    const float4 d = __ldcs(reinterpret_cast<const float4 *>(dr + l));
                    const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float xh = (xv[e] - mean) * is;
                        const float dz = dv[e] * act_grad(xh * g + b, slope);
                        o[e] = sc * (dz - k1 - xh * k2);
                    }
                }
                *reinterpret_cast<float4 *>(orow + l) = make_float4(o[0], o[1], o[2], o[3]);
            }
        } else {
            for (int l = threadIdx.x; l < L; l += BN_THREADS) {
                const float xv = xr[l];
                if (MODE == 0) {
                    orow[l] = act_fwd(fmaf(xv, sc, sh), slope);
                } else {
                    const float xh = (xv - mean) * is;
                    const float dz = dr[l] * act_grad(xh * g + b, slope);
                    orow[l] = sc * (dz - k1 - xh * k2);
                }
            }
        }
    }
}

static int bn_check(const void *x, const void *y, long long R, int C, int L, int layout, const void *acc, const char *who)
{
    MLSP_REQUIRE(x && y && acc, MLSP_EINVAL, "%s: null pointer", who);
    MLSP_REQUIRE(R > 0 && C > 0 && L > 0, MLSP_EINVAL, "%s: bad shape", who);
    MLSP_REQUIRE(layout == 0 || layout == 1, MLSP_EINVAL, "%s: layout %d", who, layout);
    if (layout == 0) {
        MLSP_REQUIRE(C % 4 == 0 && C <= 4 * BN_THREADS, MLSP_EUNSUPPORTED, "%s: rows layout needs C %% 4 == 0 and C <= %d (C=%d)", who, 4 * BN_THREADS, C);
        MLSP_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, MLSP_EINVAL, "%s: 16-byte alignment", who);
    }
    MLSP_REQUIRE((reinterpret_cast<uintptr_t>(acc) & 7) == 0, MLSP_EINVAL, "%s: acc must be 8-byte aligned", who);
    return MLSP_OK;
}

}  // namespace mlsp

// y = act(batch_norm(x)) in training mode.  layout 0 (ROWS): x (R, C) row-major, L ignored (pass 1); layout 1 (NCL): x (R = B, C, L)
// with batch strides x_batch_stride for x and y_batch_stride for y / dy / dx (in floats; 0 = C * L: a channel slice of a wider
// map -- the heads' merged first layer -- is normalised in place of its view).
// gamma / beta / running_mean / running_var may be NULL; save_mean / save_invstd (C) are written for the backward;
// acc = 2 C doubles of scratch.
extern "C" int mlsp_bn_act_fwd(const float *x, float *y, long long R, int C, int L, int layout, long long x_batch_stride,
                               long long y_batch_stride, const float *gamma, const float *beta,
                               float *running_mean, float *running_var, float momentum, float eps, float slope, float *save_mean,
                               float *save_invstd, double *acc, void *stream)
{
    using namespace mlsp;
    int rc = bn_check(x, y, R, C, L, layout, acc, "bn_act_fwd");
    if (rc) return rc;
    MLSP_REQUIRE(save_mean && save_invstd, MLSP_EINVAL, "bn_act_fwd: null pointer");
    cudaStream_t st = as_stream(stream);
    MLSP_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * (size_t)C, st));
    const int sms = sm_count();
    if (layout == 0) {
        const int C4 = C / 4, rpi = BN_THREADS / C4;
        const long long want = (R + rpi - 1) / rpi;
        const int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
        bn_rows_reduce_kernel<0><<<grid, BN_THREADS, 0, st>>>(reinterpret_cast<const float4 *>(x), nullptr, R, C4, nullptr, nullptr, nullptr,
                                                            nullptr, slope, acc);
        MLSP_LAUNCH_CHECK("bn_rows_reduce_kernel");
        const long long tot = R * C4, want2 = (tot + BN_THREADS - 1) / BN_THREADS;
        const int grid2 = (int)(want2 < (long long)sms * 8 ? want2 : (long long)sms * 8);
        bn_rows_apply_kernel<<<grid2, BN_THREADS, sizeof(float) * 2 * C, st>>>(reinterpret_cast<const float4 *>(x), reinterpret_cast<float4 *>(y), R, C4,
                                                                             acc, gamma, beta, eps, momentum, slope, running_mean, running_var,
                                                                             save_mean, save_invstd);
        MLSP_LAUNCH_CHECK("bn_rows_apply_kernel");
    } else {
        MLSP_REQUIRE(R <= 0x7fffffff, MLSP_EUNSUPPORTED, "bn_act_fwd: B too large");
        const long long rows = R * C, want = (rows + BN_THREADS / 32 - 1) / (BN_THREADS / 32);
        const int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
        const long long xbs = x_batch_stride ? x_batch_stride : (long long)C * L, obs = y_batch_stride ? y_batch_stride : (long long)C * L;
        MLSP_REQUIRE(xbs >= (long long)C * L && obs >= (long long)C * L, MLSP_EINVAL, "bn_act_fwd: batch stride smaller than C * L");
        bn_ncl_reduce_kernel<0><<<grid, BN_THREADS, 0, st>>>(x, nullptr, (int)R, C, L, xbs, obs, nullptr, nullptr, nullptr, nullptr, slope, acc);
        MLSP_LAUNCH_CHECK("bn_ncl_reduce_kernel");
        const int grid2 = (int)(rows < (long long)sms * 16 ? rows : (long long)sms * 16);
        bn_ncl_apply_kernel<0><<<grid2, BN_THREADS, 0, st>>>(x, nullptr, y, (int)R, C, L, xbs, obs, acc, gamma, beta, eps, momentum, slope, running_mean,
                                                           running_var, save_mean, save_invstd, nullptr, nullptr);
        MLSP_LAUNCH_CHECK("bn_ncl_apply_kernel");
    }
    return MLSP_OK;
}

// dx, dgamma, dbeta of the above from x (the layer's input), dy and the saved statistics.  dgamma / dbeta may be NULL.
extern "C" int mlsp_bn_act_bwd(const float *x, const float *dy, float *dx, long long R, int C, int L, int layout,
                               long long x_batch_stride, long long y_batch_stride, const float *gamma,
                               const float *beta, const float *save_mean, const float *save_invstd, float slope, float *dgamma,
                               float *dbeta, double *acc, void *stream)
{
    using namespace mlsp;
    int rc = bn_check(x, dx, R, C, L, layout, acc, "bn_act_bwd");
    if (rc) return rc;
    MLSP_REQUIRE(dy && save_mean && save_invstd, MLSP_EINVAL, "bn_act_bwd: null pointer");
    MLSP_REQUIRE(layout == 1 || (reinterpret_cast<uintptr_t>(dy) & 15) == 0, MLSP_EINVAL, "bn_act_bwd: 16-byte alignment");
    cudaStream_t st = as_stream(stream);
    MLSP_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * (size_t)C, st));
    const int sms = sm_count();
    if (layout == 0) {
        const int C4 = C / 4, rpi = BN_THREADS / C4;
        const long long want = (R + rpi - 1) / rpi;
        const int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
        bn_rows_reduce_kernel<1><<<grid, BN_THREADS, 0, st>>>(reinterpret_cast<const float4 *>(x), reinterpret_cast<const float4 *>(dy), R, C4,
                                                            save_mean, save_invstd, gamma, beta, slope, acc);
        MLSP_LAUNCH_CHECK("bn_rows_reduce_kernel");
        const long long tot = R * C4, want2 = (tot + BN_THREADS - 1) / BN_THREADS;
        const int grid2 = (int)(want2 < (long long)sms * 8 ? want2 : (long long)sms * 8);
        bn_rows_bwd_apply_kernel<<<grid2, BN_THREADS, sizeof(float) * 6 * C, st>>>(
            reinterpret_cast<const float4 *>(x), reinterpret_cast<const float4 *>(dy), reinterpret_cast<float4 *>(dx), R, C4, acc, save_mean,
            save_invstd, gamma, beta, slope, dgamma, dbeta);
        MLSP_LAUNCH_CHECK("bn_rows_bwd_apply_kernel");
    } else {
        MLSP_REQUIRE(R <= 0x7fffffff, MLSP_EUNSUPPORTED, "bn_act_bwd: B too large");
        const long long rows = R * C, want = (rows + BN_THREADS / 32 - 1) / (BN_THREADS / 32);
        const int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
        const long long xbs = x_batch_stride ? x_batch_stride : (long long)C * L, obs = y_batch_stride ? y_batch_stride : (long long)C * L;
        MLSP_REQUIRE(xbs >= (long long)C * L && obs >= (long long)C * L, MLSP_EINVAL, "bn_act_bwd: batch stride smaller than C * L");
        bn_ncl_reduce_kernel<1><<<grid, BN_THREADS, 0, st>>>(x, dy, (int)R, C, L, xbs, obs, save_mean, save_invstd, gamma, beta, slope, acc);
        MLSP_LAUNCH_CHECK("bn_ncl_reduce_kernel");
        const int grid2 = (int)(rows < (long long)sms * 16 ? rows : (long long)sms * 16);
        bn_ncl_apply_kernel<1><<<grid2, BN_THREADS, 0, st>>>(x, dy, dx, (int)R, C, L, xbs, obs, acc, gamma, beta, 0.0f, 0.0f, slope, nullptr, nullptr,
                                                           const_cast<float *>(save_mean), const_cast<float *>(save_invstd), dgamma, dbeta);
        MLSP_LAUNCH_CHECK("bn_ncl_apply_kernel");
    }
    return MLSP_OK;
}
