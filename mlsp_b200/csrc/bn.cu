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

// Statistics are accumulated about a per-channel PIVOT p (the channel's first element): d = x - p, S = sum d, T = sum d^2 in fp32
// over chunks of BN_UNROLL elements, the chunks in fp64; mean = p + S/M, var = T/M - (S/M)^2.  With the pivot inside the data the
// subtraction of the two moments loses nothing (a channel with mean 100 and spread 0.5 is the test case), and the fp64 work is
// one conversion per chunk instead of per element.
constexpr int BN_UNROLL = 4;

struct BnCoef {        // per-channel coefficient block (4 C floats) in the scratch buffer, after the partial sums
    float *a, *b, *c, *d;
};
__host__ __device__ __forceinline__ BnCoef bn_coef(float *f, int C) { return BnCoef{f, f + C, f + 2 * C, f + 3 * C}; }

// No atomics and no zeroing: every CTA of a reduction writes its partial sums to part[2 C][cta] (fp64) and the finalize kernel
// adds the G partials of a channel with one warp (deterministic; with atomicAdd(double) onto 2 C addresses from 1184 CTAs the
// reduction pass ran at 44 % of the HBM peak, ncu).
constexpr int BN_MAX_PARTS = 592;                         // 4 CTAs per SM on 148 SMs

// ---------------------------------------------------------------------------------------------------------- ROWS layout
// thread = (row residue rr, float4 channel group q); rpi = BN_THREADS / C4 consecutive rows (one contiguous 4 KB span) per CTA and
// iteration, BN_UNROLL iterations in flight.
// MODE 0: acc[c] += x - p, acc[C + c] += (x - p)^2          MODE 1: acc[c] += dz, acc[C + c] += dz * xhat
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
        float m[4], is[4] = {1, 1, 1, 1}, g[4] = {1, 1, 1, 1}, b[4] = {0, 0, 0, 0};
        if (MODE == 0) {
            const float4 p = __ldg(x + q);                  // pivot: row 0
            m[0] = p.x; m[1] = p.y; m[2] = p.z; m[3] = p.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                m[e] = mean[4 * q + e];
                is[e] = invstd[4 * q + e];
                g[e] = gamma ? gamma[4 * q + e] : 1.0f;
                b[e] = beta ? beta[4 * q + e] : 0.0f;
            }
        }
        const long long step = (long long)gridDim.x * rpi;
        for (long long r0 = (long long)blockIdx.x * rpi + rr; r0 < R; r0 += step * BN_UNROLL) {
            float4 v[BN_UNROLL], d[BN_UNROLL];
#pragma unroll
            for (int u = 0; u < BN_UNROLL; ++u) {
                const long long r = r0 + u * step;
                const bool ok = r < R;
                v[u] = ok ? __ldg(x + r * C4 + q) : make_float4(m[0], m[1], m[2], m[3]);      // contributes 0 in both modes
                if (MODE == 1) d[u] = ok ? __ldg(dy + r * C4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float fs[4] = {0, 0, 0, 0}, ft[4] = {0, 0, 0, 0};
#pragma unroll
            for (int u = 0; u < BN_UNROLL; ++u) {
                const float xv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                const float dv[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (MODE == 0) {
                        const float dd = xv[e] - m[e];
                        fs[e] += dd;
                        ft[e] = fmaf(dd, dd, ft[e]);
                    } else {
                        const float xh = (xv[e] - m[e]) * is[e];
                        const float dz = dv[e] * act_grad(fmaf(xh, g[e], b[e]), slope);
                        fs[e] += dz;
                        ft[e] = fmaf(dz, xh, ft[e]);
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                s[e] += (double)fs[e];
                t[e] += (double)ft[e];
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
        const int C = 4 * C4, G = gridDim.x;                 // part[stat][channel][cta]: the finalize warp reads a channel's row coalesced
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            acc[(size_t)(4 * q + e) * G + blockIdx.x] = s[e];
            acc[(size_t)(C + 4 * q + e) * G + blockIdx.x] = t[e];
        }
    }
}

// Between the reduction and the apply pass: one warp per channel adds the G partial sums and writes the channel's coefficients.
//   forward : mean, invstd -> save_*, running statistics, a = scale, b = shift
//   backward: a = mean(dz), b = mean(dz xhat), c = gamma * invstd; dgamma, dbeta
__device__ __forceinline__ void bn_sum_parts(const double *__restrict__ part, int G, int C, int c, double &s, double &t)
{
    s = 0.0;
    t = 0.0;
    const double *ps = part + (size_t)c * G, *pt = part + (size_t)(C + c) * G;
    for (int g = threadIdx.x & 31; g < G; g += 32) {
        s += ps[g];
        t += pt[g];
    }
    s = warp_sum_d(s);
    t = warp_sum_d(t);
}

__global__ void __launch_bounds__(BN_THREADS)
bn_finalize_fwd_kernel(const double *__restrict__ part, int G, float *coef, int C, double M, const float *__restrict__ pivot,
                       long long pivot_stride, const float *__restrict__ gamma, const float *__restrict__ beta, float eps, float momentum,
                       float *running_mean, float *running_var, float *save_mean, float *save_invstd)
{
    const BnCoef k = bn_coef(coef, C);
    const int c = blockIdx.x * (BN_THREADS / 32) + (threadIdx.x >> 5);
    if (c >= C) return;
    double S, T;
    bn_sum_parts(part, G, C, c, S, T);
    if ((threadIdx.x & 31) != 0) return;
    const double md = S / M;
    const double mean = (double)pivot[c * pivot_stride] + md;
    double var = T / M - md * md;
    var = var > 0.0 ? var : 0.0;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[c] : 1.0f, b = beta ? beta[c] : 0.0f;
    k.a[c] = g * is;
    k.b[c] = (float)((double)b - mean * (double)(g * is));
    save_mean[c] = (float)mean;
    save_invstd[c] = is;
    if (running_mean) running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
    if (running_var) running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)(var * (M > 1.0 ? M / (M - 1.0) : 1.0));
}

__global__ void __launch_bounds__(BN_THREADS)
bn_finalize_bwd_kernel(const double *__restrict__ part, int G, float *coef, int C, double M, const float *__restrict__ gamma,
                       const float *__restrict__ invstd, float *dgamma, float *dbeta)
{
    const BnCoef k = bn_coef(coef, C);
    const int c = blockIdx.x * (BN_THREADS / 32) + (threadIdx.x >> 5);
    if (c >= C) return;
    double S, T;
    bn_sum_parts(part, G, C, c, S, T);
    if ((threadIdx.x & 31) != 0) return;
    k.a[c] = (float)(S / M);                           // mean(dz)
    k.b[c] = (float)(T / M);                           // mean(dz * xhat)
    k.c[c] = (gamma ? gamma[c] : 1.0f) * invstd[c];    // gamma * invstd
    if (dgamma) dgamma[c] = (float)T;
    if (dbeta) dbeta[c] = (float)S;
}

// MODE 0: y = act(x * a + b)      MODE 1: dx = c * (dz - a - xhat * b), dz = dy * act'(xhat * gamma + beta)
template <int MODE>
__global__ void __launch_bounds__(BN_THREADS)
bn_rows_apply_kernel(const float4 *__restrict__ x, const float4 *__restrict__ dy, float4 *__restrict__ out, long long R, int C4,
                     float *__restrict__ coef, const float *__restrict__ mean, const float *__restrict__ invstd,
                     const float *__restrict__ gamma, const float *__restrict__ beta, float slope)
{
    const int C = 4 * C4;
    const BnCoef k = bn_coef(coef, C);
    const int rpi = BN_THREADS / C4;
    const int q = threadIdx.x % C4, rr = threadIdx.x / C4;
    if (rr >= rpi) return;
    float ka[4], kb[4], kc[4], m[4], is[4], g[4], b[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        ka[e] = k.a[4 * q + e];
        kb[e] = k.b[4 * q + e];
        if (MODE == 1) {
            kc[e] = k.c[4 * q + e];
            m[e] = mean[4 * q + e];
            is[e] = invstd[4 * q + e];
            g[e] = gamma ? gamma[4 * q + e] : 1.0f;
            b[e] = beta ? beta[4 * q + e] : 0.0f;
        }
    }
    const long long step = (long long)gridDim.x * rpi;
    for (long long r0 = (long long)blockIdx.x * rpi + rr; r0 < R; r0 += step * BN_UNROLL) {
        float4 v[BN_UNROLL], d[BN_UNROLL];
#pragma unroll
        for (int u = 0; u < BN_UNROLL; ++u) {
            const long long r = r0 + u * step;
            if (r < R) {
                v[u] = __ldcs(x + r * C4 + q);
                if (MODE == 1) d[u] = __ldcs(dy + r * C4 + q);
            }
        }
#pragma unroll
        for (int u = 0; u < BN_UNROLL; ++u) {
            const long long r = r0 + u * step;
            if (r >= R) break;
            const float xv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            float o[4];
            if (MODE == 0) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = act_fwd(fmaf(xv[e], ka[e], kb[e]), slope);
            } else {
                const float dv[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float xh = (xv[e] - m[e]) * is[e];
                    const float dz = dv[e] * act_grad(fmaf(xh, g[e], b[e]), slope);
                    o[e] = kc[e] * (dz - ka[e] - xh * kb[e]);
                }
            }
            out[r * C4 + q] = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ----------------------------------------------------------------------------------------------------------- NCL layout
// CTA = (channel c, batch residue): the CTA walks the rows (b, c), b = blockIdx.y, blockIdx.y + gridDim.y, ..., all threads on one
// row at a time (L = 1024: one float4 per thread), BN_UNROLL rows in flight; per-channel values are loaded once per CTA.
template <int MODE>
__global__ void __launch_bounds__(BN_THREADS)
bn_ncl_reduce_kernel(const float *__restrict__ x, const float *__restrict__ dy, int B, int C, int L, long long xbs, long long obs,
                     const float *__restrict__ mean, const float *__restrict__ invstd, const float *__restrict__ gamma,
                     const float *__restrict__ beta, float slope, double *__restrict__ acc)
{
    __shared__ double red[2][BN_THREADS / 32];
    const int c = blockIdx.x;
    float m, is = 1, g = 1, bt = 0;
    if (MODE == 0) {
        m = __ldg(x + (long long)c * L);                    // pivot: (b = 0, c, l = 0)
    } else {
        m = mean[c];
        is = invstd[c];
        g = gamma ? gamma[c] : 1.0f;
        bt = beta ? beta[c] : 0.0f;
    }
    const bool vec = (L & 3) == 0 && (xbs & 3) == 0 && (obs & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                     (MODE == 0 || (reinterpret_cast<uintptr_t>(dy) & 15) == 0);
    double s = 0, t = 0;
    if (vec) {
        for (int b0 = blockIdx.y; b0 < B; b0 += gridDim.y * BN_UNROLL) {
            for (int l = threadIdx.x * 4; l < L; l += BN_THREADS * 4) {
                float4 v[BN_UNROLL], d[BN_UNROLL];
#pragma unroll
                for (int u = 0; u < BN_UNROLL; ++u) {
                    const int b = b0 + u * gridDim.y;
                    const bool ok = b < B;
                    v[u] = ok ? __ldg(reinterpret_cast<const float4 *>(x + b * xbs + (long long)c * L + l)) : make_float4(m, m, m, m);
                    if (MODE == 1) d[u] = ok ? __ldg(reinterpret_cast<const float4 *>(dy + b * obs + (long long)c * L + l)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                float fs = 0, ft = 0;
#pragma unroll
                for (int u = 0; u < BN_UNROLL; ++u) {
                    const float xv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                    const float dv[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (MODE == 0) {
                            const float dd = xv[e] - m;
                            fs += dd;
                            ft = fmaf(dd, dd, ft);
                        } else {
                            const float xh = (xv[e] - m) * is;
                            const float dz = dv[e] * act_grad(fmaf(xh, g, bt), slope);
                            fs += dz;
                            ft = fmaf(dz, xh, ft);
                        }
                    }
                }
                s += (double)fs;
                t += (double)ft;
            }
        }
    } else {
        for (int b = blockIdx.y; b < B; b += gridDim.y)
            for (int l = threadIdx.x; l < L; l += BN_THREADS) {
                const float xv = x[b * xbs + (long long)c * L + l];
                if (MODE == 0) {
                    const float dd = xv - m;
                    s += (double)dd;
                    t += (double)dd * (double)dd;
                } else {
                    const float xh = (xv - m) * is;
                    const float dz = dy[b * obs + (long long)c * L + l] * act_grad(fmaf(xh, g, bt), slope);
                    s += (double)dz;
                    t += (double)dz * (double)xh;
                }
            }
    }
    s = warp_sum_d(s);
    t = warp_sum_d(t);
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = s;
        red[1][threadIdx.x >> 5] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < BN_THREADS / 32; ++w) {
            s += red[0][w];
            t += red[1][w];
        }
        const int G = gridDim.y;
        acc[(size_t)c * G + blockIdx.y] = s;
        acc[(size_t)(C + c) * G + blockIdx.y] = t;
    }
}

template <int MODE>
__global__ void __launch_bounds__(BN_THREADS)
bn_ncl_apply_kernel(const float *__restrict__ x, const float *__restrict__ dy, float *__restrict__ out, int B, int C, int L,
                    long long xbs, long long obs, float *__restrict__ coef, const float *__restrict__ mean,
                    const float *__restrict__ invstd, const float *__restrict__ gamma, const float *__restrict__ beta, float slope)
{
    const int c = blockIdx.x;
    const BnCoef k = bn_coef(coef, C);
    const float ka = k.a[c], kb = k.b[c];
    float kc = 0, m = 0, is = 1, g = 1, bt = 0;
    if (MODE == 1) {
        kc = k.c[c];
        m = mean[c];
        is = invstd[c];
        g = gamma ? gamma[c] : 1.0f;
        bt = beta ? beta[c] : 0.0f;
    }
    const bool vec = (L & 3) == 0 && (xbs & 3) == 0 && (obs & 3) == 0 &&
                     ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
                     (MODE == 0 || (reinterpret_cast<uintptr_t>(dy) & 15) == 0);
    if (vec) {
        for (int b0 = blockIdx.y; b0 < B; b0 += gridDim.y * BN_UNROLL) {
            for (int l = threadIdx.x * 4; l < L; l += BN_THREADS * 4) {
                float4 v[BN_UNROLL], d[BN_UNROLL];
#pragma unroll
                for (int u = 0; u < BN_UNROLL; ++u) {
                    const int b = b0 + u * gridDim.y;
                    if (b < B) {
                        v[u] = __ldcs(reinterpret_cast<const float4 *>(x + b * xbs + (long long)c * L + l));
                        if (MODE == 1) d[u] = __ldcs(reinterpret_cast<const float4 *>(dy + b * obs + (long long)c * L + l));
                    }
                }
#pragma unroll
                for (int u = 0; u < BN_UNROLL; ++u) {
                    const int b = b0 + u * gridDim.y;
                    if (b >= B) break;
                    const float xv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                    float o[4];
                    if (MODE == 0) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) o[e] = act_fwd(fmaf(xv[e], ka, kb), slope);
                    } else {
                        const float dv[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float xh = (xv[e] - m) * is;
                            const float dz = dv[e] * act_grad(fmaf(xh, g, bt), slope);
                            o[e] = kc * (dz - ka - xh * kb);
                        }
                    }
                    *reinterpret_cast<float4 *>(out + b * obs + (long long)c * L + l) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
    } else {
        for (int b = blockIdx.y; b < B; b += gridDim.y)
            for (int l = threadIdx.x; l < L; l += BN_THREADS) {
                const float xv = x[b * xbs + (long long)c * L + l];
                float o;
                if (MODE == 0) {
                    o = act_fwd(fmaf(xv, ka, kb), slope);
                } else {
                    const float xh = (xv - m) * is;
                    const float dz = dy[b * obs + (long long)c * L + l] * act_grad(fmaf(xh, g, bt), slope);
                    o = kc * (dz - ka - xh * kb);
                }
                out[b * obs + (long long)c * L + l] = o;
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
// acc = mlsp_bn_scratch_bytes(C) bytes of scratch (per-CTA partial sums + the per-channel coefficient block; no zeroing needed).
// grid of the ROWS kernels / batch residues of the NCL kernels = number of partial-sum rows in the scratch buffer
static inline int bn_rows_grid(long long R, int C4)
{
    const int rpi = mlsp::BN_THREADS / C4;
    const long long want = (R + (long long)rpi * mlsp::BN_UNROLL - 1) / ((long long)rpi * mlsp::BN_UNROLL);
    const long long cap = (long long)mlsp::sm_count() * 4 < mlsp::BN_MAX_PARTS ? (long long)mlsp::sm_count() * 4 : mlsp::BN_MAX_PARTS;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}
static inline int bn_ncl_gy(int B, int C)
{
    int gy = (mlsp::sm_count() * 8 + C - 1) / C;                                 // ~8 CTAs per SM in total
    const int gy_max = (B + mlsp::BN_UNROLL - 1) / mlsp::BN_UNROLL;
    gy = gy > gy_max ? gy_max : gy;
    gy = gy > mlsp::BN_MAX_PARTS ? mlsp::BN_MAX_PARTS : gy;
    return gy < 1 ? 1 : gy;
}

// bytes of scratch (`acc`) a call on C channels needs: BN_MAX_PARTS x 2 C doubles of partial sums + 4 C floats of coefficients
extern "C" size_t mlsp_bn_scratch_bytes(int C)
{
    return C > 0 ? sizeof(double) * 2 * (size_t)C * mlsp::BN_MAX_PARTS + sizeof(float) * 4 * (size_t)C : 0;
}

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
    const int fin_grid = (C + BN_THREADS / 32 - 1) / (BN_THREADS / 32);
    if (layout == 0) {
        const int C4 = C / 4;
        const int G = bn_rows_grid(R, C4);
        float *coef = reinterpret_cast<float *>(acc + (size_t)G * 2 * C);
        bn_rows_reduce_kernel<0><<<G, BN_THREADS, 0, st>>>(reinterpret_cast<const float4 *>(x), nullptr, R, C4, nullptr, nullptr, nullptr,
                                                         nullptr, slope, acc);
        MLSP_LAUNCH_CHECK("bn_rows_reduce_kernel");
        bn_finalize_fwd_kernel<<<fin_grid, BN_THREADS, 0, st>>>(acc, G, coef, C, (double)R, x, 1, gamma, beta, eps, momentum, running_mean,
                                                              running_var, save_mean, save_invstd);
        MLSP_LAUNCH_CHECK("bn_finalize_fwd_kernel");
        bn_rows_apply_kernel<0><<<2 * G, BN_THREADS, 0, st>>>(reinterpret_cast<const float4 *>(x), nullptr, reinterpret_cast<float4 *>(y), R, C4,
                                                            coef, nullptr, nullptr, nullptr, nullptr, slope);
        MLSP_LAUNCH_CHECK("bn_rows_apply_kernel");
    } else {
        MLSP_REQUIRE(R <= 0x7fffffff, MLSP_EUNSUPPORTED, "bn_act_fwd: B too large");
        const long long xbs = x_batch_stride ? x_batch_stride : (long long)C * L, obs = y_batch_stride ? y_batch_stride : (long long)C * L;
        MLSP_REQUIRE(xbs >= (long long)C * L && obs >= (long long)C * L, MLSP_EINVAL, "bn_act_fwd: batch stride smaller than C * L");
        const int B = (int)R, gy = bn_ncl_gy(B, C);
        float *coef = reinterpret_cast<float *>(acc + (size_t)gy * 2 * C);
        const dim3 grid(C, gy);
        bn_ncl_reduce_kernel<0><<<grid, BN_THREADS, 0, st>>>(x, nullptr, B, C, L, xbs, obs, nullptr, nullptr, nullptr, nullptr, slope, acc);
        MLSP_LAUNCH_CHECK("bn_ncl_reduce_kernel");
        bn_finalize_fwd_kernel<<<fin_grid, BN_THREADS, 0, st>>>(acc, gy, coef, C, (double)B * (double)L, x, L, gamma, beta, eps, momentum,
                                                              running_mean, running_var, save_mean, save_invstd);
        MLSP_LAUNCH_CHECK("bn_finalize_fwd_kernel");
        bn_ncl_apply_kernel<0><<<grid, BN_THREADS, 0, st>>>(x, nullptr, y, B, C, L, xbs, obs, coef, nullptr, nullptr, nullptr, nullptr, slope);
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
    const int fin_grid = (C + BN_THREADS / 32 - 1) / (BN_THREADS / 32);
    if (layout == 0) {
        const int C4 = C / 4;
        const int G = bn_rows_grid(R, C4);
        float *coef = reinterpret_cast<float *>(acc + (size_t)G * 2 * C);
        bn_rows_reduce_kernel<1><<<G, BN_THREADS, 0, st>>>(reinterpret_cast<const float4 *>(x), reinterpret_cast<const float4 *>(dy), R, C4,
                                                         save_mean, save_invstd, gamma, beta, slope, acc);
        MLSP_LAUNCH_CHECK("bn_rows_reduce_kernel");
        bn_finalize_bwd_kernel<<<fin_grid, BN_THREADS, 0, st>>>(acc, G, coef, C, (double)R, gamma, save_invstd, dgamma, dbeta);
        MLSP_LAUNCH_CHECK("bn_finalize_bwd_kernel");
        bn_rows_apply_kernel<1><<<2 * G, BN_THREADS, 0, st>>>(reinterpret_cast<const float4 *>(x), reinterpret_cast<const float4 *>(dy),
                                                            reinterpret_cast<float4 *>(dx), R, C4, coef, save_mean, save_invstd, gamma, beta, slope);
        MLSP_LAUNCH_CHECK("bn_rows_apply_kernel");
    } else {
        MLSP_REQUIRE(R <= 0x7fffffff, MLSP_EUNSUPPORTED, "bn_act_bwd: B too large");
        const long long xbs = x_batch_stride ? x_batch_stride : (long long)C * L, obs = y_batch_stride ? y_batch_stride : (long long)C * L;
        MLSP_REQUIRE(xbs >= (long long)C * L && obs >= (long long)C * L, MLSP_EINVAL, "bn_act_bwd: batch stride smaller than C * L");
        const int B = (int)R, gy = bn_ncl_gy(B, C);
        float *coef = reinterpret_cast<float *>(acc + (size_t)gy * 2 * C);
        const dim3 grid(C, gy);
        bn_ncl_reduce_kernel<1><<<grid, BN_THREADS, 0, st>>>(x, dy, B, C, L, xbs, obs, save_mean, save_invstd, gamma, beta, slope, acc);
        MLSP_LAUNCH_CHECK("bn_ncl_reduce_kernel");
        bn_finalize_bwd_kernel<<<fin_grid, BN_THREADS, 0, st>>>(acc, gy, coef, C, (double)B * (double)L, gamma, save_invstd, dgamma, dbeta);
        MLSP_LAUNCH_CHECK("bn_finalize_bwd_kernel");
        bn_ncl_apply_kernel<1><<<grid, BN_THREADS, 0, st>>>(x, dy, dx, B, C, L, xbs, obs, coef, save_mean, save_invstd, gamma, beta, slope);
        MLSP_LAUNCH_CHECK("bn_ncl_apply_kernel");
    }
    return MLSP_OK;
}
