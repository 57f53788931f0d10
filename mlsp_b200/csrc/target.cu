// target.cu -- the masked-local-structure target builder (a4, a5, a6, a7, a8):
//   voxel regions + histogram + first-fit choice   utils/pc_utils.py:33-73, MLSP/mlsp.py:28-50
//   in-place deformation + mask                    MLSP/mlsp.py:44-48, utils/pc_utils.py:105-110
//   ball membership counts (collapse_to_point)      utils/pc_utils.py:86-99
//   per-point ball cardinality + soft labels        MLSP/mlsp.py:240-272 (python-pcl radius search)
//   PCA normals                                     PointDA/trainer.py:173-188 (python-pcl NormalEstimation)
// Arithmetic pinned to oracle/mlsp_oracle.c and oracle/np_ops.py.  All of these move a few hundred KB
// at the PointDA shape, so each is a single launch with the cloud staged once in shared memory.
#include "common.cuh"
#include "topk.cuh"

namespace mlsp {

// ---------------------------------------------------------------------------------------------------
// a4: region id per point, per-cloud histogram, first-fit choice.  One CTA per cloud.
// Thresholds are the python doubles -1 + q*(2/3) rounded to float32 (what torch compares against);
// clamp(+-0.99999999) is clamp(+-1.0f) in float32; all six comparisons are strict.
struct RegionOrder {
    int32_t id[27];
};

// X (B,C,N) addressed X[b*sb + c*sc + n*sn]: the reference's trainers hand deform_input the permuted view
// `data.permute(0,2,1)` of a (B,N,3) batch (PointDA/trainer.py:380-387), i.e. strides (3N, 1, 3); it is deformed in place.
struct CloudView {
    long long sb, sc, sn;
};

__device__ __forceinline__ int voxel_axis(float v)
{
    const float t1 = -0.3333333432674408f, t2 = 0.3333333432674408f;
    v = fminf(fmaxf(v, -1.0f), 1.0f);
    if (-1.0f < v && v < t1) return 0;
    if (t1 < v && v < t2) return 1;
    if (t2 < v && v < 1.0f) return 2;
    return -1;
}

__global__ void __launch_bounds__(256)
region_assign_select_kernel(const float *__restrict__ X, CloudView V, int N, RegionOrder order, int min_pts,
                            int64_t *__restrict__ region, int32_t *__restrict__ counts,
                            int32_t *__restrict__ chosen, int32_t *__restrict__ nsel)
{
    __shared__ int hist[27];
    const int b = blockIdx.x;
    if (threadIdx.x < 27) hist[threadIdx.x] = 0;
    __syncthreads();
    const float *Xb = X + b * V.sb;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float *pn = Xb + n * V.sn;
        const int qx = voxel_axis(pn[0]), qy = voxel_axis(pn[V.sc]), qz = voxel_axis(pn[2 * V.sc]);
        const int r = (qx >= 0 && qy >= 0 && qz >= 0) ? 9 * qx + 3 * qy + qz : 0;
        region[(size_t)b * N + n] = r;
        atomicAdd(&hist[r], 1);
    }
    __syncthreads();
    if (threadIdx.x < 27) counts[b * 27 + threadIdx.x] = hist[threadIdx.x];
    if (threadIdx.x == 0) {
        int c = -1, n = 0;
        for (int t = 0; t < 27; ++t) {
            const int r = order.id[t];
            if (hist[r] >= min_pts) {
                c = r;
                n = hist[r];
                break;
            }
        }
        chosen[b] = c;
        nsel[b] = n;
    }
}

// ---------------------------------------------------------------------------------------------------
// Shared tail of a4/a5: points flagged for cloud b receive consecutive noise rows in ascending point
// order (the order of a boolean-mask assignment in torch), mask is written for every element.
// One CTA per cloud; block-wide exclusive scan over chunks of blockDim points.
template <typename FlagFn>
__device__ __forceinline__ void scatter_flagged(float *Xb, CloudView V, float *Mb, int C, int N, const float *noise,
                                                FlagFn flag)
{
    __shared__ int warp_tot[32];
    __shared__ int running;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int n0 = 0; n0 < N; n0 += blockDim.x) {
        const int n = n0 + threadIdx.x;
        const bool f = (n < N) && flag(n);
        const unsigned bal = __ballot_sync(MLSP_FULL, f);
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int before = running;
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
        const int rank = before + __popc(bal & ((1u << lane) - 1u));
        if (n < N) {
            for (int c = 0; c < C; ++c) Mb[(size_t)c * N + n] = (f && c < 3) ? 1.0f : 0.0f;
            if (f && noise) {
                float *pn = Xb + n * V.sn;
                pn[0] = noise[(size_t)rank * 3 + 0];
                pn[V.sc] = noise[(size_t)rank * 3 + 1];
                pn[2 * V.sc] = noise[(size_t)rank * 3 + 2];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = running;
            for (int w = 0; w < nwarps; ++w) t += warp_tot[w];
            running = t;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
region_mask_scatter_kernel(float *__restrict__ X, CloudView V, int C, int N, const int64_t *__restrict__ region,
                           const int32_t *__restrict__ chosen, const float *__restrict__ noise,
                           const int32_t *__restrict__ offset, float *__restrict__ mask)
{
    const int b = blockIdx.x;
    const int sel = chosen[b];
    const int64_t *rb = region + (size_t)b * N;
    const float *nz = noise ? noise + (size_t)offset[b] * 3 : nullptr;
    scatter_flagged(X + b * V.sb, V, mask + (size_t)b * C * N, C, N, nz,
                    [&](int n) { return sel >= 0 && rb[n] == (int64_t)sel; });
}

// ---------------------------------------------------------------------------------------------------
// a5: pd(i,j) = rn(rn(xx_j - 2 dot) + xx_i), dot = fma chain, xx = (x^2 + y^2) + z^2.
__device__ __forceinline__ float sq3(float x, float y, float z)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}
__device__ __forceinline__ float ball_pd(float4 pi, float4 pj)
{
    const float dot = __fmaf_rn(pi.z, pj.z, __fmaf_rn(pi.y, pj.y, __fmul_rn(pi.x, pj.x)));
    return __fadd_rn(__fmaf_rn(-2.0f, dot, pj.w), pi.w);
}

constexpr int BALL_ROWS_PER_WARP = 4;
constexpr int BALL_THREADS = 256;
constexpr int BALL_ROWS = (BALL_THREADS / 32) * BALL_ROWS_PER_WARP;

// stage cloud b as float4 (x,y,z,xx) into shared memory, N points, from (C,N) storage with strides (sc, sn)
__device__ __forceinline__ void stage_cloud_soa(const float *Xb, CloudView V, int N, float4 *s)
{
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float *pn = Xb + n * V.sn;
        const float x = pn[0], y = pn[V.sc], z = pn[2 * V.sc];
        s[n] = make_float4(x, y, z, sq3(x, y, z));
    }
}

__global__ void __launch_bounds__(BALL_THREADS)
ball_count_kernel(const float *__restrict__ X, CloudView V, int N, float r2, int32_t *__restrict__ cnt)
{
    extern __shared__ float4 cloud[];
    const int b = blockIdx.y;
    stage_cloud_soa(X + b * V.sb, V, N, cloud);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = blockIdx.x * BALL_ROWS + warp * BALL_ROWS_PER_WARP;
    float4 pi[BALL_ROWS_PER_WARP];
    int c[BALL_ROWS_PER_WARP];
#pragma unroll
    for (int r = 0; r < BALL_ROWS_PER_WARP; ++r) {
        pi[r] = cloud[min(i0 + r, N - 1)];
        c[r] = 0;
    }
    for (int j = lane; j < N; j += 32) {
        const float4 pj = cloud[j];
#pragma unroll
        for (int r = 0; r < BALL_ROWS_PER_WARP; ++r) c[r] += (ball_pd(pi[r], pj) <= r2) ? 1 : 0;
    }
#pragma unroll
    for (int r = 0; r < BALL_ROWS_PER_WARP; ++r) {
        const int tot = __reduce_add_sync(MLSP_FULL, c[r]);
        if (lane == 0 && i0 + r < N) cnt[(size_t)b * N + i0 + r] = tot;
    }
}

__global__ void __launch_bounds__(256)
ball_mask_scatter_kernel(float *__restrict__ X, CloudView V, int C, int N, float r2, const int32_t *__restrict__ centre,
                         const float *__restrict__ noise, const int32_t *__restrict__ offset,
                         float *__restrict__ mask)
{
    const int b = blockIdx.x;
    float *Xb = X + b * V.sb;
    const int ci = centre[b];
    float4 pc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ci >= 0 && ci < N) {
        const float *pn = Xb + ci * V.sn;
        const float x = pn[0], y = pn[V.sc], z = pn[2 * V.sc];
        pc = make_float4(x, y, z, sq3(x, y, z));
    }
    __syncthreads();  // every thread has read the centre before any thread overwrites it
    const float *nz = noise ? noise + (size_t)offset[b] * 3 : nullptr;
    scatter_flagged(Xb, V, mask + (size_t)b * C * N, C, N, nz, [&](int n) {
        if (ci < 0 || ci >= N) return false;
        const float *pn = Xb + n * V.sn;
        const float x = pn[0], y = pn[V.sc], z = pn[2 * V.sc];
        return ball_pd(pc, make_float4(x, y, z, sq3(x, y, z))) <= r2;
    });
}

// ---------------------------------------------------------------------------------------------------
// a6: cardinality with python-pcl radius_search_for_cloud semantics (restated, parity unpinned):
//   d = (rn(dx^2) + rn(dy^2)) + rn(dz^2); c1 = #{d < r2}; c2 = #{d < d_i0};
//   cnt = min(c1, K) - [d_i0 < r2 && c2 < K];  row = clip(cnt - shift, 0, (num_cls-1)*pergroup)
//   label = (onehot(floor(row/pg)) + onehot(ceil(row/pg))) / 2
__device__ __forceinline__ float direct_d2(float4 a, float4 q)
{
    const float dx = __fsub_rn(a.x, q.x), dy = __fsub_rn(a.y, q.y), dz = __fsub_rn(a.z, q.z);
    return sq3(dx, dy, dz);
}

__global__ void __launch_bounds__(BALL_THREADS)
ball_count_labels_kernel(const float *__restrict__ pts, int N, float r2, int K, int shift, int pergroup,
                         int num_cls, float *__restrict__ labels, int64_t *__restrict__ row)
{
    extern __shared__ float4 cloud[];
    const int b = blockIdx.y;
    const float *P = pts + (size_t)b * N * 3;
    for (int n = threadIdx.x; n < N; n += blockDim.x) cloud[n] = make_float4(P[3 * n], P[3 * n + 1], P[3 * n + 2], 0.f);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = blockIdx.x * BALL_ROWS + warp * BALL_ROWS_PER_WARP;
    const float4 p0 = cloud[0];
    float4 pi[BALL_ROWS_PER_WARP];
    float d0[BALL_ROWS_PER_WARP];
    int c1[BALL_ROWS_PER_WARP], c2[BALL_ROWS_PER_WARP];
#pragma unroll
    for (int r = 0; r < BALL_ROWS_PER_WARP; ++r) {
        pi[r] = cloud[min(i0 + r, N - 1)];
        d0[r] = direct_d2(pi[r], p0);
        c1[r] = c2[r] = 0;
    }
    for (int j = lane; j < N; j += 32) {
        const float4 pj = cloud[j];
#pragma unroll
        for (int r = 0; r < BALL_ROWS_PER_WARP; ++r) {
            const float d = direct_d2(pi[r], pj);
            c1[r] += (d < r2) ? 1 : 0;
            c2[r] += (d < d0[r]) ? 1 : 0;
        }
    }
    const int top = (num_cls - 1) * pergroup;
#pragma unroll
    for (int r = 0; r < BALL_ROWS_PER_WARP; ++r) {
        const int t1 = __reduce_add_sync(MLSP_FULL, c1[r]);
        const int t2 = __reduce_add_sync(MLSP_FULL, c2[r]);
        const int i = i0 + r;
        if (i >= N) continue;
        const int in0 = (d0[r] < r2 && t2 < K) ? 1 : 0;
        int v = min(t1, K) - in0 - shift;
        v = max(0, min(v, top));
        const int lo = v / pergroup, hi = (v + pergroup - 1) / pergroup;
        if (lane == 0) row[(size_t)b * N + i] = v;
        float *L = labels + ((size_t)b * N + i) * num_cls;
        for (int c = lane; c < num_cls; c += 32) L[c] = 0.5f * (float)(c == lo) + 0.5f * (float)(c == hi);
    }
}

// ---------------------------------------------------------------------------------------------------
// a6, list form: python-pcl KdTreeFLANN.radius_search_for_cloud(cloud, r, K) (MLSP/mlsp.py:250), restated
// like the count above: for every point the neighbours with d < r2, nearest first (ties by lowest index),
// at most K of them, rows zero-padded.  One warp per query row: the in-ball points enter the streaming
// selection of topk.cuh (value -d, so "largest" = nearest), which keeps the K nearest; distances of the
// kept entries are recomputed with the same pinned expression.
template <int KSLOTS>
__global__ void __launch_bounds__(BALL_THREADS)
radius_search_kernel(const float *__restrict__ pts, int N, float r2, int K, int32_t *__restrict__ ind,
                     float *__restrict__ sqd)
{
    extern __shared__ float4 cloud[];
    const int b = blockIdx.y;
    const float *P = pts + (size_t)b * N * 3;
    for (int n = threadIdx.x; n < N; n += blockDim.x) cloud[n] = make_float4(P[3 * n], P[3 * n + 1], P[3 * n + 2], 0.f);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = blockIdx.x * BALL_ROWS + warp * BALL_ROWS_PER_WARP;
    for (int r = 0; r < BALL_ROWS_PER_WARP; ++r) {
        const int i = i0 + r;
        if (i >= N) break;                                   // warp-uniform
        const float4 pi = cloud[i];
        TopK<KSLOTS> top;
        top.init(K);
        for (int j0 = 0; j0 < N; j0 += 32) {
            const int j = j0 + lane;
            const float d = (j < N) ? direct_d2(pi, cloud[j]) : INFINITY;
            top.offer(-d, j, d < r2);
        }
        top.finish(K);
#pragma unroll
        for (int s = 0; s < KSLOTS; ++s) {
            const int e = s * 32 + lane;
            if (e >= K) continue;
            const int j = top.j[s];
            const bool hit = j >= 0 && j < N;                // unfilled slots keep their huge placeholder index
            const size_t o = ((size_t)b * N + i) * K + e;
            ind[o] = hit ? j : 0;
            sqd[o] = hit ? direct_d2(pi, cloud[j]) : 0.0f;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// a7: PCA normals.  One thread per point, fp64 covariance about the neighbourhood mean and a cyclic
// Jacobi eigen-solve of the symmetric 3x3 (robust for the near-degenerate planar patches of CAD scans).
__device__ __forceinline__ void jacobi_rotate(double &app, double &aqq, double &apq, double &arp, double &arq,
                                              double (&v)[3][3], int p, int q)
{
    if (apq == 0.0) return;
    const double theta = (aqq - app) / (2.0 * apq);
    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
    const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
    app -= t * apq;
    aqq += t * apq;
    apq = 0.0;
    const double rp = arp, rq = arq;  // the remaining off-diagonal pair (r,p), (r,q)
    arp = c * rp - s * rq;
    arq = s * rp + c * rq;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double vp = v[i][p], vq = v[i][q];
        v[i][p] = c * vp - s * vq;
        v[i][q] = s * vp + c * vq;
    }
}

__global__ void __launch_bounds__(128)
pca_normals_kernel(const float *__restrict__ pts, const int64_t *__restrict__ idx, int N, int k,
                   float *__restrict__ normals, float *__restrict__ curvature, long long total)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    const long long b = p / N;
    const float *P = pts + b * (long long)N * 3;
    const int64_t *nb = idx + p * k;
    double mx = 0, my = 0, mz = 0;
    for (int j = 0; j < k; ++j) {
        const float *q = P + nb[j] * 3;
        mx += q[0];
        my += q[1];
        mz += q[2];
    }
    mx /= k;
    my /= k;
    mz /= k;
    double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0;
    for (int j = 0; j < k; ++j) {
        const float *q = P + nb[j] * 3;
        const double dx = q[0] - mx, dy = q[1] - my, dz = q[2] - mz;
        a00 += dx * dx;
        a01 += dx * dy;
        a02 += dx * dz;
        a11 += dy * dy;
        a12 += dy * dz;
        a22 += dz * dz;
    }
    const double inv = 1.0 / k;
    a00 *= inv; a01 *= inv; a02 *= inv; a11 *= inv; a12 *= inv; a22 *= inv;
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(a01) + fabs(a02) + fabs(a12);
        const double diag = fabs(a00) + fabs(a11) + fabs(a22);
        if (off <= 1e-18 * diag || off == 0.0) break;
        jacobi_rotate(a00, a11, a01, a02, a12, v, 0, 1);   // zero (0,1); r = 2: pairs (2,0),(2,1)
        jacobi_rotate(a00, a22, a02, a01, a12, v, 0, 2);   // zero (0,2); r = 1: pairs (1,0),(1,2)
        jacobi_rotate(a11, a22, a12, a01, a02, v, 1, 2);   // zero (1,2); r = 0: pairs (0,1),(0,2)
    }
    int m = 0;
    double lam = a00;
    if (a11 < lam) { lam = a11; m = 1; }
    if (a22 < lam) { lam = a22; m = 2; }
    double nx = v[0][m], ny = v[1][m], nz = v[2][m];
    const double nn = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
    nx *= nn; ny *= nn; nz *= nn;
    const float *self = P + (p - b * N) * 3;
    if (nx * self[0] + ny * self[1] + nz * self[2] > 0.0) { nx = -nx; ny = -ny; nz = -nz; }  // towards the origin
    normals[p * 3 + 0] = (float)nx;
    normals[p * 3 + 1] = (float)ny;
    normals[p * 3 + 2] = (float)nz;
    if (curvature) {
        const double tr = a00 + a11 + a22;
        curvature[p] = tr > 0.0 ? (float)(fabs(lam) / tr) : 0.0f;
    }
}

}  // namespace mlsp

// =====================================================================================================
// strides (in floats) of a (B,C,N) view; all zero = dense
static inline mlsp::CloudView cloud_view(int64_t sb, int64_t sc, int64_t sn, int C, int N)
{
    mlsp::CloudView V;
    if (sb == 0 && sc == 0 && sn == 0) { V.sb = (long long)C * N; V.sc = N; V.sn = 1; }
    else { V.sb = sb; V.sc = sc; V.sn = sn; }
    return V;
}

extern "C" int mlsp_region_assign_select(const float *X, int64_t xs_b, int64_t xs_c, int64_t xs_n, int B, int C, int N, const int32_t *order_host,
                                         int min_pts, int64_t *region, int32_t *counts, int32_t *chosen,
                                         int32_t *nsel, void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(X && order_host && region && counts && chosen && nsel, MLSP_EINVAL, "region_assign_select: null pointer");
    MLSP_REQUIRE(B > 0 && C >= 3 && N > 0, MLSP_EINVAL, "region_assign_select: bad shape B=%d C=%d N=%d", B, C, N);
    RegionOrder ord;
    for (int t = 0; t < 27; ++t) {
        MLSP_REQUIRE(order_host[t] >= 0 && order_host[t] < 27, MLSP_EINVAL, "region_assign_select: bad region id");
        ord.id[t] = order_host[t];
    }
    region_assign_select_kernel<<<B, 256, 0, as_stream(stream)>>>(X, cloud_view(xs_b, xs_c, xs_n, C, N), N, ord, min_pts, region, counts, chosen, nsel);
    MLSP_LAUNCH_CHECK("region_assign_select_kernel");
    return MLSP_OK;
}

extern "C" int mlsp_region_mask_scatter(float *X, int64_t xs_b, int64_t xs_c, int64_t xs_n, int B, int C, int N, const int64_t *region, const int32_t *chosen,
                                        const float *noise, const int32_t *offset, float *mask, void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(X && region && chosen && mask, MLSP_EINVAL, "region_mask_scatter: null pointer");
    MLSP_REQUIRE(!noise || offset, MLSP_EINVAL, "region_mask_scatter: noise without offsets");
    MLSP_REQUIRE(B > 0 && C >= 3 && N > 0, MLSP_EINVAL, "region_mask_scatter: bad shape");
    region_mask_scatter_kernel<<<B, 256, 0, as_stream(stream)>>>(X, cloud_view(xs_b, xs_c, xs_n, C, N), C, N, region, chosen, noise, offset, mask);
    MLSP_LAUNCH_CHECK("region_mask_scatter_kernel");
    return MLSP_OK;
}

extern "C" int mlsp_ball_count(const float *x, int64_t xs_b, int64_t xs_c, int64_t xs_n, int B, int C, int N, float r2, int32_t *cnt, void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(x && cnt, MLSP_EINVAL, "ball_count: null pointer");
    MLSP_REQUIRE(B > 0 && C >= 3 && N > 0, MLSP_EINVAL, "ball_count: bad shape");
    const size_t smem = sizeof(float4) * (size_t)N;
    MLSP_REQUIRE(smem <= 200 * 1024, MLSP_EUNSUPPORTED, "ball_count: N=%d too large", N);
    MLSP_CUDA(cudaFuncSetAttribute(ball_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ball_count_kernel<<<dim3((N + BALL_ROWS - 1) / BALL_ROWS, B), BALL_THREADS, smem, as_stream(stream)>>>(x, cloud_view(xs_b, xs_c, xs_n, C, N), N, r2, cnt);
    MLSP_LAUNCH_CHECK("ball_count_kernel");
    return MLSP_OK;
}

extern "C" int mlsp_ball_mask_scatter(float *X, int64_t xs_b, int64_t xs_c, int64_t xs_n, int B, int C, int N, float r2, const int32_t *centre,
                                      const float *noise, const int32_t *offset, float *mask, void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(X && centre && mask, MLSP_EINVAL, "ball_mask_scatter: null pointer");
    MLSP_REQUIRE(!noise || offset, MLSP_EINVAL, "ball_mask_scatter: noise without offsets");
    MLSP_REQUIRE(B > 0 && C >= 3 && N > 0, MLSP_EINVAL, "ball_mask_scatter: bad shape");
    ball_mask_scatter_kernel<<<B, 256, 0, as_stream(stream)>>>(X, cloud_view(xs_b, xs_c, xs_n, C, N), C, N, r2, centre, noise, offset, mask);
    MLSP_LAUNCH_CHECK("ball_mask_scatter_kernel");
    return MLSP_OK;
}

extern "C" int mlsp_ball_count_labels(const float *pts, int B, int N, float r2, int K, int shift, int pergroup,
                                      int num_cls, float *labels, int64_t *row, void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(pts && labels && row, MLSP_EINVAL, "ball_count_labels: null pointer");
    MLSP_REQUIRE(B > 0 && N > 0 && K > 0 && pergroup > 0 && num_cls > 0, MLSP_EINVAL, "ball_count_labels: bad arguments");
    const size_t smem = sizeof(float4) * (size_t)N;
    MLSP_REQUIRE(smem <= 200 * 1024, MLSP_EUNSUPPORTED, "ball_count_labels: N=%d too large", N);
    MLSP_CUDA(cudaFuncSetAttribute(ball_count_labels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ball_count_labels_kernel<<<dim3((N + BALL_ROWS - 1) / BALL_ROWS, B), BALL_THREADS, smem, as_stream(stream)>>>(
        pts, N, r2, K, shift, pergroup, num_cls, labels, row);
    MLSP_LAUNCH_CHECK("ball_count_labels_kernel");
    return MLSP_OK;
}

extern "C" int mlsp_radius_search(const float *pts, int B, int N, float r2, int K, int32_t *ind, float *sqdist,
                                  void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(pts && ind && sqdist, MLSP_EINVAL, "radius_search: null pointer");
    MLSP_REQUIRE(B > 0 && N > 0 && K > 0, MLSP_EINVAL, "radius_search: bad arguments");
    MLSP_REQUIRE(K <= 128, MLSP_EUNSUPPORTED, "radius_search: K=%d > 128", K);
    const size_t smem = sizeof(float4) * (size_t)N;
    MLSP_REQUIRE(smem <= 200 * 1024, MLSP_EUNSUPPORTED, "radius_search: N=%d too large", N);
    const dim3 grid((N + BALL_ROWS - 1) / BALL_ROWS, B);
#define MLSP_RS_LAUNCH(S)                                                                                             \
    do {                                                                                                              \
        MLSP_CUDA(cudaFuncSetAttribute(radius_search_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        radius_search_kernel<S><<<grid, BALL_THREADS, smem, as_stream(stream)>>>(pts, N, r2, K, ind, sqdist);          \
    } while (0)
    if (K <= 32) MLSP_RS_LAUNCH(1);
    else if (K <= 64) MLSP_RS_LAUNCH(2);
    else MLSP_RS_LAUNCH(4);
#undef MLSP_RS_LAUNCH
    MLSP_LAUNCH_CHECK("radius_search_kernel");
    return MLSP_OK;
}

extern "C" int mlsp_pca_normals(const float *pts, const int64_t *idx, int B, int N, int k, float *normals,
                                float *curvature, void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(pts && idx && normals, MLSP_EINVAL, "pca_normals: null pointer");
    MLSP_REQUIRE(B > 0 && N > 0 && k >= 1, MLSP_EINVAL, "pca_normals: bad shape");
    const long long total = (long long)B * N;
    pca_normals_kernel<<<(unsigned)((total + 127) / 128), 128, 0, as_stream(stream)>>>(pts, idx, N, k, normals, curvature, total);
    MLSP_LAUNCH_CHECK("pca_normals_kernel");
    return MLSP_OK;
}
