// knn.cu -- a1: knn(x, k)  (PointDA/model_utils.py:9-16 == PointSegDA/Models.py:8-15)
//
// Exact fp32 path on the CUDA cores.  Arithmetic is pinned to oracle/mlsp_oracle.c:orc_knn:
//   xx_j = sum_c rn(x_cj^2)              (sequential adds, channel order)
//   dot  = 8 interleaved fmaf chains (channel groups of 4, round robin) + fixed butterfly tree (dot_tree);
//          a single sequential chain when C <= 4
//   pd   = rn( rn(2 dot - xx_j) - xx_i ) == ((-xx_j) - (-2 dot)) - xx_i of the reference
// and the ranking is (pd descending, index ascending).  Nothing of size (B,N,N) is materialised: a CTA
// keeps its query rows in shared memory, streams candidate tiles through shared memory, and each warp
// maintains the running top-k of its rows in registers (topk.cuh), so the HBM traffic is the
// algorithmic 4BCN + 8BNk bytes plus L2-resident re-reads of the cloud.
#include "common.cuh"
#include "topk.cuh"

namespace mlsp {

constexpr int KNN_THREADS = 256;
constexpr int KNN_TJ = 128;                               // candidates per tile (4 per lane)
constexpr size_t KNN_WS_HEADER = 256;                     // stats live in the first bytes of the workspace

bool knn_tensor_supported(int B, int C, int N, int k);
size_t knn_tensor_workspace_bytes(int B, int C, int N, int k);
int knn_tensor_run(const float *x, int B, int C, int N, int k, int64_t *idx, void *ws, float *dump, float *edge_out, int stages_mask,
                   cudaStream_t st, long long *tstamp = nullptr, int cluster = 0, bool want_stats = false);
bool knn3_supported(int C, int N, int k);
int knn3_run(const float *x, int B, int N, int k, int64_t *idx, int *stats, float *edge_out, cudaStream_t st);

size_t knn_workspace_bytes(int B, int C, int N, int k)
{
    // k > 64: the exact kernel runs in rounds of 64 ranks; a (pd, index) bound per row carries over between rounds
    const size_t bounds = k > 64 ? align_up(sizeof(float2) * (size_t)B * N, 256) : 0;
    const size_t exact = KNN_WS_HEADER + align_up(sizeof(float) * (size_t)B * N, 256) + bounds;
    if (knn_tensor_supported(B, C, N, k)) return exact > knn_tensor_workspace_bytes(B, C, N, k) ? exact : knn_tensor_workspace_bytes(B, C, N, k);
    return exact;
}

__global__ void sq_norms_kernel(const float *__restrict__ x, int C, int N, float *__restrict__ xx)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (j >= N) return;
    const float *xb = x + (size_t)b * C * N;
    float v = xb[j];
    float s = __fmul_rn(v, v);
    for (int c = 1; c < C; ++c) {
        v = xb[(size_t)c * N + j];
        s = __fadd_rn(s, __fmul_rn(v, v));
    }
    xx[(size_t)b * N + j] = s;
}

// NPART = number of interleaved partial chains of the pinned dot product (oracle dot_tree):
//   C <= 4 : one chain (NPART = 1, 8 rows per warp);  otherwise 8 chains + butterfly (NPART = 8, 2 rows per warp).
//
// Any k <= N (the reference's args.k is a free flag): the kernel selects `k` ranks starting at rank `rank0` of a ranking whose
// rows are `kout` long.  Round 0 is the plain top-k; a later round reads the row's bound (pd, j) = the last entry of the
// previous round and only considers candidates that rank strictly after it -- (pd < bound) or (pd == bound and j > bound.j)
// -- so ceil(kout / 64) launches produce the exact ranking of any length with the same arithmetic and tie rule.
template <int KSLOTS, int NPART, int R>
__global__ void __launch_bounds__(KNN_THREADS, 2)
knn_exact_kernel(const float *__restrict__ x, const float *__restrict__ xx, int C, int N, int k,
                 int64_t *__restrict__ idx, int rank0, int kout, float2 *__restrict__ bound)
{
    constexpr int ROWS = (KNN_THREADS / 32) * R;
    constexpr int CK = (NPART == 1) ? 4 : 32;          // channels per shared-memory chunk (32 = one round of the 8 chains)
    extern __shared__ __align__(16) float smem[];
    float *rows_s = smem;                              // [C][ROWS]
    float *cand_s = rows_s + (size_t)C * ROWS;         // [CK][KNN_TJ]
    float *cn_s = cand_s + CK * KNN_TJ;                // [KNN_TJ] candidate norms

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int i0 = blockIdx.x * ROWS;
    const float *xb = x + (size_t)b * C * N;
    const float *xxb = xx + (size_t)b * N;

    for (int e = tid; e < C * ROWS; e += KNN_THREADS) {
        const int c = e / ROWS, r = e % ROWS;
        const int i = i0 + r;
        rows_s[e] = (i < N) ? xb[(size_t)c * N + i] : 0.0f;
    }
    float xxi[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
        const int i = i0 + warp * R + rr;
        xxi[rr] = (i < N) ? xxb[i] : 0.0f;
    }
    TopK<KSLOTS> top[R];
    float bpd[R];
    int bj[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
        top[rr].init(k);
        const int i = i0 + warp * R + rr;
        bpd[rr] = INFINITY;
        bj[rr] = -1;
        if (rank0 > 0 && i < N) {
            const float2 bd = bound[(size_t)b * N + i];
            bpd[rr] = bd.x;
            bj[rr] = __float_as_int(bd.y);
        }
    }

    for (int j0 = 0; j0 < N; j0 += KNN_TJ) {
        float acc[R][4][NPART];
#pragma unroll
        for (int rr = 0; rr < R; ++rr)
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int t = 0; t < NPART; ++t) acc[rr][s][t] = 0.0f;
        for (int c0 = 0; c0 < C; c0 += CK) {
            __syncthreads();  // previous chunk (and, first time, rows_s) settled
            for (int e = tid; e < CK * KNN_TJ; e += KNN_THREADS) {
                const int cc = e / KNN_TJ, jj = e % KNN_TJ;
                const int c = c0 + cc, j = j0 + jj;
                cand_s[e] = (c < C && j < N) ? xb[(size_t)c * N + j] : 0.0f;
            }
            if (c0 == 0 && tid < KNN_TJ) cn_s[tid] = (j0 + tid < N) ? xxb[j0 + tid] : 0.0f;
            __syncthreads();
#pragma unroll
            for (int cc = 0; cc < CK; ++cc) {
                const int c = c0 + cc;
                if (c < C) {
                    float rv[R];
#pragma unroll
                    for (int rr = 0; rr < R; ++rr) rv[rr] = rows_s[c * ROWS + warp * R + rr];
                    float cv[4];
#pragma unroll
                    for (int s = 0; s < 4; ++s) cv[s] = cand_s[cc * KNN_TJ + s * 32 + lane];
                    constexpr int dummy = 0;
                    const int t = (NPART == 1) ? dummy : (cc >> 2);      // (c/4) mod 8, static: c0 is a multiple of 32
#pragma unroll
                    for (int rr = 0; rr < R; ++rr)
#pragma unroll
                        for (int s = 0; s < 4; ++s)
#pragma unroll
                            for (int tt = 0; tt < NPART; ++tt)
                                if (tt == t) acc[rr][s][tt] = __fmaf_rn(rv[rr], cv[s], acc[rr][s][tt]);
                }
            }
        }
        float cn[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) cn[s] = cn_s[s * 32 + lane];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int j = j0 + s * 32 + lane;
                float dot;
                if (NPART == 1) {
                    dot = acc[rr][s][0];
                } else {
                    const float *p = acc[rr][s];
                    const float q0 = __fadd_rn(p[0], p[4 % NPART]), q1 = __fadd_rn(p[1 % NPART], p[5 % NPART]);
                    const float q2 = __fadd_rn(p[2 % NPART], p[6 % NPART]), q3 = __fadd_rn(p[3 % NPART], p[7 % NPART]);
                    dot = __fadd_rn(__fadd_rn(q0, q2), __fadd_rn(q1, q3));
                }
                const float t = __fmaf_rn(2.0f, dot, -cn[s]);
                const float pd = __fsub_rn(t, xxi[rr]);
                const bool after = pd < bpd[rr] || (pd == bpd[rr] && j > bj[rr]);     // always true in round 0
                top[rr].offer(pd, j, j < N && after);
            }
        }
    }
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
        const int i = i0 + warp * R + rr;
        // the worst kept entry (lowest value, largest index among equals) is the last of this round's ranking
        if (bound != nullptr && i < N && lane == 0) bound[(size_t)b * N + i] = make_float2(top[rr].wmin, __int_as_float(top[rr].wj));
        top[rr].finish(k);
        if (i < N) {
#pragma unroll
            for (int s = 0; s < KSLOTS; ++s) {
                const int e = s * 32 + lane;
                if (e < k) idx[((size_t)b * N + i) * kout + rank0 + e] = (int64_t)top[rr].j[s];
            }
        }
    }
}

template <int KSLOTS, int NPART, int R>
static int launch_exact_t(const float *x, const float *xx, int B, int C, int N, int k, int64_t *idx, cudaStream_t st,
                          int rank0 = 0, int kout = 0, float2 *bound = nullptr)
{
    constexpr int ROWS = (KNN_THREADS / 32) * R;
    constexpr int CK = (NPART == 1) ? 4 : 32;
    const size_t smem = sizeof(float) * ((size_t)C * ROWS + CK * KNN_TJ + KNN_TJ);
    MLSP_REQUIRE(smem <= 200 * 1024, MLSP_EUNSUPPORTED, "knn: C=%d too large for the exact kernel", C);
    dim3 grid((N + ROWS - 1) / ROWS, B);
    MLSP_CUDA(cudaFuncSetAttribute(knn_exact_kernel<KSLOTS, NPART, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_exact_kernel<KSLOTS, NPART, R><<<grid, KNN_THREADS, smem, st>>>(x, xx, C, N, k, idx, rank0, kout ? kout : k, bound);
    MLSP_LAUNCH_CHECK("knn_exact_kernel");
    return MLSP_OK;
}

static int launch_exact(const float *x, const float *xx, int B, int C, int N, int k, int64_t *idx,
                        cudaStream_t st, float2 *bound = nullptr)
{
    if (k > 64) {                                       // rounds of 64 ranks (see knn_exact_kernel)
        for (int rank0 = 0; rank0 < k; rank0 += 64) {
            const int kr = k - rank0 < 64 ? k - rank0 : 64;
            const int rc = (C <= 4) ? launch_exact_t<2, 1, 8>(x, xx, B, C, N, kr, idx, st, rank0, k, bound)
                                    : launch_exact_t<2, 8, 2>(x, xx, B, C, N, kr, idx, st, rank0, k, bound);
            if (rc) return rc;
        }
        return MLSP_OK;
    }
    if (C <= 4) return k <= 32 ? launch_exact_t<1, 1, 8>(x, xx, B, C, N, k, idx, st) : launch_exact_t<2, 1, 8>(x, xx, B, C, N, k, idx, st);
    return k <= 32 ? launch_exact_t<1, 8, 2>(x, xx, B, C, N, k, idx, st) : launch_exact_t<2, 8, 2>(x, xx, B, C, N, k, idx, st);
}

}  // namespace mlsp

extern "C" int mlsp_knn_f32(const float *x, int B, int C, int N, int k, int64_t *idx, void *ws, size_t ws_bytes,
                            int flags, void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(x && idx && ws, MLSP_EINVAL, "knn: null pointer");
    MLSP_REQUIRE(B > 0 && C > 0 && N > 0, MLSP_EINVAL, "knn: bad shape B=%d C=%d N=%d", B, C, N);
    MLSP_REQUIRE(k >= 1 && k <= N, MLSP_EINVAL, "knn: k=%d out of range for N=%d", k, N);
    MLSP_REQUIRE(B <= 65535, MLSP_EUNSUPPORTED, "knn: B=%d > 65535", B);
    MLSP_REQUIRE(ws_bytes >= knn_workspace_bytes(B, C, N, k), MLSP_EWORKSPACE, "knn: workspace too small");
    cudaStream_t st = as_stream(stream);
    const bool tensor_ok = knn_tensor_supported(B, C, N, k);
    const bool want_stats = (flags & MLSP_KNN_STATS) != 0;
    flags &= ~MLSP_KNN_STATS;
    MLSP_REQUIRE(flags != MLSP_KNN_TENSOR_ONLY || tensor_ok, MLSP_EUNSUPPORTED,
                 "knn: tensor path not available for B=%d C=%d N=%d k=%d", B, C, N, k);
    if (tensor_ok && flags != MLSP_KNN_EXACT_ONLY) return knn_tensor_run(x, B, C, N, k, idx, ws, nullptr, nullptr, 7, st, nullptr, 0, want_stats);
    if (flags != MLSP_KNN_EXACT_ONLY && knn3_supported(C, N, k)) return knn3_run(x, B, N, k, idx, static_cast<int *>(ws), nullptr, st);
    MLSP_CUDA(cudaMemsetAsync(ws, 0, KNN_WS_HEADER, st));
    float *xx = reinterpret_cast<float *>(static_cast<char *>(ws) + KNN_WS_HEADER);
    sq_norms_kernel<<<dim3((N + 255) / 256, B), 256, 0, st>>>(x, C, N, xx);
    MLSP_LAUNCH_CHECK("sq_norms_kernel");
    float2 *bound = reinterpret_cast<float2 *>(static_cast<char *>(ws) + KNN_WS_HEADER + align_up(sizeof(float) * (size_t)B * N, 256));
    return launch_exact(x, xx, B, C, N, k, idx, st, k > 64 ? bound : nullptr);
}

// Test hook: the tensor path with a dump of the approximate filter values v = |x_j|^2 - 2 dot~ (B,N,N).
// Profiling hook: the tensor path with phase marks of the filter kernel; tstamp (CTAs, 16) int64, CTA = b * ceil(N/128) + row block:
// %globaltimer (ns) at [0] kernel entry, [1] prologue done, [2] first accumulator ready, [3] pass 1 done, [4] class maxima
// sorted, [5] exchanged, [6] threshold ready, [7] pass 2 done, [8] lists complete, [9] lists copied out; [15] = SM id.
extern "C" int mlsp_knn_tensor_timeline(const float *x, int B, int C, int N, int k, int64_t *idx, void *ws, size_t ws_bytes,
                                        long long *tstamp, int cluster, void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(x && idx && ws && tstamp, MLSP_EINVAL, "knn_tensor_timeline: null pointer");
    MLSP_REQUIRE(k >= 1 && k <= N, MLSP_EINVAL, "knn_tensor_timeline: k out of range");
    MLSP_REQUIRE(knn_tensor_supported(B, C, N, k), MLSP_EUNSUPPORTED, "knn_tensor_timeline: shape not supported");
    MLSP_REQUIRE(ws_bytes >= knn_workspace_bytes(B, C, N, k), MLSP_EWORKSPACE, "knn_tensor_timeline: workspace too small");
    return knn_tensor_run(x, B, C, N, k, idx, ws, nullptr, nullptr, 7, as_stream(stream), tstamp, cluster);
}

extern "C" int mlsp_knn_tensor_debug(const float *x, int B, int C, int N, int k, int64_t *idx, void *ws, size_t ws_bytes,
                                     float *dump, void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(x && idx && ws, MLSP_EINVAL, "knn_tensor_debug: null pointer");
    MLSP_REQUIRE(k >= 1 && k <= N, MLSP_EINVAL, "knn_tensor_debug: k out of range");
    MLSP_REQUIRE(knn_tensor_supported(B, C, N, k), MLSP_EUNSUPPORTED, "knn_tensor_debug: shape not supported");
    MLSP_REQUIRE(ws_bytes >= knn_workspace_bytes(B, C, N, k), MLSP_EWORKSPACE, "knn_tensor_debug: workspace too small");
    return knn_tensor_run(x, B, C, N, k, idx, ws, dump, nullptr, 7, as_stream(stream));
}
