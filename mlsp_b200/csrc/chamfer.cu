// chamfer.cu -- a9/a10: masked Chamfer position loss, forward and backward
// (chamfer_distance MLSP/mlsp.py:115-153, reconstruction_loss :156-182, findneareat_index :196-220).
//
// The reference materialises two (B,N,N,3) repeats and a (B,N,N) matrix and lets autograd keep them.
// Only rows with mask_i != 0 reach the loss (`dist * mask_cord`, :151), so the forward evaluates just
// those rows (about 40-70 of 1024 at the PointDA shape) against all N columns: the column cloud is
// staged once per CTA in shared memory as (x,y,z,penalty), one warp per row, lanes over columns,
// two redux.sync for (min, lowest argmin).  Arithmetic pinned to oracle/mlsp_oracle.c:orc_chamfer_dir:
//   s = (rn(dx^2)+rn(dy^2))+rn(dz^2);  D = rn(sqrt(s))^2 + pen_j   (sqrt-then-square as in :138)
// The backward uses the saved argmin: d/dp1_i = 2 (p1_i - p2_j*) w, d/dp2_j* = -that.
#include "common.cuh"

namespace mlsp {

constexpr int CH_THREADS = 256;
constexpr int CH_WARPS = CH_THREADS / 32;
constexpr int CH_MAX_TILES = 64;

size_t chamfer_workspace_bytes(int B, int N)
{
    (void)N;   // (sum, count) partials per CTA; x2: both directions of mlsp_reconstruction_loss_fwd
    return align_up(sizeof(float) * 4 * (size_t)B * CH_MAX_TILES, 256);
}

struct PtView {
    const float *p;
    long long bs, ps, cs;
    __device__ __forceinline__ float3 get(int b, int i) const
    {
        const float *q = p + b * bs + i * ps;
        return make_float3(q[0], q[cs], q[2 * cs]);
    }
};

__device__ __forceinline__ void chamfer_fwd_body(const PtView &p1, const PtView &p2, const float *__restrict__ mask,
                                                 long long mask_bs, int N, int all_rows, int rows_per_cta,
                                                 float *__restrict__ rowmin, int64_t *__restrict__ argmin,
                                                 float *__restrict__ part_s, float *__restrict__ part_c)
{
    extern __shared__ float4 cols[];            // [N] (x,y,z,pen)
    int *list = reinterpret_cast<int *>(cols + N);  // [rows_per_cta] compacted row ids
    __shared__ int nlist;
    __shared__ float wsum[CH_WARPS], wcnt[CH_WARPS];

    const int b = blockIdx.y, t = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *mb = mask + b * mask_bs;
    const int r0 = t * rows_per_cta, r1 = min(N, r0 + rows_per_cta);

    if (threadIdx.x == 0) nlist = 0;
    __syncthreads();
    // ordered compaction of the rows to evaluate (ascending row id keeps the summation order fixed)
    {
        __shared__ int wtot[CH_WARPS];
        for (int base = r0; base < r1; base += CH_THREADS) {
            const int i = base + threadIdx.x;
            float m = 0.f;
            bool take = false;
            if (i < r1) {
                m = mb[i];
                take = all_rows || (m != 0.0f);
                if (!take) {
                    if (rowmin) rowmin[(size_t)b * N + i] = 0.0f;
                    argmin[(size_t)b * N + i] = -1;
                }
            }
            const unsigned bal = __ballot_sync(MLSP_FULL, take);
            if (lane == 0) wtot[warp] = __popc(bal);
            __syncthreads();
            int before = nlist;
            for (int w = 0; w < warp; ++w) before += wtot[w];
            if (take) list[before + __popc(bal & ((1u << lane) - 1u))] = i;
            __syncthreads();
            if (threadIdx.x == 0) {
                int s = nlist;
                for (int w = 0; w < CH_WARPS; ++w) s += wtot[w];
                nlist = s;
            }
            __syncthreads();
        }
    }
    const int n_rows = nlist;
    float sum = 0.f, cnt = 0.f;
    if (n_rows > 0) {
        for (int j = threadIdx.x; j < N; j += CH_THREADS) {
            const float3 q = p2.get(b, j);
            const float m = mb[j];
            const float pen = (m == 0.0f) ? 100.0f : ((m == 1.0f) ? 0.0f : m);
            cols[j] = make_float4(q.x, q.y, q.z, pen);
        }
        __syncthreads();
        for (int r = warp; r < n_rows; r += CH_WARPS) {
            const int i = list[r];
            const float3 a = p1.get(b, i);
            float best = INFINITY;
            int bj = 0x7fffffff;
            for (int j = lane; j < N; j += 32) {
                const float4 q = cols[j];
                const float dx = __fsub_rn(a.x, q.x), dy = __fsub_rn(a.y, q.y), dz = __fsub_rn(a.z, q.z);
                const float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                const float n = __fsqrt_rn(s);
                const float D = __fadd_rn(__fmul_rn(n, n), q.w);
                // torch.min propagates NaN (a diverged step): the first NaN of a lane sticks, like the first minimum
                if (D < best || (D != D && best == best)) {
                    best = D;
                    bj = j;
                }
            }
            const uint32_t vb = __float_as_uint(best);  // D >= 0: bit pattern orders like the value
            uint32_t wv = __reduce_min_sync(MLSP_FULL, vb);
            int wi = (int)__reduce_min_sync(MLSP_FULL, (vb == wv) ? (uint32_t)bj : 0x7fffffffu);
            const unsigned nan_lanes = __ballot_sync(MLSP_FULL, best != best);
            if (nan_lanes) {                            // non-finite input: NaN row minimum at the first NaN column
                wv = 0x7fc00000u;
                wi = (int)__reduce_min_sync(MLSP_FULL, (best != best) ? (uint32_t)bj : 0x7fffffffu);
            }
            if (wi >= N) wi = 0;                        // every D was +inf: torch.min returns index 0
            if (lane == 0) {
                const float m = mb[i];
                if (rowmin) rowmin[(size_t)b * N + i] = __uint_as_float(wv);
                argmin[(size_t)b * N + i] = wi;
                sum += __uint_as_float(wv) * m;
                cnt += m;
            }
        }
    }
    if (lane == 0) {
        wsum[warp] = sum;
        wcnt[warp] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f, c = 0.f;
        for (int w = 0; w < CH_WARPS; ++w) {
            s += wsum[w];
            c += wcnt[w];
        }
        part_s[b * gridDim.x + t] = s;
        part_c[b * gridDim.x + t] = c;
    }
}

__global__ void __launch_bounds__(CH_THREADS)
chamfer_fwd_kernel(PtView p1, PtView p2, const float *__restrict__ mask, long long mask_bs, int N, int all_rows,
                   int rows_per_cta, float *__restrict__ rowmin, int64_t *__restrict__ argmin,
                   float *__restrict__ part_s, float *__restrict__ part_c)
{
    chamfer_fwd_body(p1, p2, mask, mask_bs, N, all_rows, rows_per_cta, rowmin, argmin, part_s, part_c);
}

// both directions of reconstruction_loss in one launch: blockIdx.z = 0: rows of gold against pred, 1: the converse
__global__ void __launch_bounds__(CH_THREADS)
chamfer_pair_fwd_kernel(PtView pred, PtView gold, const float *__restrict__ mask, long long mask_bs, int B, int N,
                        int rows_per_cta, int64_t *__restrict__ argmin, float *__restrict__ part_s,
                        float *__restrict__ part_c)
{
    const int dir = blockIdx.z;
    const size_t po = (size_t)dir * B * gridDim.x;
    chamfer_fwd_body(dir == 0 ? gold : pred, dir == 0 ? pred : gold, mask, mask_bs, N, 0, rows_per_cta, nullptr,
                     argmin + (size_t)dir * B * N, part_s + po, part_c + po);
}

// loss = (1/B) (sum_b s0_b/c_b + sum_b s1_b/c_b)   (mlsp.py:151-153, :175-180); one warp
__global__ void chamfer_pair_finalize_kernel(const float *__restrict__ part_s, const float *__restrict__ part_c, int B,
                                             int T, float *__restrict__ loss)
{
    float acc[2] = {0.f, 0.f};
    for (int dir = 0; dir < 2; ++dir)
        for (int b = threadIdx.x; b < B; b += 32) {
            float s = 0.f, c = 0.f;
            for (int t = 0; t < T; ++t) {
                s += part_s[((size_t)dir * B + b) * T + t];
                c += part_c[((size_t)dir * B + b) * T + t];
            }
            acc[dir] += s / c;                   // empty mask: 0/0 = NaN, like the reference
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc[0] += __shfl_xor_sync(MLSP_FULL, acc[0], o);
        acc[1] += __shfl_xor_sync(MLSP_FULL, acc[1], o);
    }
    if (threadIdx.x == 0) *loss = (1.0f / (float)B) * (acc[0] + acc[1]);
}

// In all_rows mode unmasked rows contribute m=0 to both sums, so the count is still sum_i mask_i.
__global__ void chamfer_finalize_kernel(const float *__restrict__ part_s, const float *__restrict__ part_c, int B,
                                        int T, float *__restrict__ partial)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float s = 0.f, c = 0.f;
    for (int t = 0; t < T; ++t) {
        s += part_s[b * T + t];
        c += part_c[b * T + t];
    }
    partial[b] = s / c;  // empty mask: 0/0 = NaN, like the reference
}

__device__ __forceinline__ void chamfer_bwd_body(const PtView &p1, const PtView &p2, const float *__restrict__ mask,
                                                 long long mask_bs, const int64_t *__restrict__ argmin, int N,
                                                 const float *__restrict__ scale_dev, float scale_host,
                                                 float *__restrict__ g1, float *__restrict__ g2)
{
    __shared__ float wpart[8];
    __shared__ float total;
    const int b = blockIdx.x;
    const float *mb = mask + b * mask_bs;
    float c = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) c += mb[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(MLSP_FULL, c, o);
    if ((threadIdx.x & 31) == 0) wpart[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += wpart[w];
        total = s;
    }
    __syncthreads();
    const float up = scale_host * (scale_dev ? *scale_dev : 1.0f) / total;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float m = mb[i];
        if (m == 0.0f) continue;
        const long long j = argmin[(size_t)b * N + i];
        if (j < 0 || j >= N) continue;                 // unmasked row / corrupted index: never address outside the cloud
        const float3 a = p1.get(b, i), q = p2.get(b, (int)j);
        const float w = 2.0f * up * m;
        const float gx = (a.x - q.x) * w, gy = (a.y - q.y) * w, gz = (a.z - q.z) * w;
        if (g1) {
            float *d = g1 + ((size_t)b * N + i) * 3;
            atomicAdd(d + 0, gx);
            atomicAdd(d + 1, gy);
            atomicAdd(d + 2, gz);
        }
        if (g2) {
            float *d = g2 + ((size_t)b * N + j) * 3;
            atomicAdd(d + 0, -gx);
            atomicAdd(d + 1, -gy);
            atomicAdd(d + 2, -gz);
        }
    }
}

__global__ void __launch_bounds__(256)
chamfer_bwd_kernel(PtView p1, PtView p2, const float *__restrict__ mask, long long mask_bs,
                   const int64_t *__restrict__ argmin, int N, const float *__restrict__ scale_dev, float scale_host,
                   float *__restrict__ g1, float *__restrict__ g2)
{
    chamfer_bwd_body(p1, p2, mask, mask_bs, argmin, N, scale_dev, scale_host, g1, g2);
}

// gradient of reconstruction_loss w.r.t. pred, both directions (blockIdx.y), accumulated into a zeroed buffer
__global__ void __launch_bounds__(256)
chamfer_pair_bwd_kernel(PtView pred, PtView gold, const float *__restrict__ mask, long long mask_bs,
                        const int64_t *__restrict__ argmin, int B, int N, const float *__restrict__ scale_dev,
                        float scale_host, float *__restrict__ grad_pred)
{
    if (blockIdx.y == 0)      // rows of gold matched into pred: pred is p2
        chamfer_bwd_body(gold, pred, mask, mask_bs, argmin, N, scale_dev, scale_host, nullptr, grad_pred);
    else                      // rows of pred matched into gold: pred is p1
        chamfer_bwd_body(pred, gold, mask, mask_bs, argmin + (size_t)B * N, N, scale_dev, scale_host, grad_pred, nullptr);
}

}  // namespace mlsp

extern "C" int mlsp_reconstruction_loss_fwd(const float *pred, int64_t pred_bstride, int64_t pred_pstride,
                                            int64_t pred_cstride, const float *gold, int64_t gold_bstride,
                                            int64_t gold_pstride, int64_t gold_cstride, const float *mask,
                                            int64_t mask_bstride, int B, int N, int64_t *argmin, float *loss, void *ws,
                                            size_t ws_bytes, void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(pred && gold && mask && argmin && loss && ws, MLSP_EINVAL, "reconstruction_loss_fwd: null pointer");
    MLSP_REQUIRE(B > 0 && N > 0, MLSP_EINVAL, "reconstruction_loss_fwd: bad shape");
    MLSP_REQUIRE(ws_bytes >= chamfer_workspace_bytes(B, N), MLSP_EWORKSPACE, "reconstruction_loss_fwd: workspace too small");
    cudaStream_t st = as_stream(stream);
    int T = (N + 511) / 512;
    if (T > CH_MAX_TILES) T = CH_MAX_TILES;
    const int rows_per_cta = (N + T - 1) / T;
    const size_t smem = sizeof(float4) * (size_t)N + sizeof(int) * (size_t)rows_per_cta;
    MLSP_REQUIRE(smem <= 200 * 1024, MLSP_EUNSUPPORTED, "reconstruction_loss_fwd: N=%d too large", N);
    float *part_s = static_cast<float *>(ws);
    float *part_c = part_s + 2 * (size_t)B * CH_MAX_TILES;
    PtView vp{pred, pred_bstride, pred_pstride, pred_cstride}, vg{gold, gold_bstride, gold_pstride, gold_cstride};
    MLSP_CUDA(cudaFuncSetAttribute(chamfer_pair_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    chamfer_pair_fwd_kernel<<<dim3(T, B, 2), CH_THREADS, smem, st>>>(vp, vg, mask, mask_bstride, B, N, rows_per_cta,
                                                                    argmin, part_s, part_c);
    MLSP_LAUNCH_CHECK("chamfer_pair_fwd_kernel");
    chamfer_pair_finalize_kernel<<<1, 32, 0, st>>>(part_s, part_c, B, T, loss);
    MLSP_LAUNCH_CHECK("chamfer_pair_finalize_kernel");
    return MLSP_OK;
}

extern "C" int mlsp_reconstruction_loss_bwd(const float *pred, int64_t pred_bstride, int64_t pred_pstride,
                                            int64_t pred_cstride, const float *gold, int64_t gold_bstride,
                                            int64_t gold_pstride, int64_t gold_cstride, const float *mask,
                                            int64_t mask_bstride, const int64_t *argmin, int B, int N,
                                            const float *grad_loss, float *grad_pred, void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(pred && gold && mask && argmin && grad_pred, MLSP_EINVAL, "reconstruction_loss_bwd: null pointer");
    MLSP_REQUIRE(B > 0 && N > 0, MLSP_EINVAL, "reconstruction_loss_bwd: bad shape");
    cudaStream_t st = as_stream(stream);
    MLSP_CUDA(cudaMemsetAsync(grad_pred, 0, sizeof(float) * 3 * (size_t)B * N, st));
    PtView vp{pred, pred_bstride, pred_pstride, pred_cstride}, vg{gold, gold_bstride, gold_pstride, gold_cstride};
    chamfer_pair_bwd_kernel<<<dim3(B, 2), 256, 0, st>>>(vp, vg, mask, mask_bstride, argmin, B, N, grad_loss,
                                                       1.0f / (float)B, grad_pred);
    MLSP_LAUNCH_CHECK("chamfer_pair_bwd_kernel");
    return MLSP_OK;
}

extern "C" int mlsp_chamfer_dir_fwd(const float *p1, int64_t p1_bstride, int64_t p1_pstride, int64_t p1_cstride,
                                    const float *p2, int64_t p2_bstride, int64_t p2_pstride, int64_t p2_cstride,
                                    const float *mask, int64_t mask_bstride, int B, int N, int all_rows,
                                    float *rowmin, int64_t *argmin, float *partial, void *ws, size_t ws_bytes,
                                    void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(p1 && p2 && mask && rowmin && argmin && partial && ws, MLSP_EINVAL, "chamfer_dir_fwd: null pointer");
    MLSP_REQUIRE(B > 0 && N > 0, MLSP_EINVAL, "chamfer_dir_fwd: bad shape");
    MLSP_REQUIRE(ws_bytes >= chamfer_workspace_bytes(B, N), MLSP_EWORKSPACE, "chamfer_dir_fwd: workspace too small");
    cudaStream_t st = as_stream(stream);
    // masked mode: few rows per cloud -> few CTAs per cloud; all-rows mode: 32 rows per CTA
    int T = all_rows ? (N + 31) / 32 : (N + 511) / 512;
    if (T > CH_MAX_TILES) T = CH_MAX_TILES;
    const int rows_per_cta = (N + T - 1) / T;
    const size_t smem = sizeof(float4) * (size_t)N + sizeof(int) * (size_t)rows_per_cta;
    MLSP_REQUIRE(smem <= 200 * 1024, MLSP_EUNSUPPORTED, "chamfer_dir_fwd: N=%d too large", N);
    float *part_s = static_cast<float *>(ws);
    float *part_c = part_s + (size_t)B * CH_MAX_TILES;
    PtView v1{p1, p1_bstride, p1_pstride, p1_cstride}, v2{p2, p2_bstride, p2_pstride, p2_cstride};
    MLSP_CUDA(cudaFuncSetAttribute(chamfer_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    chamfer_fwd_kernel<<<dim3(T, B), CH_THREADS, smem, st>>>(v1, v2, mask, mask_bstride, N, all_rows, rows_per_cta,
                                                          rowmin, argmin, part_s, part_c);
    MLSP_LAUNCH_CHECK("chamfer_fwd_kernel");
    chamfer_finalize_kernel<<<(B + 127) / 128, 128, 0, st>>>(part_s, part_c, B, T, partial);
    MLSP_LAUNCH_CHECK("chamfer_finalize_kernel");
    return MLSP_OK;
}

extern "C" int mlsp_chamfer_dir_bwd(const float *p1, int64_t p1_bstride, int64_t p1_pstride, int64_t p1_cstride,
                                    const float *p2, int64_t p2_bstride, int64_t p2_pstride, int64_t p2_cstride,
                                    const float *mask, int64_t mask_bstride, const int64_t *argmin, int B, int N,
                                    const float *scale_dev, float scale_host, float *grad_p1, float *grad_p2,
                                    void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(p1 && p2 && mask && argmin, MLSP_EINVAL, "chamfer_dir_bwd: null pointer");
    MLSP_REQUIRE(B > 0 && N > 0, MLSP_EINVAL, "chamfer_dir_bwd: bad shape");
    if (!grad_p1 && !grad_p2) return MLSP_OK;
    PtView v1{p1, p1_bstride, p1_pstride, p1_cstride}, v2{p2, p2_bstride, p2_pstride, p2_cstride};
    chamfer_bwd_kernel<<<B, 256, 0, as_stream(stream)>>>(v1, v2, mask, mask_bstride, argmin, N, scale_dev, scale_host,
                                                         grad_p1, grad_p2);
    MLSP_LAUNCH_CHECK("chamfer_bwd_kernel");
    return MLSP_OK;
}
