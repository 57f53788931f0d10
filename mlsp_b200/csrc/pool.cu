// pool.cu -- the max-pooling reductions of the DGCNN around the point-wise layers, forward with argmax and backward:
//   x.max(dim=-1)[0] over the k neighbours of a channels-last edge tensor (transform_net, PointDA/model_utils.py:116-118),
//   torch.max(x, dim=2)[0] over the points of a channels-last map (model_utils.py:121) and
//   F.adaptive_max_pool1d(x5, 1) over the points of a channel-major map (PointDA/Models.py:133).
// torch runs the first two as generic strided reductions (0.5-0.9 ms each at 32 x 1024 x 20); they are plain streaming
// passes: "mid" = reduce the middle dimension of (R, K, C) with C contiguous (coalesced float4 over channels, 8 warps
// share the K rows of one output row and combine in shared memory), "row" = reduce the contiguous dimension of (R, K)
// (one warp per row).  Ties: the first index wins; NaN propagates (like torch.max).  HBM-bound: one read of the input.
#include "common.cuh"

namespace mlsp {

constexpr int POOL_WARPS = 8;

__device__ __forceinline__ bool pool_better(float v, float best) { return v > best || (v != v && best == best); }

// in (R,K,C) -> val (R,C), arg (R,C);  grid (C/128 rounded up, R), block 8 warps: lane = float4 of channels, warp = K residue
__global__ void __launch_bounds__(32 * POOL_WARPS)
max_mid_fwd_kernel(const float4 *__restrict__ in, int K, int C4, float4 *__restrict__ val, int4 *__restrict__ arg)
{
    __shared__ float4 sv[POOL_WARPS][32];
    __shared__ int4 si[POOL_WARPS][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = blockIdx.x * 32 + lane;
    const long long r = blockIdx.y;
    float4 b = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    int4 bi = make_int4(0, 0, 0, 0);
    if (q < C4) {
        const float4 *p = in + (r * K) * C4 + q;
#pragma unroll 4
        for (int j = w; j < K; j += POOL_WARPS) {
            const float4 v = __ldcs(p + (long long)j * C4);
            if (pool_better(v.x, b.x) || j == w) { b.x = v.x; bi.x = j; }
            if (pool_better(v.y, b.y) || j == w) { b.y = v.y; bi.y = j; }
            if (pool_better(v.z, b.z) || j == w) { b.z = v.z; bi.z = j; }
            if (pool_better(v.w, b.w) || j == w) { b.w = v.w; bi.w = j; }
        }
    }
    sv[w][lane] = b;
    si[w][lane] = (w < K) ? bi : make_int4(-1, -1, -1, -1);
    __syncthreads();
    if (w == 0 && q < C4) {
#pragma unroll
        for (int o = 1; o < POOL_WARPS; ++o) {            // ascending residues: on ties the smaller index stays
            const float4 v = sv[o][lane];
            const int4 vi = si[o][lane];
            if (vi.x >= 0 && (pool_better(v.x, b.x) || (v.x == b.x && vi.x < bi.x))) { b.x = v.x; bi.x = vi.x; }
            if (vi.y >= 0 && (pool_better(v.y, b.y) || (v.y == b.y && vi.y < bi.y))) { b.y = v.y; bi.y = vi.y; }
            if (vi.z >= 0 && (pool_better(v.z, b.z) || (v.z == b.z && vi.z < bi.z))) { b.z = v.z; bi.z = vi.z; }
            if (vi.w >= 0 && (pool_better(v.w, b.w) || (v.w == b.w && vi.w < bi.w))) { b.w = v.w; bi.w = vi.w; }
        }
        val[r * C4 + q] = b;
        arg[r * C4 + q] = bi;
    }
}

// g (R,C), arg (R,C) -> gin (R,K,C): the gradient goes to the selected row, zeros elsewhere (one streaming write)
__global__ void __launch_bounds__(256)
max_mid_bwd_kernel(const float4 *__restrict__ g, const int4 *__restrict__ arg, int K, int C4, float4 *__restrict__ gin, long long total)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // (r, j, q)
    if (t >= total) return;
    const int q = (int)(t % C4);
    const long long rj = t / C4;
    const int j = (int)(rj % K);
    const long long r = rj / K;
    const float4 gv = g[r * C4 + q];
    const int4 a = arg[r * C4 + q];
    __stcs(gin + t, make_float4(a.x == j ? gv.x : 0.f, a.y == j ? gv.y : 0.f, a.z == j ? gv.z : 0.f, a.w == j ? gv.w : 0.f));
}

// in (R,K) rows contiguous -> val (R), arg (R): one warp per row
__global__ void __launch_bounds__(256)
max_row_fwd_kernel(const float *__restrict__ in, int K, float *__restrict__ val, int *__restrict__ arg, long long R)
{
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= R) return;
    const float *p = in + r * K;
    float b = -INFINITY;
    int bi = 0x7fffffff;
    if ((K & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const float4 *p4 = reinterpret_cast<const float4 *>(p);
        for (int j = lane; j < K / 4; j += 32) {
            const float4 v = __ldcs(p4 + j);
            if (pool_better(v.x, b) || bi == 0x7fffffff) { b = v.x; bi = 4 * j; }
            if (pool_better(v.y, b)) { b = v.y; bi = 4 * j + 1; }
            if (pool_better(v.z, b)) { b = v.z; bi = 4 * j + 2; }
            if (pool_better(v.w, b)) { b = v.w; bi = 4 * j + 3; }
        }
    } else {
        for (int j = lane; j < K; j += 32) {
            const float v = p[j];
            if (pool_better(v, b) || bi == 0x7fffffff) { b = v; bi = j; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(MLSP_FULL, b, o);
        const int oi = __shfl_xor_sync(MLSP_FULL, bi, o);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || pool_better(ov, b) || (ov == b && oi < bi) || (ov != ov && b != b && oi < bi))) {
            b = ov;
            bi = oi;
        }
    }
    if (lane == 0) {
        val[r] = b;
        arg[r] = bi;
    }
}

// g (R), arg (R) -> gin (R,K)
__global__ void __launch_bounds__(256)
max_row_bwd_kernel(const float *__restrict__ g, const int *__restrict__ arg, int K, float *__restrict__ gin, long long R)
{
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= R) return;
    const float gv = g[r];
    const int a = arg[r];
    float *p = gin + r * K;
    if ((K & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        float4 *p4 = reinterpret_cast<float4 *>(p);
        for (int j = lane; j < K / 4; j += 32) {
            const int j0 = 4 * j;
            __stcs(p4 + j, make_float4(a == j0 ? gv : 0.f, a == j0 + 1 ? gv : 0.f, a == j0 + 2 ? gv : 0.f, a == j0 + 3 ? gv : 0.f));
        }
    } else {
        for (int j = lane; j < K; j += 32) p[j] = (a == j) ? gv : 0.f;
    }
}

}  // namespace mlsp

using namespace mlsp;

extern "C" {

int mlsp_max_mid_fwd(const float *in, long long R, int K, int C, float *val, int *arg, void *stream)
{
    MLSP_REQUIRE(in && val && arg, MLSP_EINVAL, "max_mid_fwd: null pointer");
    MLSP_REQUIRE(R > 0 && K > 0 && C > 0 && C % 4 == 0 && R < 65536ll * 32768, MLSP_EINVAL, "max_mid_fwd: bad shape R=%lld K=%d C=%d (C %% 4 == 0)", R, K, C);
    MLSP_REQUIRE(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(val) | reinterpret_cast<uintptr_t>(arg)) & 15) == 0, MLSP_EINVAL,
                 "max_mid_fwd: 16-byte alignment");
    MLSP_REQUIRE(R <= 0x7fffffffll, MLSP_EUNSUPPORTED, "max_mid_fwd: too many rows");
    const int C4 = C / 4;
    for (long long r0 = 0; r0 < R; r0 += 65535) {              // grid.y limit: slabs of 65535 output rows
        const long long rows = R - r0 < 65535 ? R - r0 : 65535;
        dim3 g2((unsigned)((C4 + 31) / 32), (unsigned)rows, 1);
        max_mid_fwd_kernel<<<g2, 32 * POOL_WARPS, 0, as_stream(stream)>>>(reinterpret_cast<const float4 *>(in) + r0 * K * C4, K, C4,
                                                                         reinterpret_cast<float4 *>(val) + r0 * C4,
                                                                         reinterpret_cast<int4 *>(arg) + r0 * C4);
    }
    MLSP_LAUNCH_CHECK("max_mid_fwd_kernel");
    return MLSP_OK;
}

int mlsp_max_mid_bwd(const float *g, const int *arg, long long R, int K, int C, float *gin, void *stream)
{
    MLSP_REQUIRE(g && arg && gin, MLSP_EINVAL, "max_mid_bwd: null pointer");
    MLSP_REQUIRE(R > 0 && K > 0 && C > 0 && C % 4 == 0, MLSP_EINVAL, "max_mid_bwd: bad shape");
    MLSP_REQUIRE(((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(gin) | reinterpret_cast<uintptr_t>(arg)) & 15) == 0, MLSP_EINVAL,
                 "max_mid_bwd: 16-byte alignment");
    const long long total = R * K * (C / 4);
    MLSP_REQUIRE((total + 255) / 256 < (1ll << 31), MLSP_EUNSUPPORTED, "max_mid_bwd: too many elements");
    max_mid_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4 *>(g),
                                                                                       reinterpret_cast<const int4 *>(arg), K, C / 4,
                                                                                       reinterpret_cast<float4 *>(gin), total);
    MLSP_LAUNCH_CHECK("max_mid_bwd_kernel");
    return MLSP_OK;
}

int mlsp_max_row_fwd(const float *in, long long R, int K, float *val, int *arg, void *stream)
{
    MLSP_REQUIRE(in && val && arg, MLSP_EINVAL, "max_row_fwd: null pointer");
    MLSP_REQUIRE(R > 0 && K > 0 && (R + 7) / 8 < (1ll << 31), MLSP_EINVAL, "max_row_fwd: bad shape");
    max_row_fwd_kernel<<<(unsigned)((R + 7) / 8), 256, 0, as_stream(stream)>>>(in, K, val, arg, R);
    MLSP_LAUNCH_CHECK("max_row_fwd_kernel");
    return MLSP_OK;
}

int mlsp_max_row_bwd(const float *g, const int *arg, long long R, int K, float *gin, void *stream)
{
    MLSP_REQUIRE(g && arg && gin, MLSP_EINVAL, "max_row_bwd: null pointer");
    MLSP_REQUIRE(R > 0 && K > 0 && (R + 7) / 8 < (1ll << 31), MLSP_EINVAL, "max_row_bwd: bad shape");
    max_row_bwd_kernel<<<(unsigned)((R + 7) / 8), 256, 0, as_stream(stream)>>>(g, arg, K, gin, R);
    MLSP_LAUNCH_CHECK("max_row_bwd_kernel");
    return MLSP_OK;
}
}
