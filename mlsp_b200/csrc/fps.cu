// fps.cu -- a3: farthest_point_sample(args, xyz, npoint)  (utils/pc_utils.py:137-161)
//
// The reference runs npoint dependent iterations of ~8 tiny ATen kernels.  Here one CTA owns one cloud
// for the whole loop: point coordinates and running min-distances live in registers, the cloud is also
// kept SoA in shared memory so the winner's coordinates are one broadcast read, and each round costs a
// single __syncthreads: per-warp argmax with two redux.sync instructions (distances are >= +0, so the
// float bit pattern orders as an unsigned integer), one shared-memory slot per warp (double buffered
// by round parity), and every warp redundantly reduces the slots.
// Arithmetic pinned to oracle/mlsp_oracle.c:orc_fps: d = (rn(dx^2) + rn(dy^2)) + rn(dz^2), distance =
// min(distance, d) starting from 1e10, argmax with the lowest index on ties (torch.max semantics).
#include "common.cuh"

namespace mlsp {

template <int PPT, int THREADS>
__global__ void __launch_bounds__(THREADS)
fps_kernel(const float *__restrict__ xyz, int N, int npoint, const int64_t *__restrict__ start,
           int64_t *__restrict__ centroids, float *__restrict__ vals)
{
    extern __shared__ float smem[];
    float *sx = smem, *sy = smem + N, *sz = smem + 2 * N;
    __shared__ uint32_t slot_v[2][32];
    __shared__ uint32_t slot_i[2][32];
    constexpr int WARPS = THREADS / 32;

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *X = xyz + (size_t)b * 3 * N;
    float px[PPT], py[PPT], pz[PPT], dist[PPT];
#pragma unroll
    for (int r = 0; r < PPT; ++r) {
        const int p = r * THREADS + tid;
        const bool ok = p < N;
        px[r] = ok ? X[p] : 0.0f;
        py[r] = ok ? X[N + p] : 0.0f;
        pz[r] = ok ? X[2 * N + p] : 0.0f;
        dist[r] = 1e10f;
        if (ok) {
            sx[p] = px[r];
            sy[p] = py[r];
            sz[p] = pz[r];
        }
    }
    int far = (int)start[b];
    __syncthreads();

    for (int s = 0; s < npoint; ++s) {
        const float cx = sx[far], cy = sy[far], cz = sz[far];
        if (tid == 0) {
            centroids[(size_t)b * npoint + s] = far;
            vals[((size_t)b * 3 + 0) * npoint + s] = cx;
            vals[((size_t)b * 3 + 1) * npoint + s] = cy;
            vals[((size_t)b * 3 + 2) * npoint + s] = cz;
        }
        if (s + 1 == npoint) break;
        float best = 0.0f;
        uint32_t bi = 0xffffffffu;
#pragma unroll
        for (int r = 0; r < PPT; ++r) {
            const int p = r * THREADS + tid;
            const float dx = __fsub_rn(px[r], cx), dy = __fsub_rn(py[r], cy), dz = __fsub_rn(pz[r], cz);
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            dist[r] = (d < dist[r]) ? d : dist[r];
            const bool ok = p < N;
            // strict > keeps the lowest index inside the thread (p grows with r); first valid point seeds it
            if (ok && (bi == 0xffffffffu || dist[r] > best)) {
                best = dist[r];
                bi = (uint32_t)p;
            }
        }
        const uint32_t vb = __float_as_uint(best);
        const uint32_t wv = __reduce_max_sync(MLSP_FULL, vb);
        const uint32_t wi = __reduce_min_sync(MLSP_FULL, (vb == wv) ? bi : 0xffffffffu);
        const int par = s & 1;
        if (lane == 0) {
            slot_v[par][warp] = wv;
            slot_i[par][warp] = wi;
        }
        __syncthreads();
        const uint32_t v2 = (lane < WARPS) ? slot_v[par][lane] : 0u;
        const uint32_t i2 = (lane < WARPS) ? slot_i[par][lane] : 0xffffffffu;
        const uint32_t gv = __reduce_max_sync(MLSP_FULL, v2);
        far = (int)__reduce_min_sync(MLSP_FULL, (v2 == gv) ? i2 : 0xffffffffu);
    }
}

template <int PPT, int THREADS>
static int launch_fps(const float *xyz, int B, int N, int npoint, const int64_t *start, int64_t *centroids,
                      float *vals, cudaStream_t st)
{
    const size_t smem = sizeof(float) * 3 * (size_t)N;
    MLSP_CUDA(cudaFuncSetAttribute(fps_kernel<PPT, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fps_kernel<PPT, THREADS><<<B, THREADS, smem, st>>>(xyz, N, npoint, start, centroids, vals);
    MLSP_LAUNCH_CHECK("fps_kernel");
    return MLSP_OK;
}

}  // namespace mlsp

extern "C" int mlsp_fps(const float *xyz, int B, int N, int npoint, const int64_t *start, int64_t *centroids,
                        float *vals, void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(xyz && start && centroids && vals, MLSP_EINVAL, "fps: null pointer");
    MLSP_REQUIRE(B > 0 && N > 0 && npoint >= 0, MLSP_EINVAL, "fps: bad shape");
    if (npoint == 0) return MLSP_OK;
    cudaStream_t st = as_stream(stream);
    // a valid point with the lowest index wins all-zero rounds; invalid start indices are the caller's bug
    if (N <= 256) return launch_fps<1, 256>(xyz, B, N, npoint, start, centroids, vals, st);
    if (N <= 512) return launch_fps<2, 256>(xyz, B, N, npoint, start, centroids, vals, st);
    if (N <= 1024) return launch_fps<4, 256>(xyz, B, N, npoint, start, centroids, vals, st);
    if (N <= 2048) return launch_fps<8, 256>(xyz, B, N, npoint, start, centroids, vals, st);
    if (N <= 4096) return launch_fps<8, 512>(xyz, B, N, npoint, start, centroids, vals, st);
    if (N <= 8192) return launch_fps<16, 512>(xyz, B, N, npoint, start, centroids, vals, st);
    if (N <= 16384) return launch_fps<16, 1024>(xyz, B, N, npoint, start, centroids, vals, st);
    set_error("fps: N=%d > 16384 not supported", N);
    return MLSP_EUNSUPPORTED;
}
