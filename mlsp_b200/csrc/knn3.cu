// knn3.cu -- a1 for 3-D clouds (C = 3: the first two DGCNN layers and the PCA-normal neighbourhoods).
//
// Same specification as knn.cu (oracle orc_knn with C <= 4: one fmaf chain from +0, pd = rn(rn(2 dot - xx_j) - xx_i),
// rank pd descending / index ascending), different selection: with K = 3 the distance costs 5 instructions, so the
// per-candidate insertion of topk.cuh (about 20 instructions, ~100 insertions per row) dominated.  Here a warp owns
// R rows and makes two passes over the cloud (staged once per CTA in shared memory as (x,y,z,|x|^2) float4):
//   pass 1: every lane keeps, per row, the best pd among ITS candidates (j = lane mod 32: a "class" maximum).
//           The class maxima are 32*NC distinct candidates, so tau = their k-th largest value is a lower bound of
//           the k-th best pd of the row.  (warp bitonic sort of orderable 32-bit keys, one shuffle per stage)
//   pass 2: pd is recomputed and every candidate with pd >= tau (ties included -> superset of the exact top-k,
//           about 1.5 k entries expected) is appended to the row's list in shared memory by ballot compaction.
//   final : the list is sorted by (pd desc, index asc) with 64-bit keys; the first k are the answer.
// A row whose list overflows (heavy duplicates / degenerate clouds) is re-done with the streaming selection.
#include "common.cuh"
#include "topk.cuh"

namespace mlsp {

constexpr int K3_THREADS = 256;
constexpr int K3_R = 4;                                   // rows per warp
constexpr int K3_ROWS = (K3_THREADS / 32) * K3_R;         // 32 rows per CTA

__device__ __forceinline__ float pd3(float4 a, float4 q)
{
    float d = __fmaf_rn(a.x, q.x, 0.0f);
    d = __fmaf_rn(a.y, q.y, d);
    d = __fmaf_rn(a.z, q.z, d);
    return __fsub_rn(__fmaf_rn(2.0f, d, -q.w), a.w);
}

// descending bitonic sort of 32*NC 32-bit keys across the warp (key[s] on lane l <-> element s*32+l)
template <int NC>
__device__ __forceinline__ void warp_sort_desc_u32(uint32_t (&key)[NC])
{
    const int lane = lane_id();
#pragma unroll
    for (int size = 2; size <= 32 * NC; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                const int ds = stride / 32;
#pragma unroll
                for (int s = 0; s < NC; ++s)
                    if ((s & ds) == 0) {
                        const bool down = ((s * 32) & size) == 0;      // "down" block: larger first
                        const uint32_t a = key[s], b = key[s | ds];
                        const bool a_big = a > b;
                        key[s] = (a_big == down) ? a : b;
                        key[s | ds] = (a_big == down) ? b : a;
                    }
            } else {
#pragma unroll
                for (int s = 0; s < NC; ++s) {
                    const uint32_t other = __shfl_xor_sync(MLSP_FULL, key[s], stride);
                    const int e = s * 32 + lane;
                    const bool down = (e & size) == 0;
                    const bool lower = (lane & stride) == 0;
                    const bool take_max = (lower == down);
                    key[s] = take_max ? max(key[s], other) : min(key[s], other);
                }
            }
        }
    }
}

// NC = classes per lane (1: k <= 32, 2: k <= 64);  list capacity CAPL = 64 * NC
template <int NC>
__global__ void __launch_bounds__(K3_THREADS)
knn3_kernel(const float *__restrict__ x, int N, int k, int64_t *__restrict__ idx)
{
    constexpr int CAPL = 64 * NC;
    constexpr int SL = CAPL / 32;
    extern __shared__ float4 cloud[];                                  // [N]
    uint2 *lists = reinterpret_cast<uint2 *>(cloud + N);               // [warps][R][CAPL] (pd bits, j)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const float *xb = x + (size_t)b * 3 * N;
    for (int n = tid; n < N; n += K3_THREADS) {
        const float px = xb[n], py = xb[N + n], pz = xb[2 * N + n];
        const float xx = __fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz));
        cloud[n] = make_float4(px, py, pz, xx);
    }
    __syncthreads();

    const int i0 = blockIdx.x * K3_ROWS + warp * K3_R;
    if (i0 >= N) return;
    float4 xi[K3_R];
#pragma unroll
    for (int rr = 0; rr < K3_R; ++rr) xi[rr] = cloud[min(i0 + rr, N - 1)];

    // ---- pass 1: class maxima
    float cmax[K3_R][NC];
#pragma unroll
    for (int rr = 0; rr < K3_R; ++rr)
#pragma unroll
        for (int c = 0; c < NC; ++c) cmax[rr][c] = -INFINITY;
    for (int j0 = 0; j0 < N; j0 += 32 * NC) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int j = j0 + c * 32 + lane;
            if (j < N) {
                const float4 q = cloud[j];
#pragma unroll
                for (int rr = 0; rr < K3_R; ++rr) cmax[rr][c] = fmaxf(cmax[rr][c], pd3(xi[rr], q));
            }
        }
    }
    // ---- tau = k-th largest class maximum (orderable keys; -inf classes sort last)
    float tau[K3_R];
#pragma unroll
    for (int rr = 0; rr < K3_R; ++rr) {
        uint32_t key[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) key[c] = f32_orderable(__fadd_rn(cmax[rr][c], 0.0f));
        warp_sort_desc_u32<NC>(key);
        uint32_t kth = 0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const uint32_t v = __shfl_sync(MLSP_FULL, key[c], (k - 1) & 31);
            if (c == (k - 1) / 32) kth = v;
        }
        // invert f32_orderable
        const uint32_t u = (kth & 0x80000000u) ? (kth ^ 0x80000000u) : ~kth;
        tau[rr] = __uint_as_float(u);
    }

    // ---- pass 2: collect every candidate with pd >= tau
    uint2 *my = lists + (size_t)warp * K3_R * CAPL;
    int cnt[K3_R];
#pragma unroll
    for (int rr = 0; rr < K3_R; ++rr) cnt[rr] = 0;
    for (int j0 = 0; j0 < N; j0 += 32) {
        const int j = j0 + lane;
        const float4 q = cloud[min(j, N - 1)];
#pragma unroll
        for (int rr = 0; rr < K3_R; ++rr) {
            const float pd = pd3(xi[rr], q);
            const bool pass = (j < N) && (pd >= tau[rr]);
            const unsigned m = __ballot_sync(MLSP_FULL, pass);
            if (m) {
                const int pos = cnt[rr] + __popc(m & ((1u << lane) - 1u));
                if (pass && pos < CAPL) my[rr * CAPL + pos] = make_uint2(__float_as_uint(pd), (uint32_t)j);
                cnt[rr] += __popc(m);
            }
        }
    }
    __syncwarp();

    // ---- final: exact sort of the list (or streaming selection if it overflowed)
#pragma unroll
    for (int rr = 0; rr < K3_R; ++rr) {
        const int i = i0 + rr;
        if (i >= N) break;
        int64_t *out = idx + ((size_t)b * N + i) * k;
        if (cnt[rr] <= CAPL) {
            unsigned long long key[SL];
#pragma unroll
            for (int s = 0; s < SL; ++s) {
                const int e = s * 32 + lane;
                const bool live = e < cnt[rr];
                const uint2 v = live ? my[rr * CAPL + e] : make_uint2(0u, 0x7fffffffu);
                key[s] = rank_key(__uint_as_float(v.x), (int)v.y, live);
            }
            warp_sort_u64<SL>(key);
#pragma unroll
            for (int s = 0; s < SL; ++s) {
                const int e = s * 32 + lane;
                if (e < k) out[e] = (int64_t)(uint32_t)(key[s] & 0xffffffffull);
            }
        } else {
            TopK<NC> top;
            top.init(k);
            for (int j0 = 0; j0 < N; j0 += 32) {
                const int j = j0 + lane;
                const float pd = pd3(xi[rr], cloud[min(j, N - 1)]);
                top.offer(pd, j, j < N);
            }
            top.finish(k);
#pragma unroll
            for (int s = 0; s < NC; ++s) {
                const int e = s * 32 + lane;
                if (e < k) out[e] = (int64_t)top.j[s];
            }
        }
    }
}

bool knn3_supported(int C, int N, int k) { return C == 3 && k <= 64 && N >= 1 && N <= 8192; }

int knn3_run(const float *x, int B, int N, int k, int64_t *idx, cudaStream_t st)
{
    const int NC = (k <= 32) ? 1 : 2;
    const size_t smem = sizeof(float4) * (size_t)N + sizeof(uint2) * (size_t)(K3_THREADS / 32) * K3_R * 64 * NC;
    dim3 grid((N + K3_ROWS - 1) / K3_ROWS, B);
    if (NC == 1) {
        MLSP_CUDA(cudaFuncSetAttribute(knn3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn3_kernel<1><<<grid, K3_THREADS, smem, st>>>(x, N, k, idx);
    } else {
        MLSP_CUDA(cudaFuncSetAttribute(knn3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn3_kernel<2><<<grid, K3_THREADS, smem, st>>>(x, N, k, idx);
    }
    MLSP_LAUNCH_CHECK("knn3_kernel");
    return MLSP_OK;
}

}  // namespace mlsp
