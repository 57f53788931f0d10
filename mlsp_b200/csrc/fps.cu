// fps.cu -- a3: farthest_point_sample(args, xyz, npoint)  (utils/pc_utils.py:137-161)
//
// The reference runs npoint dependent iterations of ~8 tiny ATen kernels.  Here one CTA owns one cloud
// for the whole loop: point coordinates and running min-distances live in registers, the cloud is also
// kept as float4 in shared memory so the winner's coordinates are one broadcast LDS.128, and each round
// costs one redux.sync, one ballot and one __syncthreads (see fps_kernel).  The loop is a serial
// dependency chain of npoint rounds; the kernel is latency-bound by construction (12 KB of input).
// Arithmetic pinned to oracle/mlsp_oracle.c:orc_fps: d = (rn(dx^2) + rn(dy^2)) + rn(dz^2), distance =
// min(distance, d) starting from 1e10, argmax with the lowest index on ties (torch.max semantics).
#include "common.cuh"

namespace mlsp {

// W warps per cloud, P points per lane.  Lane l of warp w owns the CONTIGUOUS points (w*32 + l)*P .. +P-1, so
// "lowest (warp, lane, register)" is "lowest index": ties resolve without reducing indices --
//   in-thread : compare tree over the P registers (strict > towards the higher index)
//   in-warp   : one redux.sync.max.f32, one ballot; the first lane holding the maximum is the leader and
//               publishes (value, index) for its warp
//   in-CTA    : one __syncthreads per round; every thread reads the W slots and picks the first maximum.
// Points beyond N carry distance -1 forever (d >= 0 never undercuts it, and it never wins a maximum).
// PCM mix-up (MLSP/PCM.py:26-38) fused with its two FPS calls: with mix.out set the grid is 2B CTAs -- CTA b < B samples
// npoint_a points of cloud b, CTA B + b samples N - npoint_a points of cloud index[b] -- and every sampled point goes straight
// to its place in the mixed cloud: slot s of cat(vals_a, vals_b) lands at out[b][:, inv_perm[s]], which is
// cat(...)[:, :, points_perm] of PCM.py:31-33 without the two value tensors, the cat and the permuting gather.
struct FpsMix {
    const int64_t *index;    // (B) the batch permutation of PCM.py:20
    const int32_t *inv_perm; // (N) inverse of points_perm (PCM.py:32)
    float *out;              // (B,3,N) mixed clouds, or NULL: plain farthest_point_sample
    int npoint_a, B;
};

//
// G clouds per CTA (G groups of W warps, each with its own named barrier and its own slice of shared memory): the loop is a
// latency chain that leaves most of an SM's issue slots idle, and a resident FPS CTA slows whatever shares its SM for the
// whole 100-300 us it lives (measured: one 32-CTA FPS call beside the DGCNN layers' kernels costs them +0.08 ms).  Packing
// G chains onto one SM disturbs B/G SMs instead of B -- an experiment kept behind mlsp_fps_set_groups: it did not pay (see
// launch_fps), the default stays one cloud per CTA.
template <int W>
__device__ __forceinline__ void group_barrier(int group)
{
    if (W == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(32 * W) : "memory");
}

template <int P, int W, int G>
__global__ void __launch_bounds__(32 * W * G)
fps_kernel(const float *__restrict__ xyz, int N, int npoint, const int64_t *__restrict__ start,
           int64_t *__restrict__ centroids, float *__restrict__ vals, FpsMix mix, int clouds, int np_max)
{
    extern __shared__ float4 spt_all[];                   // per group: [N] (x, y, z, 0), then the winners of all rounds [np_max]
    __shared__ __align__(16 * W) uint2 slot_all[G][2][W]; // (value bits, index), double buffered by round parity

    const int group = (G == 1) ? 0 : (int)threadIdx.x / (32 * W);
    const int cta_cloud = blockIdx.x * G + group;         // index into `start` (and the cloud, before the mix mapping)
    if (cta_cloud >= clouds) return;                      // whole groups leave: the named barriers are per group
    float4 *spt = spt_all + (size_t)group * (N + (np_max + 3) / 4);
    int *hist = reinterpret_cast<int *>(spt + N);         // [npoint]
    uint2 (*slot)[W] = slot_all[group];
    int b = cta_cloud;
    const int tid = (int)threadIdx.x - group * 32 * W, lane = tid & 31, warp = tid >> 5;
    int slot0 = 0, out_b = b;                             // mix: first slot of this CTA's samples, destination cloud
    if (mix.out) {
        const bool second = b >= mix.B;
        out_b = second ? b - mix.B : b;
        npoint = second ? N - mix.npoint_a : mix.npoint_a;
        slot0 = second ? mix.npoint_a : 0;
        b = second ? (int)mix.index[out_b] : b;
        if (npoint <= 0) return;
    }
    const float *X = xyz + (size_t)b * 3 * N;
    for (int p = tid; p < N; p += 32 * W) spt[p] = make_float4(X[p], X[N + p], X[2 * N + p], 0.0f);
    group_barrier<W>(group);
    const int p0 = tid * P;
    float px[P], py[P], pz[P], dist[P];
#pragma unroll
    for (int r = 0; r < P; ++r) {
        const bool ok = p0 + r < N;
        const float4 q = spt[ok ? p0 + r : 0];
        px[r] = q.x;
        py[r] = q.y;
        pz[r] = q.z;
        dist[r] = ok ? 1e10f : -1.0f;
    }
    // shared-window addresses computed once (the compiler otherwise rebuilds them every round)
    const uint32_t spt_a = (uint32_t)__cvta_generic_to_shared(spt);
    const uint32_t slot_a = (uint32_t)__cvta_generic_to_shared(&slot[0][0]);
    uint32_t my_slot = slot_a + warp * 8, all_slots = slot_a;
    unsigned lt_mask;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
    int far = (int)start[cta_cloud];

    for (int s = 0; s + 1 < npoint; ++s) {
        float cx, cy, cz, cw;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(cx), "=f"(cy), "=f"(cz), "=f"(cw) : "r"(spt_a + far * 16));
        if (tid == 0) hist[s] = far;
        float bv[P];
        int br[P];
#pragma unroll
        for (int r = 0; r < P; ++r) {
            const float dx = __fsub_rn(px[r], cx), dy = __fsub_rn(py[r], cy), dz = __fsub_rn(pz[r], cz);
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            dist[r] = (d < dist[r]) ? d : dist[r];
            bv[r] = dist[r];
            br[r] = r;
        }
#pragma unroll
        for (int w = 1; w < P; w <<= 1)
#pragma unroll
            for (int r = 0; r + w < P; r += 2 * w) {
                const bool hi = bv[r + w] > bv[r];       // strict: the lower index keeps ties
                bv[r] = hi ? bv[r + w] : bv[r];
                br[r] = hi ? br[r + w] : br[r];
            }
        const float wv = warp_max_f32(bv[0]);
        const bool mine = bv[0] == wv;
        const unsigned holders = __ballot_sync(MLSP_FULL, mine);
        // the first lane holding the maximum publishes it (predicated store, no branch)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.u32 p, %0, 0;\n\t"
            "@p st.shared.v2.u32 [%1], {%2, %3};\n\t}" ::"r"((unsigned)(mine && (holders & lt_mask) == 0u)),
            "r"(my_slot), "r"(__float_as_uint(wv)), "r"((uint32_t)(p0 + br[0]))
            : "memory");
        group_barrier<W>(group);
        uint32_t sv[W], si[W];
        if (W == 1) {
            asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(sv[0]), "=r"(si[0]) : "r"(all_slots));
        } else {
#pragma unroll
            for (int w = 0; w < W; w += 2)
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(sv[w]), "=r"(si[w]), "=r"(sv[w + 1]), "=r"(si[w + 1])
                             : "r"(all_slots + w * 8));
        }
#pragma unroll
        for (int w = 1; w < W; w <<= 1)
#pragma unroll
            for (int r = 0; r + w < W; r += 2 * w) {
                const bool hi = __uint_as_float(sv[r + w]) > __uint_as_float(sv[r]);   // lower warp keeps ties
                sv[r] = hi ? sv[r + w] : sv[r];
                si[r] = hi ? si[r + w] : si[r];
            }
        far = (int)si[0];
        my_slot ^= W * 8;                                // other parity buffer (slot is aligned to its size)
        all_slots ^= W * 8;
    }
    if (tid == 0 && npoint > 0) hist[npoint - 1] = far;
    group_barrier<W>(group);
    if (mix.out) {                                        // straight into the mixed cloud (scattered by the point permutation)
        float *o = mix.out + (size_t)out_b * 3 * N;
        for (int s = tid; s < npoint; s += 32 * W) {
            const float4 c = spt[hist[s]];
            const int p = mix.inv_perm[slot0 + s];
            o[p] = c.x;
            o[N + p] = c.y;
            o[2 * N + p] = c.z;
        }
        return;
    }
    // ---- results, written once and coalesced: centroids (B,npoint) int64, centroids_vals (B,3,npoint)
    int64_t *cen = centroids + (size_t)b * npoint;
    float *vx = vals + (size_t)b * 3 * npoint;
    for (int s = tid; s < npoint; s += 32 * W) {
        const int f = hist[s];
        const float4 c = spt[f];
        cen[s] = f;
        vx[s] = c.x;
        vx[npoint + s] = c.y;
        vx[2 * npoint + s] = c.z;
    }
}

// tuning hook (mlsp_fps_set_groups): 0 = automatic, else the clouds-per-CTA count to use where it fits
static int g_fps_groups = 0;
// tuning hook: 1 = every FPS CTA asks for the whole shared memory of its SM, so that no kernel with a shared-memory footprint
// shares the SM with it (the chain is latency-bound: co-resident warps stretch every round)
// -1 = automatic: on for N > 1024 (16 points per thread: a round is long and every co-resident warp stretches it -- at
// 16 x 2048 the call takes 0.28 ms alone and 0.9 ms beside the DGCNN layers' kernels, and the step waits for it: measured
// 0.87 -> 0.76 ms per step with the SM reserved; at 32 x 1024 the call is off the critical path and sharing is better)
static int g_fps_exclusive = -1;

template <int P, int W, int G>
static int launch_fps_g(const float *xyz, int B, int N, int npoint, const int64_t *start, int64_t *centroids,
                        float *vals, cudaStream_t st, FpsMix mix)
{
    const int np_max = mix.out ? N : npoint;
    const int clouds = mix.out ? 2 * B : B;
    const size_t smem = G * (sizeof(float4) * (size_t)N + sizeof(int) * (size_t)((np_max + 3) / 4 * 4));
    MLSP_REQUIRE(smem <= 227 * 1024, MLSP_EUNSUPPORTED, "fps: N=%d, npoint=%d needs %zu bytes of shared memory", N, np_max, smem);
    const bool exclusive = g_fps_exclusive < 0 ? N > 1024 : g_fps_exclusive != 0;
    const size_t ask = exclusive ? (size_t)(227 * 1024 - 1024) : smem;
    MLSP_CUDA(cudaFuncSetAttribute(fps_kernel<P, W, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ask));
    fps_kernel<P, W, G><<<(clouds + G - 1) / G, 32 * W * G, ask, st>>>(xyz, N, npoint, start, centroids, vals, mix, clouds, np_max);
    MLSP_LAUNCH_CHECK("fps_kernel");
    return MLSP_OK;
}

template <int P, int W>
static int launch_fps(const float *xyz, int B, int N, int npoint, const int64_t *start, int64_t *centroids,
                      float *vals, cudaStream_t st, FpsMix mix = FpsMix{nullptr, nullptr, nullptr, 0, 0})
{
    // clouds per CTA: up to 1024 threads and the shared memory of one SM; small batches stay one cloud per CTA
    const int clouds = mix.out ? 2 * B : B;
    const size_t per = sizeof(float4) * (size_t)N + sizeof(int) * (size_t)(mix.out ? N : npoint) + 16;
    // automatic = 1: measured on B200 (tools/fps_victims.py, tools/step_trace.py conc), packing makes the step SLOWER -- alone a
    // call takes 0.17 (1) / 0.21 (2) / 0.31 ms (4 clouds per CTA) at 32 x 1024, and kernels that share the saturated SMs wait for
    // their slowest CTA (kNN filter path x1.8, scatter x1.65 beside a 4-per-CTA call, x1.3 / x1.09 beside 1-per-CTA)
    // ... except with the SM reserved (N > 1024), where two chains per SM interleave well (0.78 -> 0.76 ms per step at 16 x 2048)
    const bool exclusive = g_fps_exclusive < 0 ? N > 1024 : g_fps_exclusive != 0;
    int G = g_fps_groups > 0 ? g_fps_groups : (exclusive && clouds >= 16 ? 2 : 1);
    constexpr int MAXT = (P > 4) ? 512 : 1024;          // P = 16 keeps 64 coordinates / distances in registers: <= 128 registers per thread
    while (G > 1 && (32 * W * G > MAXT || per * G > 200 * 1024)) G >>= 1;
    if constexpr (32 * W * 4 <= MAXT) {
        if (G >= 4) return launch_fps_g<P, W, 4>(xyz, B, N, npoint, start, centroids, vals, st, mix);
    }
    if constexpr (32 * W * 2 <= MAXT) {
        if (G >= 2) return launch_fps_g<P, W, 2>(xyz, B, N, npoint, start, centroids, vals, st, mix);
    }
    return launch_fps_g<P, W, 1>(xyz, B, N, npoint, start, centroids, vals, st, mix);
}

// Large clouds (8192 < N <= 16384): SoA coordinates (12 N bytes of shared memory), strided ownership, two
// redux.sync per level; results written per round.  Same arithmetic and tie rule.
template <int PPT, int THREADS>
__global__ void __launch_bounds__(THREADS)
fps_soa_kernel(const float *__restrict__ xyz, int N, int npoint, const int64_t *__restrict__ start,
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
static int launch_fps_soa(const float *xyz, int B, int N, int npoint, const int64_t *start, int64_t *centroids,
                      float *vals, cudaStream_t st)
{
    const size_t smem = sizeof(float) * 3 * (size_t)N;
    MLSP_CUDA(cudaFuncSetAttribute(fps_soa_kernel<PPT, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fps_soa_kernel<PPT, THREADS><<<B, THREADS, smem, st>>>(xyz, N, npoint, start, centroids, vals);
    MLSP_LAUNCH_CHECK("fps_soa_kernel");
    return MLSP_OK;
}

}  // namespace mlsp

extern "C" void mlsp_fps_set_groups(int groups) { mlsp::g_fps_groups = groups; }
extern "C" void mlsp_fps_set_exclusive(int on) { mlsp::g_fps_exclusive = on; }

extern "C" int mlsp_fps(const float *xyz, int B, int N, int npoint, const int64_t *start, int64_t *centroids,
                        float *vals, void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(xyz && start && centroids && vals, MLSP_EINVAL, "fps: null pointer");
    MLSP_REQUIRE(B > 0 && N > 0 && npoint >= 0, MLSP_EINVAL, "fps: bad shape");
    if (npoint == 0) return MLSP_OK;
    cudaStream_t st = as_stream(stream);
    // a valid point with the lowest index wins all-zero rounds; invalid start indices are the caller's bug
    if (N <= 128) return launch_fps<4, 1>(xyz, B, N, npoint, start, centroids, vals, st);
    if (N <= 512) return launch_fps<4, 4>(xyz, B, N, npoint, start, centroids, vals, st);
    if (N <= 1024) return launch_fps<4, 8>(xyz, B, N, npoint, start, centroids, vals, st);    // measured: 103 us / 512 rounds
    if (N <= 2048) return launch_fps<16, 4>(xyz, B, N, npoint, start, centroids, vals, st);   // 128 us (8,8: 145 us)
    if (N <= 4096) return launch_fps<16, 8>(xyz, B, N, npoint, start, centroids, vals, st);
    if (N <= 8192) return launch_fps<16, 16>(xyz, B, N, npoint, start, centroids, vals, st);
    if (N <= 16384) return launch_fps_soa<16, 1024>(xyz, B, N, npoint, start, centroids, vals, st);
    set_error("fps: N=%d > 16384 not supported", N);
    return MLSP_EUNSUPPORTED;
}

// 8f rank 3 -- PCM.mix_shapes (MLSP/PCM.py:6-38) in one launch: the two farthest_point_sample calls (npoint_a points of every
// cloud, N - npoint_a points of its partner index[b]) run side by side in 2B CTAs and write the mixed, point-permuted cloud
// directly.  start (2B): the torch.randint draws of the two FPS calls, a-half first.
extern "C" int mlsp_pcm_mix(const float *xyz, int B, int N, int npoint_a, const int64_t *index, const int64_t *start,
                            const int32_t *inv_perm, float *out, void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(xyz && index && start && inv_perm && out, MLSP_EINVAL, "pcm_mix: null pointer");
    MLSP_REQUIRE(B > 0 && N > 0 && npoint_a >= 0 && npoint_a <= N, MLSP_EINVAL, "pcm_mix: bad shape B=%d N=%d npoint_a=%d", B, N, npoint_a);
    MLSP_REQUIRE(N <= 8192, MLSP_EUNSUPPORTED, "pcm_mix: N=%d > 8192", N);
    cudaStream_t st = as_stream(stream);
    FpsMix mix{index, inv_perm, out, npoint_a, B};
    if (N <= 128) return launch_fps<4, 1>(xyz, B, N, 0, start, nullptr, nullptr, st, mix);
    if (N <= 512) return launch_fps<4, 4>(xyz, B, N, 0, start, nullptr, nullptr, st, mix);
    if (N <= 1024) return launch_fps<4, 8>(xyz, B, N, 0, start, nullptr, nullptr, st, mix);
    if (N <= 2048) return launch_fps<16, 4>(xyz, B, N, 0, start, nullptr, nullptr, st, mix);
    if (N <= 4096) return launch_fps<16, 8>(xyz, B, N, 0, start, nullptr, nullptr, st, mix);
    return launch_fps<16, 16>(xyz, B, N, 0, start, nullptr, nullptr, st, mix);
}
