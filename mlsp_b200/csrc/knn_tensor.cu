// knn_tensor.cu -- a1 for the DGCNN feature layers (C = 64, 128): the pairwise-distance contraction on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA) with a fused
// selection epilogue, and results that are STILL bit-exact with the fp32 specification of knn.cu.
//
// Filter (tensor cores) + refine (exact fp32) + certificate:
//   1. prep: x (B,C,N) fp32 -> point-major fp32 rows xt (B,N,C) and an error-compensated bf16 split
//      x = hi + lo (+ 2^-16 |x|).  With A' = [hi|hi|lo], B' = [hi|lo|hi] (K' = 3C) one UMMA chain gives
//      dot~ = hi.hi + hi.lo + lo.hi, |dot~ - dot| <= ~1e-4 |x_i||x_j|.
//   2. main kernel, one CTA per 128 query rows, candidate tiles of 128 (UMMA 128x128x16, kind::f16):
//        warp 0   : TMA producer (A' tile once, B' K-blocks through a 4-stage mbarrier ring)
//        warp 1   : TMEM allocator + single-thread MMA issuer, accumulators double-buffered in TMEM
//        warps 2-5: epilogue, one thread per query row (tcgen05.ld 32x32b: TMEM lane == row)
//      pass 1: v = |x_j|^2 - 2 dot~ ; per row the minimum of every column class (j mod NG) is tracked in
//              registers; tau = k-th smallest class minimum is an upper bound of the k-th distance.
//      pass 2: the same tiles again (the MMA is cheap); columns with v <= tau + 2 eps are appended to the
//              row's candidate list in shared memory (about 1.5 k entries expected, capacity 2 NG).
//      refine: one warp per row recomputes the candidates' distances with the exact fp32 chain of the
//              specification and sorts them (value desc, index asc); the first k are the answer.
//      eps bounds |v - exact| so every member of the exact top-k (ties included) is in the list; a row whose
//      list overflows is not certified and goes to
//   3. a fallback kernel (exact streaming top-k, one warp per listed row).
// SASS evidence: UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA) -- profiles/.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "topk.cuh"

namespace mlsp {

constexpr int KT_ROWS = 128;     // query rows per CTA   (UMMA M, TMEM lanes)
constexpr int KT_COLS = 128;     // candidates per tile  (UMMA N, TMEM columns per accumulator buffer)
constexpr int KT_KBLK = 64;      // bf16 per K block = one 128-byte swizzle span
constexpr int KT_STAGES = 4;
constexpr int KT_THREADS = 192;
constexpr uint32_t KT_BLK_BYTES = KT_COLS * KT_KBLK * 2;  // 16 KiB per (128 x 64) bf16 block
constexpr float KT_EPS_REL = 4.8828125e-4f;               // 2^-11: |v - exact| <= KT_EPS_REL |x_i| max|x_j|

// ---------------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&r)[32])
{
    uint32_t u[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
          "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
          "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
          "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// K-major, 128-byte swizzled operand block (rows 128 B apart, 8-row groups 1024 B apart), sm_100 version bit
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16: D = f32, A = B = bf16, both K-major, N = 128, M = 128
constexpr uint32_t KT_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((KT_COLS >> 3) << 17) | ((KT_ROWS >> 4) << 24);

// ------------------------------------------------------------------------------------------- prep kernels
// norms (exact, spec order) + per-cloud max norm (uint bit pattern max is valid for non-negative floats)
__global__ void sq_norms_max_kernel(const float *__restrict__ x, int C, int N, float *__restrict__ xx,
                                    unsigned int *__restrict__ maxbits)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    float s = 0.0f;
    if (j < N) {
        const float *xb = x + (size_t)b * C * N;
        float v = xb[j];
        s = __fmul_rn(v, v);
        for (int c = 1; c < C; ++c) {
            v = xb[(size_t)c * N + j];
            s = __fadd_rn(s, __fmul_rn(v, v));
        }
        xx[(size_t)b * N + j] = s;
    }
    const unsigned int m = __reduce_max_sync(MLSP_FULL, __float_as_uint(s));
    if ((threadIdx.x & 31) == 0) atomicMax(maxbits + b, m);
}

// (B,C,N) fp32 -> xt (B,N,C) fp32, hi/lo (B*N, C) bf16 ; 32x32 tiles through shared memory
__global__ void knn_prep_kernel(const float *__restrict__ x, int C, int N, float *__restrict__ xt,
                                __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo)
{
    __shared__ float tile[32][33];
    const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float *xb = x + (size_t)b * C * N;
    for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
        const int c = c0 + cc, n = n0 + threadIdx.x;
        tile[cc][threadIdx.x] = (c < C && n < N) ? xb[(size_t)c * N + n] : 0.0f;
    }
    __syncthreads();
    for (int nn = threadIdx.y; nn < 32; nn += blockDim.y) {
        const int n = n0 + nn, c = c0 + threadIdx.x;
        if (n < N && c < C) {
            const float v = tile[threadIdx.x][nn];
            const size_t o = ((size_t)b * N + n) * C + c;
            xt[o] = v;
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            hi[o] = h;
            lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
    }
}

// ------------------------------------------------------------------------------------------- helpers
// exact specification distance of (i, j) from point-major rows: fmul then fmaf chain in channel order
__device__ __forceinline__ float exact_pd(const float4 *__restrict__ xi, const float4 *__restrict__ xj, int C4,
                                          float xxi, float xxj)
{
    float4 a = xi[0], q = xj[0];
    float acc = __fmul_rn(a.x, q.x);
    acc = __fmaf_rn(a.y, q.y, acc);
    acc = __fmaf_rn(a.z, q.z, acc);
    acc = __fmaf_rn(a.w, q.w, acc);
    for (int c = 1; c < C4; ++c) {
        a = xi[c];
        q = xj[c];
        acc = __fmaf_rn(a.x, q.x, acc);
        acc = __fmaf_rn(a.y, q.y, acc);
        acc = __fmaf_rn(a.z, q.z, acc);
        acc = __fmaf_rn(a.w, q.w, acc);
    }
    return __fsub_rn(__fmaf_rn(2.0f, acc, -xxj), xxi);
}

// thread-local bitonic sort of NG registers, ascending
template <int NG>
__device__ __forceinline__ void reg_sort(float (&v)[NG])
{
#pragma unroll
    for (int size = 2; size <= NG; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int e = 0; e < NG; ++e) {
                const int p = e ^ stride;
                if (p > e) {
                    const bool up = (e & size) == 0;
                    const float a = v[e], b = v[p];
                    const float mn = fminf(a, b), mx = fmaxf(a, b);
                    v[e] = up ? mn : mx;
                    v[p] = up ? mx : mn;
                }
            }
        }
    }
}

struct KtParams {
    const float *xx;         // (B,N) exact squared norms
    const float *maxxx;      // (B) max squared norm per cloud
    const float *xt;         // (B,N,C) fp32 point-major
    int64_t *idx;            // (B,N,k)
    int *fb_count;           // fallback row counter
    int *fb_rows;            // fallback rows (b*N + i)
    int *stats;              // [0] rows certified by the tensor path
    float *dump;             // optional (B,N,N) approximate values (tests only)
    int N, C, k, T;          // T = candidate tiles per cloud
};

// ------------------------------------------------------------------------------------------- main kernel
template <int NG>
__global__ void __launch_bounds__(KT_THREADS, 1)
knn_tensor_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, KtParams P)
{
    constexpr int CAP = 2 * NG;
    extern __shared__ uint8_t smem_dyn[];
    uint8_t *smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // swizzle atoms need 1 KiB alignment
    const int KB = 3 * P.C / KT_KBLK;       // K blocks per tile (3 or 6)
    const int SEG = P.C / KT_KBLK;          // K blocks per hi / lo segment
    uint8_t *sA = smem_raw;                                   // KB blocks, resident
    uint8_t *sB = sA + (size_t)KB * KT_BLK_BYTES;             // KT_STAGES blocks, ring
    uint16_t *lists = reinterpret_cast<uint16_t *>(sB + (size_t)KT_STAGES * KT_BLK_BYTES);   // [128][CAP]
    float *nrm_s = reinterpret_cast<float *>(lists + KT_ROWS * CAP);                            // [2][128]
    int *cnt_s = reinterpret_cast<int *>(nrm_s + 2 * KT_COLS);                                 // [128]
    uint64_t *bars = reinterpret_cast<uint64_t *>(cnt_s + KT_ROWS);
    uint64_t *full = bars, *empty = bars + KT_STAGES, *a_full = bars + 2 * KT_STAGES;
    uint64_t *tm_full = a_full + 1, *tm_empty = tm_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tm_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y, i0 = blockIdx.x * KT_ROWS;
    const int N = P.N, T = P.T;
    const int rowbase = b * N;

    if (threadIdx.x == 0) {
        for (int s = 0; s < KT_STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(a_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(tm_full + s, 1);
            mbar_init(tm_empty + s, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            mbar_expect_tx(a_full, (uint32_t)KB * KT_BLK_BYTES);
            for (int kb = 0; kb < KB; ++kb) {        // A' = [hi | hi | lo]
                const int seg = kb / SEG, within = kb - seg * SEG;
                tma_load_2d(sA + (size_t)kb * KT_BLK_BYTES, seg == 2 ? &map_lo : &map_hi, within * KT_KBLK, rowbase + i0, a_full);
            }
            int it = 0;
            for (int g = 0; g < 2 * T; ++g) {
                const int j0 = (g % T) * KT_COLS;
                for (int kb = 0; kb < KB; ++kb, ++it) {   // B' = [hi | lo | hi]
                    const int stage = it % KT_STAGES;
                    const uint32_t ph = (it / KT_STAGES) & 1;
                    mbar_wait(empty + stage, ph ^ 1);
                    mbar_expect_tx(full + stage, KT_BLK_BYTES);
                    const int seg = kb / SEG, within = kb - seg * SEG;
                    tma_load_2d(sB + (size_t)stage * KT_BLK_BYTES, seg == 1 ? &map_lo : &map_hi, within * KT_KBLK,
                                rowbase + j0, full + stage);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            mbar_wait(a_full, 0);
            int it = 0;
            for (int g = 0; g < 2 * T; ++g) {
                const int buf = g & 1;
                mbar_wait(tm_empty + buf, ((g >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)buf * KT_COLS;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int stage = it % KT_STAGES;
                    mbar_wait(full + stage, (it / KT_STAGES) & 1);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(smem_u32(sA + (size_t)kb * KT_BLK_BYTES));
                    const uint64_t db = umma_desc_sw128(smem_u32(sB + (size_t)stage * KT_BLK_BYTES));
#pragma unroll
                    for (int k16 = 0; k16 < KT_KBLK / 16; ++k16)   // +32 bytes per K=16 step inside the swizzle span
                        tc_mma_bf16(d, da + 2 * k16, db + 2 * k16, KT_IDESC, (kb | k16) != 0);
                    tc_commit(empty + stage);                       // smem stage reusable when these MMAs retire
                }
                tc_commit(tm_full + buf);                           // accumulator of tile g complete
            }
        }
        __syncwarp();
    } else {
        // ================================ epilogue: one thread per query row ===========
        const int q = warp & 3;                  // TMEM lane quadrant this warp may access
        const int r = q * 32 + lane;             // row within the tile
        const int i = i0 + r;
        const int et = threadIdx.x - 64;         // 0..127 among the epilogue threads
        const float xxi = (i < N) ? P.xx[(size_t)rowbase + i] : 0.0f;
        const float eps = KT_EPS_REL * sqrtf(xxi) * sqrtf(P.maxxx[b]);
        float gmin[NG];
#pragma unroll
        for (int e = 0; e < NG; ++e) gmin[e] = INFINITY;
        float thr = 0.0f;
        uint16_t *my_list = lists + r * CAP;
        int cnt = 0;
        for (int g = 0; g < 2 * T; ++g) {
            const int buf = g & 1, t = g % T, j0 = t * KT_COLS;
            if (g == T) {
                // ---- between the passes: tau = k-th smallest class minimum (thread-local sorting network)
                reg_sort<NG>(gmin);
                float tau = gmin[0];
#pragma unroll
                for (int e = 1; e < NG; ++e) tau = (e == P.k - 1) ? gmin[e] : tau;
                thr = tau + 2.0f * eps;
            }
            {
                const int j = j0 + et;
                nrm_s[buf * KT_COLS + et] = (j < N) ? P.xx[(size_t)rowbase + j] : INFINITY;
            }
            epi_bar_sync();
            mbar_wait(tm_full + buf, (g >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * KT_COLS;
            const float4 *nrm4 = reinterpret_cast<const float4 *>(nrm_s + buf * KT_COLS);
#pragma unroll
            for (int ch = 0; ch < KT_COLS / 32; ++ch) {
                float acc[32];
                tc_ld32(taddr + ch * 32, acc);
                float v[32];
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float4 nj = nrm4[ch * 8 + c4];
                    v[4 * c4 + 0] = __fmaf_rn(-2.0f, acc[4 * c4 + 0], nj.x);
                    v[4 * c4 + 1] = __fmaf_rn(-2.0f, acc[4 * c4 + 1], nj.y);
                    v[4 * c4 + 2] = __fmaf_rn(-2.0f, acc[4 * c4 + 2], nj.z);
                    v[4 * c4 + 3] = __fmaf_rn(-2.0f, acc[4 * c4 + 3], nj.w);
                }
                if (g < T) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int e = (NG == 32) ? c : ((ch & 1) * 32 + c);   // column class j mod NG (static)
                        gmin[e] = fminf(gmin[e], v[c]);
                    }
                    if (P.dump && i < N) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const int j = j0 + ch * 32 + c;
                            if (j < N) P.dump[((size_t)rowbase + i) * N + j] = v[c];
                        }
                    }
                } else {
                    const int jb = j0 + ch * 32;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        if (v[c] <= thr) {
                            if (cnt < CAP) my_list[cnt] = (uint16_t)(jb + c);
                            ++cnt;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tm_empty + buf);
        }
        cnt_s[r] = cnt;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
    }

    // ================================ refine: exact fp32 re-rank, one warp per row ==========
    const int C4 = P.C / 4;
    int certified = 0;
    for (int r = warp; r < KT_ROWS; r += KT_THREADS / 32) {
        const int i = i0 + r;
        if (i >= N) break;
        const int cnt = cnt_s[r];
        if (cnt > CAP || cnt < P.k) {               // not certified: exact fallback kernel takes the row
            if (lane == 0) P.fb_rows[atomicAdd(P.fb_count, 1)] = rowbase + i;
            continue;
        }
        const float4 *xi = reinterpret_cast<const float4 *>(P.xt + ((size_t)rowbase + i) * P.C);
        const float xxi = P.xx[(size_t)rowbase + i];
        unsigned long long key[CAP / 32];
#pragma unroll
        for (int s = 0; s < CAP / 32; ++s) {
            const int e = s * 32 + lane;
            float pd = -INFINITY;
            int j = 0x7fffffff;
            if (e < cnt) {
                j = lists[r * CAP + e];
                const float4 *xj = reinterpret_cast<const float4 *>(P.xt + ((size_t)rowbase + j) * P.C);
                pd = exact_pd(xi, xj, C4, xxi, P.xx[(size_t)rowbase + j]);
            }
            key[s] = rank_key(pd, j, e < cnt);
        }
        warp_sort_u64<CAP / 32>(key);
#pragma unroll
        for (int s = 0; s < CAP / 32; ++s) {
            const int e = s * 32 + lane;
            if (e < P.k) P.idx[((size_t)rowbase + i) * P.k + e] = (int64_t)(uint32_t)(key[s] & 0xffffffffull);
        }
        ++certified;
    }
    if (lane == 0 && certified) atomicAdd(P.stats, certified);
}

// ------------------------------------------------------------------------------------------- fallback
// exact streaming top-k for the (rare) uncertified rows: one warp per listed row, point-major rows
template <int KSLOTS>
__global__ void __launch_bounds__(256)
knn_fallback_kernel(const float *__restrict__ xt, const float *__restrict__ xx, const int *__restrict__ fb_count,
                    const int *__restrict__ fb_rows, int N, int C, int k, int64_t *__restrict__ idx)
{
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int count = *fb_count;
    const int C4 = C / 4;
    for (int e = gw; e < count; e += nw) {
        const int row = fb_rows[e];
        const int b = row / N;
        const size_t base = (size_t)b * N;
        const float4 *xi = reinterpret_cast<const float4 *>(xt + (size_t)row * C);
        const float xxi = xx[row];
        TopK<KSLOTS> top;
        top.init(k);
        for (int j0 = 0; j0 < N; j0 += 32) {
            const int j = j0 + lane;
            float pd = -INFINITY;
            if (j < N) pd = exact_pd(xi, reinterpret_cast<const float4 *>(xt + (base + j) * C), C4, xxi, xx[base + j]);
            top.offer(pd, j, j < N);
        }
        top.finish(k);
#pragma unroll
        for (int s = 0; s < KSLOTS; ++s) {
            const int r = s * 32 + lane;
            if (r < k) idx[(size_t)row * k + r] = (int64_t)top.j[s];
        }
    }
}

// ------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static int make_map(CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols)
{
    EncodeTiledFn fn = encode_fn();
    MLSP_REQUIRE(fn, MLSP_ECUDA, "knn: cuTensorMapEncodeTiled not available");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)KT_KBLK, (cuuint32_t)KT_COLS};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MLSP_REQUIRE(r == CUDA_SUCCESS, MLSP_ECUDA, "knn: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return MLSP_OK;
}

struct KtLayout {
    size_t off_counters, off_max, off_xx, off_hi, off_lo, off_xt, off_rows, total;
};

static KtLayout kt_layout(int B, int C, int N)
{
    KtLayout L;
    size_t o = 0;
    L.off_counters = o; o += 256;                                            // [0] fb_count, [1] certified rows
    L.off_max = o;      o += align_up(sizeof(float) * (size_t)B, 256);
    L.off_xx = o;       o += align_up(sizeof(float) * (size_t)B * N, 256);
    L.off_hi = o;       o += align_up(2 * (size_t)B * N * C, 1024);
    L.off_lo = o;       o += align_up(2 * (size_t)B * N * C, 1024);
    L.off_xt = o;       o += align_up(sizeof(float) * (size_t)B * N * C, 256);
    L.off_rows = o;     o += align_up(sizeof(int) * (size_t)B * N, 256);
    L.total = o;
    return L;
}

bool knn_tensor_supported(int B, int C, int N, int k)
{
    return (C == 64 || C == 128) && N >= 256 && N <= 65535 && k <= 64 && (long long)B * N < (1ll << 31) && B <= 65535;
}

size_t knn_tensor_workspace_bytes(int B, int C, int N) { return kt_layout(B, C, N).total; }

int knn_tensor_run(const float *x, int B, int C, int N, int k, int64_t *idx, void *ws, float *dump, cudaStream_t st)
{
    const KtLayout L = kt_layout(B, C, N);
    char *w = static_cast<char *>(ws);
    int *counters = reinterpret_cast<int *>(w + L.off_counters);
    float *maxxx = reinterpret_cast<float *>(w + L.off_max);
    float *xx = reinterpret_cast<float *>(w + L.off_xx);
    __nv_bfloat16 *hi = reinterpret_cast<__nv_bfloat16 *>(w + L.off_hi);
    __nv_bfloat16 *lo = reinterpret_cast<__nv_bfloat16 *>(w + L.off_lo);
    float *xt = reinterpret_cast<float *>(w + L.off_xt);
    int *rows = reinterpret_cast<int *>(w + L.off_rows);

    MLSP_CUDA(cudaMemsetAsync(w, 0, L.off_xx, st));    // counters + per-cloud max
    sq_norms_max_kernel<<<dim3((N + 255) / 256, B), 256, 0, st>>>(x, C, N, xx, reinterpret_cast<unsigned int *>(maxxx));
    MLSP_LAUNCH_CHECK("sq_norms_max_kernel");
    knn_prep_kernel<<<dim3((N + 31) / 32, (C + 31) / 32, B), dim3(32, 8), 0, st>>>(x, C, N, xt, hi, lo);
    MLSP_LAUNCH_CHECK("knn_prep_kernel");

    CUtensorMap map_hi, map_lo;
    int rc = make_map(&map_hi, hi, (uint64_t)B * N, (uint64_t)C);
    if (rc) return rc;
    rc = make_map(&map_lo, lo, (uint64_t)B * N, (uint64_t)C);
    if (rc) return rc;

    KtParams P;
    P.xx = xx; P.maxxx = maxxx; P.xt = xt; P.idx = idx; P.fb_count = counters; P.fb_rows = rows;
    P.stats = counters + 1; P.dump = dump; P.N = N; P.C = C; P.k = k; P.T = (N + KT_COLS - 1) / KT_COLS;
    const int NG = (k <= 32) ? 32 : 64;
    const int KB = 3 * C / KT_KBLK;
    const size_t smem = (size_t)(KB + KT_STAGES) * KT_BLK_BYTES + (size_t)KT_ROWS * 2 * NG * 2 + 2 * KT_COLS * 4 +
                        KT_ROWS * 4 + 16 * 8 + 16 + 1024;
    dim3 grid((N + KT_ROWS - 1) / KT_ROWS, B);
    if (NG == 32) {
        MLSP_CUDA(cudaFuncSetAttribute(knn_tensor_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_tensor_kernel<32><<<grid, KT_THREADS, smem, st>>>(map_hi, map_lo, P);
    } else {
        MLSP_CUDA(cudaFuncSetAttribute(knn_tensor_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_tensor_kernel<64><<<grid, KT_THREADS, smem, st>>>(map_hi, map_lo, P);
    }
    MLSP_LAUNCH_CHECK("knn_tensor_kernel");
    const int fb_blocks = 2 * sm_count();
    if (k <= 32)
        knn_fallback_kernel<1><<<fb_blocks, 256, 0, st>>>(xt, xx, counters, rows, N, C, k, idx);
    else
        knn_fallback_kernel<2><<<fb_blocks, 256, 0, st>>>(xt, xx, counters, rows, N, C, k, idx);
    MLSP_LAUNCH_CHECK("knn_fallback_kernel");
    return MLSP_OK;
}

}  // namespace mlsp
