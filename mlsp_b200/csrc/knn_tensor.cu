// knn_tensor.cu -- a1 for the DGCNN feature layers (C = 64, 128): the pairwise-distance contraction on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA) with a fused
// selection epilogue, and results that are STILL bit-exact with the fp32 specification of knn.cu.
//
// Filter (tensor cores) + refine (exact fp32) + certificate:
//   1. prep: x (B,C,N) fp32 -> point-major fp32 rows xt (B,N,C) and an error-compensated bf16 split
//      x = hi + lo (+ 2^-16 |x|).  The A operand [hi|lo] of the CTA's 128 query rows stays in shared memory; every
//      candidate block B_hi is multiplied with A_hi and A_lo, every B_lo block with A_hi, all into one TMEM
//      accumulator: dot~ = hi.hi + lo.hi + hi.lo, |dot~ - dot| <= ~1e-4 |x_i||x_j|, and each B byte is fetched once.
//   2. main kernel, one CTA per 128 query rows, candidate tiles of 128 (UMMA 128x128x16, kind::f16):
//        warp 0   : TMA producer (A tile once, B K-blocks through an mbarrier ring)
//        warp 1   : TMEM allocator + single-thread MMA issuer, accumulators double-buffered in TMEM
//        warps 2-9: epilogue, two threads per query row (tcgen05.ld 32x32b: TMEM lane == row)
//      pass 1: v = |x_j|^2 - 2 dot~ ; per row the minimum of every column class (j mod NG) is tracked in
//              registers; tau = k-th smallest class minimum is an upper bound of the k-th distance.
//      pass 2: the same tiles again (the MMA is cheap); columns with v <= tau + 2 eps are written as (v, j) pairs
//              to the row's candidate list in global memory (L2): the two threads of a row fill it from both ends
//              with private cursors -- predicated stores, no atomics, no votes (about 1.2 k entries expected,
//              capacity 2 NG).  eps bounds |v - exact| so every member of the exact top-k (ties included) is listed.
//   3. refine kernel (one warp per row): sorts the list by the approximate value.  Two neighbours of that order
//      whose values differ by more than 2 eps are provably in the same order in exact fp32 arithmetic, so only
//      runs of near-ties ("clusters") that start inside the first k positions need the exact value: for those the
//      distance is recomputed with the pinned fp32 chain of the specification (four lanes per candidate, coalesced
//      reads of the point-major rows) and the cluster is re-sorted by (value desc, index asc).  Typical rows need
//      no or a few exact distances instead of one 4C-byte gather per candidate.
//      A row whose list overflowed is not certified and goes to
//      the exact streaming top-k (topk.cuh) in the same warp.
// SASS evidence: UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA) -- profiles/.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "topk.cuh"

namespace mlsp {

constexpr int KT_ROWS = 128;     // query rows per CTA   (UMMA M, TMEM lanes)
constexpr int KT_COLS = 128;     // candidates per tile  (UMMA N, TMEM columns per accumulator buffer)
constexpr int KT_KBLK = 64;      // bf16 per K block = one 128-byte swizzle span
constexpr int KT_MAX_STAGES = 8;
constexpr int KT_EPI_WARPS = 8;  // two per TMEM lane quadrant: each takes one 64-column half of every tile
constexpr int KT_THREADS = 64 + 32 * KT_EPI_WARPS;
constexpr uint32_t KT_BLK_BYTES = KT_COLS * KT_KBLK * 2;  // 16 KiB per (128 x 64) bf16 block
// |v - exact| <= KT_EPS_REL |x_i| max_j|x_j| :  dropped lo.lo and split residuals are <= 3*2^-16 |x_i||x_j| on the
// dot product (Cauchy-Schwarz), i.e. 9.2e-5 on v; 2^-12 = 2.4e-4 leaves > 2.5x for the tensor core's fp32
// accumulation and the specification's own roundings.  Measured worst case: 0.12 of this bound (tests).
constexpr float KT_EPS_REL = 2.44140625e-4f;
// Pass 1 multiplies the bf16 heads only (dot~1 = hi.hi): x = hi + r with |r_c| <= 2^-9 |x_c|, so
// |hi_i.hi_j - x_i.x_j| <= (2^-8 + 2^-17 + 2^-18) |x_i||x_j| and the pass-1 value v1 = |x_j|^2 - 2 dot~1 is within
// KT_EPS1_REL |x_i||x_j| (= 2^-7 + 2^-14) of the three-term value; tau1 + eps1 still upper-bounds the k-th distance.
constexpr float KT_EPS1_REL = 0.00787353515625f;

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
// multicast variant: the box lands at the same CTA-relative offset in every CTA of `mask`, and complete_tx is signalled
// on the mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_load_2d_mc(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar, uint16_t mask)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the same arrive, delivered to the mbarrier at this offset in every CTA of `mask` (a stage shared by a CTA pair is
// free only when both CTAs' MMAs have read it)
__device__ __forceinline__ void tc_commit_mc(uint64_t *bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
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
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&r)[16])
{
    uint32_t u[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
          "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}
// asynchronous 16-column load: the registers are valid only after tc_wait16 on the same array (the "+r"
// operands make every later use of the values depend on the wait, so the compiler cannot hoist them above it)
__device__ __forceinline__ void tc_ld16_issue(uint32_t taddr, uint32_t (&u)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
          "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_wait16(uint32_t (&u)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]),
                   "+r"(u[8]), "+r"(u[9]), "+r"(u[10]), "+r"(u[11]), "+r"(u[12]), "+r"(u[13]), "+r"(u[14]), "+r"(u[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * KT_EPI_WARPS) : "memory"); }

// K-major, 128-byte swizzled operand block (rows 128 B apart, 8-row groups 1024 B apart), sm_100 version bit
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16: D = f32, A = B = bf16, both K-major, N = 128, M = 128
constexpr uint32_t KT_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((KT_COLS >> 3) << 17) | ((KT_ROWS >> 4) << 24);

// ------------------------------------------------------------------------------------------- prep kernel
// One pass over x (B,C,N): exact norms in the specification order (sequential adds over the channels), the
// point-major fp32 rows xt (B,N,C) and the bf16 split hi/lo (B*N, C); also zeroes the two diagnostic counters.  A CTA stages a [C][PREP_PTS] slab in shared memory
// (coalesced 128-byte reads along n), then warps write whole point rows (coalesced along c).
constexpr int PREP_PTS = 32;
constexpr int PREP_THREADS = 256;

__global__ void __launch_bounds__(PREP_THREADS)
knn_prep_kernel(const float *__restrict__ x, int C, int N, float *__restrict__ xx, int *__restrict__ counters,
                float *__restrict__ xt, __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo,
                int *__restrict__ cloud_mode)
{
    __shared__ float mode_acc[2];
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 2) counters[threadIdx.x] = 0;
    if (threadIdx.x < 2) mode_acc[threadIdx.x] = 0.0f;
    extern __shared__ float slab[];                       // [C][PREP_PTS + 1]
    const int b = blockIdx.y, n0 = blockIdx.x * PREP_PTS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *xb = x + (size_t)b * C * N;
    for (int c = warp; c < C; c += PREP_THREADS / 32) {
        const int n = n0 + lane;
        slab[c * (PREP_PTS + 1) + lane] = (n < N) ? xb[(size_t)c * N + n] : 0.0f;
    }
    __syncthreads();
    if (warp == 0) {                                       // norms: one point per lane, channels in order
        float v = slab[lane];
        float s = __fmul_rn(v, v);
        for (int c = 1; c < C; ++c) {
            v = slab[c * (PREP_PTS + 1) + lane];
            s = __fadd_rn(s, __fmul_rn(v, v));
        }
        const int n = n0 + lane;
        if (n < N) xx[(size_t)b * N + n] = s;
    }
    // Which pass 1 for this cloud?  The bf16-head product is within KT_EPS1_REL |x_i||x_j| of the distance: fine for
    // activations whose spread is comparable to their norm (BatchNorm + LeakyReLU outputs: E|x|^2 ~ 1.3-3 x the variance),
    // hopeless for a tight cluster far from the origin (BatchNorm-free layers with a bias: the bound exceeds the k-th
    // distance, every list overflows and the rows fall back to the exact streaming selection, 50x slower).  The first
    // CTA of a cloud estimates E|x|^2 and the total variance from its 32 staged points and flags the cloud for the
    // three-term pass 1 (error KT_EPS_REL, 32x tighter, 13 % more filter time) when E|x|^2 > 4 Var.  Either choice is
    // certified; the flag only picks the cheaper one that works.
    if (blockIdx.x == 0) {
        const int np = min(PREP_PTS, N - n0);
        if ((int)threadIdx.x < C) {
            float sm = 0.0f, sq = 0.0f;
            for (int pt = 0; pt < np; ++pt) {
                const float v = slab[threadIdx.x * (PREP_PTS + 1) + pt];
                sm += v;
                sq = fmaf(v, v, sq);
            }
            const float m2 = sq / (float)np, mu = sm / (float)np;
            atomicAdd(&mode_acc[0], m2);                      // E|x|^2 summed over channels
            atomicAdd(&mode_acc[1], fmaxf(m2 - mu * mu, 0.0f));   // total variance
        }
        __syncthreads();
        if (threadIdx.x == 0) cloud_mode[b] = (mode_acc[0] > 4.0f * mode_acc[1]) ? 1 : 0;
    }
    // rows: thread t handles channel pair (2t mod C ...) of point rows; consecutive threads -> consecutive channels
    const int pairs = C / 2;                               // C is even (64 or 128)
    for (int e = threadIdx.x; e < PREP_PTS * pairs; e += PREP_THREADS) {
        const int pt = e / pairs, c = 2 * (e - pt * pairs);
        const int n = n0 + pt;
        if (n >= N) continue;
        const float v0 = slab[c * (PREP_PTS + 1) + pt], v1 = slab[(c + 1) * (PREP_PTS + 1) + pt];
        const size_t o = ((size_t)b * N + n) * C + c;
        *reinterpret_cast<float2 *>(xt + o) = make_float2(v0, v1);
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
        __nv_bfloat162 hv, lv;
        hv.x = h0; hv.y = h1;
        lv.x = __float2bfloat16_rn(v0 - __bfloat162float(h0));
        lv.y = __float2bfloat16_rn(v1 - __bfloat162float(h1));
        *reinterpret_cast<__nv_bfloat162 *>(hi + o) = hv;
        *reinterpret_cast<__nv_bfloat162 *>(lo + o) = lv;
    }
}

// ------------------------------------------------------------------------------------------- helpers
// exact specification distance of (i, j) from point-major rows (oracle dot_tree): 8 interleaved fmaf chains
// over groups of 4 channels, fixed butterfly.  One lane does the whole candidate; requires C % 32 == 0.
__device__ __forceinline__ float exact_pd(const float4 *__restrict__ xi, const float4 *__restrict__ xj, int C4,
                                          float xxi, float xxj)
{
    float p[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) p[t] = 0.0f;
    for (int f0 = 0; f0 < C4; f0 += 8) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const float4 a = xi[f0 + t], q = xj[f0 + t];
            p[t] = __fmaf_rn(a.x, q.x, p[t]);
            p[t] = __fmaf_rn(a.y, q.y, p[t]);
            p[t] = __fmaf_rn(a.z, q.z, p[t]);
            p[t] = __fmaf_rn(a.w, q.w, p[t]);
        }
    }
    const float q0 = __fadd_rn(p[0], p[4]), q1 = __fadd_rn(p[1], p[5]), q2 = __fadd_rn(p[2], p[6]), q3 = __fadd_rn(p[3], p[7]);
    const float dot = __fadd_rn(__fadd_rn(q0, q2), __fadd_rn(q1, q3));
    return __fsub_rn(__fmaf_rn(2.0f, dot, -xxj), xxi);
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

// per-pair error bound of the filter value: |v~_ij - exact_ij| <= KT_EPS_REL |x_i||x_j| + 2^-21 (|x_i|^2 + |x_j|^2)
// (Cauchy-Schwarz on the dropped split terms and the accumulation, plus the specification's own roundings)
__device__ __forceinline__ float pair_eps(float ni, float xxi, float xxj)
{
    return __fmaf_rn(KT_EPS_REL * ni, sqrtf(xxj), 4.76837158203125e-7f * (xxi + xxj));
}

// ---- epilogue pieces ---------------------------------------------------------------------------------
// pass-2 cursor of one thread: its end of the row's global candidate list
struct Cursor {
    uint2 *base;             // the thread's end of the row's list
    int step;                // +1 (columns 0..63 of every tile: from the front) or -1 (columns 64..127: from the back)
    int off;                 // step * (entries written so far)
    bool ovf;
};

// one 16-column piece of one query row: v = |x_j|^2 - 2 dot~
//   PASS 1: class minima (class = column within the thread's 64-column half, mod NG)
//   PASS 2: every column with v <= thr is stored as (v, j) at the cursor -- predicated, branch-free
template <int NG, int PASS>
__device__ __forceinline__ void epi_piece(float (&gmin)[NG], Cursor &cur, const uint32_t (&u)[16], const float4 *nrm4,
                                          int ch, float thr, uint32_t jcol)
{
    if (PASS == 2) {
        // a piece can add 16 entries: without room for them nothing is stored any more and the row is flagged
        const bool room = abs(cur.off) <= 2 * NG - 16;
        cur.ovf |= !room;
        thr = room ? thr : -INFINITY;
    }
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
        const float4 nj = nrm4[ch * 4 + c4];
        const float nv[4] = {nj.x, nj.y, nj.z, nj.w};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int c = 4 * c4 + w;
            const float v = __fmaf_rn(-2.0f, __uint_as_float(u[c]), nv[w]);
            if (PASS == 1) {
                const int e = (ch * 16 + c) % NG;
                gmin[e] = fminf(gmin[e], v);
            } else {
                // if (v <= thr) { list[off] = (v, j); off += step; }: one predicated store + one predicated add, no branch
                uint2 *a = cur.base + cur.off;
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "setp.le.f32 p, %1, %2;\n\t"
                    "@p st.global.v2.b32 [%3], {%4, %5};\n\t"
                    "@p add.s32 %0, %0, %6;\n\t}"
                    : "+r"(cur.off)
                    : "f"(v), "f"(thr), "l"(a), "r"(__float_as_uint(v)), "r"(jcol + (uint32_t)c), "r"(cur.step));
            }
        }
    }
}

// one thread, one query row, 64 columns in four pieces (tcgen05.ld 32x32b.x16), the load of piece p+1 in flight
// while piece p is processed
template <int NG, int PASS>
__device__ __forceinline__ void epi_tile(float (&gmin)[NG], Cursor &cur, uint32_t taddr, const float *nrm, float thr,
                                         uint32_t jbase)
{
    const float4 *nrm4 = reinterpret_cast<const float4 *>(nrm);
    uint32_t ua[16], ub[16];
    tc_ld16_issue(taddr, ua);
    tc_wait16(ua);
    tc_ld16_issue(taddr + 16, ub);
    epi_piece<NG, PASS>(gmin, cur, ua, nrm4, 0, thr, jbase);
    tc_wait16(ub);
    tc_ld16_issue(taddr + 32, ua);
    epi_piece<NG, PASS>(gmin, cur, ub, nrm4, 1, thr, jbase + 16);
    tc_wait16(ua);
    tc_ld16_issue(taddr + 48, ub);
    epi_piece<NG, PASS>(gmin, cur, ua, nrm4, 2, thr, jbase + 32);
    tc_wait16(ub);
    epi_piece<NG, PASS>(gmin, cur, ub, nrm4, 3, thr, jbase + 48);
}

// test hook: write the approximate values of this thread's 64 columns
// (tcgen05.ld is .sync.aligned: the whole warp must execute it, so row validity only predicates the stores)
__device__ __noinline__ void dump_tile(float *row_out, bool row_valid, uint32_t taddr, const float *nrm, int jbase, int N)
{
    for (int ch = 0; ch < 2; ++ch) {
        float acc[32];
        tc_ld32(taddr + ch * 32, acc);
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const int j = jbase + ch * 32 + c;
            if (row_valid && j < N) row_out[j] = __fmaf_rn(-2.0f, acc[c], nrm[ch * 32 + c]);
        }
    }
}

struct KtParams {
    const float *xx;         // (B,N) exact squared norms
    const float *xt;         // (B,N,C) fp32 point-major
    int64_t *idx;            // (B,N,k)
    int *fb_count;           // rows whose list overflowed (re-done with the exact streaming selection)
    int *stats;              // rows certified by the tensor path
    uint2 *cand;             // (B*N, CAP) candidate lists (v bits, j): main kernel -> refine kernel
    int *cand_cnt;           // (B*N, 2) entries written from the front / from the back (> CAP: overflowed)
    float *dump;             // optional (2,B,N,N) filter values of pass 1 and pass 2 (tests only)
    float4 *edge_out;        // optional (B,N,k,2C): the refine kernel also writes the row's edge features (a2 fused)
    int N, C, k, T;          // T = candidate tiles per cloud
    int stages;              // depth of the B-operand smem ring
    int pair;                // launched in clusters of two CTAs sharing the candidate blocks by TMA multicast
    const int *cloud_mode;   // (B) written by the prep kernel: 1 = this cloud's pass 1 uses the three-term product (below)
    int mode;                // tuning hook (MLSP_KT_MODE): bit 0 skips the pass-1 math, bit 1 the pass-2 math, bit 2: three-term pass 1
                             // for every cloud, bit 3: bf16-head pass 1 for every cloud (ignore cloud_mode)
};

// ------------------------------------------------------------------------------------------- main kernel
// shared memory: A (2*SEG blocks) | B ring (STAGES blocks) | xchg [min(k,NG)][128] | nrm [2][128] | thr [128] | max [4] | barriers
__host__ __device__ inline size_t kt_smem_bytes(int C, int k, int NG, int stages)
{
    const int kx = k < NG ? k : NG;
    return (size_t)(2 * C / KT_KBLK + stages) * KT_BLK_BYTES + (size_t)kx * KT_ROWS * 4 + 2 * KT_COLS * 4 + KT_ROWS * 4 + 16 +
           32 * 8 + 16 + 1024;
}

template <int NG>
__global__ void __launch_bounds__(KT_THREADS, (NG == 32) ? 2 : 1)   // NG = 32: two CTAs per SM (one wave at 32 x 1024), <= 102 registers
knn_tensor_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                  const __grid_constant__ CUtensorMap half_hi, const __grid_constant__ CUtensorMap half_lo, KtParams P)
{
    // P.pair: the kernel was launched in clusters of two CTAs = two adjacent row blocks of one cloud.  Both stream the
    // same candidate blocks, so each CTA fetches HALF of every block (64 of its 128 rows) and TMA multicasts it into
    // both shared memories: the L2 -> SM operand traffic, which bounds the MMA pipeline, halves.
    const bool pair = P.pair != 0;
    const uint32_t crank = pair ? cluster_ctarank() : 0u;
    constexpr int CAP = 2 * NG;
    extern __shared__ uint8_t smem_dyn[];
    uint8_t *smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // swizzle atoms need 1 KiB alignment
    const int SEG = P.C / KT_KBLK;          // K blocks per hi / lo segment (1 or 2)
    const int KB = 2 * SEG;                 // B blocks streamed per tile: hi blocks, then lo blocks
    const int STAGES = P.stages;
    const int KX = min(P.k, NG);
    uint8_t *sA = smem_raw;                                   // [hi blocks | lo blocks], resident
    uint8_t *sB = sA + (size_t)KB * KT_BLK_BYTES;             // STAGES blocks, ring
    float *xchg = reinterpret_cast<float *>(sB + (size_t)STAGES * KT_BLK_BYTES);   // [KX][128] sorted class minima of the upper half
    float *nrm_s = xchg + (size_t)KX * KT_ROWS;                                     // [2][128]
    float *thr_s = nrm_s + 2 * KT_COLS;                                             // [128]
    uint32_t *max_s = reinterpret_cast<uint32_t *>(thr_s + KT_ROWS);                // [4] per-warp maxima of the cloud's norms
    uint64_t *bars = reinterpret_cast<uint64_t *>(max_s + 4);
    uint64_t *full = bars, *empty = bars + KT_MAX_STAGES, *a_full = bars + 2 * KT_MAX_STAGES;
    uint64_t *tm_full = a_full + 1, *tm_empty = tm_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tm_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y, i0 = blockIdx.x * KT_ROWS;
    const int N = P.N, T = P.T;
    const int rowbase = b * N;
    // pass 1 on the three-term product (as tight as pass 2) instead of the bf16 heads: forced by the tuning hook, or
    // chosen per cloud by the prep kernel for activations far from the origin (uniform over the CTA and its pair)
    const bool three = (P.mode & 4) || (!(P.mode & 8) && P.cloud_mode[b] != 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < KT_MAX_STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, pair ? 2 : 1);                   // pair: one commit-arrive from each CTA's MMA issuer
        }
        mbar_init(a_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(tm_full + s, 1);
            mbar_init(tm_empty + s, KT_EPI_WARPS);
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
    if (pair) cluster_sync_all();                                   // the peer's barriers exist before anything signals them
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            mbar_expect_tx(a_full, (uint32_t)KB * KT_BLK_BYTES);
            for (int kb = 0; kb < KB; ++kb)          // A = [hi | lo]
                tma_load_2d(sA + (size_t)kb * KT_BLK_BYTES, kb < SEG ? &map_hi : &map_lo, (kb % SEG) * KT_KBLK,
                            rowbase + i0, a_full);
            int stage = 0;
            uint32_t ph = 0;
            for (int g = 0; g < 2 * T; ++g) {
                const int j0 = (g % T) * KT_COLS;
                const int nkb = (g < T && !three) ? SEG : KB;          // pass 1 multiplies hi.hi only: no lo blocks
                for (int kb = 0; kb < nkb; ++kb) {   // B blocks: hi ..., lo ...
                    mbar_wait(empty + stage, ph ^ 1);
                    mbar_expect_tx(full + stage, KT_BLK_BYTES);
                    if (pair)                                       // my half of the block, into both CTAs
                        tma_load_2d_mc(sB + (size_t)stage * KT_BLK_BYTES + (size_t)crank * (KT_BLK_BYTES / 2),
                                       kb < SEG ? &half_hi : &half_lo, (kb % SEG) * KT_KBLK,
                                       rowbase + j0 + (int)crank * (KT_COLS / 2), full + stage, (uint16_t)3);
                    else
                        tma_load_2d(sB + (size_t)stage * KT_BLK_BYTES, kb < SEG ? &map_hi : &map_lo, (kb % SEG) * KT_KBLK,
                                    rowbase + j0, full + stage);
                    if (++stage == STAGES) { stage = 0; ph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            mbar_wait(a_full, 0);
            int stage = 0;
            uint32_t ph = 0;
            for (int g = 0; g < 2 * T; ++g) {
                const int buf = g & 1;
                mbar_wait(tm_empty + buf, ((g >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)buf * KT_COLS;
                const bool first = g < T && !three;                 // pass 1: dot~ = hi.hi (a looser, cheaper bound)
                const int nkb = first ? SEG : KB;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(full + stage, ph);
                    tc_fence_after();
                    const uint64_t db = umma_desc_sw128(smem_u32(sB + (size_t)stage * KT_BLK_BYTES));
                    const int ka = kb % SEG;                        // matching K block of A
                    const uint64_t da_hi = umma_desc_sw128(smem_u32(sA + (size_t)ka * KT_BLK_BYTES));
#pragma unroll
                    for (int k16 = 0; k16 < KT_KBLK / 16; ++k16)   // +32 bytes per K=16 step inside the swizzle span
                        tc_mma_bf16(d, da_hi + 2 * k16, db + 2 * k16, KT_IDESC, (kb | k16) != 0);   // hi.hi | hi.lo
                    if (kb < SEG && !first) {                       // pass 2: a B_hi block also meets A_lo:  lo.hi
                        const uint64_t da_lo = umma_desc_sw128(smem_u32(sA + (size_t)(SEG + ka) * KT_BLK_BYTES));
#pragma unroll
                        for (int k16 = 0; k16 < KT_KBLK / 16; ++k16)
                            tc_mma_bf16(d, da_lo + 2 * k16, db + 2 * k16, KT_IDESC, 1u);
                    }
                    if (pair) tc_commit_mc(empty + stage, (uint16_t)3);   // stage reusable when BOTH CTAs' MMAs retired
                    else tc_commit(empty + stage);                  // smem stage reusable when these MMAs retire
                    if (++stage == STAGES) { stage = 0; ph ^= 1; }
                }
                tc_commit(tm_full + buf);                           // accumulator of tile g complete
            }
        }
        __syncwarp();
    } else {
        // ================================ epilogue: two threads per query row ===========
        // warp w (2..9): TMEM lane quadrant q = w & 3 (hardware rule), column half h = (w - 2) >> 2.
        const int q = warp & 3;
        const int h = (warp - 2) >> 2;
        const int r = q * 32 + lane;             // row within the tile
        const int i = i0 + r;
        const int et = threadIdx.x - 64;         // 0..255 among the epilogue threads
        const float xxi = (i < N) ? P.xx[(size_t)rowbase + i] : 0.0f;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)h * 64;
        float gmin[NG];
#pragma unroll
        for (int e = 0; e < NG; ++e) gmin[e] = INFINITY;
        Cursor cur;
        {
            uint2 *row_list = P.cand + ((size_t)rowbase + min(i, N - 1)) * CAP;
            cur.step = h ? -1 : 1;
            cur.base = h ? row_list + (CAP - 1) : row_list;
            cur.off = 0;
            cur.ovf = false;
        }
        const float *xxb = P.xx + rowbase;
        // norms of tile 0; afterwards the norms of tile g+1 are fetched while tile g is processed.  The same threads
        // see every norm of the cloud once during pass 1: their maximum gives the error bound of the row.
        float nmax = 0.0f;
        if (et < KT_COLS) {
            const float n0 = (et < N) ? xxb[et] : INFINITY;
            nrm_s[et] = n0;
            if (et < N) nmax = n0;
        }

        // ---- pass 1: class minima
        for (int g = 0; g < T; ++g) {
            const int buf = g & 1;
            float nxt = INFINITY;
            const int jn = ((g + 1) % T) * KT_COLS + et;       // tile g+1 (pass 2 restarts at tile 0)
            if (et < KT_COLS && jn < N) nxt = xxb[jn];           // consumed at the bottom of the iteration: the L2 latency
                                                                // hides behind the tile (a use here would stall all 8 warps at the barrier)
            epi_bar_sync();                                     // norms of tile g visible
            mbar_wait(tm_full + buf, (g >> 1) & 1);
            tc_fence_after();
            if (!(P.mode & 1)) epi_tile<NG, 1>(gmin, cur, tlane + (uint32_t)buf * KT_COLS, nrm_s + buf * KT_COLS + h * 64, 0.0f, 0u);
            if (P.dump) dump_tile(P.dump + ((size_t)rowbase + min(i, N - 1)) * N, i < N, tlane + (uint32_t)buf * KT_COLS,
                                  nrm_s + buf * KT_COLS + h * 64, g * KT_COLS + h * 64, N);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tm_empty + buf);
            if (et < KT_COLS) {
                nrm_s[(buf ^ 1) * KT_COLS + et] = nxt;
                if (jn < N) nmax = fmaxf(nmax, nxt);
            }
        }
        // ---- between the passes: the row has 2 NG class minima (NG per thread).  Each thread sorts its own;
        // the k-th smallest of the union of two sorted lists A, B is max_{i<k} min(A[i], B[k-1-i]).
        reg_sort<NG>(gmin);
        if (h == 1) {
#pragma unroll
            for (int e = 0; e < NG; ++e)
                if (e < KX) xchg[e * KT_ROWS + r] = gmin[e];
        }
        if (et < KT_COLS) {                                     // warps 2..5 hold the norms
            const uint32_t m = __reduce_max_sync(MLSP_FULL, __float_as_uint(nmax));   // non-negative floats order as uints
            if (lane == 0) max_s[warp - 2] = m;
        }
        epi_bar_sync();
        if (h == 0) {
            const float maxxx = __uint_as_float(max(max(max_s[0], max_s[1]), max(max_s[2], max_s[3])));
            const float eps = pair_eps(sqrtf(xxi), xxi, maxxx);   // bound for every candidate of the cloud
            // pass 1 saw dot~ = hi.hi only: |v1 - v| <= KT_EPS1_REL |x_i| max_j |x_j|
            const float eps1 = three ? 0.0f : KT_EPS1_REL * sqrtf(xxi) * sqrtf(maxxx);
            float tau = -INFINITY;
#pragma unroll
            for (int e = 0; e < NG; ++e)
                if (e < P.k) tau = fmaxf(tau, fminf(gmin[e], xchg[(P.k - 1 - e) * KT_ROWS + r]));
            thr_s[r] = tau + eps1 + 2.0f * eps;
        }
        epi_bar_sync();
        const float thr = (i < N) ? thr_s[r] : -INFINITY;    // rows beyond the cloud collect nothing
        // ---- pass 2: collect the candidates
        for (int g = T; g < 2 * T; ++g) {
            const int buf = g & 1;
            float nxt = INFINITY;
            const int jn = (g + 1 - T) * KT_COLS + et;
            if (et < KT_COLS && g + 1 < 2 * T && jn < N) nxt = xxb[jn];
            epi_bar_sync();
            mbar_wait(tm_full + buf, (g >> 1) & 1);
            tc_fence_after();
            if (!(P.mode & 2)) epi_tile<NG, 2>(gmin, cur, tlane + (uint32_t)buf * KT_COLS, nrm_s + buf * KT_COLS + h * 64, thr,
                            (uint32_t)((g - T) * KT_COLS + h * 64));
            if (P.dump) dump_tile(P.dump + ((size_t)gridDim.y * N + rowbase + min(i, N - 1)) * N, i < N, tlane + (uint32_t)buf * KT_COLS,
                                  nrm_s + buf * KT_COLS + h * 64, (g - T) * KT_COLS + h * 64, N);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tm_empty + buf);
            if (et < KT_COLS) nrm_s[(buf ^ 1) * KT_COLS + et] = nxt;
        }
        if (i < N) {
            P.cand_cnt[2 * ((size_t)rowbase + i) + h] = cur.ovf ? CAP + 1 : abs(cur.off);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (pair) cluster_sync_all();            // no CTA leaves while its peer can still signal its barriers / write its smem
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
    }
}

// ------------------------------------------------------------------------------------------- refine
// One warp per query row.  S slots of 32 lanes hold the row's candidates sorted by (v~ asc, j asc).
//   link[e]    : v~[e] - v~[e-1] <= 2 eps  (the exact order of e-1 and e is not certified)
//   cluster    : maximal run of linked positions; cs[e] = its first position
//   amb[e]     : e belongs to a cluster of >= 2 members that starts inside the first k positions
// Clusters are ordered among themselves by certificate, so only amb entries need the exact fp32 distance:
// FOUR lanes per candidate, eight sorted positions per pass (passes without an amb entry are skipped): lane u
// of a group owns chains t = u and t = u + 4 of the pinned dot product (float4 pieces f = u, u+4, u+8, ... of the
// point-major rows), so a warp-wide load touches 8 candidates x 64 contiguous bytes; p_u + p_{u+4} is a register
// add, two xor-shuffles finish the tree of oracle dot_tree: ((p0+p4)+(p2+p6)) + ((p1+p5)+(p3+p7)).  A second
// sort on (cluster start, exact value desc, index asc) then gives the specification's order.
constexpr int RF_WARPS = 8;

// Keys arrive as (orderable v~ << 32 | j << 16 | list slot): the slot names where the candidate's norm |x_j|^2 --
// fetched before the sort, so that its L2 latency hides behind the sorting network -- waits in shared memory.
template <int S, int C, int KS>
__device__ __forceinline__ void refine_sorted(const KtParams &P, uint32_t row, uint32_t base, unsigned long long (&key)[S],
                                              const float (&xpre)[S], uint16_t *sj, float *se, uint32_t (&nbr)[KS])
{
    constexpr int M = C / 16;                             // float4 pieces per lane: 4 (C = 64) or 8 (C = 128)
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, u = lane & 3;
    const int k = P.k;
    const float xxi = P.xx[row];
    const float ni = sqrtf(xxi);
    warp_sort_u64<S>(key);
#pragma unroll
    for (int s = 0; s < S; ++s) se[s * 32 + lane] = xpre[s];             // norms by list slot
    __syncwarp();
    float xs[S];                                                         // |x_j|^2 of the candidate now at (s, lane)
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const uint32_t low = (uint32_t)key[s];
        const bool live = key[s] != ~0ull;
        xs[s] = live ? se[low & 0xffffu] : 0.0f;
        if (live) key[s] = (key[s] & 0xffffffff00000000ull) | (low >> 16);   // -> (orderable v~ << 32 | j)
    }
    // ---- links, clusters
    // A position e that starts a new cluster must certify EVERY later position m >= e against EVERY earlier one
    // t < e: exact_m - exact_t >= (v_e - v_{e-1}) - eps_m - eps_t.  The per-pair bounds differ with the norms, so the
    // test uses the largest bound of the whole list (one redux per slot): gap > 2 max eps.
    float v[S];
    bool link[S];
    int cs[S];
    int carry = 0;
    float emax = 0.0f;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const bool live = key[s] != ~0ull;
        const uint32_t ob = (uint32_t)(key[s] >> 32);                    // f32_orderable(v~)
        v[s] = live ? __uint_as_float((ob & 0x80000000u) ? (ob ^ 0x80000000u) : ~ob) : INFINITY;
        emax = fmaxf(emax, warp_max_f32(live ? pair_eps(ni, xxi, xs[s]) : 0.0f));
    }
    const float gap_max = 1.0009765625f * (emax + emax);
#pragma unroll
    for (int s = 0; s < S; ++s) {
        float prev = __shfl_up_sync(MLSP_FULL, v[s], 1);
        if (s > 0) {
            const float last = __shfl_sync(MLSP_FULL, v[s - 1], 31);
            if (lane == 0) prev = last;
        }
        // order across position e not certified: the gap does not exceed the error bounds (inf - x, inf - inf: false)
        link[s] = (s > 0 || lane > 0) && (v[s] - prev <= gap_max);
        int m = link[s] ? 0 : s * 32 + lane;                             // inclusive max-scan = cluster start
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(MLSP_FULL, m, o);
            if (lane >= o) m = max(m, t);
        }
        cs[s] = max(m, carry);
        carry = __shfl_sync(MLSP_FULL, cs[s], 31);
    }
    bool amb[S];
    unsigned need[S];
    int total = 0;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        bool nxt = __shfl_down_sync(MLSP_FULL, link[s], 1);
        if (s + 1 < S) {
            const bool first = __shfl_sync(MLSP_FULL, link[s + 1], 0);
            if (lane == 31) nxt = first;
        } else if (lane == 31) {
            nxt = false;
        }
        amb[s] = (link[s] || nxt) && cs[s] < k;
        need[s] = __ballot_sync(MLSP_FULL, amb[s]);
        total += __popc(need[s]);
    }
    if (total) {                                                         // warp-uniform
        // compact the amb entries: rank -> candidate index in shared memory, eight per pass, results back by rank
        int rank[S], off = 0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            rank[s] = off + __popc(need[s] & ((1u << lane) - 1u));
            if (amb[s]) sj[rank[s]] = (uint16_t)((uint32_t)key[s] & 0xffffu);
            off += __popc(need[s]);
        }
        __syncwarp();
        const float4 *xig = reinterpret_cast<const float4 *>(P.xt + (size_t)row * C);
        for (int t0 = 0; t0 < total; t0 += 8) {
            const int t = min(t0 + g, total - 1);
            const int j = (int)sj[t];
            const float4 *xj = reinterpret_cast<const float4 *>(P.xt + ((size_t)base + j) * C);
            const float xxj = P.xx[base + j];
            float pa = 0.0f, pb = 0.0f;                // chains t = u and t = u + 4
            // four float4 pieces of both rows at a time (C = 128: two rounds) keeps the kernel at 64 registers;
            // each chain still sees its channels in ascending order
#pragma unroll
            for (int m0 = 0; m0 < M; m0 += 4) {
                float4 xr[4], q[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    xr[m] = __ldg(xig + 4 * (m0 + m) + u);
                    q[m] = __ldg(xj + 4 * (m0 + m) + u);
                }
#pragma unroll
                for (int m = 0; m < 4; m += 2) {
                    pa = __fmaf_rn(xr[m].x, q[m].x, pa);
                    pa = __fmaf_rn(xr[m].y, q[m].y, pa);
                    pa = __fmaf_rn(xr[m].z, q[m].z, pa);
                    pa = __fmaf_rn(xr[m].w, q[m].w, pa);
                    pb = __fmaf_rn(xr[m + 1].x, q[m + 1].x, pb);
                    pb = __fmaf_rn(xr[m + 1].y, q[m + 1].y, pb);
                    pb = __fmaf_rn(xr[m + 1].z, q[m + 1].z, pb);
                    pb = __fmaf_rn(xr[m + 1].w, q[m + 1].w, pb);
                }
            }
            float acc = __fadd_rn(pa, pb);                                        // q_u = p_u + p_{u+4}
            acc = __fadd_rn(acc, __shfl_xor_sync(MLSP_FULL, acc, 2));             // q0+q2 | q1+q3
            acc = __fadd_rn(acc, __shfl_xor_sync(MLSP_FULL, acc, 1));             // (q0+q2) + (q1+q3)
            if (u == 0 && t0 + g < total) se[t] = __fsub_rn(__fmaf_rn(2.0f, acc, -xxj), xxi);
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < S; ++s) {
            // (cluster start | exact value, best first | index); entries outside amb clusters keep their position
            const uint32_t j = (uint32_t)key[s] & 0xffffu;
            const uint32_t sub = amb[s] ? ~f32_orderable(__fadd_rn(se[rank[s]], 0.0f)) : 0u;
            key[s] = (key[s] == ~0ull) ? ~0ull : ((unsigned long long)cs[s] << 48) | ((unsigned long long)sub << 16) | j;
        }
        warp_sort_u64<S>(key);
        __syncwarp();
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int e = s * 32 + lane;
        if (e < k) P.idx[(size_t)row * k + e] = (int64_t)((uint32_t)key[s] & 0xffffu);
        if (s < KS) nbr[s] = (uint32_t)key[s] & 0xffffu;                  // rank e = s*32 + lane, for the fused gather
    }
}

// a2 fused into the refine kernel (get_graph_feature with idx=None): the warp that ranked row i writes the row's
// k x 2C edge features [x_j - x_i | x_i] straight away -- the latency-bound ranking of some warps overlaps the
// write stream of others, and idx is not read back.  Same lane layout as edge_fwd_vec_kernel (edge.cu): lanes span
// the 2C channels in float4, four neighbour rows in flight, evict-first 128-bit stores.  nbr[s] on lane l = the
// neighbour of rank s*32 + l.
template <int C, int KS>
__device__ __forceinline__ void gather_row(const KtParams &P, uint32_t row, uint32_t base, const uint32_t (&nbr)[KS])
{
    constexpr int Q4 = C / 4, W = 2 * Q4;                 // float4 per point row / per output row
    static_assert(C == 64 || C == 128, "gather_row: C");
    const int lane = threadIdx.x & 31, k = P.k;
    const float4 *xtb = reinterpret_cast<const float4 *>(P.xt) + (size_t)base * Q4;
    const float4 *ctr_row = reinterpret_cast<const float4 *>(P.xt) + (size_t)row * Q4;
    float4 *o = P.edge_out + (size_t)row * k * W + lane;
    // C = 64: one output row is 32 float4 -- lanes 0..15 hold the difference half, 16..31 the centre half.
    // C = 128: 64 float4 -- every lane holds one difference slot (q = lane) and one centre slot (q = 32 + lane).
    const bool is_diff = (C == 128) || lane < Q4;
    const uint32_t qc = (C == 128) ? lane : (lane & (Q4 - 1));
    const float4 ctr = __ldg(ctr_row + qc);
    // out = nb - sub with (nb, sub) = (x_j, x_i) on difference lanes and (x_i, 0) on centre lanes: x - 0 == x exactly
    const float4 sub = is_diff ? ctr : make_float4(0.f, 0.f, 0.f, 0.f);
    // row (within the cloud) a lane reads for the neighbour of rank e (warp-uniform e, a register shuffle): x_j on
    // difference lanes, x_i again on centre lanes (an L1 hit) -- one unconditional load, no per-lane copies of ctr
    const uint32_t self = row - base;
    auto rank = [&](int e) -> uint32_t {
        uint32_t n = __shfl_sync(MLSP_FULL, nbr[0], e & 31);
        if (KS > 1) {
            const uint32_t n1 = __shfl_sync(MLSP_FULL, nbr[KS - 1], e & 31);
            if (e >> 5) n = n1;
        }
        return is_diff ? n : self;
    };
    int j = 0;
    for (; j + 4 <= k; j += 4, o += 4 * W) {               // four neighbour rows in flight, no predicates
        float4 nb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            nb[u] = __ldg(xtb + (rank(j + u) * Q4 + qc));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            st_stream_f4(o + u * W, make_float4(__fsub_rn(nb[u].x, sub.x), __fsub_rn(nb[u].y, sub.y),
                                                __fsub_rn(nb[u].z, sub.z), __fsub_rn(nb[u].w, sub.w)));
            if (C == 128) st_stream_f4(o + u * W + 32, ctr);
        }
    }
    for (; j < k; ++j, o += W) {
        const float4 nb = __ldg(xtb + (rank(j) * Q4 + qc));
        st_stream_f4(o, make_float4(__fsub_rn(nb.x, sub.x), __fsub_rn(nb.y, sub.y), __fsub_rn(nb.z, sub.z),
                                    __fsub_rn(nb.w, sub.w)));
        if (C == 128) st_stream_f4(o + 32, ctr);
    }
}

// exact streaming top-k of one row by one warp, for the (rare) rows whose candidate list overflowed
// (duplicate-heavy or degenerate clouds): point-major rows, the pinned dot order, selection of topk.cuh
template <int KSLOTS>
__device__ __noinline__ void knn_row_exact(const float *__restrict__ xt, const float *__restrict__ xx, long long row,
                                           int N, int C, int k, int64_t *__restrict__ idx)
{
    const int lane = threadIdx.x & 31;
    const size_t base = (size_t)(row / N) * N;
    const int C4 = C / 4;
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

template <int NG, int C>
__global__ void __launch_bounds__(32 * RF_WARPS, 4)   // <= 64 registers: four CTAs per SM
knn_refine_kernel(KtParams P, long long total_rows)
{
    constexpr int CAP = 2 * NG;
    constexpr int SLOTS = CAP / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long row64 = (long long)blockIdx.x * RF_WARPS + warp;   // b*N + i
    if (row64 >= total_rows) return;
    const uint32_t row = (uint32_t)row64;                              // B*N < 2^31 (knn_tensor_supported)
    const uint32_t base = row / (uint32_t)P.N * (uint32_t)P.N;         // first row of the cloud
    // The two ends of the row's list are fetched SPECULATIVELY together with the two counters (one L2 round trip
    // instead of two dependent ones): slot (h, lane) of the front end is valid iff h*32+lane < c0, of the back end iff
    // h*32+lane < c1.  Counts above CAP/2 at one end (rare) re-read that end below.
    constexpr int HS = CAP / 64;
    const uint2 *list = P.cand + (size_t)row * CAP;
    const int2 cc = *reinterpret_cast<const int2 *>(P.cand_cnt + 2 * (size_t)row);
    uint2 fa[HS], fb[HS];
#pragma unroll
    for (int h = 0; h < HS; ++h) {
        fa[h] = list[h * 32 + lane];
        fb[h] = list[CAP - 1 - (h * 32 + lane)];
    }
    const int c0 = cc.x, c1 = cc.y;
    const int cnt = c0 + c1;
    constexpr int KS = NG / 32;                            // slots holding the k <= NG ranked neighbours
    uint32_t nbr[KS];
    if (c0 > CAP || c1 > CAP || cnt > CAP || cnt < P.k) {
        knn_row_exact<NG / 32>(P.xt, P.xx, row64, P.N, C, P.k, P.idx);   // warp-uniform
        if (lane == 0) atomicAdd(P.fb_count, 1);
        if (P.edge_out) {
            __syncwarp();                                  // the row's idx, written by this warp, is visible to it
#pragma unroll
            for (int s = 0; s < KS; ++s) nbr[s] = (s * 32 + lane < P.k) ? (uint32_t)P.idx[(size_t)row * P.k + s * 32 + lane] : 0u;
            gather_row<C, KS>(P, row, base, nbr);
        }
        return;
    }
    __shared__ uint16_t sj_all[RF_WARPS][CAP];
    __shared__ float se_all[RF_WARPS][CAP];
    uint16_t *sj = sj_all[warp];
    float *se = se_all[warp];
    __shared__ uint2 sc_all[RF_WARPS][CAP];
    uint2 *sc = sc_all[warp];
    if (c0 <= 32 * HS && c1 <= 32 * HS) {                  // warp-uniform: compact the two ends into list order
#pragma unroll
        for (int h = 0; h < HS; ++h) {
            const int e = h * 32 + lane;
            if (e < c0) sc[e] = fa[h];
            if (e < c1) sc[c0 + e] = fb[h];
        }
    } else {
        for (int e = lane; e < cnt; e += 32) sc[e] = list[e < c0 ? e : CAP - 1 - (e - c0)];
    }
    __syncwarp();
    unsigned long long key[SLOTS];
    float xpre[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int e = s * 32 + lane;
        key[s] = ~0ull;
        xpre[s] = 0.0f;
        if (e < cnt) {
            const uint2 ent = sc[e];
            key[s] = ((unsigned long long)f32_orderable(__fadd_rn(__uint_as_float(ent.x), 0.0f)) << 32) | (ent.y << 16) | (uint32_t)e;
            xpre[s] = P.xx[base + ent.y];                  // consumed after the sort
        }
    }
    if (cnt <= 32) {                                       // warp-uniform, the usual case for k <= 20
        unsigned long long k1[1] = {key[0]};
        const float x1[1] = {xpre[0]};
        refine_sorted<1, C, KS>(P, row, base, k1, x1, sj, se, nbr);
    } else if (SLOTS > 2 && cnt <= 64) {
        unsigned long long k2[2] = {key[0], key[1]};
        const float x2[2] = {xpre[0], xpre[1]};
        refine_sorted<2, C, KS>(P, row, base, k2, x2, sj, se, nbr);
    } else {
        refine_sorted<SLOTS, C, KS>(P, row, base, key, xpre, sj, se, nbr);
    }
    if (lane == 0) atomicAdd(P.stats, 1);
    if (P.edge_out) gather_row<C, KS>(P, row, base, nbr);
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

static int make_map(CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows)
{
    EncodeTiledFn fn = encode_fn();
    MLSP_REQUIRE(fn, MLSP_ECUDA, "knn: cuTensorMapEncodeTiled not available");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)KT_KBLK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MLSP_REQUIRE(r == CUDA_SUCCESS, MLSP_ECUDA, "knn: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return MLSP_OK;
}

struct KtLayout {
    size_t off_counters, off_mode, off_xx, off_hi, off_lo, off_xt, off_cand, off_cnt, total;
};

static KtLayout kt_layout(int B, int C, int N, int k)
{
    const int CAP = (k <= 32) ? 64 : 128;
    KtLayout L;
    size_t o = 0;
    L.off_counters = o; o += 256;                                            // [0] fb_count, [1] certified rows
    L.off_mode = o;     o += align_up(sizeof(int) * (size_t)B, 256);         // per-cloud pass-1 choice
    L.off_xx = o;       o += align_up(sizeof(float) * (size_t)B * N, 256);
    L.off_hi = o;       o += align_up(2 * (size_t)B * N * C, 1024);
    L.off_lo = o;       o += align_up(2 * (size_t)B * N * C, 1024);
    L.off_xt = o;       o += align_up(sizeof(float) * (size_t)B * N * C, 256);
    L.off_cand = o;     o += align_up(sizeof(uint2) * (size_t)B * N * CAP, 256);
    L.off_cnt = o;      o += align_up(2 * sizeof(int) * (size_t)B * N, 256);
    L.total = o;
    return L;
}

bool knn_tensor_supported(int B, int C, int N, int k)
{
    return (C == 64 || C == 128) && N >= 256 && N <= 65535 && k <= 64 && (long long)B * N < (1ll << 31) && B <= 65535;
}

size_t knn_tensor_workspace_bytes(int B, int C, int N, int k) { return kt_layout(B, C, N, k).total; }

// the point-major copy xt (B,N,C) the tensor path leaves in its workspace (reused by the fused edge gather)
const float *knn_tensor_xt(const void *ws, int B, int C, int N, int k)
{
    return reinterpret_cast<const float *>(static_cast<const char *>(ws) + kt_layout(B, C, N, k).off_xt);
}

// measurement hook (mlsp_graph_feature_fwd_stage): which of the three kernels a call launches; 7 = all
thread_local int g_kt_stages = 7;

int knn_tensor_run(const float *x, int B, int C, int N, int k, int64_t *idx, void *ws, float *dump, float *edge_out,
                   cudaStream_t st)
{
    const KtLayout L = kt_layout(B, C, N, k);
    char *w = static_cast<char *>(ws);
    int *counters = reinterpret_cast<int *>(w + L.off_counters);
    int *cloud_mode = reinterpret_cast<int *>(w + L.off_mode);
    float *xx = reinterpret_cast<float *>(w + L.off_xx);
    __nv_bfloat16 *hi = reinterpret_cast<__nv_bfloat16 *>(w + L.off_hi);
    __nv_bfloat16 *lo = reinterpret_cast<__nv_bfloat16 *>(w + L.off_lo);
    float *xt = reinterpret_cast<float *>(w + L.off_xt);

    if (g_kt_stages & 1) {
        knn_prep_kernel<<<dim3((N + PREP_PTS - 1) / PREP_PTS, B), PREP_THREADS, sizeof(float) * C * (PREP_PTS + 1), st>>>(
            x, C, N, xx, counters, xt, hi, lo, cloud_mode);
        MLSP_LAUNCH_CHECK("knn_prep_kernel");
    }

    CUtensorMap map_hi, map_lo, half_hi, half_lo;       // boxes of 128 rows (A tiles, unpaired B blocks) and of 64 rows (paired)
    int rc = make_map(&map_hi, hi, (uint64_t)B * N, (uint64_t)C, KT_COLS);
    if (rc) return rc;
    rc = make_map(&map_lo, lo, (uint64_t)B * N, (uint64_t)C, KT_COLS);
    if (rc) return rc;
    rc = make_map(&half_hi, hi, (uint64_t)B * N, (uint64_t)C, KT_COLS / 2);
    if (rc) return rc;
    rc = make_map(&half_lo, lo, (uint64_t)B * N, (uint64_t)C, KT_COLS / 2);
    if (rc) return rc;

    KtParams P;
    P.xx = xx; P.xt = xt; P.idx = idx; P.fb_count = counters; P.cloud_mode = cloud_mode;
    P.stats = counters + 1; P.dump = dump; P.edge_out = reinterpret_cast<float4 *>(edge_out);
    P.cand = reinterpret_cast<uint2 *>(w + L.off_cand); P.cand_cnt = reinterpret_cast<int *>(w + L.off_cnt); P.N = N; P.C = C; P.k = k; P.T = (N + KT_COLS - 1) / KT_COLS;
    const int NG = (k <= 32) ? 32 : 64;
    // ring depth: as deep as two CTAs per SM allow (NG = 32), else one CTA per SM with up to 8 stages
    const size_t per_sm = 227 * 1024, reserved = 1024;
    const size_t budget2 = per_sm / 2 - reserved;
    int stages = 0;
    if (NG == 32)
        for (int s_ = 2; s_ <= KT_MAX_STAGES; ++s_)
            if (kt_smem_bytes(C, k, NG, s_) <= budget2) stages = s_;
    if (stages == 0)
        for (int s_ = 2; s_ <= KT_MAX_STAGES; ++s_)
            if (kt_smem_bytes(C, k, NG, s_) <= per_sm - reserved) stages = s_;
    P.stages = stages;
    if (const char *e = getenv("MLSP_KT_STAGES")) P.stages = atoi(e);   // tuning hook
    P.mode = 0;
    if (const char *e = getenv("MLSP_KT_MODE")) P.mode = atoi(e);
    MLSP_REQUIRE(P.stages >= 2 && P.stages <= KT_MAX_STAGES && kt_smem_bytes(C, k, NG, P.stages) <= per_sm - reserved,
                 MLSP_EUNSUPPORTED, "knn: no shared-memory configuration for C=%d k=%d", C, k);
    const size_t smem = kt_smem_bytes(C, k, NG, P.stages);
    dim3 grid((N + KT_ROWS - 1) / KT_ROWS, B);
    // MLSP_KT_CLUSTER=1 (tuning hook): clusters of two adjacent row blocks of a cloud share every candidate block by TMA
    // multicast (needs an even number of row blocks per cloud).  Off by default: halving the L2 -> SM operand traffic
    // changed nothing (profiles/kt_ablate_r1h.log), i.e. the MMA pipeline is not fed-bound.
    P.pair = 0;
    if (const char *e = getenv("MLSP_KT_CLUSTER")) P.pair = (atoi(e) != 0 && grid.x % 2 == 0) ? 1 : 0;
    if (g_kt_stages & 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(KT_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = P.pair ? 2 : 1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (NG == 32) {
            MLSP_CUDA(cudaFuncSetAttribute(knn_tensor_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            MLSP_CUDA(cudaLaunchKernelEx(&cfg, knn_tensor_kernel<32>, map_hi, map_lo, half_hi, half_lo, P));
        } else {
            MLSP_CUDA(cudaFuncSetAttribute(knn_tensor_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            MLSP_CUDA(cudaLaunchKernelEx(&cfg, knn_tensor_kernel<64>, map_hi, map_lo, half_hi, half_lo, P));
        }
    }
    MLSP_LAUNCH_CHECK("knn_tensor_kernel");
    if (g_kt_stages & 4) {
        const long long rows_total = (long long)B * N;
        const unsigned rblocks = (unsigned)((rows_total + RF_WARPS - 1) / RF_WARPS);
        if (NG == 32 && C == 64)
            knn_refine_kernel<32, 64><<<rblocks, 32 * RF_WARPS, 0, st>>>(P, rows_total);
        else if (NG == 32)
            knn_refine_kernel<32, 128><<<rblocks, 32 * RF_WARPS, 0, st>>>(P, rows_total);
        else if (C == 64)
            knn_refine_kernel<64, 64><<<rblocks, 32 * RF_WARPS, 0, st>>>(P, rows_total);
        else
            knn_refine_kernel<64, 128><<<rblocks, 32 * RF_WARPS, 0, st>>>(P, rows_total);
        MLSP_LAUNCH_CHECK("knn_refine_kernel");
    }
    return MLSP_OK;
}

}  // namespace mlsp
