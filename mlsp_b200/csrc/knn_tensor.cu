// knn_tensor.cu -- a1 for the DGCNN feature layers (C = 64, 128): the pairwise-distance contraction on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA) with a fused
// selection epilogue, and results that are STILL bit-exact with the fp32 specification of knn.cu.
//
// Filter (tensor cores) + refine (exact fp32) + certificate:
//   1. prep: x (B,C,N) fp32 -> point-major fp32 rows xt (B,N,C); the cloud is re-centred (y = x - c, c = mean of a strided
//      sample of its points: distances are translation invariant, the filter's error bounds scale with |y_i||y_j|) and split
//      into bf16 pieces y = hi + lo (+ 2^-18 |y|).  Per point also the column term t_j = |x_j|^2 - 2 c.y_j - |c|^2 (the
//      specification's own norm, so its rounding cancels), stored as three bf16 pieces of -t_j/2.
//   2. filter kernel, one CTA per 128 query rows (one CTA per SM), candidate tiles of 128 (UMMA 128x128x16, kind::f16):
//        warp 0    : TMA producer (A = [hi|lo] of the rows once; per candidate block 16 KiB + the 2 KiB norm tile)
//        warp 1    : TMEM allocator (all 512 columns = four accumulator buffers) + single-thread MMA issuer
//        warps 2-17: epilogue, FOUR threads per query row (tcgen05.ld 32x32b.x32: TMEM lane == row, 32 columns each)
//      Every tile starts with one extra K=16 MMA  ones(128x16) . norms(128x16)^T  that initialises the accumulator
//      with -t_j/2, so the accumulator IS the ranking value: acc = y_i.y_j - t_j/2 = -(|x_j|^2 - 2 x_i.x_j)/2 + row const.
//      pass 1 (hi.hi only): per thread the maximum of every column class in registers -- one FMNMX3 per TWO elements;
//              tau = k-th largest class maximum of the row (thread-local sort + 4-way merge) bounds the k-th distance.
//      pass 2 (hi.hi + lo.hi + hi.lo): columns with acc >= tau - slack are appended to the row's candidate list in
//              SHARED memory (32-bit addresses): setp + lop3 + predicated st.shared + predicated add per element, the
//              column index packed into the 5 low mantissa bits, the tile recovered from per-tile cursor snapshots.
//              The lists are copied to global memory (L2) once at the end.
//   3. refine kernel (one warp per row): decodes the list, sorts it by the approximate value.  Two neighbours of that
//      order whose values differ by more than 2 eps are provably in the same order in exact fp32 arithmetic, so only
//      runs of near-ties ("clusters") that start inside the first k positions need the exact value: for those the
//      distance is recomputed with the pinned fp32 chain of the specification (four lanes per candidate, coalesced
//      reads of the point-major rows) and the cluster is re-sorted by (value desc, index asc).
//      A row whose list overflowed is not certified and goes to the exact streaming top-k (topk.cuh) in the same warp.
// SASS evidence: UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA) -- profiles/.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "topk.cuh"

namespace mlsp {

constexpr int KT_ROWS = 128;     // query rows per CTA   (UMMA M, TMEM lanes)
constexpr int KT_COLS = 256;     // candidates per tile  (UMMA N, TMEM columns per accumulator buffer): at N = 128 an MMA
                                 // retires in 64 clocks, about what its ISSUE costs the single issuing thread (measured: the
                                 // issue section of a 5-MMA tile took 650 clocks) -- N = 256 makes the tensor pipe the bound
constexpr int KT_KBLK = 64;      // bf16 per K block = one 128-byte swizzle span
constexpr int KT_MAX_STAGES = 8;
constexpr int KT_TPR = 4;        // epilogue threads per query row: each owns a 32-column quarter of every tile
constexpr int KT_EPI_WARPS = 4 * KT_TPR;
constexpr int KT_EPI_THREADS = 32 * KT_EPI_WARPS;
constexpr int KT_THREADS = 64 + KT_EPI_THREADS;
constexpr int KT_NBUF = 2;       // accumulator buffers: 2 x 256 columns = the whole TMEM of the SM
constexpr uint32_t KT_A_BYTES = KT_ROWS * KT_KBLK * 2;    // 16 KiB per (128 x 64) bf16 block of the resident A operand
constexpr uint32_t KT_BLK_BYTES = KT_COLS * KT_KBLK * 2;  // 32 KiB per (256 x 64) bf16 candidate block
constexpr uint32_t KT_NB_BYTES = KT_COLS * 16;            // 4 KiB: one K-group (8 bf16) of the norm operand per candidate
constexpr uint32_t KT_ONES_BYTES = 2048 + KT_NB_BYTES;    // [ones: 128 rows x 16 B | zeros: the second K group of both norm operands]
constexpr uint32_t KT_STAGE_BYTES = KT_BLK_BYTES + KT_NB_BYTES;
constexpr int KT_MAX_TILES = 32; // snapshots in shared memory: N <= 8192
// Error model of the filter value v~ = -2 acc against the specification value:
//   |v~_ij - (row const_i) - (-spec_ij)| <= eps_ij = KT_EPS_REL |y_i||y_j| + KT_EPS_YY (|y_i|^2+|y_j|^2) + KT_RND(C) (|x_i|^2+|x_j|^2)
// KT_EPS_REL: dropped lo.lo and split residuals 3*2^-16 on the dot product (Cauchy-Schwarz) = 9.2e-5 on v, the 5 mantissa
//   bits given to the column index 2^-16, the tensor core's fp32 accumulation; 2^-12 = 2.4e-4 leaves > 2x.
// KT_EPS_YY: rounding / packing of the column term (magnitude |y_j|^2/2 in the accumulator).
// KT_RND(C) = (C/8 + 8) 2^-24: the specification's own roundings (8 fma chains of C/8 terms, the butterfly, two
//   subtractions), which act on the UNCENTRED norms.
constexpr float KT_EPS_REL = 2.44140625e-4f;
constexpr float KT_EPS_YY = 3.0517578125e-5f;             // 2^-15
// Pass 1 multiplies the bf16 heads only (dot~1 = hi.hi): y = hi + r with |r_c| <= 2^-9 |y_c|, so the pass-1 value is
// within KT_EPS1_REL |y_i||y_j| (= 2^-7 + 2^-14) of the pass-2 value; tau1 + eps1 still upper-bounds the k-th distance.
constexpr float KT_EPS1_REL = 0.00787353515625f;
__host__ __device__ __forceinline__ float kt_rnd(int C) { return (float)(C / 8 + 8) * 5.9604644775390625e-8f; }

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
// ---- CTA pair (cta_group::2): the two CTAs of a cluster run ONE MMA of M = 256 -- each holds its own 128 rows of A and
// HALF of every B block (64 of the 128 candidates), so the operand stream every SM has to ingest halves.  Only the
// leader (cluster rank 0) issues MMAs and owns the full / tm_empty barriers; addresses of its barriers come from mapa.
__device__ __forceinline__ uint32_t mapa_rank0(const void *p)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(smem_u32(p)));
    return r;
}
// the box lands in THIS CTA's shared memory, complete_tx is signalled on the barrier `bar_cluster` (a shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *map, int c0, int c1, uint32_t bar_cluster)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster)
{
    // relaxed: the accumulator reads this orders are TMEM reads, fenced by tcgen05.fence::before_thread_sync; a release at
    // cluster scope costs a full memory fence + L1 invalidation per arrive
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t *bar)      // arrives on the barrier at this offset in BOTH CTAs
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// shared -> global bulk copy by the TMA engine (one thread), and the wait that makes the source reusable
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
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
// One lane of a converged warp (warp-uniform predicate: code under it keeps its operands in uniform registers -- under a
// plain `lane == 0` test the compiler wraps every tcgen05.mma / TMA instruction in an R2UR waterfall loop of ~20
// instructions, which made the single issuing thread the bottleneck of the whole kernel)
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
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
// asynchronous 32-column load: the registers are valid only after tc_wait32 on the same array (the "+r" operands
// make every later use of the values depend on the wait, so the compiler cannot hoist them above it)
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&u)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
          "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
          "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
          "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_wait32(uint32_t (&u)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]),
                   "+r"(u[8]), "+r"(u[9]), "+r"(u[10]), "+r"(u[11]), "+r"(u[12]), "+r"(u[13]), "+r"(u[14]), "+r"(u[15]),
                   "+r"(u[16]), "+r"(u[17]), "+r"(u[18]), "+r"(u[19]), "+r"(u[20]), "+r"(u[21]), "+r"(u[22]), "+r"(u[23]),
                   "+r"(u[24]), "+r"(u[25]), "+r"(u[26]), "+r"(u[27]), "+r"(u[28]), "+r"(u[29]), "+r"(u[30]), "+r"(u[31])
                 :
                 : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(KT_EPI_THREADS) : "memory"); }
// MMA warp + epilogue warps (everything but the TMA producer, which is already streaming)
__device__ __forceinline__ void setup_bar_sync() { asm volatile("bar.sync 2, %0;" ::"n"(KT_EPI_THREADS + 32) : "memory"); }

// K-major, 128-byte swizzled operand block (rows 128 B apart, 8-row groups 1024 B apart), sm_100 version bit
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// K-major operand without swizzle: 8-row x 16-byte core matrices, `sbo` bytes between 8-row groups, `lbo` bytes between
// the two K groups of one K = 16 instruction
__device__ __forceinline__ uint64_t umma_desc_plain(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// kind::f16: D = f32, A = B = bf16, both K-major, N = 128, M = 128
constexpr uint32_t KT_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((KT_COLS >> 3) << 17) | ((KT_ROWS >> 4) << 24);
constexpr uint32_t KT_IDESC_PAIR = (1u << 4) | (1u << 7) | (1u << 10) | ((KT_COLS >> 3) << 17) | (((2 * KT_ROWS) >> 4) << 24);   // M = 256

// ------------------------------------------------------------------------------------------- programmatic dependent launch
// The four kernels of a call (centre -> prep -> filter -> ranking) are chained with programmatic stream serialization: every
// kernel lets its successor launch as soon as all of its own CTAs have started (pdl_launch_dependents) and the successor runs
// whatever does not depend on the predecessor -- slab loads, barrier initialisation, tensor-map prefetch -- before pdl_wait,
// which returns when the predecessor grid has completed and its writes are visible.  A kernel launched without the attribute
// passes pdl_wait immediately.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// mlsp_knn_set_pdl: bit 0 centre -> prep, bit 1 prep -> filter, bit 2 filter -> ranking; 0 = plain stream order
static int g_kt_pdl = 3;
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------------------------------- centre kernel
// c (B,C) = mean of KT_CEN_SAMPLES points of every cloud, taken at a fixed stride over the cloud (a representative
// sample whatever the point order).  Any c works for correctness -- distances are translation invariant -- a c inside
// the cloud makes |y| = |x - c| comparable to the cloud's spread, which is what the filter's error bounds scale with.
constexpr int KT_CEN_SAMPLES = 64;

__global__ void __launch_bounds__(256)
knn_centre_kernel(const float *__restrict__ x, int C, int N, float *__restrict__ cen)
{
    __shared__ float part[256];
    pdl_launch_dependents();                               // the prep kernel may start staging its slabs
    const int b = blockIdx.x, t = threadIdx.x;
    const int parts = 256 / C;                             // 2 (C = 128) or 4 (C = 64)
    const int c = t % C, pr = t / C;
    const int stride = N / KT_CEN_SAMPLES;                 // N >= 256
    const float *row = x + ((size_t)b * C + c) * N;
    // all of a thread's samples are loaded before the first add: one DRAM latency instead of a chain of 16-32
    float v[KT_CEN_SAMPLES / 2];
#pragma unroll
    for (int u = 0; u < KT_CEN_SAMPLES / 2; ++u) {
        const int q = pr + u * parts;
        v[u] = (q < KT_CEN_SAMPLES) ? __ldg(row + (size_t)q * stride) : 0.0f;
    }
    float s = 0.0f;
#pragma unroll
    for (int u = 0; u < KT_CEN_SAMPLES / 2; ++u) s = __fadd_rn(s, v[u]);
    part[t] = s;
    __syncthreads();
    if (pr == 0) {
        for (int q = 1; q < parts; ++q) s = __fadd_rn(s, part[q * C + c]);
        cen[(size_t)b * C + c] = __fmul_rn(s, 1.0f / KT_CEN_SAMPLES);
    }
}

// ------------------------------------------------------------------------------------------- prep kernel
// One pass over x (B,C,N).  Per cloud the centre c comes from knn_centre_kernel.  Per point: the exact norm xx in the specification order (sequential adds over the channels), the
// centred norm yy = |y|^2 (error bounds), the row shift ss = 2 c.y + |c|^2 and the column term t = xx - ss (fp64, one
// rounding) whose half, negated, goes into the norm operand nb as three bf16 pieces; the point-major fp32 rows xt
// (B,N,C) of the UNCENTRED cloud (exact re-rank, edge gather) and the bf16 split hi/lo (B*N, C) of y.
// A CTA stages a [C][PREP_PTS] slab in shared memory (coalesced 128-byte reads along n), then warps write whole point
// rows (coalesced along c).  Also zeroes the two diagnostic counters.
constexpr int PREP_PTS = 32;
constexpr int PREP_THREADS = 256;

__global__ void __launch_bounds__(PREP_THREADS)
knn_prep_kernel(const float *__restrict__ x, const float *__restrict__ cen, int C, int N, float *__restrict__ xx, float *__restrict__ yy,
                float *__restrict__ ss, int *__restrict__ counters, float *__restrict__ xt,
                __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo, uint4 *__restrict__ nb)
{
    pdl_launch_dependents();
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 4) counters[threadIdx.x] = 0;
    extern __shared__ float slab[];                       // [C][PREP_PTS + 1] | cvec [C]
    float *cvec = slab + C * (PREP_PTS + 1);
    const int b = blockIdx.y, n0 = blockIdx.x * PREP_PTS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *xb = x + (size_t)b * C * N;
    for (int c = warp; c < C; c += PREP_THREADS / 32) {   // independent of the centre kernel: runs beside it
        const int n = n0 + lane;
        slab[c * (PREP_PTS + 1) + lane] = (n < N) ? xb[(size_t)c * N + n] : 0.0f;
    }
    pdl_wait();                                           // the centres are complete and visible
    if ((int)threadIdx.x < C) cvec[threadIdx.x] = cen[(size_t)b * C + threadIdx.x];
    __syncthreads();
    if (warp == 0) {                                       // one point per lane, channels in order
        float s = 0.0f, q = 0.0f;
        double cy = 0.0, cc = 0.0;
        for (int c = 0; c < C; ++c) {
            const float v = slab[c * (PREP_PTS + 1) + lane], cv = cvec[c];
            const float y = __fsub_rn(v, cv);
            s = (c == 0) ? __fmul_rn(v, v) : __fadd_rn(s, __fmul_rn(v, v));
            q = __fmaf_rn(y, y, q);
            cy += (double)cv * (double)y;
            cc += (double)cv * (double)cv;
        }
        const int n = n0 + lane;
        if (n < N) {
            const double shift = 2.0 * cy + cc;
            const float t = (float)((double)s - shift);
            const float hneg = -0.5f * t;
            const __nv_bfloat16 h1 = __float2bfloat16_rn(hneg);
            const float r1 = hneg - __bfloat162float(h1);
            const __nv_bfloat16 h2 = __float2bfloat16_rn(r1);
            const __nv_bfloat16 h3 = __float2bfloat16_rn(r1 - __bfloat162float(h2));
            const size_t o = (size_t)b * N + n;
            xx[o] = s;
            yy[o] = q;
            ss[o] = (float)shift;
            nb[o] = make_uint4((uint32_t)__bfloat16_as_ushort(h1) | ((uint32_t)__bfloat16_as_ushort(h2) << 16),
                               (uint32_t)__bfloat16_as_ushort(h3), 0u, 0u);
        }
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
        const float y0 = __fsub_rn(v0, cvec[c]), y1 = __fsub_rn(v1, cvec[c + 1]);
        const __nv_bfloat16 h0 = __float2bfloat16_rn(y0), h1 = __float2bfloat16_rn(y1);
        __nv_bfloat162 hv, lv;
        hv.x = h0; hv.y = h1;
        lv.x = __float2bfloat16_rn(y0 - __bfloat162float(h0));
        lv.y = __float2bfloat16_rn(y1 - __bfloat162float(h1));
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

// thread-local bitonic sort of NG registers, DESCENDING
template <int NG>
__device__ __forceinline__ void reg_sort_desc(float (&v)[NG])
{
#pragma unroll
    for (int size = 2; size <= NG; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int e = 0; e < NG; ++e) {
                const int p = e ^ stride;
                if (p > e) {
                    const bool down = (e & size) == 0;
                    const float a = v[e], b = v[p];
                    const float mn = fminf(a, b), mx = fmaxf(a, b);
                    v[e] = down ? mx : mn;
                    v[p] = down ? mn : mx;
                }
            }
        }
    }
}

// in-register bitonic merge: v is bitonic (descending then ascending) on entry, sorted descending on exit
template <int L>
__device__ __forceinline__ void bitonic_merge_desc(float (&v)[L])
{
#pragma unroll
    for (int stride = L / 2; stride > 0; stride >>= 1) {
#pragma unroll
        for (int e = 0; e < L; ++e) {
            if ((e & stride) == 0) {
                const float a = v[e], b = v[e + stride];
                v[e] = fmaxf(a, b);
                v[e + stride] = fminf(a, b);
            }
        }
    }
}

// per-pair error bound of the filter value against the specification value (constants above)
__device__ __forceinline__ float pair_eps(float yyi, float yyj, float xxi, float xxj, int C)
{
    return __fmaf_rn(KT_EPS_REL * sqrtf(yyi), sqrtf(yyj), __fmaf_rn(KT_EPS_YY, yyi + yyj, kt_rnd(C) * (xxi + xxj)));
}

// pass 2, one element: if (acc >= t) { *cur = (acc & ~63) | column; cur += step; } -- setp, lop3 (column as the immediate),
// one predicated st.shared and one predicated add on a 32-bit shared-memory address; no branch, no vote, no atomics
template <int COL, int BASE>
__device__ __forceinline__ void collect_tile(uint32_t &cur, const uint32_t (&u)[32], float t, int step, uint32_t keep)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 w;\n\t"
        "setp.ge.f32 p, %1, %2;\n\t"
        "lop3.b32 w, %3, %6, %4, 0xEA;\n\t"
        "@p st.shared.b32 [%0], w;\n\t"
        "@p add.s32 %0, %0, %5;\n\t}"
        : "+r"(cur)
        : "f"(__uint_as_float(u[COL])), "f"(t), "r"(u[COL]), "n"(BASE + COL), "r"(step), "r"(keep)
        : "memory");
    if constexpr (COL + 1 < 32) collect_tile<COL + 1, BASE>(cur, u, t, step, keep);
}

struct KtParams {
    const float *xx;         // (B,N) exact squared norms (specification order)
    const float *yy;         // (B,N) squared norms of the centred points
    const float *ss;         // (B,N) row shift 2 c.y_i + |c|^2 (only the test dump needs it)
    const float *xt;         // (B,N,C) fp32 point-major
    int64_t *idx;            // (B,N,k)
    int *fb_count;           // rows whose list overflowed (re-done with the exact streaming selection)
    int *stats;              // [0] rows certified by the tensor path, [1] total list length, [2] exact distances recomputed
    int want_stats;          // the four diagnostic counters are maintained only on request (MLSP_KNN_STATS): one same-address
                             // atomic per row serialises in L2 -- 32 k rows cost ~17 us per counter
    uint32_t *cand;          // (B*N, capw) packed candidate words: filter value with the column (mod 32) in its 5 low bits
    uint8_t *cand_cnt;       // (B*N, 4) entries written by each column quarter; 255: overflowed
    uint8_t *snap;           // (B*N, 4, TP) cursor of each quarter after every candidate tile (padding 255)
    float *dump;             // optional (2,B,N,N) filter values of pass 1 and pass 2 (tests only)
    long long *tstamp;       // optional (CTAs, 16) clock64 / globaltimer marks of the filter kernel's phases (tools/kt_timeline.py)
    float4 *edge_out;        // optional (B,N,k,2C): the refine kernel also writes the row's edge features (a2 fused)
    int N, C, k, T;          // T = candidate tiles per cloud
    int TP;                  // T rounded up to a multiple of 8
    int stages;              // depth of the B-operand smem ring
    int capw;                // list words per row: 128 (k <= 32) or 192
    int cs;                  // 2: clusters of two CTAs (adjacent row blocks of a cloud) working as a tcgen05 CTA pair; 1: unpaired
};

// ------------------------------------------------------------------------------------------- filter kernel
// shared memory: A (2*SEG blocks) | ring (STAGES x (16 KiB block + 2 KiB norm tile)) | ones (4 KiB) |
//                lists [128][capw] (aliased: sorted class maxima [4][KX][128]) | snapshots [T][512] | thr [128] | red [32] | barriers
__host__ __device__ inline size_t kt_smem_bytes(int C, int k, int T, int stages, int cs)
{
    const int capw = k <= 32 ? 128 : 192;
    return (size_t)(2 * C / KT_KBLK) * KT_A_BYTES + (size_t)stages * (KT_STAGE_BYTES / cs) + KT_ONES_BYTES + (size_t)KT_ROWS * capw * 4 +
           (size_t)T * KT_EPI_THREADS + KT_ROWS * 4 + 32 * 4 + 32 * 8 + 16 + 1024;
}

__device__ __forceinline__ void kt_mark(const KtParams &P, int slot)
{
    if (P.tstamp) {
        long long g;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
        P.tstamp[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + slot] = g;
    }
}

// NGT: column classes per thread: 16 (k <= 32) or 32 (k <= 64).  PAIR: launched in clusters of two CTAs working as a tcgen05
// CTA pair (a kernel that contains cta_group::2 instructions cannot be launched without the cluster, hence two instances).
template <int NGT, bool PAIR>
__global__ void __launch_bounds__(KT_THREADS, 1)
knn_tensor_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                  const __grid_constant__ CUtensorMap part_hi, const __grid_constant__ CUtensorMap part_lo,
                  const __grid_constant__ CUtensorMap map_nb, KtParams P)
{
    // P.cs == 2: launched in clusters of two CTAs = two adjacent row blocks of one cloud, working as a CTA pair
    // (tcgen05 cta_group::2): one MMA of M = 256 per instruction, each CTA holding its 128 rows of A and half of every
    // candidate block.  The operand stream bounds this kernel (2 x 128 MACs per streamed byte; measured 19 B/clk/SM for
    // the unpaired kernel whatever the ring depth, and TMA multicast did not change it): the pair halves it.
    constexpr bool pair = PAIR;
    const uint32_t crank = pair ? cluster_ctarank() : 0u;
    const bool leader = crank == 0;
    const uint32_t BLK = pair ? KT_BLK_BYTES / 2 : KT_BLK_BYTES;    // this CTA's part of a candidate block
    const uint32_t NBB = pair ? KT_NB_BYTES / 2 : KT_NB_BYTES;      // ... and of its norm tile
    const uint32_t STG = BLK + NBB;
    extern __shared__ uint8_t smem_dyn[];
    uint8_t *smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // swizzle atoms need 1 KiB alignment
    const int SEG = P.C / KT_KBLK;          // K blocks per hi / lo segment (1 or 2)
    const int KB = 2 * SEG;                 // B blocks streamed per pass-2 tile: hi blocks, then lo blocks
    const int STAGES = P.stages;
    const int CAPW = P.capw, HC = CAPW / 2;
    const int KX = min(P.k, NGT);
    uint8_t *sA = smem_raw;                                   // [hi blocks | lo blocks], resident
    uint8_t *sB = sA + (size_t)KB * KT_A_BYTES;               // ring: STAGES x (block | norm tile)
    uint8_t *sOnes = sB + (size_t)STAGES * (PAIR ? KT_STAGE_BYTES / 2 : KT_STAGE_BYTES);   // [K group 0: rows of (1,1,1,0,...) | K group 1: zeros]
    uint32_t *lists = reinterpret_cast<uint32_t *>(sOnes + KT_ONES_BYTES);
    float *xchg = reinterpret_cast<float *>(lists);           // between the passes only
    uint8_t *snap_s = reinterpret_cast<uint8_t *>(lists + (size_t)KT_ROWS * CAPW);
    float *thr_s = reinterpret_cast<float *>(snap_s + (((size_t)P.T * KT_EPI_THREADS + 15) & ~(size_t)15));
    float *red_s = thr_s + KT_ROWS;                           // [16] per-warp max |x|^2, [16] per-warp max |y|^2
    uint64_t *bars = reinterpret_cast<uint64_t *>(red_s + 32);
    uint64_t *full = bars, *empty = bars + KT_MAX_STAGES, *a_full = bars + 2 * KT_MAX_STAGES;
    uint64_t *tm_full = a_full + 1, *tm_empty = tm_full + KT_NBUF;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tm_empty + KT_NBUF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y, i0 = blockIdx.x * KT_ROWS;
    const int N = P.N, T = P.T;
    const int rowbase = b * N;
    if (threadIdx.x == 64) {
        kt_mark(P, 0);
        if (P.tstamp) {
            uint32_t smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            P.tstamp[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + 15] = smid;
        }
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < KT_MAX_STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(a_full, 1);
        for (int s = 0; s < KT_NBUF; ++s) {
            mbar_init(tm_full + s, 1);
            mbar_init(tm_empty + s, pair ? 2 * KT_EPI_WARPS : KT_EPI_WARPS);   // pair: both CTAs' epilogues release the leader's buffer
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&map_hi);
        tma_prefetch_desc(&map_lo);
        tma_prefetch_desc(pair ? &part_hi : &map_nb);
    }
    if (pair) cluster_sync_all();                                 // the peers' barriers exist before anything signals them
    else __syncthreads();
    pdl_launch_dependents();
    pdl_wait();                                                   // the prep kernel's operands (hi / lo / norms) are complete
    // From here the TMA producer streams; the MMA warp allocates TMEM and the epilogue warps build the constant operand
    // and reduce the cloud's norms meanwhile; those 17 warps meet at setup_bar_sync.
    if (warp == 1) {
        if constexpr (pair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        tc_fence_before();
        setup_bar_sync();
        tc_fence_after();
    }
    float xxi = 0.0f, yyi = 0.0f;                                   // epilogue threads: the norms of their row
    if (warp >= 2) {
        const int et = threadIdx.x - 64;
        const int ii = i0 + ((warp & 3) * 32 + lane);
        if (ii < N) {
            xxi = P.xx[(size_t)rowbase + ii];
            yyi = P.yy[(size_t)rowbase + ii];
        }
        if (et < (int)(KT_ONES_BYTES / 16)) {                 // the constant A operand of the norm MMA + the zero K group
            const uint4 one = make_uint4(0x3f803f80u, 0x00003f80u, 0u, 0u);   // bf16 (1, 1, 1, 0, 0, 0, 0, 0)
            reinterpret_cast<uint4 *>(sOnes)[et] = (et < KT_ROWS) ? one : make_uint4(0u, 0u, 0u, 0u);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // visible to the tensor core's reads
        }
        tc_fence_before();
        setup_bar_sync();
        tc_fence_after();
        // maxima of the cloud's norms (the row-level error bounds of the threshold, read after pass 1): their loads hide
        // behind the wait for the first accumulator
        float mx = 0.0f, my = 0.0f;
        for (int j = et; j < N; j += KT_EPI_THREADS) {
            mx = fmaxf(mx, P.xx[(size_t)rowbase + j]);
            my = fmaxf(my, P.yy[(size_t)rowbase + j]);
        }
        mx = warp_max_f32(mx);
        my = warp_max_f32(my);
        if (lane == 0) {
            red_s[warp - 2] = mx;
            red_s[16 + warp - 2] = my;
        }
    }
    const uint32_t tmem_base = (warp >= 1) ? *tmem_slot : 0u;

    if (warp == 0) {
        // ================================ TMA producer ================================
        // the whole warp runs the loop (waits included); one elected lane issues.  Pair: both CTAs load their own rows of
        // A and their half of every candidate block; every complete_tx goes to the LEADER's barrier, which expects both.
        const uint32_t a_full_c = pair ? mapa_rank0(a_full) : 0u;
        if (elect_one()) {
            if (leader) mbar_expect_tx(a_full, (uint32_t)KB * KT_A_BYTES * (pair ? 2u : 1u));
            for (int kb = 0; kb < KB; ++kb) {        // A = [hi | lo]
                if constexpr (pair)
                    tma_load_2d_pair(sA + (size_t)kb * KT_A_BYTES, kb < SEG ? &map_hi : &map_lo, (kb % SEG) * KT_KBLK,
                                     rowbase + i0, a_full_c);
                else
                    tma_load_2d(sA + (size_t)kb * KT_A_BYTES, kb < SEG ? &map_hi : &map_lo, (kb % SEG) * KT_KBLK,
                                rowbase + i0, a_full);
            }
        }
        int stage = 0;
        uint32_t ph = 0;
        for (int g = 0; g < 2 * T; ++g) {
            const int j0 = (g % T) * KT_COLS + (int)crank * (KT_COLS / 2);   // pair: my half of the tile's candidates
            const int nkb = (g < T) ? SEG : KB;                 // pass 1 multiplies hi.hi only: no lo blocks
            for (int kb = 0; kb < nkb; ++kb) {                  // B blocks: hi ..., lo ...
                uint8_t *st = sB + (size_t)stage * STG;
                mbar_wait(empty + stage, ph ^ 1);
                const uint32_t full_c = pair ? mapa_rank0(full + stage) : 0u;
                if (elect_one()) {
                    const bool nbt = kb == 0;
                    if (leader) mbar_expect_tx(full + stage, (BLK + (nbt ? NBB : 0u)) * (pair ? 2u : 1u));
                    if constexpr (pair) {
                        tma_load_2d_pair(st, kb < SEG ? &part_hi : &part_lo, (kb % SEG) * KT_KBLK, rowbase + j0, full_c);
                        if (nbt) tma_load_2d_pair(st + BLK, &map_nb, 0, rowbase + j0, full_c);
                    } else {
                        tma_load_2d(st, kb < SEG ? &part_hi : &part_lo, (kb % SEG) * KT_KBLK, rowbase + j0, full + stage);
                        if (nbt) tma_load_2d(st + BLK, &map_nb, 0, rowbase + j0, full + stage);   // the tile's norm operand
                    }
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        // the whole warp runs the loop; one elected lane issues the tcgen05 instructions (pair: the leader's warp only)
        if (leader) {
            mbar_wait(a_full, 0);
            const uint32_t ones_addr = smem_u32(sOnes);
            const uint64_t da_ones = umma_desc_plain(ones_addr, 2048u, 128u);
            const uint32_t idesc = pair ? KT_IDESC_PAIR : KT_IDESC;
            int stage = 0;
            uint32_t ph = 0;
            for (int g = 0; g < 2 * T; ++g) {
                const int buf = g & (KT_NBUF - 1);
                mbar_wait(tm_empty + buf, ((g / KT_NBUF) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)buf * KT_COLS;
                const bool first = g < T;                           // pass 1: dot~ = hi.hi (a looser, cheaper bound)
                const int nkb = first ? SEG : KB;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(full + stage, ph);
                    tc_fence_after();
                    const uint32_t st_addr = smem_u32(sB + (size_t)stage * STG);
                    const uint64_t db = umma_desc_sw128(st_addr);
                    const int ka = kb % SEG;                        // matching K block of A
                    const uint64_t da_hi = umma_desc_sw128(smem_u32(sA + (size_t)ka * KT_A_BYTES));
                    const uint64_t da_lo = umma_desc_sw128(smem_u32(sA + (size_t)(SEG + ka) * KT_A_BYTES));
                    // acc = ones . norms^T = -t_j / 2 in every row: K group 0 of B is the tile the TMA just wrote,
                    // K group 1 is the zero half of the ones block (its A counterpart is zero as well)
                    const uint32_t nb_addr = st_addr + BLK;
                    const uint64_t db_nb = umma_desc_plain(nb_addr, ones_addr + 2048u - nb_addr, 128u);
                    if (elect_one()) {
                        if constexpr (pair) {
                            if (kb == 0) tc_mma_bf16_pair(d, da_ones, db_nb, idesc, 0u);
#pragma unroll
                            for (int k16 = 0; k16 < KT_KBLK / 16; ++k16)
                                tc_mma_bf16_pair(d, da_hi + 2 * k16, db + 2 * k16, idesc, 1u);
                            if (kb < SEG && !first) {
#pragma unroll
                                for (int k16 = 0; k16 < KT_KBLK / 16; ++k16)
                                    tc_mma_bf16_pair(d, da_lo + 2 * k16, db + 2 * k16, idesc, 1u);
                            }
                            tc_commit_pair(empty + stage);              // both CTAs' producers may refill the stage
                            if (kb == nkb - 1) tc_commit_pair(tm_full + buf);   // both CTAs' epilogues may read tile g
                        } else {
                            if (kb == 0) tc_mma_bf16(d, da_ones, db_nb, idesc, 0u);
#pragma unroll
                            for (int k16 = 0; k16 < KT_KBLK / 16; ++k16)   // +32 bytes per K=16 step inside the swizzle span
                                tc_mma_bf16(d, da_hi + 2 * k16, db + 2 * k16, idesc, 1u);   // hi.hi | hi.lo
                            if (kb < SEG && !first) {                   // pass 2: a B_hi block also meets A_lo:  lo.hi
#pragma unroll
                                for (int k16 = 0; k16 < KT_KBLK / 16; ++k16)
                                    tc_mma_bf16(d, da_lo + 2 * k16, db + 2 * k16, idesc, 1u);
                            }
                            tc_commit(empty + stage);                   // smem stage reusable when these MMAs retire
                            if (kb == nkb - 1) tc_commit(tm_full + buf);    // accumulator of tile g complete
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; ph ^= 1; }
                }
            }
        }
    } else {
        // ================================ epilogue: four threads per query row =========
        // warp w (2..17): TMEM lane quadrant q = w & 3 (hardware rule), column quarter h = (w - 2) >> 2.
        const int q = warp & 3;
        const int h = (warp - 2) >> 2;
        const int r = q * 32 + lane;             // row within the tile
        const int i = i0 + r;
        const int et = threadIdx.x - 64;         // 0..511 among the epilogue threads
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)h * 64;
        const int last_valid = N - (T - 1) * KT_COLS - h * 64;    // valid columns of this quarter in the last tile
        const bool ragged = last_valid < 64;
        const float shift = (P.dump && i < N) ? P.ss[(size_t)rowbase + i] : 0.0f;
        const uint32_t tm_empty_c = pair ? mapa_rank0(tm_empty) : 0u;   // pair: the LEADER's buffer-free barriers
        float gmax[NGT];
#pragma unroll
        for (int e = 0; e < NGT; ++e) gmax[e] = -INFINITY;
        if (et == 0) kt_mark(P, 1);

        // A thread owns 64 columns of every tile: two tcgen05.ld of 32 columns.  With 16 classes per thread (k <= 32) both
        // are in flight before the wait (tcgen05.wait::ld waits for all of a thread's loads anyway); with 32 classes the
        // registers only allow one at a time.
        constexpr bool BOTH = false;
        auto release = [&](int g) {                                  // the accumulator is in registers: free the buffer
            const int buf = g & (KT_NBUF - 1);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (pair) mbar_arrive_cluster(tm_empty_c + 8u * (uint32_t)buf);
                else mbar_arrive(tm_empty + buf);
            }
        };
        auto ready = [&](int g) {
            const int buf = g & (KT_NBUF - 1);
            mbar_wait(tm_full + buf, (g / KT_NBUF) & 1);
            tc_fence_after();
        };
        auto load = [&](int g, int half, uint32_t (&u)[32]) {
            tc_ld32_issue(tlane + (uint32_t)(g & (KT_NBUF - 1)) * KT_COLS + (uint32_t)half * 32, u);
        };
        auto mask_and_dump = [&](int g, int half, uint32_t (&u)[32]) {   // g: tile index over both passes
            const int tile = g % T;
            if (ragged && tile == T - 1) {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (half * 32 + c >= last_valid) u[c] = 0xff800000u;   // columns beyond the cloud: -inf
            }
            if (P.dump && i < N) {
                const int j0 = tile * KT_COLS + h * 64 + half * 32;
                float *drow = P.dump + ((size_t)(g / T) * gridDim.y * N + rowbase + i) * N + j0;
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (j0 + c < N) drow[c] = __fmaf_rn(-2.0f, __uint_as_float(u[c]), -shift);
            }
        };
        auto classes = [&](const uint32_t (&u)[32]) {
            if (NGT == 16) {
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    gmax[c] = fmaxf(fmaxf(gmax[c], __uint_as_float(u[c])), __uint_as_float(u[c + 16]));   // FMNMX3
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) gmax[c % NGT] = fmaxf(gmax[c % NGT], __uint_as_float(u[c]));
            }
        };

        // ---- pass 1: class maxima of acc = hi.hi - t_j/2
        for (int g = 0; g < T; ++g) {
            uint32_t ua[32];
            ready(g);
            if (et == 0 && g == 0) kt_mark(P, 2);
            load(g, 0, ua);
            if constexpr (BOTH) {
                uint32_t ub[32];
                load(g, 1, ub);
                tc_wait32(ua);
                tc_wait32(ub);
                release(g);
                mask_and_dump(g, 0, ua);
                classes(ua);
                mask_and_dump(g, 1, ub);
                classes(ub);
            } else {
                tc_wait32(ua);
                mask_and_dump(g, 0, ua);
                classes(ua);
                load(g, 1, ua);
                tc_wait32(ua);
                release(g);
                mask_and_dump(g, 1, ua);
                classes(ua);
            }
        }
        // ---- between the passes: the row has 4 NGT class maxima (NGT per thread); tau = their k-th largest.  Every thread
        // sorts its own (descending, in registers); quarters 0 and 2 merge their list with their neighbour's (bitonic merge
        // of [own | reversed neighbour]); quarter 0 then takes the k-th largest of the union of the two merged lists A, B:
        // max over i of min(A[i-1], B[k-i-1]) (i values from A, k - i from B).
        if (et == 0) kt_mark(P, 3);
        reg_sort_desc<NGT>(gmax);
        float *xchg2 = xchg + (size_t)4 * NGT * KT_ROWS;           // [2 NGT][128]: the merged list of quarters 2 | 3
#pragma unroll
        for (int e = 0; e < NGT; ++e) xchg[((size_t)h * NGT + e) * KT_ROWS + r] = gmax[e];
        if (et == 0) kt_mark(P, 4);
        epi_bar_sync();
        if (et == 0) kt_mark(P, 5);
        float mg[2 * NGT];
        if ((h & 1) == 0) {
#pragma unroll
            for (int e = 0; e < NGT; ++e) {
                mg[e] = gmax[e];
                mg[NGT + e] = xchg[((size_t)(h + 1) * NGT + (NGT - 1 - e)) * KT_ROWS + r];
            }
            bitonic_merge_desc<2 * NGT>(mg);
            if (h == 2) {
#pragma unroll
                for (int e = 0; e < 2 * NGT; ++e) xchg2[(size_t)e * KT_ROWS + r] = mg[e];
            }
        }
        epi_bar_sync();
        if (h == 0) {
            float mxx = red_s[0], myy = red_s[16];
#pragma unroll
            for (int w = 1; w < 16; ++w) {
                mxx = fmaxf(mxx, red_s[w]);
                myy = fmaxf(myy, red_s[16 + w]);
            }
            const float eps = pair_eps(yyi, myy, xxi, mxx, P.C);    // bound for every candidate of the cloud
            const float eps1 = KT_EPS1_REL * sqrtf(yyi) * sqrtf(myy);   // pass 1 saw dot~ = hi.hi only
            const int k = P.k;
            float tau = -INFINITY;
#pragma unroll
            for (int t = 0; t <= 2 * NGT; ++t) {                    // t values from A = mg, k - t from B = xchg2
                const int bi = k - t - 1;
                if (t <= k && bi < 2 * NGT) {
                    const float av = (t == 0) ? INFINITY : mg[t == 0 ? 0 : t - 1];
                    const float bv = (bi < 0) ? INFINITY : xchg2[(size_t)bi * KT_ROWS + r];
                    tau = fmaxf(tau, fminf(av, bv));
                }
            }
            // k columns have pass-1 values >= tau, so the k-th best exact value is >= tau - (eps1 + eps)/2 on the accumulator
            // scale (v = -2 acc) and every member of the exact top-k has a pass-2 accumulator >= tau - (eps1 + 2 eps)/2
            const float slack = 0.5f * 1.0009765625f * (eps1 + 2.0f * eps);
            thr_s[r] = (i < N && tau > -INFINITY) ? tau - slack : INFINITY;   // rows beyond the cloud collect nothing
        }
        epi_bar_sync();                                             // thr_s written, xchg (aliases the lists) no longer read
        if (et == 0) kt_mark(P, 6);
        const float thr = thr_s[r];
        // ---- pass 2: collect the candidates into the row's list in shared memory.  Quarters 0/1 share the first half of
        // the row's words (0 from the front, 1 from the back), quarters 2/3 the second half.
        // all ones above bit 4; kept opaque to the compiler so that it stays in a register and the COLUMN is the immediate of
        // the packing LOP3 (the other way round costs a MOV per element)
        const uint32_t keep = 0xffffffc0u | ((uint32_t)N >> 31);
        const uint32_t cur0 = smem_u32(lists + (size_t)r * CAPW + (h >> 1) * HC + ((h & 1) ? HC - 1 : 0));
        const int step = (h & 1) ? -4 : 4;
        uint32_t cur = cur0;
        bool ovf = false;
        auto collect = [&](auto half_tag, const uint32_t (&u)[32]) {
            // a half tile can add 32 entries: without room for them nothing is stored any more and the row is flagged
            const int cnt = abs((int)(cur - cur0)) >> 2;
            const bool room = cnt <= HC - 32;
            ovf |= !room;
            const float t = room ? thr : INFINITY;
            collect_tile<0, decltype(half_tag)::value * 32>(cur, u, t, step, keep);
        };
        for (int g = T; g < 2 * T; ++g) {
            uint32_t ua[32];
            ready(g);
            load(g, 0, ua);
            if constexpr (BOTH) {
                uint32_t ub[32];
                load(g, 1, ub);
                tc_wait32(ua);
                tc_wait32(ub);
                release(g);
                mask_and_dump(g, 0, ua);
                collect(std::integral_constant<int, 0>{}, ua);
                mask_and_dump(g, 1, ub);
                collect(std::integral_constant<int, 1>{}, ub);
            } else {
                tc_wait32(ua);
                mask_and_dump(g, 0, ua);
                collect(std::integral_constant<int, 0>{}, ua);
                load(g, 1, ua);
                tc_wait32(ua);
                release(g);
                mask_and_dump(g, 1, ua);
                collect(std::integral_constant<int, 1>{}, ua);
            }
            snap_s[(size_t)(g - T) * KT_EPI_THREADS + et] = (uint8_t)(abs((int)(cur - cur0)) >> 2);
        }
        const int cnt_end = abs((int)(cur - cur0)) >> 2;
        if (et == 0) kt_mark(P, 7);
        if (i < N) {
            P.cand_cnt[4 * ((size_t)rowbase + i) + h] = (uint8_t)(ovf ? 255 : cnt_end);
            // snapshots of this quarter: TP bytes, padding 255 (never <= an entry position)
            uint8_t *sg = P.snap + (4 * ((size_t)rowbase + i) + h) * P.TP;
            for (int t0 = 0; t0 < P.TP; t0 += 8) {
                uint32_t w0 = 0, w1 = 0;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const uint32_t a = (t0 + t < T) ? snap_s[(size_t)(t0 + t) * KT_EPI_THREADS + et] : 255u;
                    const uint32_t c = (t0 + 4 + t < T) ? snap_s[(size_t)(t0 + 4 + t) * KT_EPI_THREADS + et] : 255u;
                    w0 |= a << (8 * t);
                    w1 |= c << (8 * t);
                }
                *reinterpret_cast<uint2 *>(sg + t0) = make_uint2(w0, w1);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this thread's list words, for the TMA engine
        epi_bar_sync();                                             // every list of the CTA is complete
        if (et == 0) kt_mark(P, 8);
        if (et == 0) {
            // one bulk copy of the CTA's lists (contiguous in global memory: rows i0 .. of the cloud) by the TMA engine.
            // Compacting first was measured and dropped: scattered or per-row stores cost 2.3 - 3.4 us against 1.3 us for
            // this L2 write burst of 64 KiB per CTA.
            const int rows_valid = min(KT_ROWS, N - i0);
            if (rows_valid > 0) {
                bulk_store(P.cand + ((size_t)rowbase + i0) * CAPW, lists, (uint32_t)rows_valid * CAPW * 4u);
                bulk_store_wait();                                  // shared memory is released when the CTA exits
            }
        }
        if (et == 0) kt_mark(P, 9);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (pair) cluster_sync_all();            // no CTA leaves while its peer can still signal its barriers / read its smem
    if (warp == 1) {
        if constexpr (pair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
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

// Keys arrive as (orderable v~ << 32 | j); xxj / yyj: the candidates' norms in LIST order, fetched before the sort so
// that their L2 latency hides behind the sorting network (only their list-wide maximum bound is needed).
template <int S, int C, int KS>
__device__ __forceinline__ void refine_sorted(const KtParams &P, uint32_t row, uint32_t base, unsigned long long (&key)[S],
                                              const float (&xxj)[S], const float (&yyj)[S], uint16_t *sj, float *se,
                                              uint32_t (&nbr)[KS])
{
    constexpr int M = C / 16;                             // float4 pieces per lane: 4 (C = 64) or 8 (C = 128)
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, u = lane & 3;
    const int k = P.k;
    const float xxi = P.xx[row], yyi = P.yy[row];
    warp_sort_u64<S>(key);
    // ---- links, clusters
    // A position e that starts a new cluster must certify EVERY later position m >= e against EVERY earlier one
    // t < e: exact_m - exact_t >= (v_e - v_{e-1}) - eps_m - eps_t.  The per-pair bounds differ with the norms, so the
    // test uses the largest bound of the whole list (one redux): gap > 2 max eps.
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
        emax = fmaxf(emax, pair_eps(yyi, yyj[s], xxi, xxj[s], C));       // list order: only the maximum matters
    }
    emax = warp_max_f32(emax);
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
    if (P.want_stats && total && lane == 0) atomicAdd(P.stats + 2, total);   // diagnostics: exact distances recomputed
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

// list capacity per row (words) -> sorting slots: 128 words (k <= 32): up to 64 candidates; 192 words: up to 128
template <int CAPW, int C>
__global__ void __launch_bounds__(32 * RF_WARPS, 4)   // <= 64 registers: four CTAs per SM
knn_refine_kernel(KtParams P, long long total_rows)
{
    constexpr int HC = CAPW / 2;
    constexpr int SLOTS = (CAPW == 128) ? 2 : 4;
    constexpr int KS = SLOTS / 2;                          // slots holding the k ranked neighbours (k <= 32 / 64)
    constexpr int SW = 4 * KT_MAX_TILES / 4;               // snapshot words per row, at most
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pdl_wait();                                                        // the filter's candidate lists are complete
    const long long row64 = (long long)blockIdx.x * RF_WARPS + warp;   // b*N + i
    if (row64 >= total_rows) return;
    const uint32_t row = (uint32_t)row64;                              // B*N < 2^31 (knn_tensor_supported)
    const uint32_t base = row / (uint32_t)P.N * (uint32_t)P.N;         // first row of the cloud
    __shared__ __align__(16) uint32_t sl_all[RF_WARPS][CAPW];
    __shared__ uint32_t sn_all[RF_WARPS][SW];
    __shared__ uint16_t sj_all[RF_WARPS][32 * SLOTS];
    __shared__ float se_all[RF_WARPS][32 * SLOTS];
    uint32_t *sl = sl_all[warp], *sn = sn_all[warp];
    uint16_t *sj = sj_all[warp];
    float *se = se_all[warp];
    // the row's list words, its four counters and its snapshots in ONE L2 round trip
    const uint4 *l4 = reinterpret_cast<const uint4 *>(P.cand + (size_t)row * CAPW);
    const uint32_t *g_sn = reinterpret_cast<const uint32_t *>(P.snap + (size_t)row * 4 * P.TP);
    const uint4 w0 = l4[lane];
    uint4 w1 = make_uint4(0u, 0u, 0u, 0u);
    if (CAPW > 128 && lane < CAPW / 4 - 32) w1 = l4[32 + lane];
    const uint32_t cc = *reinterpret_cast<const uint32_t *>(P.cand_cnt + 4 * (size_t)row);
    const int TPW = P.TP;                                  // 4 quarters x TP bytes = TP words
    const uint32_t s0 = (lane < TPW) ? g_sn[lane] : 0u;
    const uint32_t s1 = (32 + lane < TPW) ? g_sn[32 + lane] : 0u;
    reinterpret_cast<uint4 *>(sl)[lane] = w0;
    if (CAPW > 128 && lane < CAPW / 4 - 32) reinterpret_cast<uint4 *>(sl)[32 + lane] = w1;
    if (lane < TPW) sn[lane] = s0;
    if (32 + lane < TPW) sn[32 + lane] = s1;
    __syncwarp();
    const int c0 = cc & 255, c1 = (cc >> 8) & 255, c2 = (cc >> 16) & 255, c3 = cc >> 24;
    const int cnt = c0 + c1 + c2 + c3;
    uint32_t nbr[KS];
    if (c0 == 255 || c1 == 255 || c2 == 255 || c3 == 255 || c0 + c1 > HC || c2 + c3 > HC || cnt > 32 * SLOTS || cnt < P.k) {
        knn_row_exact<KS>(P.xt, P.xx, row64, P.N, C, P.k, P.idx);        // warp-uniform
        if (P.want_stats && lane == 0) atomicAdd(P.fb_count, 1);
        if (P.edge_out) {
            __syncwarp();                                  // the row's idx, written by this warp, is visible to it
#pragma unroll
            for (int s = 0; s < KS; ++s) nbr[s] = (s * 32 + lane < P.k) ? (uint32_t)P.idx[(size_t)row * P.k + s * 32 + lane] : 0u;
            gather_row<C, KS>(P, row, base, nbr);
        }
        return;
    }
    // decode: entry e of the concatenated quarters -> (filter value, candidate index)
    unsigned long long key[SLOTS];
    float xxj[SLOTS], yyj[SLOTS];
    const int TQ = P.TP / 4;                               // snapshot words per quarter
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int e = s * 32 + lane;
        key[s] = ~0ull;
        xxj[s] = 0.0f;
        yyj[s] = 0.0f;
        if (e < cnt) {
            int seg = 0, p = e;
            if (p >= c0) { p -= c0; seg = 1;
                if (p >= c1) { p -= c1; seg = 2;
                    if (p >= c2) { p -= c2; seg = 3; } } }
            const uint32_t w = sl[(seg >> 1) * HC + ((seg & 1) ? HC - 1 - p : p)];
            // the tile the entry was written in = number of tiles whose closing cursor is <= its position
            const uint32_t pp = (uint32_t)p * 0x01010101u;
            int tile = 0;
            for (int t = 0; t < TQ; ++t) tile += __popc(__vcmpleu4(sn[seg * TQ + t], pp)) >> 3;
            const uint32_t j = (uint32_t)tile * KT_COLS + (uint32_t)seg * 64u + (w & 63u);
            const float vf = __fmul_rn(-2.0f, __uint_as_float(w & 0xffffffc0u));   // v~ = -2 acc (row constant dropped)
            key[s] = ((unsigned long long)f32_orderable(__fadd_rn(vf, 0.0f)) << 32) | j;
            xxj[s] = P.xx[base + j];                       // consumed after the sort
            yyj[s] = P.yy[base + j];
        }
    }
    if (cnt <= 32) {                                       // warp-uniform, the usual case for k <= 20
        unsigned long long k1[1] = {key[0]};
        const float x1[1] = {xxj[0]}, y1[1] = {yyj[0]};
        refine_sorted<1, C, KS>(P, row, base, k1, x1, y1, sj, se, nbr);
    } else if (SLOTS > 2 && cnt <= 64) {
        unsigned long long k2[2] = {key[0], key[1]};
        const float x2[2] = {xxj[0], xxj[1]}, y2[2] = {yyj[0], yyj[1]};
        refine_sorted<2, C, KS>(P, row, base, k2, x2, y2, sj, se, nbr);
    } else {
        refine_sorted<SLOTS, C, KS>(P, row, base, key, xxj, yyj, sj, se, nbr);
    }
    if (P.want_stats && lane == 0) {
        atomicAdd(P.stats, 1);
        atomicAdd(P.stats + 1, cnt);                       // diagnostics: total length of the certified rows' lists
    }
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

// bf16 matrix (rows, cols) row-major, boxes of (128 rows, box_cols): the operand blocks (64 columns, 128-byte swizzle)
// and the norm operand (8 columns = one 16-byte K group per row, no swizzle: rows land 16 bytes apart)
static int make_map(CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows,
                    bool swizzle)
{
    EncodeTiledFn fn = encode_fn();
    MLSP_REQUIRE(fn, MLSP_ECUDA, "knn: cuTensorMapEncodeTiled not available");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MLSP_REQUIRE(r == CUDA_SUCCESS, MLSP_ECUDA, "knn: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return MLSP_OK;
}

struct KtLayout {
    size_t off_counters, off_cen, off_xx, off_yy, off_ss, off_hi, off_lo, off_nb, off_xt, off_cand, off_cnt, off_snap, total;
    int capw, TP;
};

static KtLayout kt_layout(int B, int C, int N, int k)
{
    KtLayout L;
    L.capw = (k <= 32) ? 128 : 192;
    const int T = (N + KT_COLS - 1) / KT_COLS;
    L.TP = (T + 7) / 8 * 8;
    const size_t rows = (size_t)B * N;
    size_t o = 0;
    L.off_counters = o; o += 256;                                            // [0] fb_count, [1] certified rows
    L.off_cen = o;      o += align_up(sizeof(float) * (size_t)B * C, 256);
    L.off_xx = o;       o += align_up(sizeof(float) * rows, 256);
    L.off_yy = o;       o += align_up(sizeof(float) * rows, 256);
    L.off_ss = o;       o += align_up(sizeof(float) * rows, 256);
    L.off_hi = o;       o += align_up(2 * rows * C, 1024);
    L.off_lo = o;       o += align_up(2 * rows * C, 1024);
    L.off_nb = o;       o += align_up(16 * rows, 1024);
    L.off_xt = o;       o += align_up(sizeof(float) * rows * C, 256);
    L.off_cand = o;     o += align_up(sizeof(uint32_t) * rows * L.capw, 256);
    L.off_cnt = o;      o += align_up(4 * rows, 256);
    L.off_snap = o;     o += align_up(4 * (size_t)L.TP * rows, 256);
    L.total = o;
    return L;
}

// ring depth: as deep as the 227 KiB of one SM allow (one CTA per SM)
static int kt_stages(int C, int k, int N, int cs)
{
    const int T = (N + KT_COLS - 1) / KT_COLS;
    const size_t per_sm = 227 * 1024;
    int stages = 0;
    for (int s_ = 2; s_ <= KT_MAX_STAGES; ++s_)
        if (kt_smem_bytes(C, k, T, s_, cs) <= per_sm) stages = s_;
    return stages;
}

// Limits of the tensor path (anything else takes the fp32 kernels of knn3.cu / knn.cu): C = 64 or 128; 256 <= N <= 8192
// (per-tile cursor snapshots live in shared memory); k <= 64; a shared-memory configuration with >= 2 ring stages
// exists (it does not for C = 128, k > 32 beyond N = 6144).
bool knn_tensor_supported(int B, int C, int N, int k)
{
    return (C == 64 || C == 128) && N >= 256 && N <= KT_MAX_TILES * KT_COLS && k >= 1 && k <= 64 && (long long)B * N < (1ll << 31) &&
           B <= 65535 && kt_stages(C, k, N, 2) >= 2;
}

size_t knn_tensor_workspace_bytes(int B, int C, int N, int k) { return kt_layout(B, C, N, k).total; }

// the point-major copy xt (B,N,C) the tensor path leaves in its workspace (reused by the fused edge gather)
const float *knn_tensor_xt(const void *ws, int B, int C, int N, int k)
{
    return reinterpret_cast<const float *>(static_cast<const char *>(ws) + kt_layout(B, C, N, k).off_xt);
}

// `stages_mask` (measurement hook of mlsp_graph_feature_fwd_stage): bit 0 prep, bit 1 filter, bit 2 ranking (+ gather); 7 = all
int knn_tensor_run(const float *x, int B, int C, int N, int k, int64_t *idx, void *ws, float *dump, float *edge_out,
                   int stages_mask, cudaStream_t st, long long *tstamp, int cluster, bool want_stats)
{
    const KtLayout L = kt_layout(B, C, N, k);
    char *w = static_cast<char *>(ws);
    int *counters = reinterpret_cast<int *>(w + L.off_counters);
    float *cen = reinterpret_cast<float *>(w + L.off_cen);
    float *xx = reinterpret_cast<float *>(w + L.off_xx);
    float *yy = reinterpret_cast<float *>(w + L.off_yy);
    float *ss = reinterpret_cast<float *>(w + L.off_ss);
    __nv_bfloat16 *hi = reinterpret_cast<__nv_bfloat16 *>(w + L.off_hi);
    __nv_bfloat16 *lo = reinterpret_cast<__nv_bfloat16 *>(w + L.off_lo);
    uint4 *nb = reinterpret_cast<uint4 *>(w + L.off_nb);
    float *xt = reinterpret_cast<float *>(w + L.off_xt);

    if (stages_mask & 1) {
        knn_centre_kernel<<<B, 256, 0, st>>>(x, C, N, cen);
        MLSP_LAUNCH_CHECK("knn_centre_kernel");
        cudaLaunchConfig_t pcfg = {};
        pcfg.gridDim = dim3((N + PREP_PTS - 1) / PREP_PTS, B);
        pcfg.blockDim = dim3(PREP_THREADS);
        pcfg.dynamicSmemBytes = sizeof(float) * (C * (PREP_PTS + 1) + C);
        pcfg.stream = st;
        cudaLaunchAttribute pattr[1];
        pattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        pattr[0].val.programmaticStreamSerializationAllowed = 1;
        pcfg.attrs = pattr;
        pcfg.numAttrs = (g_kt_pdl & 1) ? 1 : 0;
        MLSP_CUDA(cudaLaunchKernelEx(&pcfg, knn_prep_kernel, x, (const float *)cen, C, N, xx, yy, ss, counters, xt, hi, lo, nb));
        MLSP_LAUNCH_CHECK("knn_prep_kernel");
    }

    // CTA pairs (cta_group::2); an odd number of row blocks per cloud gets a trailing CTA without valid rows, which still
    // fetches its half of every candidate block.  `cluster` = 1 (profiling hook) forces the unpaired kernel.
    const int cs = (cluster == 1) ? 1 : 2;
    const unsigned rblk = (unsigned)((N + KT_ROWS - 1) / KT_ROWS);
    dim3 grid(cs == 2 ? (rblk + 1) / 2 * 2 : rblk, B);
    CUtensorMap map_hi, map_lo, part_hi, part_lo, map_nb;  // boxes of 128 rows (A tiles, unshared B blocks) and of 128 / cs rows
    int rc = make_map(&map_hi, hi, (uint64_t)B * N, (uint64_t)C, KT_KBLK, KT_ROWS, true);
    if (rc) return rc;
    rc = make_map(&map_lo, lo, (uint64_t)B * N, (uint64_t)C, KT_KBLK, KT_ROWS, true);
    if (rc) return rc;
    rc = make_map(&part_hi, hi, (uint64_t)B * N, (uint64_t)C, KT_KBLK, KT_COLS / cs, true);
    if (rc) return rc;
    rc = make_map(&part_lo, lo, (uint64_t)B * N, (uint64_t)C, KT_KBLK, KT_COLS / cs, true);
    if (rc) return rc;
    rc = make_map(&map_nb, nb, (uint64_t)B * N, 8, 8, KT_COLS / cs, false);
    if (rc) return rc;

    KtParams P;
    P.xx = xx; P.yy = yy; P.ss = ss; P.xt = xt; P.idx = idx; P.fb_count = counters; P.stats = counters + 1;
    P.cand = reinterpret_cast<uint32_t *>(w + L.off_cand);
    P.cand_cnt = reinterpret_cast<uint8_t *>(w + L.off_cnt);
    P.snap = reinterpret_cast<uint8_t *>(w + L.off_snap);
    P.want_stats = (want_stats || dump) ? 1 : 0;
    P.dump = dump; P.tstamp = tstamp; P.edge_out = reinterpret_cast<float4 *>(edge_out);
    P.N = N; P.C = C; P.k = k; P.T = (N + KT_COLS - 1) / KT_COLS; P.TP = L.TP; P.capw = L.capw;
    P.stages = kt_stages(C, k, N, cs);
    MLSP_REQUIRE(P.stages >= 2, MLSP_EUNSUPPORTED, "knn: no shared-memory configuration for C=%d k=%d N=%d", C, k, N);
    P.cs = cs;
    const size_t smem = kt_smem_bytes(C, k, P.T, P.stages, cs);
    if (stages_mask & 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(KT_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = (g_kt_pdl & 2) ? 2 : 1;
#define MLSP_KT_LAUNCH(NGT_, PAIR_)                                                                                              \
    do {                                                                                                                         \
        MLSP_CUDA(cudaFuncSetAttribute(knn_tensor_kernel<NGT_, PAIR_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        MLSP_CUDA(cudaLaunchKernelEx(&cfg, knn_tensor_kernel<NGT_, PAIR_>, map_hi, map_lo, part_hi, part_lo, map_nb, P));        \
    } while (0)
        if (k <= 32 && cs == 2) MLSP_KT_LAUNCH(16, true);
        else if (k <= 32) MLSP_KT_LAUNCH(16, false);
        else if (cs == 2) MLSP_KT_LAUNCH(32, true);
        else MLSP_KT_LAUNCH(32, false);
#undef MLSP_KT_LAUNCH
        MLSP_LAUNCH_CHECK("knn_tensor_kernel");
    }
    if (stages_mask & 4) {
        const long long rows_total = (long long)B * N;
        const unsigned rblocks = (unsigned)((rows_total + RF_WARPS - 1) / RF_WARPS);
        cudaLaunchConfig_t rcfg = {};
        rcfg.gridDim = dim3(rblocks);
        rcfg.blockDim = dim3(32 * RF_WARPS);
        rcfg.stream = st;
        cudaLaunchAttribute rattr[1];
        rattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        rattr[0].val.programmaticStreamSerializationAllowed = 1;
        rcfg.attrs = rattr;
        rcfg.numAttrs = (g_kt_pdl & 4) ? 1 : 0;
        if (k <= 32 && C == 64)
            MLSP_CUDA(cudaLaunchKernelEx(&rcfg, knn_refine_kernel<128, 64>, P, rows_total));
        else if (k <= 32)
            MLSP_CUDA(cudaLaunchKernelEx(&rcfg, knn_refine_kernel<128, 128>, P, rows_total));
        else if (C == 64)
            MLSP_CUDA(cudaLaunchKernelEx(&rcfg, knn_refine_kernel<192, 64>, P, rows_total));
        else
            MLSP_CUDA(cudaLaunchKernelEx(&rcfg, knn_refine_kernel<192, 128>, P, rows_total));
        MLSP_LAUNCH_CHECK("knn_refine_kernel");
    }
    return MLSP_OK;
}

}  // namespace mlsp

// Measurement hook (process-wide): 1 (default) = the four kernels of the tcgen05 kNN path are chained with programmatic
// dependent launch, 0 = plain stream order.  Results do not depend on it.
extern "C" void mlsp_knn_set_pdl(int on) { mlsp::g_kt_pdl = on; }

