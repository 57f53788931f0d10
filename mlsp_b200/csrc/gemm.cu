// gemm.cu -- the point-wise (1x1) convolutions around the neighbourhood engine as ONE hand-written tensor-core GEMM
// (SURVEY.md 8f ranks 1 and 4): the layer product [Y|Z] = x . [Wa ; Wb-Wa]^T of EdgeConv (PointDA/Models.py:114-128 with
// conv_2d of PointDA/model_utils.py:45-63), its two backward products, the heads' shared first layer
// (PointDA/Models.py:156-160) and every other nn.Conv1d / nn.Conv2d(kernel 1) / nn.Linear of the DGCNN.
//
//     D[z] (M x N) = A[z] (M x K) . B[z]^T (N x K)  (+ bias[n]),   z = 0 .. batch-1,   fp32 in, fp32 out
//
// Arithmetic: every fp32 operand is split on the fly into three bf16 pieces x = x1 + x2 + x3 (24 significant bits, i.e. the
// fp32 value itself) and the product is accumulated in fp32 (TMEM) from the six piece products of weight >= 2^-16
// (x1w1, x1w2, x2w1, x1w3, x2w2, x3w1); what is dropped is below 2^-24 |x||w| per term, the size of fp32's own rounding.
// The result is an fp32 GEMM up to summation order -- not a bf16 / tf32 approximation (tests: 1e-5 against fp64, measured ~1e-7).
//
// Persistent CTAs, one per SM, K in chunks of 64.  Two instances: single CTAs on 128 x 128 output tiles (small products), and
// CTA PAIRS (clusters of two, tcgen05 cta_group::2) on 256 x 256 tiles: one MMA of M = 256, N = 256 per instruction, each CTA
// loading its own 128 rows of A and HALF of the tile's B rows -- the operand stream every SM pulls through L2 per flop halves,
// and that stream (fp32 operands, 4 bytes per element) is what bounds this kernel.
//   warps 0-3  : epilogue -- tcgen05.ld (TMEM lane == output row), + bias, coalesced stores in either output orientation
//   warp  4    : TMEM allocation + the single-thread tcgen05.mma issuer (kind::f16, bf16 x bf16 -> f32, M = 128, N <= 128)
//   warps 5-20 : loaders -- fp32 straight from global memory (either operand K-major or MN-major, any leading dimension,
//                ragged edges zero-filled), split into pieces in registers, written to shared memory in the 128-byte-swizzled
//                canonical UMMA layout of the operand's own majorness (no transposition: MN-major operands use MN-major
//                descriptors); the loads of chunk i+1 are in flight while chunk i is converted.
// Two shared-memory stages (2 x 96 KiB) and two accumulator buffers (2 x 128 or 2 x 256 TMEM columns): the loaders, the tensor pipe
// and the epilogue of consecutive tiles overlap.  HBM-bound by design on the skinny products of this model (K = 64..512):
// each operand byte is read once per tile row/column from HBM or L2 and nothing but D is written.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"

namespace mlsp {
namespace gemm {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int STAGES = 2;
constexpr int NBUF = 2;
constexpr int EPI_WARPS = 4, LOAD_WARPS = 16;
constexpr int LOAD_THREADS = 32 * LOAD_WARPS;
constexpr int THREADS = 32 * (EPI_WARPS + 1 + LOAD_WARPS);       // 672
constexpr uint32_t PIECE_BYTES = BM * BK * 2;                    // 16 KiB: one bf16 piece of a 128 x 64 operand chunk
constexpr uint32_t OPER_BYTES = 3 * PIECE_BYTES;                 // 48 KiB
constexpr uint32_t STAGE_BYTES = 2 * OPER_BYTES;                 // 96 KiB: A pieces | B pieces
constexpr int EPI_PITCH = 33;
constexpr uint32_t EPI_BYTES = EPI_WARPS * 32 * EPI_PITCH * 4;
constexpr uint32_t SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + EPI_BYTES + 256;

struct Params {
    const float *A, *B, *bias;
    float *D;
    long long lda, ldb, ldd, sa, sb, sd;
    int M, N, K, batch;
    int a_kmajor, b_kmajor, d_rowmajor;
    int a_vec, b_vec;                 // 16-byte loads allowed (pointer, leading dimension and batch stride aligned)
    int mt, nt, kc;                   // tiles along M, N; K chunks
    int tiles;
    long long *tstamp;                // measurement hook (mlsp_gemm_f32_timeline): SM-clock stamps of CTA 0's first 64 iterations
};

// stamps per iteration: [0] loader: stage free, [1] loader: pieces stored, [2] loader: arrived, [3] MMA: stage full, [4] MMA: issued + committed
__device__ __forceinline__ void stamp(const Params &P, int it, int what)
{
    if (P.tstamp && blockIdx.x == 0 && it < 64) P.tstamp[it * 8 + what] = clock64();
}

// ---------------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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
        "GW_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra GW_DONE;\n\t"
        "bra GW_LOOP;\n\t"
        "GW_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
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
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&u)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
          "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
          "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
          "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]),
                   "+r"(u[8]), "+r"(u[9]), "+r"(u[10]), "+r"(u[11]), "+r"(u[12]), "+r"(u[13]), "+r"(u[14]), "+r"(u[15]),
                   "+r"(u[16]), "+r"(u[17]), "+r"(u[18]), "+r"(u[19]), "+r"(u[20]), "+r"(u[21]), "+r"(u[22]), "+r"(u[23]),
                   "+r"(u[24]), "+r"(u[25]), "+r"(u[26]), "+r"(u[27]), "+r"(u[28]), "+r"(u[29]), "+r"(u[30]), "+r"(u[31])
                 :
                 : "memory");
}

// ---- CTA pair (cta_group::2): the leader (cluster rank 0) issues the MMAs and owns the full / tm_empty barriers
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
__device__ __forceinline__ uint32_t mapa_rank0(const void *p)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(smem_u32(p)));
    return r;
}
// Arrive on the leader's barrier.  Default semantics (release at CTA scope): the operand pieces were already made visible to
// the async proxy by fence.proxy.async and are shared memory, not global -- an explicit cluster-scope release compiles to
// MEMBAR.ALL.GPU + ERRBAR per arrive (measured: 19 % of the kernel's stall samples).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar_cluster)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
// relaxed: what this orders are TMEM reads, fenced by tcgen05.fence::before_thread_sync
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t bar_cluster)
{
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
template <bool PAIR>
__device__ __forceinline__ void tc_commit_to(uint64_t *bar)      // pair: arrives on the barrier at this offset in BOTH CTAs
{
    if constexpr (PAIR)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                     "h"((uint16_t)3)
                     : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <bool PAIR>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    if constexpr (PAIR)
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(d_tmem),
            "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(d_tmem),
            "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
}

// Shared-memory matrix descriptors (sm_100 version bit 46, 128-byte swizzle = layout type 2 in bits 61..63).
// K-major: rows of 64 bf16 (128 B), 8-row swizzle atoms 1024 B apart (SBO); a K = 16 step is +32 B inside the row.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major: rows of 64 consecutive M (or N) indices at one k (128 B), 8 k rows per swizzle atom, atoms of the next 8 k
// 1024 B apart (SBO), the next 64 M/N indices 8192 B apart (LBO: 64 k rows x 128 B); a K = 16 step is +2048 B.
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (512ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// ---------------------------------------------------------------------------------------------------- split
// (a, b) -> three packed bf16 pairs with a = a1 + a2 + a3 (+ < 2^-25 |a|); the low half is the lower address (a).
__device__ __forceinline__ void split3(float a, float b, uint32_t &p0, uint32_t &p1, uint32_t &p2)
{
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p0) : "f"(b), "f"(a));
    const float a1 = __fsub_rn(a, __uint_as_float(p0 << 16)), b1 = __fsub_rn(b, __uint_as_float(p0 & 0xffff0000u));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(b1), "f"(a1));
    const float a2 = __fsub_rn(a1, __uint_as_float(p1 << 16)), b2 = __fsub_rn(b1, __uint_as_float(p1 & 0xffff0000u));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(b2), "f"(a2));
}

// One loader work item = 8 consecutive elements along the operand's contiguous dimension = one 16-byte chunk of a
// 128-byte swizzled row, per piece.  1024 items per operand chunk, two per loader thread.
struct Item {
    float4 lo, hi;
};

// K-major operand: item (row r = 0..127, chunk c = 0..7) = elements (row0 + r, k0 + 8c .. +7)
// MN-major operand: item (k row kr = 0..63, chunk mc = 0..15) = elements (row0 + 8mc .. +7, k0 + kr)
// `rows` = exclusive limit of the row index (rows at or beyond it read as zero).
__device__ __forceinline__ void load_item(Item &it, const float *__restrict__ base, long long ld, int kmajor, int vec, int g,
                                          int row0, int rows, int k0, int K)
{
    it.lo = make_float4(0.f, 0.f, 0.f, 0.f);
    it.hi = it.lo;
    long long off;
    int n_valid;                                   // elements of the run that exist
    if (kmajor) {
        const int r = g >> 3, c = g & 7;
        const int row = row0 + r, k = k0 + 8 * c;
        if (row >= rows || k >= K) return;
        off = (long long)row * ld + k;
        n_valid = min(8, K - k);
    } else {
        const int kr = g >> 4, mc = g & 15;
        const int k = k0 + kr, row = row0 + 8 * mc;
        if (k >= K || row >= rows) return;
        off = (long long)k * ld + row;
        n_valid = min(8, rows - row);
    }
    const float *p = base + off;
    if (vec && n_valid == 8) {
        it.lo = __ldg(reinterpret_cast<const float4 *>(p));
        it.hi = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    } else {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (e < n_valid) ? __ldg(p + e) : 0.0f;
        it.lo = make_float4(v[0], v[1], v[2], v[3]);
        it.hi = make_float4(v[4], v[5], v[6], v[7]);
    }
}

// byte offset of item g's 16-byte chunk inside one piece (128-byte swizzle: chunk index XOR row-in-atom)
__device__ __forceinline__ uint32_t item_smem_offset(int kmajor, int g)
{
    if (kmajor) {
        const int r = g >> 3, c = g & 7;
        return (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((c ^ (r & 7)) << 4);
    }
    const int kr = g >> 4, mc = g & 15;
    return (uint32_t)(mc >> 3) * 8192u + (uint32_t)(kr >> 3) * 1024u + (uint32_t)(kr & 7) * 128u + (uint32_t)(((mc & 7) ^ (kr & 7)) << 4);
}

__device__ __forceinline__ void store_item(const Item &it, uint8_t *oper, uint32_t off)
{
    uint4 q0, q1, q2;
    split3(it.lo.x, it.lo.y, q0.x, q1.x, q2.x);
    split3(it.lo.z, it.lo.w, q0.y, q1.y, q2.y);
    split3(it.hi.x, it.hi.y, q0.z, q1.z, q2.z);
    split3(it.hi.z, it.hi.w, q0.w, q1.w, q2.w);
    *reinterpret_cast<uint4 *>(oper + off) = q0;
    *reinterpret_cast<uint4 *>(oper + PIECE_BYTES + off) = q1;
    *reinterpret_cast<uint4 *>(oper + 2 * PIECE_BYTES + off) = q2;
}

struct StageRegs {
    Item a[2], b[2];
};

// One operand as one loader thread sees it.  Item 1 is item 0 shifted by 64 rows (K-major) or 32 k rows (MN-major): a fixed
// element offset `d1` in global memory and a fixed byte offset in the stage.  p0 runs along K (+ `adv` elements per chunk).
// mode bits (per item u: bits 2u, 2u+1) say how the item is fetched in a K chunk that lies wholly inside K:
// 0 = out of range (zeros), 1 = two 16-byte loads, 2 = the guarded element-wise path, 3 = a row of the stage that no MMA
// reads (beyond this CTA's share of the tile's B rows): neither fetched nor stored.
struct OperandLane {
    const float *p0;
    uint32_t soff0;
    int modes;
};

__device__ __forceinline__ void lane_begin_tile(OperandLane &L, const float *base, long long ld, int kmajor, int vec, int lt,
                                                int row0, int rows, int read_rows)
{
    L.modes = 0;
    if (kmajor) {
        const int r = lt >> 3, c = lt & 7;
        L.p0 = base + (long long)(row0 + r) * ld + 8 * c;
#pragma unroll
        for (int u = 0; u < 2; ++u)
            L.modes |= ((r + 64 * u >= read_rows) ? 3 : (row0 + r + 64 * u >= rows) ? 0 : (vec ? 1 : 2)) << (2 * u);
    } else {
        const int kr = lt >> 4, mc = lt & 15;
        const int row = row0 + 8 * mc;
        L.p0 = base + (long long)kr * ld + row;
        const int m = (8 * mc >= read_rows) ? 3 : (row >= rows) ? 0 : ((vec && row + 8 <= rows) ? 1 : 2);
        L.modes = m | (m << 2);
    }
}

// fetch item u of the chunk p0 currently points at
__device__ __forceinline__ void lane_load(Item &it, const OperandLane &L, int u, long long d1, bool kfull, const float *base,
                                          long long ld, int kmajor, int vec, int lt, int row0, int rows, int k0, int K)
{
    const int mode = (L.modes >> (2 * u)) & 3;
    if (mode == 3) return;
    if (kfull && mode == 1) {
        const float4 *q = reinterpret_cast<const float4 *>(L.p0 + (u ? d1 : 0));
        it.lo = __ldg(q);
        it.hi = __ldg(q + 1);
    } else if (kfull && mode == 0) {
        it.lo = make_float4(0.f, 0.f, 0.f, 0.f);
        it.hi = it.lo;
    } else {
        load_item(it, base, ld, kmajor, vec, lt + u * LOAD_THREADS, row0, rows, k0, K);
    }
}

// What one CTA does for output tile `tile`: rows [m0, m0+128) of A, rows [nb0, nb_end) of B (its share of the tile's
// columns), and the accumulator it drains: 128 rows x ncols columns starting at (m0, n0).
struct TileWork {
    int z, m0, n0, ncols, nb0, nb_end;
};

template <bool PAIR>
__device__ __forceinline__ TileWork tile_work(const Params &P, int tile, uint32_t crank)
{
    constexpr int TM = PAIR ? 2 * BM : BM, TN = PAIR ? 2 * BN : BN;
    TileWork w;
    const int ni = tile % P.nt;
    const int rest = tile / P.nt;
    const int mi = rest % P.mt;
    w.z = rest / P.mt;
    w.m0 = mi * TM + (PAIR ? (int)crank * BM : 0);
    w.n0 = ni * TN;
    w.ncols = min(TN, (P.N - w.n0 + 15) & ~15);            // UMMA N: a multiple of 16 (M = 128 and M = 256)
    const int share = PAIR ? w.ncols / 2 : w.ncols;        // cta_group::2: each CTA supplies half of the B rows
    w.nb0 = w.n0 + (PAIR ? (int)crank * share : 0);
    w.nb_end = min(P.N, w.nb0 + share);
    return w;
}

template <bool PAIR>
__global__ void __launch_bounds__(THREADS, 1)
gemm3_kernel(const Params P)
{
    constexpr int TN = PAIR ? 2 * BN : BN;                   // accumulator columns per buffer
    extern __shared__ uint8_t smem_dyn[];
    uint8_t *smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);      // swizzle atoms need 1 KiB alignment
    uint8_t *stages = smem;
    float *epi = reinterpret_cast<float *>(smem + STAGES * STAGE_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES + EPI_BYTES);
    uint64_t *full = bars, *empty = bars + STAGES, *tm_full = bars + 2 * STAGES, *tm_empty = tm_full + NBUF;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tm_empty + NBUF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = crank == 0;
    const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;          // CTA (pair) index
    const int units = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int my_tiles = (P.tiles - unit + units - 1) / units;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + s, (PAIR ? 2 : 1) * LOAD_WARPS);     // pair: both CTAs' loaders fill the leader's barrier
            mbar_init(empty + s, 1);
        }
        for (int s = 0; s < NBUF; ++s) {
            mbar_init(tm_full + s, 1);
            mbar_init(tm_empty + s, (PAIR ? 2 : 1) * EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == EPI_WARPS) {
        if constexpr (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(NBUF * TN));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(NBUF * TN));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all();                  // the peer's barriers exist before anything signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp > EPI_WARPS) {
        // ================================ loaders ================================
        // One register set: as soon as an item has been split and stored, the same registers receive the item's successor
        // from the next K chunk, so every load has (almost) a whole iteration to arrive.
        const int lt = threadIdx.x - 32 * (EPI_WARPS + 1);
        const uint32_t full_c = PAIR ? mapa_rank0(full) : 0u;
        OperandLane LA, LB;
        LA.soff0 = item_smem_offset(P.a_kmajor, lt);
        LB.soff0 = item_smem_offset(P.b_kmajor, lt);
        const uint32_t sda = P.a_kmajor ? 8192u : 4096u, sdb = P.b_kmajor ? 8192u : 4096u;      // stage offset of item 1
        const long long da1 = (P.a_kmajor ? 64 : 32) * P.lda, db1 = (P.b_kmajor ? 64 : 32) * P.ldb;
        const long long adv_a = P.a_kmajor ? (long long)BK : (long long)BK * P.lda;
        const long long adv_b = P.b_kmajor ? (long long)BK : (long long)BK * P.ldb;
        const int n_it = my_tiles * P.kc;
        Item ra[2], rb[2];
        TileWork w;
        const float *Ab = P.A, *Bb = P.B;
        int j = 0, kc = 0;                                    // the chunk being fetched
        bool kfull = BK <= P.K;
        if (my_tiles > 0) {
            w = tile_work<PAIR>(P, unit, crank);
            Ab = P.A + (long long)w.z * P.sa;
            Bb = P.B + (long long)w.z * P.sb;
            lane_begin_tile(LA, Ab, P.lda, P.a_kmajor, P.a_vec, lt, w.m0, P.M, BM);
            lane_begin_tile(LB, Bb, P.ldb, P.b_kmajor, P.b_vec, lt, w.nb0, w.nb_end, PAIR ? w.ncols / 2 : w.ncols);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                lane_load(ra[u], LA, u, da1, kfull, Ab, P.lda, P.a_kmajor, P.a_vec, lt, w.m0, P.M, 0, P.K);
                lane_load(rb[u], LB, u, db1, kfull, Bb, P.ldb, P.b_kmajor, P.b_vec, lt, w.nb0, w.nb_end, 0, P.K);
            }
        }
        for (int it = 0; it < n_it; ++it) {
            const int store_b = LB.modes;                     // of the chunk in the registers (the state below moves on)
            // step the fetch state to chunk it + 1
            bool more = true;
            if (++kc == P.kc) {
                kc = 0;
                if (++j < my_tiles) {
                    w = tile_work<PAIR>(P, unit + j * units, crank);
                    Ab = P.A + (long long)w.z * P.sa;
                    Bb = P.B + (long long)w.z * P.sb;
                    lane_begin_tile(LA, Ab, P.lda, P.a_kmajor, P.a_vec, lt, w.m0, P.M, BM);
                    lane_begin_tile(LB, Bb, P.ldb, P.b_kmajor, P.b_vec, lt, w.nb0, w.nb_end, PAIR ? w.ncols / 2 : w.ncols);
                } else {
                    more = false;
                }
            } else {
                LA.p0 += adv_a;
                LB.p0 += adv_b;
            }
            const int k0 = kc * BK;
            kfull = k0 + BK <= P.K;
            const int slot = it % STAGES;
            mbar_wait(empty + slot, ((it / STAGES) & 1) ^ 1);
            if (lt == 0) stamp(P, it, 0);
            uint8_t *st = stages + (size_t)slot * STAGE_BYTES;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                store_item(ra[u], st, LA.soff0 + (u ? sda : 0u));
                if (more) lane_load(ra[u], LA, u, da1, kfull, Ab, P.lda, P.a_kmajor, P.a_vec, lt, w.m0, P.M, k0, P.K);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (((store_b >> (2 * u)) & 3) != 3) store_item(rb[u], st + OPER_BYTES, LB.soff0 + (u ? sdb : 0u));
                if (more) lane_load(rb[u], LB, u, db1, kfull, Bb, P.ldb, P.b_kmajor, P.b_vec, lt, w.nb0, w.nb_end, k0, P.K);
            }
            if (lt == 0) stamp(P, it, 1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> the tensor core's reads
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) mbar_arrive_remote(full_c + 8u * (uint32_t)slot);
                else mbar_arrive(full + slot);
            }
            if (lt == 0) stamp(P, it, 2);
        }
    } else if (warp == EPI_WARPS) {
        // ================================ MMA issuer (pair: the leader's warp only) ================================
        if (leader) {
            int it = 0;
            const uint32_t adv_a = P.a_kmajor ? 2u : 128u, adv_b = P.b_kmajor ? 2u : 128u;   // descriptor advance per K = 16 step
            for (int j = 0; j < my_tiles; ++j) {
                const TileWork w = tile_work<PAIR>(P, unit + j * units, 0u);
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((P.a_kmajor ? 0u : 1u) << 15) | ((P.b_kmajor ? 0u : 1u) << 16) |
                                       ((uint32_t)(w.ncols >> 3) << 17) | ((uint32_t)((PAIR ? 2 * BM : BM) >> 4) << 24);
                const int buf = j & (NBUF - 1);
                mbar_wait(tm_empty + buf, ((j / NBUF) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)buf * TN;
                for (int kc = 0; kc < P.kc; ++kc, ++it) {
                    const int slot = it % STAGES;
                    mbar_wait(full + slot, (it / STAGES) & 1);
                    if (lane == 0) stamp(P, it, 3);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(stages + (size_t)slot * STAGE_BYTES), sb = sa + OPER_BYTES;
                    const int ksteps = (min(BK, P.K - kc * BK) + 15) >> 4;
                    uint64_t da[3], db[3];
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
                        da[p] = P.a_kmajor ? desc_kmajor(sa + p * PIECE_BYTES) : desc_mnmajor(sa + p * PIECE_BYTES);
                        db[p] = P.b_kmajor ? desc_kmajor(sb + p * PIECE_BYTES) : desc_mnmajor(sb + p * PIECE_BYTES);
                    }
                    if (elect_one()) {
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint64_t oa = (uint64_t)(adv_a * ks), ob = (uint64_t)(adv_b * ks);
                            // smallest products first
                            tc_mma<PAIR>(d, da[2] + oa, db[0] + ob, idesc, (kc | ks) != 0);
                            tc_mma<PAIR>(d, da[1] + oa, db[1] + ob, idesc, 1u);
                            tc_mma<PAIR>(d, da[0] + oa, db[2] + ob, idesc, 1u);
                            tc_mma<PAIR>(d, da[1] + oa, db[0] + ob, idesc, 1u);
                            tc_mma<PAIR>(d, da[0] + oa, db[1] + ob, idesc, 1u);
                            tc_mma<PAIR>(d, da[0] + oa, db[0] + ob, idesc, 1u);
                        }
                        tc_commit_to<PAIR>(empty + slot);                       // pair: both CTAs' loaders may refill the stage
                        if (kc == P.kc - 1) tc_commit_to<PAIR>(tm_full + buf);  // ... and both epilogues may drain the tile
                    }
                    __syncwarp();
                    if (lane == 0) stamp(P, it, 4);
                }
            }
        }
    } else {
        // ================================ epilogue ================================
        float *stg = epi + warp * 32 * EPI_PITCH;
        const uint32_t tm_empty_c = PAIR ? mapa_rank0(tm_empty) : 0u;
        for (int j = 0; j < my_tiles; ++j) {
            const TileWork w = tile_work<PAIR>(P, unit + j * units, crank);
            const int buf = j & (NBUF - 1);
            mbar_wait(tm_full + buf, (j / NBUF) & 1);
            tc_fence_after();
            float *Dz = P.D + (long long)w.z * P.sd;
            const int ncols = min(w.ncols, P.N - w.n0);
            const int mrow = w.m0 + warp * 32;                            // first row of this warp's lane quadrant
            if (mrow < P.M) {
                for (int cb = 0; cb * 32 < ncols; ++cb) {
                    uint32_t v[32];
                    tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * TN + cb * 32), v);
                    const int nb = w.n0 + cb * 32;
                    const int nv = min(32, P.N - nb);
                    if (P.bias) {
#pragma unroll
                        for (int c = 0; c < 32; ++c)
                            if (c < nv) v[c] = __float_as_uint(__fadd_rn(__uint_as_float(v[c]), __ldg(P.bias + nb + c)));
                    }
                    if (P.d_rowmajor) {
                        // transpose the 32 x 32 block through shared memory: every row leaves as one coalesced 128-byte store
#pragma unroll
                        for (int c = 0; c < 32; ++c) stg[lane * EPI_PITCH + c] = __uint_as_float(v[c]);
                        __syncwarp();
                        const int mv = min(32, P.M - mrow);
                        float *drow = Dz + (long long)mrow * P.ldd + nb + lane;
                        if (mv == 32 && nv == 32) {
#pragma unroll
                            for (int r = 0; r < 32; ++r) drow[(long long)r * P.ldd] = stg[r * EPI_PITCH + lane];
                        } else {
                            for (int r = 0; r < mv; ++r)
                                if (lane < nv) drow[(long long)r * P.ldd] = stg[r * EPI_PITCH + lane];
                        }
                        __syncwarp();
                    } else {
                        const int m = mrow + lane;
                        if (m < P.M) {
#pragma unroll
                            for (int c = 0; c < 32; ++c)
                                if (c < nv) Dz[(long long)(nb + c) * P.ldd + m] = __uint_as_float(v[c]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) mbar_arrive_cluster_relaxed(tm_empty_c + 8u * (uint32_t)buf);
                else mbar_arrive(tm_empty + buf);
            }
        }
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all();       // no CTA leaves while its peer can still signal its barriers / read its smem
    else __syncthreads();
    if (warp == EPI_WARPS) {
        tc_fence_after();
        if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(NBUF * TN));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(NBUF * TN));
    }
}

}  // namespace gemm
}  // namespace mlsp

static int gemm_impl(const float *A, int a_kmajor, long long lda, long long a_batch_stride, const float *B, int b_kmajor,
                     long long ldb, long long b_batch_stride, float *D, int d_rowmajor, long long ldd,
                     long long d_batch_stride, const float *bias, int M, int N, int K, int batch, long long *tstamp, void *stream)
{
    using namespace mlsp;
    using namespace mlsp::gemm;
    MLSP_REQUIRE(A && B && D, MLSP_EINVAL, "mlsp_gemm_f32: null operand");
    MLSP_REQUIRE(M >= 0 && N >= 0 && K > 0 && batch >= 0, MLSP_EINVAL, "mlsp_gemm_f32: bad shape M=%d N=%d K=%d batch=%d", M, N, K, batch);
    if (M == 0 || N == 0 || batch == 0) return MLSP_OK;
    MLSP_REQUIRE(lda >= (a_kmajor ? K : M) && ldb >= (b_kmajor ? K : N) && ldd >= (d_rowmajor ? N : M), MLSP_EINVAL,
                 "mlsp_gemm_f32: leading dimension smaller than the contiguous extent");
    Params P;
    P.A = A; P.B = B; P.D = D; P.bias = bias;
    P.lda = lda; P.ldb = ldb; P.ldd = ldd;
    P.sa = a_batch_stride; P.sb = b_batch_stride; P.sd = d_batch_stride;
    P.M = M; P.N = N; P.K = K; P.batch = batch;
    P.tstamp = tstamp;
    P.a_kmajor = a_kmajor ? 1 : 0; P.b_kmajor = b_kmajor ? 1 : 0; P.d_rowmajor = d_rowmajor ? 1 : 0;
    P.a_vec = ((reinterpret_cast<uintptr_t>(A) & 15) == 0 && lda % 4 == 0 && a_batch_stride % 4 == 0) ? 1 : 0;
    P.b_vec = ((reinterpret_cast<uintptr_t>(B) & 15) == 0 && ldb % 4 == 0 && b_batch_stride % 4 == 0) ? 1 : 0;
    P.kc = (K + BK - 1) / BK;
    const int sms = sm_count();
    // CTA pairs on 256 x 256 tiles when that still fills the machine (or the product is big enough that the halved operand
    // stream matters more than a partial last wave); single CTAs on 128 x 128 tiles otherwise
    const long long pair_tiles = (long long)((M + 2 * BM - 1) / (2 * BM)) * ((N + 2 * BN - 1) / (2 * BN)) * batch;
    const bool pair = N > BN / 2 && pair_tiles >= sms / 4;
    P.mt = pair ? (M + 2 * BM - 1) / (2 * BM) : (M + BM - 1) / BM;
    P.nt = pair ? (N + 2 * BN - 1) / (2 * BN) : (N + BN - 1) / BN;
    const long long tiles = (long long)P.mt * P.nt * batch;
    MLSP_REQUIRE(tiles < (1ll << 30), MLSP_EUNSUPPORTED, "mlsp_gemm_f32: too many output tiles");
    P.tiles = (int)tiles;
    static thread_local int configured_dev = -1;
    int dev = 0;
    MLSP_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        MLSP_CUDA(cudaFuncSetAttribute(gemm3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        MLSP_CUDA(cudaFuncSetAttribute(gemm3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        configured_dev = dev;
    }
    if (!pair) {
        const int grid = (int)std::min<long long>(tiles, sms);
        gemm3_kernel<false><<<grid, THREADS, SMEM_BYTES, as_stream(stream)>>>(P);
    } else {
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.gridDim = dim3((unsigned)(2 * std::min<long long>(tiles, sms / 2)), 1, 1);
        cfg.blockDim = dim3(THREADS, 1, 1);
        cfg.dynamicSmemBytes = SMEM_BYTES;
        cfg.stream = as_stream(stream);
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        MLSP_CUDA(cudaLaunchKernelEx(&cfg, gemm3_kernel<true>, P));
    }
    MLSP_LAUNCH_CHECK("gemm3_kernel");
    return MLSP_OK;
}

extern "C" int mlsp_gemm_f32(const float *A, int a_kmajor, long long lda, long long a_batch_stride, const float *B, int b_kmajor,
                             long long ldb, long long b_batch_stride, float *D, int d_rowmajor, long long ldd,
                             long long d_batch_stride, const float *bias, int M, int N, int K, int batch, void *stream)
{
    return gemm_impl(A, a_kmajor, lda, a_batch_stride, B, b_kmajor, ldb, b_batch_stride, D, d_rowmajor, ldd, d_batch_stride, bias, M, N,
                     K, batch, nullptr, stream);
}

// Measurement hook: the same call, with CTA 0 writing SM-clock stamps of its first 64 K-chunk iterations to tstamp (64 x 8
// int64, device): [0] loaders see the stage free, [1] pieces stored, [2] arrived on `full`, [3] MMA warp sees the stage full,
// [4] MMAs issued and committed.  tools/gemm_timeline.py turns them into the per-iteration phase table of DESIGN.md section 10.
extern "C" int mlsp_gemm_f32_timeline(const float *A, int a_kmajor, long long lda, long long a_batch_stride, const float *B,
                                      int b_kmajor, long long ldb, long long b_batch_stride, float *D, int d_rowmajor, long long ldd,
                                      long long d_batch_stride, const float *bias, int M, int N, int K, int batch, long long *tstamp,
                                      void *stream)
{
    MLSP_REQUIRE(tstamp, MLSP_EINVAL, "mlsp_gemm_f32_timeline: null tstamp");
    return gemm_impl(A, a_kmajor, lda, a_batch_stride, B, b_kmajor, ldb, b_batch_stride, D, d_rowmajor, ldd, d_batch_stride, bias, M, N,
                     K, batch, tstamp, stream);
}
