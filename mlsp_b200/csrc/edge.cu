// edge.cu -- a2: get_graph_feature(x, args, k, idx) forward and backward
// (PointDA/model_utils.py:18-42 == PointSegDA/Models.py:18-45).
//
// The reference materialises transpose.contiguous -> gather -> repeat -> cat and returns a permuted view
// whose memory order is [B][N][k][2C] (exactly NCHW channels_last).  Here: one tiled transpose of x to
// point-major rows xt (B,N,C) (8-16 MB, stays in the 126 MB L2), then a single pass that reads each
// neighbour row with 128-bit loads and streams the (i,j) rows of the output with 128-bit evict-first
// stores.  HBM traffic is the algorithmic 4BCN + 8BNk + 8BCNk bytes; the kernel is bandwidth bound.
// Backward: one pass over grad_out accumulates the centre terms in registers and scatters the
// neighbour terms with vector float atomics into gxt (B,N,C) in L2, then a transpose back to (B,C,N).
// C = 3 (the two 3-D layers) has its own forward (no transposed copy at all) and backward (padded scratch).
#include <stdlib.h>

#include "common.cuh"

namespace mlsp {

size_t edge_workspace_bytes(int B, int C, int N, int k)
{
    (void)k;
    return align_up(sizeof(float) * (size_t)B * (C < 4 ? 4 : C) * N, 256);   // C = 3 backward: padded (B,N,4) scratch
}

size_t knn_workspace_bytes(int B, int C, int N, int k);
bool knn_tensor_supported(int B, int C, int N, int k);
bool knn3_supported(int C, int N, int k);
int knn3_run(const float *x, int B, int N, int k, int64_t *idx, int *stats, float *edge_out, cudaStream_t st);
const float *knn_tensor_xt(const void *ws, int B, int C, int N, int k);
int knn_tensor_run(const float *x, int B, int C, int N, int k, int64_t *idx, void *ws, float *dump, float *edge_out,
                   int stages_mask, cudaStream_t st, long long *tstamp = nullptr, int cluster = 0, bool want_stats = false);

size_t graph_feature_workspace_bytes(int B, int C, int N, int k)
{
    const size_t a = knn_workspace_bytes(B, C, N, k), b = edge_workspace_bytes(B, C, N, k);
    return a > b ? a : b;
}

// (B,R,S) -> (B,S,R), 32x32 tiles through shared memory
__global__ void transpose_kernel(const float *__restrict__ in, float *__restrict__ out, int R, int S)
{
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const float *ib = in + (size_t)b * R * S;
    float *ob = out + (size_t)b * R * S;
    const int s0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
        const int r = r0 + rr, s = s0 + threadIdx.x;
        tile[rr][threadIdx.x] = (r < R && s < S) ? ib[(size_t)r * S + s] : 0.0f;
    }
    __syncthreads();
    for (int ss = threadIdx.y; ss < 32; ss += blockDim.y) {
        const int s = s0 + ss, r = r0 + threadIdx.x;
        if (r < R && s < S) ob[(size_t)s * R + r] = tile[threadIdx.x][ss];
    }
}

// (B,R,S) -> (B,S,R) for R % 4 == 0, S % 4 == 0, 16-byte aligned: 64 x 64 tiles, 128-bit global loads and stores (the 32 x 32
// scalar kernel above ran the (B,N,C) -> (B,C,N) transpose at the end of edge_gather_bwd at a third of the HBM rate: 2048 CTAs
// of 4 KB each).  256 threads: 4 float4 in and 4 float4 out per thread.
__global__ void __launch_bounds__(256)
transpose64_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, int R, int S)
{
    __shared__ float tile[64][65];                        // [s][r]
    const int b = blockIdx.z;
    const int s0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
    const int S4 = S >> 2, R4 = R >> 2;
    const float4 *ib = in + (size_t)b * R * S4;
    float4 *ob = out + (size_t)b * S * R4;
    const int c4 = threadIdx.x & 15, row = threadIdx.x >> 4;          // 16 float4 per tile row, 16 tile rows per pass
    float4 v[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int r = r0 + row + 16 * p, sq = (s0 >> 2) + c4;
        v[p] = (r < R && sq < S4) ? __ldg(ib + (size_t)r * S4 + sq) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int rr = row + 16 * p;
        tile[4 * c4 + 0][rr] = v[p].x;
        tile[4 * c4 + 1][rr] = v[p].y;
        tile[4 * c4 + 2][rr] = v[p].z;
        tile[4 * c4 + 3][rr] = v[p].w;
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int ss = row + 16 * p, s = s0 + ss, rq = (r0 >> 2) + c4;
        if (s < S && rq < R4)
            ob[(size_t)s * R4 + rq] = make_float4(tile[ss][4 * c4 + 0], tile[ss][4 * c4 + 1], tile[ss][4 * c4 + 2], tile[ss][4 * c4 + 3]);
    }
}

static int launch_transpose(const float *in, float *out, int B, int R, int S, cudaStream_t st)
{
    if ((R & 3) == 0 && (S & 3) == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
        dim3 grid((S + 63) / 64, (R + 63) / 64, B);
        transpose64_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4 *>(in), reinterpret_cast<float4 *>(out), R, S);
        MLSP_LAUNCH_CHECK("transpose64_kernel");
        return MLSP_OK;
    }
    dim3 grid((S + 31) / 32, (R + 31) / 32, B);
    transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(in, out, R, S);
    MLSP_LAUNCH_CHECK("transpose_kernel");
    return MLSP_OK;
}

// ---- forward, C % 4 == 0: one warp per query point, 128-bit lanes over the 2C channels ------------
// Q4 = C/4 float4 per point row.  Lane l handles float4 slots q = l, l+32, ... of the 2*Q4-wide output row.
template <int UNROLL>
__global__ void __launch_bounds__(256)
edge_fwd_vec_kernel(const float4 *__restrict__ xt, const int64_t *__restrict__ idx, int N, int k, int Q4,
                    float4 *__restrict__ out, long long total_points)
{
    const int lane = threadIdx.x & 31;
    const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // b*N + i
    if (p >= total_points) return;
    const long long b = p / N;
    const float4 *xtb = xt + b * (long long)N * Q4;
    const float4 *ctr_row = xt + p * Q4;
    const int64_t *irow = idx + p * k;
    float4 *orow = out + p * (long long)k * (2 * Q4);
    const int W = 2 * Q4;  // float4 per output row

    for (int q = lane; q < W; q += 32) {
        const bool is_diff = q < Q4;
        const int qc = is_diff ? q : q - Q4;
        const float4 ctr = ctr_row[qc];
        int j = 0;
        for (; j + UNROLL <= k; j += UNROLL) {
            float4 nb[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const long long n = is_diff ? irow[j + u] : 0;
                nb[u] = is_diff ? xtb[n * Q4 + qc] : ctr;
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                float4 o = nb[u];
                if (is_diff) {
                    o.x = __fsub_rn(o.x, ctr.x);
                    o.y = __fsub_rn(o.y, ctr.y);
                    o.z = __fsub_rn(o.z, ctr.z);
                    o.w = __fsub_rn(o.w, ctr.w);
                }
                st_stream_f4(orow + (long long)(j + u) * W + q, o);
            }
        }
        for (; j < k; ++j) {
            float4 o = ctr;
            if (is_diff) {
                const float4 nb = xtb[(long long)irow[j] * Q4 + qc];
                o.x = __fsub_rn(nb.x, ctr.x);
                o.y = __fsub_rn(nb.y, ctr.y);
                o.z = __fsub_rn(nb.z, ctr.z);
                o.w = __fsub_rn(nb.w, ctr.w);
            }
            st_stream_f4(orow + (long long)j * W + q, o);
        }
    }
}

// ---- forward, C = 3 (the two 3-D DGCNN layers): one thread per 128-bit piece of the output ------------
// The output row of (point p, neighbour j) is 6 floats (d0 d1 d2 c0 c1 c2), so three float4 cover two rows:
//   piece 0 = row 2m (d0 d1 d2 c0), piece 1 = row 2m (c1 c2) + row 2m+1 (d0 d1), piece 2 = row 2m+1 (d2 c0 c1 c2).
// Consecutive threads write consecutive float4 (fully coalesced streaming stores); x is read in its own
// (B,3,N) layout (12 KB per cloud: L1 hits), so no transposed copy is needed.
__global__ void __launch_bounds__(256)
edge_fwd3_kernel(const float *__restrict__ x, const int64_t *__restrict__ idx, uint32_t N, uint32_t k,
                 float4 *__restrict__ out, uint32_t total4)
{
    const uint32_t t = blockIdx.x * 256u + threadIdx.x;
    if (t >= total4) return;
    const uint32_t m = t / 3u, s = t - 3u * m;
    const uint32_t r0 = 2u * m, r1 = r0 + 1u;
    const uint32_t ra = (s == 0u) ? r0 : r1;               // the row whose difference terms this piece holds
    const uint32_t pa = ra / k, ba = pa / N, ia = pa - ba * N;
    const float *xa = x + (size_t)ba * 3u * N;
    const uint32_t n = (uint32_t)idx[ra];
    float4 o;
    if (s == 0u) {
        const float c0 = __ldg(xa + ia), c1 = __ldg(xa + N + ia), c2 = __ldg(xa + 2u * N + ia);
        o = make_float4(__fsub_rn(__ldg(xa + n), c0), __fsub_rn(__ldg(xa + N + n), c1),
                        __fsub_rn(__ldg(xa + 2u * N + n), c2), c0);
    } else if (s == 2u) {
        const float c0 = __ldg(xa + ia), c1 = __ldg(xa + N + ia), c2 = __ldg(xa + 2u * N + ia);
        o = make_float4(__fsub_rn(__ldg(xa + 2u * N + n), c2), c0, c1, c2);
    } else {
        const uint32_t p0 = r0 / k, b0 = p0 / N, i0 = p0 - b0 * N;
        const float *x0 = x + (size_t)b0 * 3u * N;
        o = make_float4(__ldg(x0 + N + i0), __ldg(x0 + 2u * N + i0), __fsub_rn(__ldg(xa + n), __ldg(xa + ia)),
                        __fsub_rn(__ldg(xa + N + n), __ldg(xa + N + ia)));
    }
    st_stream_f4(out + t, o);
}

// ---- forward, any C (C % 4 != 0 and not the 3-D fast path): one thread per output element, fully coalesced stores ----------
__global__ void __launch_bounds__(256)
edge_fwd_scalar_kernel(const float *__restrict__ xt, const int64_t *__restrict__ idx, int N, int k, int C,
                       float *__restrict__ out, long long total)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int W = 2 * C;
    const long long row = e / W;  // (b*N + i)*k + j
    const int c = (int)(e - row * W);
    const long long p = row / k;  // b*N + i
    const long long b = p / N;
    float v;
    if (c < C) {
        const long long n = idx[row];
        v = __fsub_rn(xt[(b * N + n) * C + c], xt[p * C + c]);
    } else {
        v = xt[p * C + (c - C)];
    }
    __stcs(out + e, v);
}

// ---- backward --------------------------------------------------------------------------------------
// gxt (B,N,C) zero-initialised.  One warp per query point.  The warp is split into G = 32/Q4 groups of Q4
// lanes when a point row is narrower than the warp (C = 64: two groups), each group taking every G-th
// neighbour, so all 32 lanes issue 128-bit loads; four neighbours per lane are in flight before their
// vector atomics (red.global.add.v4.f32 into the L2-resident gxt) are issued.
template <int UNROLL>
__global__ void __launch_bounds__(256)
edge_bwd_vec_kernel(const float4 *__restrict__ g, const int64_t *__restrict__ idx, int N, int k, int Q4,
                    float4 *__restrict__ gxt, long long total_points)
{
    const int lane = threadIdx.x & 31;
    const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= total_points) return;
    const long long b = p / N;
    float4 *gb = gxt + b * (long long)N * Q4;
    const int64_t *irow = idx + p * k;
    const float4 *grow = g + p * (long long)k * (2 * Q4);
    const int W = 2 * Q4;
    const int G = (Q4 < 32 && (32 % Q4) == 0) ? 32 / Q4 : 1;      // neighbour groups per warp
    const int grp = (G > 1) ? lane / Q4 : 0;
    const int q0 = (G > 1) ? lane - grp * Q4 : lane;
    const int qstep = (G > 1) ? Q4 : 32;                          // G > 1: a single q per lane
    for (int q = q0; q < Q4; q += qstep) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int j = grp;
        for (; j + (UNROLL - 1) * G < k; j += UNROLL * G) {
            float4 gd[UNROLL], gc[UNROLL];
            long long n[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                gd[u] = __ldcs(grow + (long long)(j + u * G) * W + q);        // d out / d (nbr - ctr)
                gc[u] = __ldcs(grow + (long long)(j + u * G) * W + Q4 + q);   // d out / d ctr copy
                n[u] = irow[j + u * G];
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                acc.x += gc[u].x - gd[u].x;
                acc.y += gc[u].y - gd[u].y;
                acc.z += gc[u].z - gd[u].z;
                acc.w += gc[u].w - gd[u].w;
                atomicAdd(gb + n[u] * Q4 + q, gd[u]);
            }
        }
        for (; j < k; j += G) {
            const float4 gd = __ldcs(grow + (long long)j * W + q);
            const float4 gc = __ldcs(grow + (long long)j * W + Q4 + q);
            acc.x += gc.x - gd.x;
            acc.y += gc.y - gd.y;
            acc.z += gc.z - gd.z;
            acc.w += gc.w - gd.w;
            atomicAdd(gb + (long long)irow[j] * Q4 + q, gd);
        }
        for (int o = Q4; o < 32 && G > 1; o <<= 1) {             // fold the groups' centre terms
            acc.x += __shfl_xor_sync(MLSP_FULL, acc.x, o);
            acc.y += __shfl_xor_sync(MLSP_FULL, acc.y, o);
            acc.z += __shfl_xor_sync(MLSP_FULL, acc.z, o);
            acc.w += __shfl_xor_sync(MLSP_FULL, acc.w, o);
        }
        if (grp == 0) atomicAdd(gxt + p * Q4 + q, acc);
    }
}

// ---- backward, C = 3: one warp per query point, one lane per neighbour --------------------------------
// Lane j reads the 6 gradients of output row (i, j) as three float2 (the warp covers the point's k*24
// contiguous bytes) and adds (d0, d1, d2, 0) to the neighbour's slot of a padded point-major scratch
// gx4 (B,N,4) with ONE 128-bit reduction (red.global.add.v4.f32, L2 resident); the centre terms sum_j (c - d)
// are reduced over the warp and added with one more.  A second tiny kernel turns gx4 into grad_x (B,3,N).
// (Shared-memory accumulation is not an option: fp32 atomicAdd on shared memory is a CAS loop on sm_100.)
__global__ void __launch_bounds__(256)
edge_bwd3_kernel(const float *__restrict__ g, const int64_t *__restrict__ idx, int N, int k, float4 *__restrict__ gx4,
                 long long total_points)
{
    const int lane = threadIdx.x & 31;
    const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // b*N + i
    if (p >= total_points) return;
    const long long b = p / N;
    float4 *gb = gx4 + b * N;
    const float2 *gr = reinterpret_cast<const float2 *>(g + p * k * 6);
    const int64_t *ir = idx + p * k;
    float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f;
    for (int j = lane; j < k; j += 32) {
        const float2 a = __ldcs(gr + 3 * j), m = __ldcs(gr + 3 * j + 1), c = __ldcs(gr + 3 * j + 2);   // d0 d1 | d2 c0 | c1 c2
        const long long n = ir[j];
        atomicAdd(gb + n, make_float4(a.x, a.y, m.x, 0.0f));
        c0 += m.y - a.x;
        c1 += c.x - a.y;
        c2 += c.y - m.x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c0 += __shfl_xor_sync(MLSP_FULL, c0, o);
        c1 += __shfl_xor_sync(MLSP_FULL, c1, o);
        c2 += __shfl_xor_sync(MLSP_FULL, c2, o);
    }
    if (lane == 0) atomicAdd(gx4 + p, make_float4(c0, c1, c2, 0.0f));
}

// gx4 (B,N,4) -> grad_x (B,3,N)
__global__ void __launch_bounds__(256)
edge_bwd3_finish_kernel(const float4 *__restrict__ gx4, int N, float *__restrict__ gx, long long total_points)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total_points) return;
    const long long b = p / N, n = p - b * N;
    const float4 v = gx4[p];
    float *o = gx + b * 3 * N + n;
    o[0] = v.x;
    o[N] = v.y;
    o[2 * (long long)N] = v.z;
}

__global__ void __launch_bounds__(256)
edge_bwd_scalar_kernel(const float *__restrict__ g, const int64_t *__restrict__ idx, int N, int k, int C,
                       float *__restrict__ gxt, long long total_rows)
{
    // one thread per (row, c): row = (b*N+i)*k + j
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_rows * C) return;
    const long long row = t / C;
    const int c = (int)(t - row * C);
    const long long p = row / k;
    const long long b = p / N;
    const float gd = g[row * 2 * C + c];
    const float gc = g[row * 2 * C + C + c];
    atomicAdd(gxt + (b * N + idx[row]) * C + c, gd);
    atomicAdd(gxt + p * C + c, gc - gd);
}

}  // namespace mlsp

namespace mlsp {
static int launch_edge_fwd_vec(const float *xt, const int64_t *idx, int B, int C, int N, int k, float *out, cudaStream_t st)
{
    const long long points = (long long)B * N;
    const int warps = 8;
    const long long blocks = (points + warps - 1) / warps;
    MLSP_REQUIRE(blocks < (1ll << 31), MLSP_EUNSUPPORTED, "edge_gather_fwd: too many points");
    edge_fwd_vec_kernel<4><<<(unsigned)blocks, warps * 32, 0, st>>>(
        reinterpret_cast<const float4 *>(xt), idx, N, k, C / 4, reinterpret_cast<float4 *>(out), points);
    MLSP_LAUNCH_CHECK("edge_fwd_vec_kernel");
    return MLSP_OK;
}
}  // namespace mlsp

extern "C" int mlsp_edge_gather_fwd(const float *x, const int64_t *idx, int B, int C, int N, int k, float *out,
                                    void *ws, size_t ws_bytes, void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(x && idx && out && ws, MLSP_EINVAL, "edge_gather_fwd: null pointer");
    MLSP_REQUIRE(B > 0 && C > 0 && N > 0 && k > 0, MLSP_EINVAL, "edge_gather_fwd: bad shape");
    MLSP_REQUIRE(ws_bytes >= edge_workspace_bytes(B, C, N, k), MLSP_EWORKSPACE, "edge_gather_fwd: workspace too small");
    MLSP_REQUIRE(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(ws)) & 15) == 0 || C % 4 != 0, MLSP_EINVAL,
                 "edge_gather_fwd: out and ws must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const long long points = (long long)B * N;
    if (C == 3 && (points * k) % 2 == 0 && points * k < (1ll << 31) && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        const uint32_t total4 = (uint32_t)(points * k * 6 / 4);
        edge_fwd3_kernel<<<(total4 + 255u) / 256u, 256, 0, st>>>(x, idx, (uint32_t)N, (uint32_t)k,
                                                                  reinterpret_cast<float4 *>(out), total4);
        MLSP_LAUNCH_CHECK("edge_fwd3_kernel");
        return MLSP_OK;
    }
    float *xt = static_cast<float *>(ws);
    int rc = launch_transpose(x, xt, B, C, N, st);  // (B,C,N) -> (B,N,C)
    if (rc) return rc;
    if (C % 4 == 0) {
        return launch_edge_fwd_vec(xt, idx, B, C, N, k, out, st);
    } else {
        const long long total = points * k * 2 * C;
        const long long blocks = (total + 255) / 256;
        MLSP_REQUIRE(blocks < (1ll << 31), MLSP_EUNSUPPORTED, "edge_gather_fwd: too many elements");
        edge_fwd_scalar_kernel<<<(unsigned)blocks, 256, 0, st>>>(xt, idx, N, k, C, out, total);
        MLSP_LAUNCH_CHECK("edge_fwd_scalar_kernel");
    }
    return MLSP_OK;
}

extern "C" int mlsp_edge_gather_bwd(const float *grad_out, const int64_t *idx, int B, int C, int N, int k,
                                    float *grad_x, void *ws, size_t ws_bytes, void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(grad_out && idx && grad_x && ws, MLSP_EINVAL, "edge_gather_bwd: null pointer");
    MLSP_REQUIRE(B > 0 && C > 0 && N > 0 && k > 0, MLSP_EINVAL, "edge_gather_bwd: bad shape");
    MLSP_REQUIRE(ws_bytes >= edge_workspace_bytes(B, C, N, k), MLSP_EWORKSPACE, "edge_gather_bwd: workspace too small");
    MLSP_REQUIRE(((reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(ws)) & 15) == 0 || C % 4 != 0, MLSP_EINVAL,
                 "edge_gather_bwd: grad_out and ws must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    if (C == 3 && (long long)B * N < (1ll << 31) - 256) {
        const long long points = (long long)B * N;
        float4 *gx4 = static_cast<float4 *>(ws);
        MLSP_CUDA(cudaMemsetAsync(gx4, 0, sizeof(float4) * (size_t)points, st));
        edge_bwd3_kernel<<<(unsigned)((points + 7) / 8), 256, 0, st>>>(grad_out, idx, N, k, gx4, points);
        MLSP_LAUNCH_CHECK("edge_bwd3_kernel");
        edge_bwd3_finish_kernel<<<(unsigned)((points + 255) / 256), 256, 0, st>>>(gx4, N, grad_x, points);
        MLSP_LAUNCH_CHECK("edge_bwd3_finish_kernel");
        return MLSP_OK;
    }
    float *gxt = static_cast<float *>(ws);
    MLSP_CUDA(cudaMemsetAsync(gxt, 0, sizeof(float) * (size_t)B * C * N, st));
    const long long points = (long long)B * N;
    if (C % 4 == 0) {
        const int warps = 8;
        const long long blocks = (points + warps - 1) / warps;
        MLSP_REQUIRE(blocks < (1ll << 31), MLSP_EUNSUPPORTED, "edge_gather_bwd: too many points");
        edge_bwd_vec_kernel<4><<<(unsigned)blocks, warps * 32, 0, st>>>(
            reinterpret_cast<const float4 *>(grad_out), idx, N, k, C / 4, reinterpret_cast<float4 *>(gxt), points);
        MLSP_LAUNCH_CHECK("edge_bwd_vec_kernel");
    } else {
        const long long total = points * k * C;
        const long long blocks = (total + 255) / 256;
        MLSP_REQUIRE(blocks < (1ll << 31), MLSP_EUNSUPPORTED, "edge_gather_bwd: too many elements");
        edge_bwd_scalar_kernel<<<(unsigned)blocks, 256, 0, st>>>(grad_out, idx, N, k, C, gxt, points * k);
        MLSP_LAUNCH_CHECK("edge_bwd_scalar_kernel");
    }
    return launch_transpose(gxt, grad_x, B, N, C, st);  // (B,N,C) -> (B,C,N)
}

// get_graph_feature(x, args, k) with idx=None in one call: knn + edge gather.  On the tcgen05 path the point-major
// copy of x that the kNN prep kernel leaves in the workspace is gathered from directly (no second transpose).
// stages: which kernels of the tcgen05 path run (7 = the whole op; see mlsp_graph_feature_fwd_stage)
static int graph_feature_fwd_impl(const float *x, int B, int C, int N, int k, int64_t *idx, float *out, void *ws,
                                  size_t ws_bytes, int stages, void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(x && idx && out && ws, MLSP_EINVAL, "graph_feature_fwd: null pointer");
    MLSP_REQUIRE(B > 0 && C > 0 && N > 0 && k > 0, MLSP_EINVAL, "graph_feature_fwd: bad shape");
    MLSP_REQUIRE(k <= N, MLSP_EINVAL, "graph_feature_fwd: k=%d out of range for N=%d", k, N);
    MLSP_REQUIRE(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(ws)) & 15) == 0 || C % 4 != 0, MLSP_EINVAL,
                 "graph_feature_fwd: out and ws must be 16-byte aligned");
    MLSP_REQUIRE(ws_bytes >= graph_feature_workspace_bytes(B, C, N, k), MLSP_EWORKSPACE, "graph_feature_fwd: workspace too small");
    if (knn_tensor_supported(B, C, N, k))
        // tcgen05 path: the refine kernel that ranks a row also writes its edge features (no separate gather launch,
        // idx is not read back)
        return knn_tensor_run(x, B, C, N, k, idx, ws, nullptr, out, stages, as_stream(stream));
    if (knn3_supported(C, N, k) && B <= 65535 && (reinterpret_cast<uintptr_t>(out) & 7) == 0)
        // 3-D clouds: the two-pass kernel that ranks a row also writes its k x 6 edge features from the staged cloud
        return knn3_run(x, B, N, k, idx, static_cast<int *>(ws), out, as_stream(stream));
    int rc = mlsp_knn_f32(x, B, C, N, k, idx, ws, ws_bytes, MLSP_KNN_AUTO, stream);
    if (rc) return rc;
    return mlsp_edge_gather_fwd(x, idx, B, C, N, k, out, ws, ws_bytes, stream);   // stream order: the kNN is done with ws
}

extern "C" int mlsp_graph_feature_fwd(const float *x, int B, int C, int N, int k, int64_t *idx, float *out, void *ws,
                                      size_t ws_bytes, void *stream)
{
    return graph_feature_fwd_impl(x, B, C, N, k, idx, out, ws, ws_bytes, 7, stream);
}

// Measurement hook: mlsp_graph_feature_fwd restricted to some of the kernels of the tcgen05 path (stages: bit 0 = prep,
// bit 1 = tensor-core filter, bit 2 = ranking + fused edge gather).  Each kernel only reads what the earlier ones left
// in `ws`, so after one full call on the same arguments any single stage can be re-run -- and timed -- alone.
// Shapes that do not take the tcgen05 path run the whole op.
extern "C" int mlsp_graph_feature_fwd_stage(const float *x, int B, int C, int N, int k, int64_t *idx, float *out, void *ws,
                                            size_t ws_bytes, int stages, void *stream)
{
    return graph_feature_fwd_impl(x, B, C, N, k, idx, out, ws, ws_bytes, stages & 7, stream);
}
