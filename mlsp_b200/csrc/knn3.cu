// knn3.cu -- a1 for 3-D clouds (C = 3: the first two DGCNN layers and the PCA-normal neighbourhoods).
//
// Same specification as knn.cu (oracle orc_knn with C <= 4: one fmaf chain from +0, pd = rn(rn(2 dot - xx_j) - xx_i),
// rank pd descending / index ascending), different selection: with K = 3 the distance costs 5 instructions, so the
// per-candidate insertion of topk.cuh (about 20 instructions, ~100 insertions per row) dominated.  Here a warp owns
// R rows and makes two passes over the cloud (staged once per CTA in shared memory as (x,y,z,|x|^2) float4):
//   pass 1: every lane keeps, per row, the best pd of each of ITS NC candidate classes (class = (j/32 mod NC, j mod 32)).
//           The class maxima are 32*NC distinct candidates, so tau = their k-th largest value is a lower bound of
//           the k-th best pd of the row.  (warp bitonic sort of orderable 32-bit keys, one shuffle per stage)
//   pass 2: pd is recomputed and every candidate with pd >= tau (ties included -> superset of the exact top-k,
//           about 1.2 k entries expected) is recorded as one bit of a per-lane mask; the masks become the row's
//           index list in shared memory once per 1024 candidates (prefix sum of popcounts).
//   final : the list is sorted by (pd desc, index asc) with 64-bit keys; the first k are the answer.
// A row whose list overflows (heavy duplicates / degenerate clouds) is re-done with the streaming selection.
#include "common.cuh"
#include "topk.cuh"

namespace mlsp {

constexpr int K3_THREADS = 256;
constexpr int K3_R = 4;                                   // rows per warp
constexpr int K3_ROWS = (K3_THREADS / 32) * K3_R;         // 32 rows per CTA

// t = rn(2 dot - |x_j|^2): the candidate-dependent part of pd
__device__ __forceinline__ float t3(float4 a, float4 q)
{
    float d = __fmaf_rn(a.x, q.x, 0.0f);
    d = __fmaf_rn(a.y, q.y, d);
    d = __fmaf_rn(a.z, q.z, d);
    return __fmaf_rn(2.0f, d, -q.w);
}
__device__ __forceinline__ float pd3(float4 a, float4 q) { return __fsub_rn(t3(a, q), a.w); }

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

// a2 fused (get_graph_feature with idx=None on a 3-D cloud): the warp that ranked row i writes the row's k x 6 edge
// features [x_j - x_i | x_i] as 3k float2 (d0 d1 | d2 c0 | c1 c2 per neighbour), consecutive lanes -> consecutive
// float2, neighbours read from the staged cloud.  nbr[s] on lane l = the neighbour of rank s*32 + l.
template <int SL>
__device__ __forceinline__ void edge_row3(float2 *__restrict__ orow, float4 ctr, const float4 *cloud, const uint32_t (&nbr)[SL],
                                          int k, int N)
{
    const int lane = lane_id();
    for (int t0 = 0; t0 < 3 * k; t0 += 32) {
        const int t = t0 + lane;
        const int e = min(t / 3, k - 1), part = t - 3 * (t / 3);
        uint32_t n = 0;
#pragma unroll
        for (int s = 0; s < SL; ++s) {
            const uint32_t v = __shfl_sync(MLSP_FULL, nbr[s], e & 31);
            if ((e >> 5) == s) n = v;
        }
        const float4 q = cloud[min(n, (uint32_t)(N - 1))];
        float2 o;
        if (part == 0) o = make_float2(__fsub_rn(q.x, ctr.x), __fsub_rn(q.y, ctr.y));
        else if (part == 1) o = make_float2(__fsub_rn(q.z, ctr.z), ctr.x);
        else o = make_float2(ctr.y, ctr.z);
        if (t < 3 * k) __stcs(orow + t, o);
    }
}

// SURVEY.md 8f rank 2 -- ONE 3-D neighbourhood pass for the target builder: with STRUCT the same launch that ranks the
// k = near neighbours of every point of the (undeformed) target cloud also emits
//   * a6: the ball cardinality of mlsp.cal_density (python-pcl radius search semantics: d < r2 on the direct-difference
//         distance, at most K, neighbour index 0 dropped) and its soft labels -- counted in pass 1 on the staged cloud;
//   * a7: the PCA normal of kSearchNormalEstimation from the ranked neighbourhood while it is still in registers
//         (fp64 covariance about the neighbourhood mean by warp shuffles, Jacobi eigen-solve on one lane per row);
// idx and the edge rows stay optional outputs.  Arithmetic identical to target.cu's stand-alone kernels.
struct K3Struct {
    float *normals;          // (B,N,3)
    float *curvature;        // (B,N) or NULL
    float *labels;           // (B,N,num_cls) or NULL (then no cardinality work at all)
    int64_t *row;            // (B,N)
    float r2;
    int K, shift, pergroup, num_cls;
};

__device__ __forceinline__ float direct_d2_3(float4 a, float4 q)
{
    const float dx = __fsub_rn(a.x, q.x), dy = __fsub_rn(a.y, q.y), dz = __fsub_rn(a.z, q.z);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ double warp_sum_f64(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MLSP_FULL, v, o);
    return v;
}

// covariance (a00 a01 a02 a11 a12 a22) of the k ranked neighbours of one row, identical on every lane
template <int SLH>
__device__ __forceinline__ void row_covariance(const float4 *cloud, const uint32_t (&nbr)[SLH], int k, int N, double (&a)[6])
{
    const int lane = lane_id();
    double sx = 0, sy = 0, sz = 0;
    float4 q[SLH];
#pragma unroll
    for (int s = 0; s < SLH; ++s) {
        const bool live = s * 32 + lane < k;
        q[s] = cloud[min(nbr[s], (uint32_t)(N - 1))];
        if (live) { sx += q[s].x; sy += q[s].y; sz += q[s].z; }
    }
    const double mx = warp_sum_f64(sx) / k, my = warp_sum_f64(sy) / k, mz = warp_sum_f64(sz) / k;
#pragma unroll
    for (int t = 0; t < 6; ++t) a[t] = 0.0;
#pragma unroll
    for (int s = 0; s < SLH; ++s) {
        if (s * 32 + lane < k) {
            const double dx = q[s].x - mx, dy = q[s].y - my, dz = q[s].z - mz;
            a[0] += dx * dx; a[1] += dx * dy; a[2] += dx * dz; a[3] += dy * dy; a[4] += dy * dz; a[5] += dz * dz;
        }
    }
    const double inv = 1.0 / k;
#pragma unroll
    for (int t = 0; t < 6; ++t) a[t] = warp_sum_f64(a[t]) * inv;
}

// smallest-eigenvalue eigenvector of the symmetric 3x3 covariance, oriented towards the origin.  Closed form in fp64
// (trigonometric eigenvalue + the largest cross product of two rows of A - lambda I): ~5x fewer instructions than the
// cyclic Jacobi of target.cu's stand-alone kernel, which matters here because the solve of a row runs on ONE lane while
// the rest of its warp waits.  The two agree to ~1e-12 wherever the smallest eigenvalue is separated (tests: < 1e-6).
__device__ __forceinline__ void normal_from_cov(const double (&a)[6], float4 self, float *nout, float *curv)
{
    const double a00 = a[0], a01 = a[1], a02 = a[2], a11 = a[3], a12 = a[4], a22 = a[5];
    const double p1 = a01 * a01 + a02 * a02 + a12 * a12;
    const double tr = a00 + a11 + a22;
    double lam, nx, ny, nz;
    if (p1 <= 1e-36 * tr * tr) {                                       // already diagonal: the axis of the smallest entry
        lam = a00; nx = 1.0; ny = 0.0; nz = 0.0;
        if (a11 < lam) { lam = a11; nx = 0.0; ny = 1.0; }
        if (a22 < lam) { lam = a22; nx = 0.0; ny = 0.0; nz = 1.0; }
    } else {
        const double q = tr / 3.0;
        const double b00 = a00 - q, b11 = a11 - q, b22 = a22 - q;
        const double p = sqrt((b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * p1) / 6.0);
        const double ip = 1.0 / p;
        const double c00 = b00 * ip, c01 = a01 * ip, c02 = a02 * ip, c11 = b11 * ip, c12 = a12 * ip, c22 = b22 * ip;
        double r = 0.5 * (c00 * (c11 * c22 - c12 * c12) - c01 * (c01 * c22 - c12 * c02) + c02 * (c01 * c12 - c11 * c02));
        r = fmin(1.0, fmax(-1.0, r));
        const double phi = acos(r) / 3.0;
        lam = q + 2.0 * p * cos(phi + 2.0943951023931954923);          // the smallest of the three roots
        // rows of A - lam I; the eigenvector is orthogonal to all of them: take the best-conditioned cross product
        const double r0x = a00 - lam, r0y = a01, r0z = a02;
        const double r1x = a01, r1y = a11 - lam, r1z = a12;
        const double r2x = a02, r2y = a12, r2z = a22 - lam;
        const double ux = r0y * r1z - r0z * r1y, uy = r0z * r1x - r0x * r1z, uz = r0x * r1y - r0y * r1x;
        const double vx = r0y * r2z - r0z * r2y, vy = r0z * r2x - r0x * r2z, vz = r0x * r2y - r0y * r2x;
        const double wx = r1y * r2z - r1z * r2y, wy = r1z * r2x - r1x * r2z, wz = r1x * r2y - r1y * r2x;
        const double nu = ux * ux + uy * uy + uz * uz, nv = vx * vx + vy * vy + vz * vz, nw = wx * wx + wy * wy + wz * wz;
        nx = ux; ny = uy; nz = uz;
        double best = nu;
        if (nv > best) { best = nv; nx = vx; ny = vy; nz = vz; }
        if (nw > best) { best = nw; nx = wx; ny = wy; nz = wz; }
        const double nn = rsqrt(best);
        nx *= nn; ny *= nn; nz *= nn;
    }
    if (nx * self.x + ny * self.y + nz * self.z > 0.0) { nx = -nx; ny = -ny; nz = -nz; }  // towards the origin
    nout[0] = (float)nx;
    nout[1] = (float)ny;
    nout[2] = (float)nz;
    if (curv) *curv = tr > 0.0 ? (float)(fabs(lam) / tr) : 0.0f;
}

// NC = classes per lane (class of candidate j = (j / 32) mod NC, j mod 32): 32*NC classes.  More classes than k
// tighten tau: the expected list length is sum_{i<k} M/(M-i) for M classes (23.7 for k = 20, M = 64; 47.7 for
// k = 40, M = 128), so the final sort usually runs on HALF the capacity CAPL = 32*NC.
template <int NC, bool STRUCT>
__global__ void __launch_bounds__(K3_THREADS, STRUCT ? 4 : 1)      // STRUCT: <= 64 registers keeps four CTAs per SM
knn3_kernel(const float *__restrict__ x, long long xs_b, long long xs_c, long long xs_n, int N, int k, int64_t *__restrict__ idx,
            int *__restrict__ stats, float2 *__restrict__ edge_out, K3Struct S)
{
    if (stats && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 2) stats[threadIdx.x] = 0;   // {fallback rows, certified rows}: tensor path only
    constexpr int CAPL = 32 * NC;
    constexpr int SL = CAPL / 32;
    extern __shared__ float4 cloud[];                                  // [N]
    uint16_t *lists = reinterpret_cast<uint16_t *>(cloud + N);         // [warps][R][CAPL] candidate indices

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const float *xb = x + b * xs_b;                                    // (B,3,N) or the (B,N,3) batch the normals / cardinality take
    for (int n = tid; n < N; n += K3_THREADS) {
        const float *pn = xb + n * xs_n;
        const float px = pn[0], py = pn[xs_c], pz = pn[2 * xs_c];
        const float xx = __fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz));
        cloud[n] = make_float4(px, py, pz, xx);
    }
    __syncthreads();

    const int i0 = blockIdx.x * K3_ROWS + warp * K3_R;
    if (i0 >= N) return;
    float4 xi[K3_R];
#pragma unroll
    for (int rr = 0; rr < K3_R; ++rr) xi[rr] = cloud[min(i0 + rr, N - 1)];

    // ---- pass 1: class maxima (+ STRUCT: the ball cardinality counts of a6 on the same staged candidates)
    float cmax[K3_R][NC];
    const bool card = STRUCT && S.labels != nullptr;
    float d0[K3_R];
    int c1[K3_R], c2[K3_R];
#pragma unroll
    for (int rr = 0; rr < K3_R; ++rr) {
#pragma unroll
        for (int c = 0; c < NC; ++c) cmax[rr][c] = -INFINITY;
        d0[rr] = STRUCT ? direct_d2_3(xi[rr], cloud[0]) : 0.0f;
        c1[rr] = c2[rr] = 0;
    }
    for (int j0 = 0; j0 < N; j0 += 32 * NC) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int j = j0 + c * 32 + lane;
            if (j < N) {
                const float4 q = cloud[j];
#pragma unroll
                for (int rr = 0; rr < K3_R; ++rr) {
                    cmax[rr][c] = fmaxf(cmax[rr][c], t3(xi[rr], q));
                    if (STRUCT && card) {
                        const float d = direct_d2_3(xi[rr], q);
                        c1[rr] += (d < S.r2) ? 1 : 0;
                        c2[rr] += (d < d0[rr]) ? 1 : 0;
                    }
                }
            }
        }
    }
    if (STRUCT && card) {                                               // cal_density's row and soft labels (MLSP/mlsp.py:252-266)
        const int top = (S.num_cls - 1) * S.pergroup;
#pragma unroll
        for (int rr = 0; rr < K3_R; ++rr) {
            const int t1 = __reduce_add_sync(MLSP_FULL, c1[rr]);
            const int t2 = __reduce_add_sync(MLSP_FULL, c2[rr]);
            const int i = i0 + rr;
            if (i >= N) continue;
            const int in0 = (d0[rr] < S.r2 && t2 < S.K) ? 1 : 0;
            int v = min(t1, S.K) - in0 - S.shift;
            v = max(0, min(v, top));
            const int lo = v / S.pergroup, hi = (v + S.pergroup - 1) / S.pergroup;
            if (lane == 0) S.row[(size_t)b * N + i] = v;
            float *L = S.labels + ((size_t)b * N + i) * S.num_cls;
            for (int c = lane; c < S.num_cls; c += 32) L[c] = 0.5f * (float)(c == lo) + 0.5f * (float)(c == hi);
        }
    }
    // the maxima were taken on t = rn(2 dot - |x_j|^2); rn(. - |x_i|^2) is monotone, so max_j pd = rn(max_j t - |x_i|^2)
#pragma unroll
    for (int rr = 0; rr < K3_R; ++rr)
#pragma unroll
        for (int c = 0; c < NC; ++c) cmax[rr][c] = __fsub_rn(cmax[rr][c], xi[rr].w);
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

    // ---- pass 2: every candidate with pd >= tau (ties included -> superset of the exact top-k).  Lane l tests
    // candidates j = blk0 + 32 c + l and records hits as bit c of a per-row mask (one predicated OR per test, no
    // votes); after each block of 1024 candidates the masks are turned into list entries: warp prefix sum of the
    // popcounts, then every lane appends its own hits.  List order is irrelevant (the list is sorted below).
    uint16_t *my = lists + (size_t)warp * K3_R * CAPL;
    int cnt[K3_R];
#pragma unroll
    for (int rr = 0; rr < K3_R; ++rr) cnt[rr] = 0;
    for (int blk0 = 0; blk0 < N; blk0 += 1024) {
        uint32_t hm[K3_R];
#pragma unroll
        for (int rr = 0; rr < K3_R; ++rr) hm[rr] = 0u;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            if (blk0 + c * 32 < N) {                                    // warp-uniform
                const int j = blk0 + c * 32 + lane;
                const float4 q = cloud[min(j, N - 1)];
#pragma unroll
                for (int rr = 0; rr < K3_R; ++rr)
                    if (j < N && pd3(xi[rr], q) >= tau[rr]) hm[rr] |= 1u << c;
            }
        }
#pragma unroll
        for (int rr = 0; rr < K3_R; ++rr) {
            const int mine = __popc(hm[rr]);
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(MLSP_FULL, incl, o);
                if (lane >= o) incl += t;
            }
            int pos = cnt[rr] + incl - mine;
            cnt[rr] += __shfl_sync(MLSP_FULL, incl, 31);
            uint32_t m = hm[rr];
            while (m) {
                const int c = __ffs(m) - 1;
                m &= m - 1;
                if (pos < CAPL) my[rr * CAPL + pos] = (uint16_t)(blk0 + c * 32 + lane);
                ++pos;
            }
        }
    }
    __syncwarp();

    // ---- final: exact sort of the list by (pd desc, index asc), pd recomputed from the staged cloud
    // (or the streaming selection if the list overflowed)
    double cov[6] = {0, 0, 0, 0, 0, 0};                                   // STRUCT: lane rr keeps the covariance of row rr
#pragma unroll
    for (int rr = 0; rr < K3_R; ++rr) {
        const int i = i0 + rr;
        if (i >= N) break;
        int64_t *out = idx ? idx + ((size_t)b * N + i) * k : nullptr;
        const int n_l = cnt[rr];
        uint32_t nbr[(SL + 1) / 2];                                       // k <= 16 NC: the ranked neighbours sit in the lower slots
        if (n_l <= CAPL) {
            unsigned long long key[SL];
#pragma unroll
            for (int s = 0; s < SL; ++s) {
                const int e = s * 32 + lane;
                const bool live = e < n_l;
                const int j = live ? (int)my[rr * CAPL + e] : 0;
                key[s] = rank_key(pd3(xi[rr], cloud[j]), live ? j : 0x7fffffff, live);
            }
            if (n_l <= CAPL / 2) {                                       // warp-uniform, the usual case: the upper
                unsigned long long half[SL / 2];                         // half of the slots is dead, sort the lower
#pragma unroll
                for (int s = 0; s < SL / 2; ++s) half[s] = key[s];
                warp_sort_u64<SL / 2>(half);
#pragma unroll
                for (int s = 0; s < SL / 2; ++s) key[s] = half[s];
            } else {
                warp_sort_u64<SL>(key);
            }
#pragma unroll
            for (int s = 0; s < SL; ++s) {
                const int e = s * 32 + lane;
                if (out && e < k) out[e] = (int64_t)(uint32_t)(key[s] & 0xffffffffull);
            }
#pragma unroll
            for (int s = 0; s < (SL + 1) / 2; ++s) nbr[s] = (uint32_t)(key[s] & 0xffffffffull);
        } else {
            TopK<(SL + 1) / 2> top;
            top.init(k);
            for (int j0 = 0; j0 < N; j0 += 32) {
                const int j = j0 + lane;
                const float pd = pd3(xi[rr], cloud[min(j, N - 1)]);
                top.offer(pd, j, j < N);
            }
            top.finish(k);
#pragma unroll
            for (int s = 0; s < (SL + 1) / 2; ++s) {
                const int e = s * 32 + lane;
                if (out && e < k) out[e] = (int64_t)top.j[s];
                nbr[s] = (uint32_t)top.j[s];
            }
        }
        if (edge_out) edge_row3<(SL + 1) / 2>(edge_out + ((size_t)b * N + i) * k * 3, xi[rr], cloud, nbr, k, N);
        if (STRUCT) {
            // a7: covariance of the ranked neighbourhood by the whole warp; the eigen-solve of row rr runs on lane rr, the
            // four rows of the warp side by side
            double a[6];
            row_covariance<(SL + 1) / 2>(cloud, nbr, k, N, a);
#pragma unroll
            for (int t = 0; t < 6; ++t)
                if (lane == rr) cov[t] = a[t];
        }
    }
    if (STRUCT) {
        if (lane < K3_R && i0 + lane < N) {
            float4 self = xi[0];
#pragma unroll
            for (int rr = 1; rr < K3_R; ++rr)
                if (lane == rr) self = xi[rr];
            const size_t o = (size_t)b * N + i0 + lane;
            normal_from_cov(cov, self, S.normals + 3 * o, S.curvature ? S.curvature + o : nullptr);
        }
    }
}

bool knn3_supported(int C, int N, int k) { return C == 3 && k <= 64 && N >= 1 && N <= 8192; }

template <int NC, bool STRUCT>
static int knn3_launch(const float *x, long long sb, long long sc, long long sn, int B, int N, int k, int64_t *idx, int *stats,
                       float2 *eo, const K3Struct &S, cudaStream_t st)
{
    const size_t smem = sizeof(float4) * (size_t)N + sizeof(uint16_t) * (size_t)(K3_THREADS / 32) * K3_R * 32 * NC;
    dim3 grid((N + K3_ROWS - 1) / K3_ROWS, B);
    MLSP_CUDA(cudaFuncSetAttribute(knn3_kernel<NC, STRUCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn3_kernel<NC, STRUCT><<<grid, K3_THREADS, smem, st>>>(x, sb, sc, sn, N, k, idx, stats, eo, S);
    MLSP_LAUNCH_CHECK("knn3_kernel");
    return MLSP_OK;
}

int knn3_run(const float *x, int B, int N, int k, int64_t *idx, int *stats, float *edge_out, cudaStream_t st)
{
    float2 *eo = reinterpret_cast<float2 *>(edge_out);                  // (B,N,k,6) floats: rows are 8-byte aligned
    K3Struct S = {};
    if (k <= 32) return knn3_launch<2, false>(x, 3ll * N, N, 1, B, N, k, idx, stats, eo, S, st);   // 64 classes
    return knn3_launch<4, false>(x, 3ll * N, N, 1, B, N, k, idx, stats, eo, S, st);                 // 128 classes
}

}  // namespace mlsp

// the 3-D neighbourhood of the undeformed target cloud in ONE launch (include/mlsp_b200.h)
extern "C" int mlsp_target_structure(const float *pts, int B, int N, int near, float r2, int K, int shift, int pergroup, int num_cls,
                                     float *normals, float *curvature, float *labels, int64_t *row, int64_t *idx, void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(pts && normals, MLSP_EINVAL, "target_structure: null pointer");
    MLSP_REQUIRE(B > 0 && N > 0 && near >= 1 && near <= N, MLSP_EINVAL, "target_structure: bad shape B=%d N=%d near=%d", B, N, near);
    MLSP_REQUIRE(!labels || (row && K > 0 && pergroup > 0 && num_cls > 0), MLSP_EINVAL, "target_structure: bad cardinality arguments");
    MLSP_REQUIRE(knn3_supported(3, N, near) && B <= 65535, MLSP_EUNSUPPORTED, "target_structure: N=%d near=%d outside the 3-D kernel", N, near);
    K3Struct S;
    S.normals = normals; S.curvature = curvature; S.labels = labels; S.row = row;
    S.r2 = r2; S.K = K; S.shift = shift; S.pergroup = pergroup; S.num_cls = num_cls;
    // pts is the (B,N,3) batch the trainers hand to pcl / cal_density: point stride 3, channel stride 1
    if (near <= 32) return knn3_launch<2, true>(pts, 3ll * N, 1, 3, B, N, near, idx, nullptr, nullptr, S, as_stream(stream));
    return knn3_launch<4, true>(pts, 3ll * N, 1, 3, B, N, near, idx, nullptr, nullptr, S, as_stream(stream));
}
