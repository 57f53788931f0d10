/*
 * oracle/mlsp_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, one rounding per written operation) of the arithmetic of the
 * MLSP hot path.  It is the bit-exact specification the sm_100a kernels in
 * mlsp_b200/csrc/ are checked against; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg may load it.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference checkout, VITA-Group/MLSP).  Where the reference leaves the floating-point
 * evaluation order to a library (MKL / cuBLAS sgemm, ATen reductions) this file PINS one
 * order; the pin is validated against goldens produced by the reference's own functions
 * (tests/golden/, made by oracle/gen_golden.py): strict equality on grid-quantised inputs
 * (all orders agree there), fp64-certified near-ties on continuous inputs.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off: no implicit FMA contraction;
 * the only fused operations are the explicit fmaf() calls below).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ---- shared pieces ------------------------------------------------------------------ */

/* xx[j] = sum_c x[c][j]^2, square rounded, then sequential adds in channel order.
 * Reference: `xx = torch.sum(x**2, dim=1, keepdim=True)` PointDA/model_utils.py:11,
 * PointSegDA/Models.py:10, utils/pc_utils.py:88. */
static void sq_norms(const float *x, int C, int N, float *xx)
{
    for (int j = 0; j < N; ++j) {
        float s = x[j] * x[j];
        for (int c = 1; c < C; ++c) {
            float v = x[(size_t)c * N + j];
            float q = v * v;
            s = s + q;
        }
        xx[j] = s;
    }
}

/* dot(x_i, x_j) for the 3-D ball test: first product rounded, then an fmaf chain in channel order.
 * Reference: `torch.matmul(x.transpose(1, 0), x)` utils/pc_utils.py:87 (library sgemm, order
 * unspecified there; pinned here). */
static inline float dot_chain(const float *x, int C, int N, int i, int j)
{
    float acc = x[i] * x[j];
    for (int c = 1; c < C; ++c)
        acc = fmaf(x[(size_t)c * N + i], x[(size_t)c * N + j], acc);
    return acc;
}

/* dot(x_i, x_j) for knn: the reference leaves the order to sgemm (`torch.matmul`,
 * PointDA/model_utils.py:10); pinned here as EIGHT interleaved fmaf chains and a fixed tree:
 *   p_t = fmaf chain from +0 over the channels c with (c/4) mod 8 == t, in increasing c   (t = 0..7)
 *   dot = ((p0+p4) + (p2+p6)) + ((p1+p5) + (p3+p7))
 * (groups of 4 consecutive channels go round-robin to 8 accumulators; this is the order in which
 * 8 GPU lanes read one point-major row with coalesced 16-byte loads, and the butterfly they reduce
 * with).  For C <= 4 it degenerates to one sequential chain. */
static inline float dot_tree(const float *x, int C, int N, int i, int j)
{
    float p[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < C; ++c) {
        int t = (c >> 2) & 7;
        p[t] = fmaf(x[(size_t)c * N + i], x[(size_t)c * N + j], p[t]);
    }
    float q0 = p[0] + p[4], q1 = p[1] + p[5], q2 = p[2] + p[6], q3 = p[3] + p[7];
    float r0 = q0 + q2, r1 = q1 + q3;
    return r0 + r1;
}

/* ---- a1: knn ------------------------------------------------------------------------
 * Reference: knn(x,k) PointDA/model_utils.py:9-16 == PointSegDA/Models.py:8-15.
 *   inner = -2*matmul(x^T, x); xx = sum(x**2); pd = -xx - inner - xx^T; topk(k)
 * pd[i][j] = ((-xx[j]) - inner[i][j]) - xx[i],  inner = -2*dot (exact scaling), so
 * (-xx[j]) - inner == rn(2*dot - xx[j]) == fmaf(2, dot, -xx[j]).
 * Ranking: largest pd first, ties -> lowest j (north_star tie rule; torch.topk itself is
 * not tie-stable).
 * x: (B,C,N) contiguous.  idx: (B,N,k) int64.  pd_out: (B,N,k) or NULL. */
ORC_API int orc_knn(const float *x, int B, int C, int N, int k, int64_t *idx, float *pd_out)
{
    if (k < 1 || k > N) return 1;
    float *xx = (float *)malloc(sizeof(float) * (size_t)N);
    float *bv = (float *)malloc(sizeof(float) * (size_t)k);
    int *bj = (int *)malloc(sizeof(int) * (size_t)k);
    for (int b = 0; b < B; ++b) {
        const float *xb = x + (size_t)b * C * N;
        sq_norms(xb, C, N, xx);
        for (int i = 0; i < N; ++i) {
            int cnt = 0;
            for (int j = 0; j < N; ++j) {
                float d = dot_tree(xb, C, N, i, j);
                float t = fmaf(2.0f, d, -xx[j]);
                float pd = t - xx[i];
                /* candidates arrive in increasing j: a tie never displaces an earlier one */
                if (cnt == k && !(pd > bv[k - 1])) continue;
                int p = (cnt < k) ? cnt : k - 1;
                while (p > 0 && pd > bv[p - 1]) { bv[p] = bv[p - 1]; bj[p] = bj[p - 1]; --p; }
                bv[p] = pd; bj[p] = j;
                if (cnt < k) ++cnt;
            }
            for (int r = 0; r < k; ++r) {
                idx[((size_t)b * N + i) * k + r] = bj[r];
                if (pd_out) pd_out[((size_t)b * N + i) * k + r] = bv[r];
            }
        }
    }
    free(xx); free(bv); free(bj);
    return 0;
}

/* Same ranking in fp64 from the fp32 inputs (used to certify near-ties).
 * pd64[i][j] = 2*dot - xx[j] - xx[i] evaluated in double. Full matrix row for one (b,i). */
ORC_API int orc_knn_row_f64(const float *xb, int C, int N, int i, double *row)
{
    for (int j = 0; j < N; ++j) {
        double d = 0, xj = 0, xi = 0;
        for (int c = 0; c < C; ++c) {
            double a = xb[(size_t)c * N + i], bb = xb[(size_t)c * N + j];
            d += a * bb; xj += bb * bb; xi += a * a;
        }
        row[j] = 2 * d - xj - xi;
    }
    return 0;
}

/* ---- a2: get_graph_feature ------------------------------------------------------------
 * Reference: PointDA/model_utils.py:18-42 == PointSegDA/Models.py:18-45.
 * out logical (B,2C,N,k), stored channels-last: out[((b*N+i)*k+j)*2C + c].
 *   c<C : x[b,c,idx[b,i,j]] - x[b,c,i]      c>=C : x[b,c-C,i] */
ORC_API int orc_edge_gather(const float *x, const int64_t *idx, int B, int C, int N, int k,
                            float *out)
{
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < k; ++j) {
                int64_t n = idx[((size_t)b * N + i) * k + j];
                if (n < 0 || n >= N) return 1;
                float *o = out + (((size_t)b * N + i) * k + j) * 2 * C;
                for (int c = 0; c < C; ++c) {
                    float ctr = x[((size_t)b * C + c) * N + i];
                    o[c] = x[((size_t)b * C + c) * N + n] - ctr;
                    o[C + c] = ctr;
                }
            }
    return 0;
}

/* backward of the above w.r.t. x, accumulated in double (order-free reference).
 * g: channels-last (B,N,k,2C).  gx: (B,C,N) double. */
ORC_API int orc_edge_gather_bwd(const float *g, const int64_t *idx, int B, int C, int N, int k,
                                double *gx)
{
    memset(gx, 0, sizeof(double) * (size_t)B * C * N);
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < k; ++j) {
                int64_t n = idx[((size_t)b * N + i) * k + j];
                const float *go = g + (((size_t)b * N + i) * k + j) * 2 * C;
                for (int c = 0; c < C; ++c) {
                    gx[((size_t)b * C + c) * N + n] += go[c];
                    gx[((size_t)b * C + c) * N + i] += (double)go[C + c] - (double)go[c];
                }
            }
    return 0;
}

/* ---- a3: farthest point sampling ---------------------------------------------------------
 * Reference: farthest_point_sample utils/pc_utils.py:137-161.
 *   distance = 1e10; farthest = start (torch.randint on the CPU generator, drawn by the host)
 *   loop: record; dist = sum((xyz-centroid)**2, 1); distance = min(distance, dist);
 *         farthest = argmax(distance) (lowest index on ties)
 * xyz (B,3,N); start (B); centroids (B,npoint) int64; vals (B,3,npoint). */
ORC_API int orc_fps(const float *xyz, int B, int N, int npoint, const int64_t *start,
                    int64_t *centroids, float *vals)
{
    float *dist = (float *)malloc(sizeof(float) * (size_t)N);
    for (int b = 0; b < B; ++b) {
        const float *X = xyz + (size_t)b * 3 * N;
        for (int j = 0; j < N; ++j) dist[j] = 1e10f;
        int far = (int)start[b];
        if (far < 0 || far >= N) { free(dist); return 1; }
        for (int s = 0; s < npoint; ++s) {
            centroids[(size_t)b * npoint + s] = far;
            float cx = X[far], cy = X[N + far], cz = X[2 * N + far];
            vals[((size_t)b * 3 + 0) * npoint + s] = cx;
            vals[((size_t)b * 3 + 1) * npoint + s] = cy;
            vals[((size_t)b * 3 + 2) * npoint + s] = cz;
            float best = -1.0f; int bi = 0;
            for (int j = 0; j < N; ++j) {
                float dx = X[j] - cx, dy = X[N + j] - cy, dz = X[2 * N + j] - cz;
                float qx = dx * dx, qy = dy * dy, qz = dz * dz;
                float d = qx + qy; d = d + qz;
                if (d < dist[j]) dist[j] = d;
                if (dist[j] > best) { best = dist[j]; bi = j; }
            }
            far = bi;
        }
    }
    free(dist);
    return 0;
}

/* ---- a5: ball membership counts (collapse_to_point) --------------------------------------
 * Reference: utils/pc_utils.py:86-99.
 *   inner = -2*matmul(x^T,x); xx = sum(x**2,0); pd = xx + inner + xx^T
 *   pd[i][j] = (xx[j] + inner[i][j]) + xx[i] = rn(rn(xx[j] - 2*dot) + xx[i])
 *   in-ball <=> pd <= RADIUS**2 (0.25).   cnt[i] = #in-ball (row sum of the 0/1 mask)
 * x (B,3,N). cnt (B,N) int32.  r2 passed as float (0.25f). */
static inline float ball_pd(const float *X, int N, const float *xx, int i, int j)
{
    float d = dot_chain(X, 3, N, i, j);
    float t = fmaf(-2.0f, d, xx[j]);
    return t + xx[i];
}

ORC_API int orc_ball_count(const float *x, int B, int N, float r2, int32_t *cnt)
{
    float *xx = (float *)malloc(sizeof(float) * (size_t)N);
    for (int b = 0; b < B; ++b) {
        const float *X = x + (size_t)b * 3 * N;
        sq_norms(X, 3, N, xx);
        for (int i = 0; i < N; ++i) {
            int c = 0;
            for (int j = 0; j < N; ++j) c += (ball_pd(X, N, xx, i, j) <= r2);
            cnt[(size_t)b * N + i] = c;
        }
    }
    free(xx);
    return 0;
}

/* in-ball flags of one centre row (the `point_mask = mask[point_ind, :]` of pc_utils.py:105) */
ORC_API int orc_ball_row(const float *xb, int N, float r2, int centre, uint8_t *flag)
{
    float *xx = (float *)malloc(sizeof(float) * (size_t)N);
    sq_norms(xb, 3, N, xx);
    for (int j = 0; j < N; ++j) flag[j] = ball_pd(xb, N, xx, centre, j) <= r2;
    free(xx);
    return 0;
}

/* ---- a6: per-point ball cardinality (cal_density) -----------------------------------------
 * Reference: MLSP/mlsp.py:240-272 -> python-pcl KdTreeFLANN.radius_search_for_cloud(cloud, r, K)
 * (third party, un-vendored, no version pinned anywhere in the reference: PARITY UNPINNED).
 * Restated from the published behaviour of PCL/FLANN: squared L2 by direct differences
 * (FLANN L2_Simple: diff=a-b; result += diff*diff), strict `< r*r`, at most K nearest
 * returned, result row zero-padded; the reference then counts `ind != 0`, i.e. drops
 * neighbour index 0 when it was returned.
 *   c1 = #{j: d_ij < r2};  in0 = d_i0 < r2 && #{j: d_ij < d_i0} < K
 *   cnt = min(c1, K) - in0
 * pts (B,N,3) (the layout cal_density receives).  cnt (B,N) int32 (before shift/clip). */
ORC_API int orc_density_count(const float *pts, int B, int N, float r2, int K, int32_t *cnt)
{
    for (int b = 0; b < B; ++b) {
        const float *P = pts + (size_t)b * N * 3;
        for (int i = 0; i < N; ++i) {
            float d0;
            {
                float dx = P[3 * i] - P[0], dy = P[3 * i + 1] - P[1], dz = P[3 * i + 2] - P[2];
                float qx = dx * dx, qy = dy * dy, qz = dz * dz;
                d0 = qx + qy; d0 = d0 + qz;
            }
            int c1 = 0, c2 = 0;
            for (int j = 0; j < N; ++j) {
                float dx = P[3 * i] - P[3 * j], dy = P[3 * i + 1] - P[3 * j + 1],
                      dz = P[3 * i + 2] - P[3 * j + 2];
                float qx = dx * dx, qy = dy * dy, qz = dz * dz;
                float d = qx + qy; d = d + qz;
                c1 += (d < r2);
                c2 += (d < d0);
            }
            int in0 = (d0 < r2) && (c2 < K);
            int c = c1 < K ? c1 : K;
            cnt[(size_t)b * N + i] = c - in0;
        }
    }
    return 0;
}

/* ---- a6, list form: pcl KdTreeFLANN.radius_search_for_cloud(cloud, r, K) (MLSP/mlsp.py:250) ----------
 * Same restatement as orc_density_count (PARITY UNPINNED): per point the neighbours with d < r2, sorted by
 * (d ascending, index ascending), the first K kept, rows zero-padded.  Plain insertion into a sorted list.
 * pts (B,N,3); ind (B,N,K) int32; sqd (B,N,K) float. */
ORC_API int orc_radius_search(const float *pts, int B, int N, float r2, int K, int32_t *ind, float *sqd)
{
    for (int b = 0; b < B; ++b) {
        const float *P = pts + (size_t)b * N * 3;
        for (int i = 0; i < N; ++i) {
            int32_t *I = ind + ((size_t)b * N + i) * K;
            float *D = sqd + ((size_t)b * N + i) * K;
            int n = 0;
            for (int j = 0; j < N; ++j) {
                float dx = P[3 * i] - P[3 * j], dy = P[3 * i + 1] - P[3 * j + 1],
                      dz = P[3 * i + 2] - P[3 * j + 2];
                float qx = dx * dx, qy = dy * dy, qz = dz * dz;
                float d = qx + qy; d = d + qz;
                if (!(d < r2)) continue;
                /* position: after every kept entry with distance <= d (equal distances keep index order) */
                int pos = n;
                while (pos > 0 && D[pos - 1] > d) --pos;
                if (pos >= K) continue;
                int last = n < K ? n : K - 1;
                for (int t = last; t > pos; --t) { D[t] = D[t - 1]; I[t] = I[t - 1]; }
                D[pos] = d; I[pos] = j;
                if (n < K) ++n;
            }
            for (int t = n; t < K; ++t) { D[t] = 0.0f; I[t] = 0; }
        }
    }
    return 0;
}

/* ---- a9/a10: masked Chamfer ------------------------------------------------------------------
 * Reference: chamfer_distance MLSP/mlsp.py:115-153, findneareat_index :196-220.
 *   D[i][j] = (||p1_i - p2_j||_2)^2  (sqrt then square, :138) + (mask_j==0 ? 100 : 0)
 *   rowmin_i = min_j D, argmin lowest j;  S = sum_i rowmin_i*mask_i / sum_i mask_i;  sum_b
 * p1,p2 (B,N,3); mask (B,N) of 0/1 floats (the reference's mask[:, :, 0]).
 * rowmin (B,N) float, argmin (B,N) int64 for ALL rows; returns sum_b S_b (double accumulate). */
ORC_API double orc_chamfer_dir(const float *p1, const float *p2, const float *mask, int B, int N,
                               float *rowmin, int64_t *argmin)
{
    double total = 0;
    for (int b = 0; b < B; ++b) {
        const float *A = p1 + (size_t)b * N * 3, *Q = p2 + (size_t)b * N * 3;
        const float *m = mask + (size_t)b * N;
        double S = 0, cnt = 0;
        for (int i = 0; i < N; ++i) {
            float best = INFINITY; int bj = 0;
            for (int j = 0; j < N; ++j) {
                float dx = A[3 * i] - Q[3 * j], dy = A[3 * i + 1] - Q[3 * j + 1],
                      dz = A[3 * i + 2] - Q[3 * j + 2];
                float qx = dx * dx, qy = dy * dy, qz = dz * dz;
                float s = qx + qy; s = s + qz;
                float n = sqrtf(s);
                float D = n * n;
                float pen = (m[j] == 0.0f) ? 100.0f : ((m[j] == 1.0f) ? 0.0f : m[j]);
                D = D + pen;
                if (D < best) { best = D; bj = j; }
            }
            rowmin[(size_t)b * N + i] = best;
            argmin[(size_t)b * N + i] = bj;
            S += (double)best * (double)m[i];
            cnt += m[i];
        }
        total += S / cnt; /* empty mask -> 0/0 = NaN like the reference */
    }
    return total;
}

/* reconstruction_loss forward + closed-form gradient w.r.t. pred (MLSP/mlsp.py:156-182).
 *   loss = (1/B) * (chamfer(gold,pred,mask) + chamfer(pred,gold,mask))
 * grad (B,N,3) double, for d loss / d pred with upstream gradient 1. */
ORC_API double orc_reconstruction_loss(const float *pred, const float *gold, const float *mask,
                                       int B, int N, double *grad_pred)
{
    float *rm = (float *)malloc(sizeof(float) * (size_t)B * N);
    int64_t *am1 = (int64_t *)malloc(sizeof(int64_t) * (size_t)B * N);
    int64_t *am2 = (int64_t *)malloc(sizeof(int64_t) * (size_t)B * N);
    double d1 = orc_chamfer_dir(gold, pred, mask, B, N, rm, am1);
    double d2 = orc_chamfer_dir(pred, gold, mask, B, N, rm, am2);
    if (grad_pred) {
        memset(grad_pred, 0, sizeof(double) * (size_t)B * N * 3);
        for (int b = 0; b < B; ++b) {
            double cnt = 0;
            for (int i = 0; i < N; ++i) cnt += mask[(size_t)b * N + i];
            for (int i = 0; i < N; ++i) {
                double w = mask[(size_t)b * N + i] / (cnt * B);
                if (w == 0) continue;
                int64_t j1 = am1[(size_t)b * N + i], j2 = am2[(size_t)b * N + i];
                for (int c = 0; c < 3; ++c) {
                    /* dir 1: rows gold_i, cols pred_j1 */
                    double g1 = 2.0 * ((double)pred[((size_t)b * N + j1) * 3 + c] -
                                       (double)gold[((size_t)b * N + i) * 3 + c]);
                    grad_pred[((size_t)b * N + j1) * 3 + c] += w * g1;
                    /* dir 2: rows pred_i, cols gold_j2 */
                    double g2 = 2.0 * ((double)pred[((size_t)b * N + i) * 3 + c] -
                                       (double)gold[((size_t)b * N + j2) * 3 + c]);
                    grad_pred[((size_t)b * N + i) * 3 + c] += w * g2;
                }
            }
        }
    }
    free(rm); free(am1); free(am2);
    return (d1 + d2) / B;
}
