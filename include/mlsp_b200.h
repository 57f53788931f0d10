/*
 * mlsp_b200.h -- C ABI of libmlsp_b200.so: the B200 (sm_100a) implementation of the MLSP
 * data-parallel hot path (DGCNN neighbourhood engine, masked-local-structure target builder,
 * masked Chamfer loss).
 *
 * The reference (VITA-Group/MLSP) is pure Python/PyTorch and has no FFI; its boundary is the set
 * of module-level Python functions named below.  Each entry point cites the
 * reference function it replaces (file:line in the reference checkout).  The Python host side
 * (mlsp_b200/ops.py) keeps the reference signatures and calls these through ctypes; see
 * INTEGRATION.md for the binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless marked "host";
 *   - tensors are dense, row-major, in the layouts written next to each argument; edge tensors, their gradients
 *     and workspaces are 16-byte aligned (128-bit accesses);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued, nothing synchronises;
 *   - outputs and workspaces are caller-allocated (mlsp_workspace_bytes gives the size);
 *   - return value: 0 on success, otherwise an MLSP_E* code; mlsp_last_error() (thread-local)
 *     describes it.  There is no CPU fallback: without a CUDA device every call fails.
 *   - indices are int64 at the boundary (the reference returns torch.long);
 *   - inputs must be finite for the index-producing ops (the reference's NaN ranking is not reproduced); the Chamfer
 *     entry points propagate non-finite points to a NaN loss like torch.min does and never address outside the cloud;
 *   - caller-supplied neighbour indices (idx arguments) must lie in [0, N): they are used as raw offsets, the host
 *     layer (mlsp_b200/ops.py) range-checks them when MLSP_B200_CHECK_IDX=1 is set in the environment.
 */
#ifndef MLSP_B200_H
#define MLSP_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define MLSP_API __attribute__((visibility("default")))
#else
#define MLSP_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define MLSP_OK 0
#define MLSP_EINVAL 1      /* bad shape / null pointer / k out of range            */
#define MLSP_EUNSUPPORTED 2 /* shape outside what the kernels are built for        */
#define MLSP_ECUDA 3       /* a CUDA runtime call or launch failed                  */
#define MLSP_EWORKSPACE 4  /* workspace too small                                   */

/* operation ids for mlsp_workspace_bytes */
#define MLSP_OP_KNN 1
#define MLSP_OP_EDGE_FWD 2
#define MLSP_OP_EDGE_BWD 3
#define MLSP_OP_CHAMFER 4
#define MLSP_OP_GRAPH_FEATURE 5
#define MLSP_OP_EDGECONV_BWD 6 /* C = output channels O */

/* flags for mlsp_knn_f32 */
#define MLSP_KNN_AUTO 0        /* C=3: two-pass 3-D kernel; C in {64,128}: tcgen05 filter + exact re-rank; else streaming */
#define MLSP_KNN_EXACT_ONLY 1  /* force the generic fp32 streaming-selection kernel (cross-check path)  */
#define MLSP_KNN_TENSOR_ONLY 2 /* force the tcgen05 path (error if the shape does not allow it)  */
#define MLSP_KNN_STATS 4       /* OR-ed in: the tcgen05 path maintains four int32 diagnostic counters at offset 0 of ws --
                                  {rows re-done exactly, rows certified, total candidate-list length, exact distances recomputed};
                                  off by default: one same-address atomic per row costs ~17 us per counter at 32 k rows */

MLSP_API int mlsp_version(void);
MLSP_API const char *mlsp_last_error(void);
MLSP_API size_t mlsp_workspace_bytes(int op, int B, int C, int N, int k);

/* a1 -- knn(x, k): PointDA/model_utils.py:9-16 == PointSegDA/Models.py:8-15.
 *   pd[i][j] = ((-|x_j|^2) - (-2 x_i.x_j)) - |x_i|^2 in fp32; the k largest per row, best first,
 *   ties broken by lowest index.  x (B,C,N); idx (B,N,k) int64. 1 <= k <= min(N,64). */
MLSP_API int mlsp_knn_f32(const float *x, int B, int C, int N, int k, int64_t *idx, void *ws, size_t ws_bytes,
                 int flags, void *stream);

/* Test hook for the tcgen05 path of a1 (C in {64,128}, N >= 256): same result as mlsp_knn_f32, and when
 * `dump` is non-NULL the approximate filter values |x_j|^2 - 2 dot~(x_i,x_j) are written to dump (2,B,N,N):
 * [0] = pass 1 (bf16 heads only, sets the threshold), [1] = pass 2 (three-term split, the listed values).
 * After the call ws holds two int32 counters at offset 0: {rows re-done by the exact streaming selection, rows certified}. */
MLSP_API int mlsp_knn_tensor_debug(const float *x, int B, int C, int N, int k, int64_t *idx, void *ws, size_t ws_bytes,
                                   float *dump, void *stream);

/* Profiling hook for the tcgen05 path of a1: same result as mlsp_knn_f32; tstamp (B * ceil(N/128), 16) int64 receives, per
 * CTA of the filter kernel, %globaltimer (ns) marks of its phases ([0] entry, [1] prologue done, [2] first accumulator
 * ready, [3] pass 1 done, [4] class maxima sorted, [5] exchanged, [6] threshold ready, [7] pass 2 done, [8] lists complete,
 * [9] lists copied out) and the SM id in [15].  cluster: CTAs per cluster sharing the candidate blocks by TMA multicast (1, 2, 4;
 * 0 = the library's default).  tools/kt_timeline.py prints the phase breakdown. */
MLSP_API int mlsp_knn_tensor_timeline(const float *x, int B, int C, int N, int k, int64_t *idx, void *ws, size_t ws_bytes,
                                      long long *tstamp, int cluster, void *stream);

/* a2 -- get_graph_feature(x, args, k, idx): PointDA/model_utils.py:18-42 == PointSegDA/Models.py:18-45.
 *   out logical shape (B,2C,N,k) stored channels-last, i.e. memory order [B][N][k][2C]:
 *   out[b][i][j][c] = x[b][c][idx[b][i][j]] - x[b][c][i] (c < C),  x[b][c-C][i] (c >= C). */
MLSP_API int mlsp_edge_gather_fwd(const float *x, const int64_t *idx, int B, int C, int N, int k, float *out,
                         void *ws, size_t ws_bytes, void *stream);

/* a1 + a2 in one call -- get_graph_feature(x, args, k) with idx=None (the way DGCNN calls it: PointDA/Models.py:111-127,
 *   PointSegDA/Models.py:172-182,219): idx (B,N,k) int64 is written (kept for the backward) and out as in
 *   mlsp_edge_gather_fwd.  On the tcgen05 path the gather reads the point-major copy the kNN already made.
 *   Workspace: mlsp_workspace_bytes(MLSP_OP_GRAPH_FEATURE, ...). */
MLSP_API int mlsp_graph_feature_fwd(const float *x, int B, int C, int N, int k, int64_t *idx, float *out, void *ws,
                           size_t ws_bytes, void *stream);

/* Measurement hook for the call above on the tcgen05 path (C in {64,128}): `stages` selects which of its kernels are
 *   launched -- bit 0: knn_prep, bit 1: knn_tensor (filter), bit 2: knn_refine (ranking + fused edge gather).  Every
 *   kernel only reads what the earlier ones left in ws, so after one full call with the same arguments and the same ws
 *   a single stage can be re-run alone (bench.py times each kernel this way with CUDA events). */
MLSP_API int mlsp_graph_feature_fwd_stage(const float *x, int B, int C, int N, int k, int64_t *idx, float *out, void *ws,
                                 size_t ws_bytes, int stages, void *stream);

/* backward of a2 w.r.t. x (the reference gets it from autograd: index_put_(accumulate=True)).
 *   grad_out in the same channels-last storage [B][N][k][2C]; grad_x (B,C,N), overwritten. */
MLSP_API int mlsp_edge_gather_bwd(const float *grad_out, const int64_t *idx, int B, int C, int N, int k,
                         float *grad_x, void *ws, size_t ws_bytes, void *stream);

/* a3 -- farthest_point_sample(args, xyz, npoint): utils/pc_utils.py:137-161.
 *   xyz (B,3,N); start (B) int64 = the torch.randint draw of :150 (made by the host so the CPU RNG
 *   stream stays the reference's); centroids (B,npoint) int64; vals (B,3,npoint). */
MLSP_API int mlsp_fps(const float *xyz, int B, int N, int npoint, const int64_t *start, int64_t *centroids,
             float *vals, void *stream);
/* Measurement hook (process-wide): 1 (default) = the centre / prep / filter / ranking kernels of the tcgen05 kNN path are chained
 * with programmatic dependent launch (each starts its independent prologue while the predecessor drains), 0 = plain stream order.
 * Results do not depend on it. */
MLSP_API void mlsp_knn_set_pdl(int on);
/* Tuning hook (process-wide, not thread-safe): clouds per CTA of the FPS kernels -- 0 = automatic (1: packing measured slower on B200,
 * see fps.cu), 1/2/4 = forced where 1024 threads and the shared memory allow.  Results do not depend on it. */
MLSP_API void mlsp_fps_set_groups(int groups);
/* Tuning hook (process-wide): -1 = automatic (N > 1024), 0 = never, 1 = every FPS CTA reserves the whole shared memory of its SM so that kernels with a shared-memory
 * footprint do not share the SM with the latency-bound sampling loop.  Results do not depend on it. */
MLSP_API void mlsp_fps_set_exclusive(int on);

/* SURVEY.md 8f rank 3 -- PCM.mix_shapes (MLSP/PCM.py:6-38) fused with its two farthest_point_sample calls: 2B CTAs sample
 *   npoint_a points of cloud b and N - npoint_a points of cloud index[b] side by side and write the mixed, point-permuted
 *   cloud directly: out[b][:, inv_perm[s]] = slot s of cat(vals_a, vals_b), i.e. cat(...)[:, :, points_perm] of PCM.py:31-33.
 *   xyz (B,3,N); index (B) int64 = torch.randperm(B); start (2B) int64 = the two torch.randint draws (a-half first);
 *   inv_perm (N) int32 = inverse of torch.randperm(N); out (B,3,N).  N <= 8192. */
MLSP_API int mlsp_pcm_mix(const float *xyz, int B, int N, int npoint_a, const int64_t *index, const int64_t *start,
                          const int32_t *inv_perm, float *out, void *stream);

/* a4 -- assign_region_to_point (utils/pc_utils.py:33-73) + the per-region counts and the first-fit
 *   region choice of deform_input (MLSP/mlsp.py:28-50, groups=1).
 *   X (B,C,N), C >= 3, addressed X[b*xs_b + c*xs_c + n*xs_n] (strides in floats; all three 0 = dense): the reference's
 *   trainers pass the permuted view data.permute(0,2,1) of a (B,N,3) batch (PointDA/trainer.py:380-387), taken as is;
 *   order = host pointer to the 27 region ids of np.random.permutation(27);
 *   region (B,N) int64; counts (B,27) int32; chosen (B) int32 (-1: no region reaches min_pts);
 *   nsel (B) int32 = points in the chosen region. */
MLSP_API int mlsp_region_assign_select(const float *X, int64_t xs_b, int64_t xs_c, int64_t xs_n, int B, int C, int N,
                              const int32_t *order_host, int min_pts,
                              int64_t *region, int32_t *counts, int32_t *chosen, int32_t *nsel,
                              void *stream);

/* a4/a8 -- the in-place deformation + mask of deform_input (MLSP/mlsp.py:44-48).
 *   The r-th point (ascending index) of cloud b's chosen region receives noise[(offset[b]+r)*3 + c]
 *   (the host's np.random.multivariate_normal draws); X strided as above; mask (B,C,N) dense is fully written:
 *   1 on channels 0..2 of selected points, 0 elsewhere.  noise may be NULL (mask only). */
MLSP_API int mlsp_region_mask_scatter(float *X, int64_t xs_b, int64_t xs_c, int64_t xs_n, int B, int C, int N,
                             const int64_t *region, const int32_t *chosen,
                             const float *noise, const int32_t *offset, float *mask, void *stream);

/* a5 -- collapse_to_point (utils/pc_utils.py:76-111), part 1: per point, the number of points with
 *   pd = (|x_j|^2 - 2 x_i.x_j) + |x_i|^2 <= r2 (`pts_pass` before thresholding).  x (B,C>=3,N), strided like X above. */
MLSP_API int mlsp_ball_count(const float *x, int64_t xs_b, int64_t xs_c, int64_t xs_n, int B, int C, int N, float r2,
                    int32_t *cnt, void *stream);

/* a5 part 2: collapse the ball of `centre[b]` (host-chosen by np.random.choice, pc_utils.py:102):
 *   in-ball points (ascending index) receive the noise rows, mask as in mlsp_region_mask_scatter.
 *   centre[b] < 0 leaves the cloud untouched. */
MLSP_API int mlsp_ball_mask_scatter(float *X, int64_t xs_b, int64_t xs_c, int64_t xs_n, int B, int C, int N, float r2,
                           const int32_t *centre, const float *noise,
                           const int32_t *offset, float *mask, void *stream);

/* a6 -- cal_density (MLSP/mlsp.py:240-272): per-point ball cardinality with the python-pcl radius-search
 *   semantics (strict < r2, at most K nearest, neighbour index 0 dropped) and the soft class labels.
 *   pts (B,N,3); labels (B,N,num_cls) float32; row (B,N) int64 = clip(count-shift, 0, (num_cls-1)*pergroup). */
MLSP_API int mlsp_ball_count_labels(const float *pts, int B, int N, float r2, int K, int shift, int pergroup,
                           int num_cls, float *labels, int64_t *row, void *stream);

/* a6, list form -- pcl KdTreeFLANN.radius_search_for_cloud(cloud, radius, K) as called at MLSP/mlsp.py:250 (python-pcl;
 *   the `pcl` shim of mlsp_b200/pcl_shim.py binds it): per point the neighbours with squared distance < r2, nearest first
 *   (ties by lowest index), at most K (<= 128) of them; unused slots are 0 in both outputs, so `(ind != 0).sum(1)` is the
 *   reference's count.  pts (B,N,3); ind (B,N,K) int32; sqdist (B,N,K) float32. */
MLSP_API int mlsp_radius_search(const float *pts, int B, int N, float r2, int K, int32_t *ind, float *sqdist,
                                void *stream);

/* a7 -- kSearchNormalEstimation (PointDA/trainer.py:173-188, PointSegDA/trainer.py:73-88; python-pcl):
 *   unit eigenvector of the smallest eigenvalue of the covariance of the k nearest neighbours (idx from
 *   mlsp_knn_f32 on the same cloud, self included), oriented towards the origin (n.p <= 0).
 *   pts (B,N,3); idx (B,N,k) int64; normals (B,N,3); curvature (B,N) or NULL: pcl's 4th output column,
 *   lambda_min / (lambda_0 + lambda_1 + lambda_2). */
MLSP_API int mlsp_pca_normals(const float *pts, const int64_t *idx, int B, int N, int k, float *normals,
                              float *curvature, void *stream);

/* SURVEY.md 8f rank 2 -- one 3-D neighbourhood pass for the target builder: what PointDA/trainer.py:524-536 computes on the
 *   undeformed target batch -- kSearchNormalEstimation per cloud (a7) and mlsp.cal_density (a6) -- in ONE launch: the kernel
 *   that ranks the `near` nearest neighbours of every point (a1, 3-D) counts the ball cardinality in the same pass over the
 *   staged cloud and solves the 3x3 PCA from the ranked neighbourhood while it is in registers.  Same arithmetic and
 *   results as mlsp_knn_f32 + mlsp_pca_normals + mlsp_ball_count_labels.
 *   pts (B,N,3); normals (B,N,3); curvature (B,N) or NULL; labels (B,N,num_cls) + row (B,N) or both NULL (normals only);
 *   idx (B,N,near) int64 or NULL (the neighbourhoods, e.g. for a later mlsp_edge_gather_fwd).  N <= 8192, near <= 64. */
MLSP_API int mlsp_target_structure(const float *pts, int B, int N, int near, float r2, int K, int shift, int pergroup,
                                   int num_cls, float *normals, float *curvature, float *labels, int64_t *row,
                                   int64_t *idx, void *stream);

/* a9/a10 -- chamfer_distance(p1,p2,mask) MLSP/mlsp.py:115-153 and findneareat_index :196-220, one direction.
 *   D[i][j] = (|p1_i - p2_j|_2)^2 + (mask_j == 0 ? 100 : 0); rowmin/argmin over j (lowest j on ties).
 *   Points are addressed p[b*bstride + i*pstride + c*cstride] so both (B,N,3) and (B,3,N) tensors are
 *   accepted without a copy; mask[b*mask_bstride + i] is the 0/1 row mask (the reference's mask[:,:,0]).
 *   all_rows = 0: only rows with mask != 0 are evaluated (the others do not reach the loss; rowmin = 0,
 *   argmin = -1 there); all_rows = 1: every row (findneareat_index).
 *   partial (B) float: sum_i rowmin_i*mask_i / sum_i mask_i   (NaN for an empty mask, like the reference). */
MLSP_API int mlsp_chamfer_dir_fwd(const float *p1, int64_t p1_bstride, int64_t p1_pstride, int64_t p1_cstride,
                         const float *p2, int64_t p2_bstride, int64_t p2_pstride, int64_t p2_cstride,
                         const float *mask, int64_t mask_bstride, int B, int N, int all_rows, float *rowmin,
                         int64_t *argmin, float *partial, void *ws, size_t ws_bytes, void *stream);

/* backward of one direction: for masked rows i with j* = argmin[i],
 *   g = 2 (p1_i - p2_j*) * scale * grad_partial_sum / count_b ;  grad_p1[i] += g ; grad_p2[j*] -= g.
 *   grad_p1 / grad_p2 are (B,N,3) contiguous float32 accumulators (either may be NULL);
 *   scale_dev: device scalar (upstream gradient), scale_host multiplies it (e.g. 1/B). */
MLSP_API int mlsp_chamfer_dir_bwd(const float *p1, int64_t p1_bstride, int64_t p1_pstride, int64_t p1_cstride,
                         const float *p2, int64_t p2_bstride, int64_t p2_pstride, int64_t p2_cstride,
                         const float *mask, int64_t mask_bstride, const int64_t *argmin, int B, int N,
                         const float *scale_dev, float scale_host, float *grad_p1, float *grad_p2,
                         void *stream);

/* a9 -- reconstruction_loss(pred, gold, mask) MLSP/mlsp.py:156-182 in one call: both Chamfer directions
 *   (rows of gold against pred, rows of pred against gold), the per-cloud normalisation by the mask count and the
 *   1/B batch mean.  pred/gold are addressed like the points above (any (B,N,3)/(B,3,N) strides), mask like above.
 *   argmin (2,B,N) int64: matches of direction 0 (gold rows -> pred) then direction 1 (pred rows -> gold), -1 on
 *   unmasked rows; loss: one device float.  Workspace: mlsp_workspace_bytes(MLSP_OP_CHAMFER, ...). */
MLSP_API int mlsp_reconstruction_loss_fwd(const float *pred, int64_t pred_bstride, int64_t pred_pstride,
                         int64_t pred_cstride, const float *gold, int64_t gold_bstride, int64_t gold_pstride,
                         int64_t gold_cstride, const float *mask, int64_t mask_bstride, int B, int N,
                         int64_t *argmin, float *loss, void *ws, size_t ws_bytes, void *stream);

/* backward of the above w.r.t. pred: grad_pred (B,N,3) contiguous float32 is OVERWRITTEN with
 *   grad_loss/B * sum over both directions of +-2 (pred - match) / count_b   (closed form on the saved argmin).
 *   grad_loss: device scalar (upstream gradient) or NULL for 1. */
MLSP_API int mlsp_reconstruction_loss_bwd(const float *pred, int64_t pred_bstride, int64_t pred_pstride,
                         int64_t pred_cstride, const float *gold, int64_t gold_bstride, int64_t gold_pstride,
                         int64_t gold_cstride, const float *mask, int64_t mask_bstride, const int64_t *argmin,
                         int B, int N, const float *grad_loss, float *grad_pred, void *stream);

/* ---- EdgeConv without the edge tensor ------------------------------------------------------
 * Replaces the layer  get_graph_feature -> conv_2d (1x1 Conv2d [+ BatchNorm2d] + LeakyReLU) -> max over k
 * (PointDA/Models.py:114-128 with PointDA/model_utils.py:45-63; PointSegDA/Models.py:171-184, stacked convs without
 * BatchNorm / activation).  The 1x1 convolution is linear in [x_j - x_i | x_i]:
 *     h[b,o,i,j] = Y[b,idx[b,i,j],o] + Z[b,i,o],  Y = Wa x, Z = (Wb - Wa) x + bias,  W = [Wa | Wb]
 * so the caller makes ONE point-wise GEMM yz (B,N,2O) = [Y | Z] (a library call) and these entry points do the rest.
 * BatchNorm is a*h + c per channel with a = gamma*invstd, LeakyReLU is increasing:
 *     out[b,o,i] = lrelu(a_o * max_j h[b,o,i,j] + c_o)  for a_o >= 0 (negative scales: see mlsp_edgeconv_reduce_fwd).
 * All tensors point-major, 16-byte aligned; O % 4 == 0, O <= 1024, k <= 255. */

/* per point: hsel (B,N,O) = max_j h over the k neighbours, slot (B,N,O) uint8 = the neighbour rank that attains it
 *   (first on ties).  A channel whose scale a_o is negative needs the minimum instead: the caller folds the sign into
 *   that channel's rows of the weight (h' = -h, a' = -a; mlsp_b200/edgeconv.py does).  With stats != NULL (training-mode
 *   BatchNorm) also rowsum (B,N,O) = sum_j h and stats (2,O) double = [sum h ; sum h^2] over all B*N*k edges
 *   (zeroed by the call, accumulated with fp64 atomics). */
MLSP_API int mlsp_edgeconv_reduce_fwd(const float *yz, const int64_t *idx, int B, int N, int O, int k, float *hsel, uint8_t *slot,
                             float *rowsum, double *stats, void *stream);

/* BatchNorm2d in training mode (torch.nn.functional.batch_norm: biased variance): stats of `count` = B*N*k edges ->
 *   coef (4,O) = [a = gamma*invstd ; c = beta - a*mean ; mean ; invstd].  gamma / beta NULL = 1 / 0 (affine=False).
 *   running_mean / running_var (O, both or neither): updated in place like torch.nn.BatchNorm2d does,
 *   r = (1-momentum)*r + momentum*batch value, the variance unbiased by count/(count-1); `sign` (O, +-1, may be NULL)
 *   multiplies the mean first -- for callers that folded the sign of gamma into the rows of the weight. */
MLSP_API int mlsp_edgeconv_bn_coeffs(const double *stats, const float *gamma, const float *beta, int O, double count, float eps,
                            float *coef, float *running_mean, float *running_var, const float *sign, float momentum,
                            void *stream);

/* out (B,O,N) = lrelu_slope(a_o * hsel[b,i,o] + c_o), coef rows 0 and 1 = a, c  (slope 1: no activation, 0: ReLU) */
MLSP_API int mlsp_edgeconv_apply_fwd(const float *hsel, const float *coef, int B, int N, int O, float slope, float *out,
                            void *stream);

/* backward: g (B,O,N) = d loss / d out, rows dense, batches g_bstride floats apart (O*N when contiguous; a channel slice
 *   of a concatenated gradient, as torch.cat's backward hands it over, is used in place)  ->  dyz (B,N,2O) = [dY | dZ].
 *   bn_train != 0: exact gradient through the batch statistics (needs rowsum and coef rows 2,3 = mean, invstd);
 *   dgamma_dbeta (2,O), may be NULL when bn_train == 0: [sum dy*(hsel-mean)*invstd ; sum dy], dy = g*lrelu'(.) -- the
 *   gradients of gamma and beta, or of (a, c) when the caller sets mean = 0, invstd = 1 (fixed affine / bias).
 *   ws: mlsp_workspace_bytes(MLSP_OP_EDGECONV_BWD, B, O, N, k). */
MLSP_API int mlsp_edgeconv_bwd(const float *g, int64_t g_bstride, const float *yz, const int64_t *idx, const float *hsel, const uint8_t *slot,
                      const float *rowsum, const float *coef, int B, int N, int O, int k, float slope, int bn_train,
                      float *dyz, float *dgamma_dbeta, void *ws, size_t ws_bytes, void *stream);

/* weight side of the layer product (one launch instead of a dozen elementwise kernels):
 *   W (O,2C) = [Wa | Wb] over [x_j - x_i | x_i]; scale (O) or NULL (BatchNorm weight / fixed affine scale); bias (O) or NULL
 *   -> Wcat (2O,C) = [s Wa ; s (Wb - Wa)], sgn (O) = s = -1 where scale < 0 else +1, zb (2O, may be NULL when bias is NULL) =
 *   [0 ; s bias]: yz = x^T Wcat^T + zb is mlsp_gemm_f32's job (conv_2d's nn.Conv2d, PointDA/model_utils.py:45-63). */
MLSP_API int mlsp_edgeconv_weight_prep(const float *W, const float *scale, const float *bias, int O, int C, float *Wcat, float *sgn,
                              float *zb, void *stream);

/* part (Z,2O,C) = partial products dyz^T x  ->  gW (O,2C) = [s (gY - gZ) | s gZ], summed over Z in a fixed order */
MLSP_API int mlsp_edgeconv_weight_grad(const float *part, int Z, const float *sgn, int O, int C, float *gW, void *stream);

/* ---- max pooling around the point-wise layers, with argmax (first index on ties, NaN propagates) and backward ----
 * "mid": in (R,K,C), C contiguous and C % 4 == 0 -> val (R,C), arg (R,C) int32 = max over K.  x.max(dim=-1) over the k
 *   neighbours of a channels-last edge tensor (R = B*N, K = k) and torch.max(x, dim=2) over the points of a channels-last map
 *   (R = B, K = N): transform_net, PointDA/model_utils.py:116-121.  bwd: g (R,C), arg -> gin (R,K,C) (zeros off the argmax).
 * "row": in (R,K), K contiguous -> val (R), arg (R): F.adaptive_max_pool1d(x, 1) of a (B,C,N) map, PointDA/Models.py:133. */
MLSP_API int mlsp_max_mid_fwd(const float *in, long long R, int K, int C, float *val, int *arg, void *stream);
MLSP_API int mlsp_max_mid_bwd(const float *g, const int *arg, long long R, int K, int C, float *gin, void *stream);
MLSP_API int mlsp_max_row_fwd(const float *in, long long R, int K, float *val, int *arg, void *stream);
MLSP_API int mlsp_max_row_bwd(const float *g, const int *arg, long long R, int K, float *gin, void *stream);

/* ---- the point-wise products around the neighbourhood engine (8f ranks 1, 4): every 1x1 convolution / Linear of the DGCNN ----
 * Replaces the library GEMM behind nn.Conv2d(kernel_size=1) in conv_2d (PointDA/model_utils.py:45-63, used by the EdgeConv
 * layers PointDA/Models.py:114-128 and by transform_net model_utils.py:92-130), nn.Conv1d(kernel_size=1) of conv5 and of the
 * heads (PointDA/Models.py:131, 156-160, 165-285) and nn.Linear (fc_layer model_utils.py:65-89), forward and both backward
 * products.
 *
 *     D[z] (M x N) = A[z] (M x K) . B[z]^T (N x K)  (+ bias[n]),     z = 0 .. batch-1,   fp32 in / fp32 out
 *
 * a_kmajor != 0: A[z][m*lda + k] (K contiguous), else A[z][k*lda + m] (M contiguous); b_kmajor alike with ldb over (n, k);
 * d_rowmajor != 0: D[z][m*ldd + n], else D[z][n*ldd + m].  Batch strides in floats (0 = operand shared by all z).
 * bias (N) or NULL.  Any M, N, K >= 1, any leading dimensions (16-byte aligned operands take the vector loads).
 * Arithmetic: tcgen05 tensor cores on three bf16 pieces per fp32 operand, six piece products accumulated in fp32 -- an fp32
 * GEMM up to summation order (error ~1e-7 relative to sum |a||b|; NOT a bf16/tf32 approximation).  Non-finite inputs
 * give NaN; |values| must stay below 3.3e38 (bf16 rounding of the leading piece). */
MLSP_API int mlsp_gemm_f32(const float *A, int a_kmajor, long long lda, long long a_batch_stride, const float *B, int b_kmajor,
                  long long ldb, long long b_batch_stride, float *D, int d_rowmajor, long long ldd, long long d_batch_stride,
                  const float *bias, int M, int N, int K, int batch, void *stream);

/* ---- scan_input / p_scan, MLSP/mlsp.py:54-94 (the Scan_on_trgt branch, PointDA/trainer.py:492-503; 8f rank 3) ----
 * X (B,N,3) contiguous, mutated in place; rot (B,3,3) double: the rotation matrix of every cloud (rotate_point_cloud_3d,
 * mlsp.py:96-112 -- drawn by the caller from numpy's RNG like the reference); pixel = int(2 / pixel_size).
 * Per cloud: p' = p . rot (fp64), bin = int((p'_z+1)/2*pixel*pixel + (p'_y+1)/2*pixel) on a (pixel+5)^2 grid (negative values
 * wrap once, like the Python list index they are), the point with the largest p'_x of every bin (first index on ties) is
 * kept: mask (B,N,3) = 0 there and X unchanged; every other point: mask = 1, X = 0.
 * err_flag (device int): set to 1 if a bin index falls outside the grid (the reference raises IndexError).
 * (pixel+5)^2 <= 3072 (the reference draws pixel_size in [0.045, 0.075]: pixel <= 44). */
MLSP_API int mlsp_scan_zbuffer(float *X, int B, int N, const double *rot, int pixel, float *mask, int *err_flag, void *stream);

/* measurement hook: mlsp_gemm_f32 with CTA 0 recording SM-clock stamps of its first 64 K-chunk iterations into tstamp (64 x 8 int64,
 * device memory): [0] loaders see the stage free, [1] pieces stored, [2] arrived, [3] MMA warp sees the stage full, [4] MMAs issued */
MLSP_API int mlsp_gemm_f32_timeline(const float *A, int a_kmajor, long long lda, long long a_batch_stride, const float *B, int b_kmajor,
                           long long ldb, long long b_batch_stride, float *D, int d_rowmajor, long long ldd, long long d_batch_stride,
                           const float *bias, int M, int N, int K, int batch, long long *tstamp, void *stream);

/* ---- BatchNorm (training mode) fused with the activation after it: conv_2d / fc_layer of PointDA/model_utils.py:45-89
 * (Conv -> BatchNorm -> LeakyReLU(0.2)), the heads of PointDA/Models.py:165-285 and PointSegDA/Models.py:245-392
 * (Conv1d -> BatchNorm1d -> ReLU) ----
 * y = leaky_relu(batch_norm(x), slope): slope 0 = ReLU, slope 1 = no activation.  layout 0: x (R, C) row-major (channels-last
 * 4-D maps, 2-D fc inputs; C % 4 == 0, C <= 1024, L ignored); layout 1: x (R = B, C, L) with batch stride x_batch_stride for x and
 * y_batch_stride for y / dy / dx (floats; 0 = C * L -- a channel slice of a wider map is a valid x).  Statistics over all but
 * the channel dimension, biased variance for the normalisation, running_var updated with the unbiased one (torch semantics);
 * gamma / beta / running_* may be NULL.  save_mean / save_invstd (C) feed the backward; acc: mlsp_bn_scratch_bytes(C) bytes of
 * scratch, 8-byte aligned, contents irrelevant (per-CTA partial sums: no atomics, deterministic). */
MLSP_API size_t mlsp_bn_scratch_bytes(int C);
MLSP_API int mlsp_bn_act_fwd(const float *x, float *y, long long R, int C, int L, int layout, long long x_batch_stride,
                    long long y_batch_stride, const float *gamma, const float *beta,
                    float *running_mean, float *running_var, float momentum, float eps, float slope, float *save_mean,
                    float *save_invstd, double *acc, void *stream);
/* dx (and dgamma, dbeta (C), NULL to skip) from the layer's INPUT x, dy and the saved statistics (the activation's derivative is
 * recomputed from x, the output is not needed). */
MLSP_API int mlsp_bn_act_bwd(const float *x, const float *dy, float *dx, long long R, int C, int L, int layout,
                    long long x_batch_stride, long long y_batch_stride, const float *gamma,
                    const float *beta, const float *save_mean, const float *save_invstd, float slope, float *dgamma,
                    float *dbeta, double *acc, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MLSP_B200_H */
