// scan.cu -- scan_input / p_scan of MLSP/mlsp.py:54-94: the single-view "scan" simulation of the
// Scan_on_trgt branch (PointDA/trainer.py:492-503).  Per cloud the reference rotates the points (numpy, fp64), bins them on a
// (pixel+5)^2 grid by  int((z'+1)/2*pixel*pixel + (y'+1)/2*pixel)  and keeps per bin the point with the largest x' (the first
// one on ties) in a Python loop over points; the kept points get mask 0 and keep their ORIGINAL coordinates, all others are
// zeroed with mask 1.  Here: one CTA per cloud, the grid in shared memory, a z-buffer in two atomic passes (max of the
// order-preserving bits of x' per bin, then min index among the points that attain it), outputs written in place.
// The rotation matrices come from the host (the reference's numpy RNG stream is consumed there, mlsp_b200/ops.py).
// fp64 arithmetic in the reference's evaluation order, no contraction -- the bin index and the comparison are reproduced
// exactly up to the last-bit freedom of numpy's own BLAS dot (measure-zero effect: a value exactly on a bin boundary).
#include "common.cuh"

namespace mlsp {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_MAX_CELLS = 3072;      // pixel = int(2 / pixel_size) <= 44 for the reference's pixel sizes: (44+5)^2 = 2401

__device__ __forceinline__ unsigned long long f64_orderable(double v)
{
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return u ^ ((u >> 63) ? 0xffffffffffffffffull : 0x8000000000000000ull);
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_zbuffer_kernel(float *__restrict__ X, int N, const double *__restrict__ rot, int pixel, int cells, float *__restrict__ mask,
                    int *__restrict__ err)
{
    __shared__ unsigned long long best[SCAN_MAX_CELLS];
    __shared__ int who[SCAN_MAX_CELLS];
    __shared__ double R[9];
    const int b = blockIdx.x;
    float *xb = X + (size_t)b * N * 3;
    float *mb = mask + (size_t)b * N * 3;
    if (threadIdx.x < 9) R[threadIdx.x] = rot[(size_t)b * 9 + threadIdx.x];
    for (int c = threadIdx.x; c < cells; c += SCAN_THREADS) {
        best[c] = 0ull;                    // below the image of every double (f64_orderable(-inf) = 0x000f...)
        who[c] = 0x7fffffff;
    }
    __syncthreads();
    const double px = (double)pixel;
    // pass 1: bin index and rotated x of every point; the bin keeps the largest x
    for (int i = threadIdx.x; i < N; i += SCAN_THREADS) {
        const double p0 = (double)xb[3 * i], p1 = (double)xb[3 * i + 1], p2 = (double)xb[3 * i + 2];
        // np.dot(pc, R): row . column, k = 0, 1, 2 in order
        const double r0 = __dadd_rn(__dadd_rn(__dmul_rn(p0, R[0]), __dmul_rn(p1, R[3])), __dmul_rn(p2, R[6]));
        const double r1 = __dadd_rn(__dadd_rn(__dmul_rn(p0, R[1]), __dmul_rn(p1, R[4])), __dmul_rn(p2, R[7]));
        const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(p0, R[2]), __dmul_rn(p1, R[5])), __dmul_rn(p2, R[8]));
        const double comp = __dadd_rn(__dmul_rn(__dmul_rn(__ddiv_rn(__dadd_rn(r2, 1.0), 2.0), px), px),
                                      __dmul_rn(__ddiv_rn(__dadd_rn(r1, 1.0), 2.0), px));
        long long ci = (long long)comp;                       // astype(int): truncation toward zero
        if (ci < 0) ci += cells;                              // a Python list index: negative values wrap once
        if (ci < 0 || ci >= cells || !(comp == comp)) {
            atomicExch(err, 1);                               // the reference raises IndexError here
            continue;
        }
        atomicMax(&best[ci], f64_orderable(r0));
    }
    __syncthreads();
    // pass 2: among the points that attain their bin's maximum, the first one (the reference replaces only on strict >)
    for (int i = threadIdx.x; i < N; i += SCAN_THREADS) {
        const double p0 = (double)xb[3 * i], p1 = (double)xb[3 * i + 1], p2 = (double)xb[3 * i + 2];
        const double r0 = __dadd_rn(__dadd_rn(__dmul_rn(p0, R[0]), __dmul_rn(p1, R[3])), __dmul_rn(p2, R[6]));
        const double r1 = __dadd_rn(__dadd_rn(__dmul_rn(p0, R[1]), __dmul_rn(p1, R[4])), __dmul_rn(p2, R[7]));
        const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(p0, R[2]), __dmul_rn(p1, R[5])), __dmul_rn(p2, R[8]));
        const double comp = __dadd_rn(__dmul_rn(__dmul_rn(__ddiv_rn(__dadd_rn(r2, 1.0), 2.0), px), px),
                                      __dmul_rn(__ddiv_rn(__dadd_rn(r1, 1.0), 2.0), px));
        long long ci = (long long)comp;
        if (ci < 0) ci += cells;
        if (ci < 0 || ci >= cells || !(comp == comp)) continue;
        // -0.0 == +0.0 for the reference's `>`: compare values, not bit images
        const unsigned long long top = best[ci];
        const unsigned long long raw = top ^ ((top >> 63) ? 0x8000000000000000ull : 0xffffffffffffffffull);
        if (r0 == __longlong_as_double((long long)raw)) atomicMin(&who[ci], i);
    }
    __syncthreads();
    // pass 3: kept points stay (mask 0), everything else is zeroed (mask 1)
    for (int i = threadIdx.x; i < N; i += SCAN_THREADS) {
        const double p0 = (double)xb[3 * i], p1 = (double)xb[3 * i + 1], p2 = (double)xb[3 * i + 2];
        const double r1 = __dadd_rn(__dadd_rn(__dmul_rn(p0, R[1]), __dmul_rn(p1, R[4])), __dmul_rn(p2, R[7]));
        const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(p0, R[2]), __dmul_rn(p1, R[5])), __dmul_rn(p2, R[8]));
        const double comp = __dadd_rn(__dmul_rn(__dmul_rn(__ddiv_rn(__dadd_rn(r2, 1.0), 2.0), px), px),
                                      __dmul_rn(__ddiv_rn(__dadd_rn(r1, 1.0), 2.0), px));
        long long ci = (long long)comp;
        if (ci < 0) ci += cells;
        const bool ok = !(ci < 0 || ci >= cells || !(comp == comp));
        const bool keep = ok && who[ci] == i;
        const float m = keep ? 0.0f : 1.0f;
        mb[3 * i] = m;
        mb[3 * i + 1] = m;
        mb[3 * i + 2] = m;
        if (!keep) {
            xb[3 * i] = 0.0f;
            xb[3 * i + 1] = 0.0f;
            xb[3 * i + 2] = 0.0f;
        }
    }
}

}  // namespace mlsp

extern "C" int mlsp_scan_zbuffer(float *X, int B, int N, const double *rot, int pixel, float *mask, int *err_flag, void *stream)
{
    using namespace mlsp;
    MLSP_REQUIRE(X && rot && mask && err_flag, MLSP_EINVAL, "mlsp_scan_zbuffer: null pointer");
    MLSP_REQUIRE(B > 0 && N > 0 && pixel > 0, MLSP_EINVAL, "mlsp_scan_zbuffer: bad shape B=%d N=%d pixel=%d", B, N, pixel);
    const long long cells = (long long)(pixel + 5) * (pixel + 5);
    MLSP_REQUIRE(cells <= SCAN_MAX_CELLS, MLSP_EUNSUPPORTED, "mlsp_scan_zbuffer: (pixel+5)^2 = %lld bins exceed %d", cells, SCAN_MAX_CELLS);
    MLSP_CUDA(cudaMemsetAsync(err_flag, 0, sizeof(int), as_stream(stream)));
    scan_zbuffer_kernel<<<B, SCAN_THREADS, 0, as_stream(stream)>>>(X, N, rot, pixel, (int)cells, mask, err_flag);
    MLSP_LAUNCH_CHECK("scan_zbuffer_kernel");
    return MLSP_OK;
}
