// Micro-benchmarks that size the epilogue of knn_tensor_kernel: how fast can epilogue warps
// pull accumulators out of TMEM (tcgen05.ld 32x32b), and how fast do the ALU instructions of the selection issue?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/tmem_bench.bin tools/ubench/tmem_bench.cu
// Run on a B200: tools/ubench/tmem_bench.bin  (prints one line per configuration)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X> struct Ld;
template <> struct Ld<16> {
    static __device__ __forceinline__ void issue(uint32_t taddr, uint32_t (&u)[16]) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
              "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]) : "r"(taddr));
    }
};
template <> struct Ld<32> {
    static __device__ __forceinline__ void issue(uint32_t taddr, uint32_t (&u)[32]) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
              "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
              "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
              "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31]) : "r"(taddr));
    }
};
template <int X>
__device__ __forceinline__ void wait_ld(uint32_t (&u)[X]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < X; ++i) asm volatile("" : "+r"(u[i]));
}

// Every warp reads `span` columns of its lane quadrant per tile, in pieces of X columns; DEPTH pieces in flight.
// work: 0 = one XOR per piece; 1 = one FMNMX per element (pass-1-like); 2 = FMNMX3 on pairs
template <int X, int WORK>
__global__ void __launch_bounds__(512) tmem_ld_kernel(int iters, int span, long long *cycles, float *sink)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    const int q = warp & 3, grp = warp >> 2;
    const uint32_t t0 = base + ((uint32_t)(q * 32) << 16) + (uint32_t)((grp * span) & 127);
    float acc[X];
#pragma unroll
    for (int i = 0; i < X; ++i) acc[i] = -1e30f;
    uint32_t x = 0;
    __syncthreads();
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t tb = t0 + (uint32_t)((it & 1) * 128);
        uint32_t ua[X], ub[X];
        Ld<X>::issue(tb, ua);
        for (int p = 0; p < span; p += 2 * X) {
            wait_ld<X>(ua);
            if (p + X < span) Ld<X>::issue(tb + p + X, ub);
            if (WORK == 0) x ^= ua[0];
            if (WORK == 1) {
#pragma unroll
                for (int i = 0; i < X; ++i) acc[i] = fmaxf(acc[i], __uint_as_float(ua[i]));
            }
            if (WORK == 2) {
#pragma unroll
                for (int i = 0; i < X / 2; ++i) acc[i] = fmaxf(fmaxf(acc[i], __uint_as_float(ua[i])), __uint_as_float(ua[i + X / 2]));
            }
            if (p + X < span) {
                wait_ld<X>(ub);
                if (p + 2 * X < span) Ld<X>::issue(tb + p + 2 * X, ua);
                if (WORK == 0) x ^= ub[0];
                if (WORK == 1) {
#pragma unroll
                    for (int i = 0; i < X; ++i) acc[i] = fmaxf(acc[i], __uint_as_float(ub[i]));
                }
                if (WORK == 2) {
#pragma unroll
                    for (int i = 0; i < X / 2; ++i) acc[i] = fmaxf(fmaxf(acc[i], __uint_as_float(ub[i])), __uint_as_float(ub[i + X / 2]));
                }
            }
        }
    }
    __syncthreads();
    const long long c1 = clock64();
    float s = __uint_as_float(x);
#pragma unroll
    for (int i = 0; i < X; ++i) s += acc[i];
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = c1 - c0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(256));
}

// ALU issue rates: NCH independent chains per thread.  OP 0: FMNMX, 1: FMNMX3, 2: FSETP + predicated IADD,
// 3: FFMA + FMNMX (the current pass 1), 4: FSETP + predicated LOP3.OR (bit-mask building)
template <int OP>
__global__ void __launch_bounds__(512) alu_kernel(int iters, const float *in, long long *cycles, float *sink)
{
    constexpr int NCH = 16;
    float a[NCH], v[NCH];
    int cnt[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { a[i] = in[i]; v[i] = in[NCH + i] + threadIdx.x; cnt[i] = 0; }
    const float thr = in[40], nrm = in[41];
    __syncthreads();
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            if (OP == 0) a[i] = fmaxf(a[i], v[i]);
            if (OP == 1) a[i] = fmaxf(fmaxf(a[i], v[(i + 1) & (NCH - 1)]), v[(i + 2) & (NCH - 1)]);
            if (OP == 2) asm volatile("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(cnt[i]) : "f"(v[i]), "f"(thr));
            if (OP == 3) a[i] = fmaxf(a[i], fmaf(-2.0f, v[i], nrm));
            if (OP == 4) asm volatile("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(cnt[i]) : "f"(v[i]), "f"(thr), "r"(1 << i));
        }
        if (OP == 0 || OP == 1) {
#pragma unroll
            for (int i = 0; i < NCH; ++i) asm volatile("" : "+f"(a[i]));
        }
        if (OP == 2 || OP == 4) {
#pragma unroll
            for (int i = 0; i < NCH; ++i) asm volatile("" : "+f"(v[i]));
        }
        if (OP == 3) {
#pragma unroll
            for (int i = 0; i < NCH; ++i) asm volatile("" : "+f"(v[i]));
        }
    }
    __syncthreads();
    const long long c1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += a[i] + cnt[i] + v[i];
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = c1 - c0;
}

template <int X, int WORK>
static int run_tmem(int warps, int ctas_per_sm, int span, long long *dcyc, float *sink)
{
    const int iters = 2000, nsm = 148;
    const int grid = nsm * ctas_per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    tmem_ld_kernel<X, WORK><<<grid, warps * 32>>>(10, span, dcyc, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    tmem_ld_kernel<X, WORK><<<grid, warps * 32>>>(iters, span, dcyc, sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long h[148 * 4];
    CK(cudaMemcpy(h, dcyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    // bytes per CTA: every warp reads 32 lanes x span columns x 4 B per iteration
    const double bytes_cta = (double)iters * warps * 32.0 * span * 4.0;
    printf("tmem_ld x%-2d work=%d warps=%-2d ctas/sm=%d span=%-3d : %8.1f B/clk/SM (max-CTA clocks %lld), %.3f ms -> %.2f TB/s chip, %.0f clk per 64KB tile/SM\n",
           X, WORK, warps, ctas_per_sm, span, bytes_cta * ctas_per_sm / (double)mx, mx, ms, bytes_cta * grid / (ms * 1e-3) / 1e12,
           65536.0 / (bytes_cta * ctas_per_sm / (double)mx));
    return 0;
}

template <int OP>
static int run_alu(int warps, int ctas_per_sm, const float *din, long long *dcyc, float *sink, const char *name, int instr_per_elem)
{
    const int iters = 4000, grid = 148 * ctas_per_sm;
    alu_kernel<OP><<<grid, warps * 32>>>(10, din, dcyc, sink);
    CK(cudaDeviceSynchronize());
    alu_kernel<OP><<<grid, warps * 32>>>(iters, din, dcyc, sink);
    CK(cudaDeviceSynchronize());
    long long h[148 * 4];
    CK(cudaMemcpy(h, dcyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    const double elems = (double)iters * 16 * warps * ctas_per_sm;   // warp-level "elements groups" per SM
    printf("alu %-28s warps=%-2d ctas/sm=%d : %.3f warp-ops/clk/SM (%d instr each) -> %.1f clk per 128x128 tile\n", name, warps, ctas_per_sm,
           elems / (double)mx, instr_per_elem, 512.0 / (elems / (double)mx));
    return 0;
}

int main()
{
    long long *dcyc; float *sink, *din;
    CK(cudaMalloc(&dcyc, sizeof(long long) * 148 * 4));
    CK(cudaMalloc(&sink, 64));
    CK(cudaMalloc(&din, 256));
    float h[64]; for (int i = 0; i < 64; ++i) h[i] = 0.37f * i - 3.0f;
    CK(cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice));
    int rc = 0;
    // TMEM read-out: 4 / 8 / 16 warps, 1 / 2 CTAs per SM; a "tile" = 128 lanes x 128 columns
    rc |= run_tmem<16, 0>(4, 1, 128, dcyc, sink);
    rc |= run_tmem<32, 0>(4, 1, 128, dcyc, sink);
    rc |= run_tmem<16, 0>(8, 1, 64, dcyc, sink);
    rc |= run_tmem<32, 0>(8, 1, 64, dcyc, sink);
    rc |= run_tmem<16, 0>(16, 1, 32, dcyc, sink);
    rc |= run_tmem<32, 0>(16, 1, 32, dcyc, sink);
    rc |= run_tmem<16, 0>(8, 2, 64, dcyc, sink);
    rc |= run_tmem<32, 0>(8, 2, 64, dcyc, sink);
    rc |= run_tmem<32, 0>(4, 2, 128, dcyc, sink);
    rc |= run_tmem<16, 1>(8, 1, 64, dcyc, sink);
    rc |= run_tmem<16, 1>(8, 2, 64, dcyc, sink);
    rc |= run_tmem<32, 1>(8, 2, 64, dcyc, sink);
    rc |= run_tmem<16, 2>(8, 1, 64, dcyc, sink);
    rc |= run_tmem<16, 2>(8, 2, 64, dcyc, sink);
    rc |= run_tmem<32, 2>(8, 2, 64, dcyc, sink);
    rc |= run_tmem<32, 2>(16, 1, 32, dcyc, sink);
    rc |= run_tmem<32, 2>(4, 2, 128, dcyc, sink);
    rc |= run_alu<0>(8, 2, din, dcyc, sink, "FMNMX", 1);
    rc |= run_alu<1>(8, 2, din, dcyc, sink, "FMNMX3 (2 elements)", 1);
    rc |= run_alu<2>(8, 2, din, dcyc, sink, "FSETP+@p IADD", 2);
    rc |= run_alu<3>(8, 2, din, dcyc, sink, "FFMA+FMNMX", 2);
    rc |= run_alu<4>(8, 2, din, dcyc, sink, "FSETP+@p LOP3", 2);
    rc |= run_alu<0>(16, 1, din, dcyc, sink, "FMNMX", 1);
    rc |= run_alu<1>(16, 1, din, dcyc, sink, "FMNMX3 (2 elements)", 1);
    return rc;
}
