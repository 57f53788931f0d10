// api.cu -- error plumbing, version and workspace sizing of the C ABI (include/mlsp_b200.h)
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mlsp {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what)
{
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return MLSP_ECUDA;
}

int sm_count()
{
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

size_t knn_workspace_bytes(int B, int C, int N, int k);
size_t edge_workspace_bytes(int B, int C, int N, int k);
size_t chamfer_workspace_bytes(int B, int N);
size_t graph_feature_workspace_bytes(int B, int C, int N, int k);
size_t edgeconv_workspace_bytes(int B, int O, int N);

}  // namespace mlsp

extern "C" {

int mlsp_version(void) { return 100; }

const char *mlsp_last_error(void) { return mlsp::g_err; }

size_t mlsp_workspace_bytes(int op, int B, int C, int N, int k)
{
    if (B <= 0 || N <= 0) return 0;
    switch (op) {
        case MLSP_OP_KNN: return mlsp::knn_workspace_bytes(B, C, N, k);
        case MLSP_OP_EDGE_FWD:
        case MLSP_OP_EDGE_BWD: return mlsp::edge_workspace_bytes(B, C, N, k);
        case MLSP_OP_CHAMFER: return mlsp::chamfer_workspace_bytes(B, N);
        case MLSP_OP_GRAPH_FEATURE: return mlsp::graph_feature_workspace_bytes(B, C, N, k);
        case MLSP_OP_EDGECONV_BWD: return mlsp::edgeconv_workspace_bytes(B, C, N);
        default: return 0;
    }
}
}
