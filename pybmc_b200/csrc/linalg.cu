// C ABI: orthogonalisation / residual-setup entry points (see include/bmc_b200.h).
#include <algorithm>
#include <string>
#include "common.h"
#include "linalg_kernels.cuh"

namespace bmc {
static thread_local std::string g_last_error;
void set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}
static long long gram_chunks(long long n) {
    // enough row chunks to fill the machine for small M, few enough that partials stay small
    const long long by_rows = (n + 255) / 256;
    return std::max(1ll, std::min(by_rows, 296ll));
}
}  // namespace bmc

using namespace bmc;

extern "C" {

int bmc_version(void) { return 100; }

const char* bmc_last_error(void) { return g_last_error.c_str(); }

int bmc_device_caps(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin) {
    cudaDeviceProp p;
    BMC_CUDA(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (smem_optin) *smem_optin = p.sharedMemPerBlockOptin;
    return BMC_OK;
}

int bmc_center_rows(const double* preds, int64_t n, int m, int64_t ld, const double* truth, double* mu, double* y,
                    double* xc, int64_t ldx, void* stream) {
    BMC_REQUIRE(preds && mu, "bmc_center_rows: preds and mu are required");
    BMC_REQUIRE(n >= 0 && m >= 1 && ld >= m, "bmc_center_rows: bad shape n=%lld m=%d ld=%lld", (long long)n, m,
                (long long)ld);
    BMC_REQUIRE(!y || truth, "bmc_center_rows: y requested without truth");
    BMC_REQUIRE(!xc || ldx >= m, "bmc_center_rows: ldx < m");
    if (n == 0) return BMC_OK;
    const int threads = 256;
    const long long warps_needed = n;
    const int blocks = static_cast<int>(std::min<long long>((warps_needed * 32 + threads - 1) / threads, sm_count() * 16));
    center_rows_kernel<<<blocks, threads, 0, as_stream(stream)>>>(preds, n, m, ld, truth, mu, y, xc, ldx);
    BMC_LAUNCH_CHECK();
    return BMC_OK;
}

size_t bmc_gram_workspace_bytes(int64_t n, int m1) {
    return static_cast<size_t>(gram_chunks(n)) * m1 * m1 * sizeof(double);
}

int bmc_gram(const double* a, int64_t n, int m, int64_t ld, const double* mu, const double* extra, double* gram,
             void* workspace, size_t workspace_bytes, void* stream) {
    BMC_REQUIRE(a && gram, "bmc_gram: a and gram are required");
    BMC_REQUIRE(n >= 1 && m >= 1 && ld >= m, "bmc_gram: bad shape n=%lld m=%d ld=%lld", (long long)n, m,
                (long long)ld);
    const int m1 = m + (extra ? 1 : 0);
    const long long chunks = gram_chunks(n);
    if (workspace_bytes < static_cast<size_t>(chunks) * m1 * m1 * sizeof(double) || !workspace) {
        set_error("bmc_gram: workspace too small (%zu bytes given)", workspace_bytes);
        return BMC_ERR_WORKSPACE;
    }
    const long long rows_per_chunk = ((n + chunks - 1) / chunks + kGramRows - 1) / kGramRows * kGramRows;
    const int nt = (m1 + kGramTile - 1) / kGramTile;
    dim3 grid(static_cast<unsigned>((n + rows_per_chunk - 1) / rows_per_chunk), nt * (nt + 1) / 2);
    double* partial = static_cast<double*>(workspace);
    gram_partial_kernel<<<grid, 256, 0, as_stream(stream)>>>(a, n, m, ld, mu, extra, m1, rows_per_chunk, partial);
    BMC_LAUNCH_CHECK();
    const long long width = static_cast<long long>(m1) * m1;
    sum_partials_kernel<<<static_cast<unsigned>((width + 255) / 256), 256, 0, as_stream(stream)>>>(partial, grid.x,
                                                                                                   width, gram);
    BMC_LAUNCH_CHECK();
    return BMC_OK;
}

int bmc_project_rows(const double* a, int64_t n, int m, int64_t ld, const double* mu, const double* vt, int k,
                     double* out, int64_t ldo, void* stream) {
    BMC_REQUIRE(a && vt && out, "bmc_project_rows: null pointer");
    BMC_REQUIRE(n >= 0 && m >= 1 && k >= 1 && ld >= m && ldo >= k, "bmc_project_rows: bad shape");
    if (n == 0) return BMC_OK;
    dim3 grid(static_cast<unsigned>((n + 63) / 64), (k + 31) / 32);
    project_rows_kernel<<<grid, 256, 0, as_stream(stream)>>>(a, n, m, ld, mu, vt, k, out, ldo);
    BMC_LAUNCH_CHECK();
    return BMC_OK;
}

size_t bmc_rss_workspace_bytes(int64_t n) {
    const long long blocks = std::max(1ll, std::min<long long>((n + 255) / 256, 1184));
    return static_cast<size_t>(blocks) * sizeof(double);
}

int bmc_residual_ss(const double* x, int64_t n, int k, int64_t ld, const double* y, const double* b, double* rss,
                    void* workspace, size_t workspace_bytes, void* stream) {
    BMC_REQUIRE(x && y && b && rss, "bmc_residual_ss: null pointer");
    BMC_REQUIRE(n >= 1 && k >= 1 && ld >= k, "bmc_residual_ss: bad shape");
    const long long blocks = std::max(1ll, std::min<long long>((n + 255) / 256, 1184));
    if (!workspace || workspace_bytes < static_cast<size_t>(blocks) * sizeof(double)) {
        set_error("bmc_residual_ss: workspace too small");
        return BMC_ERR_WORKSPACE;
    }
    double* partial = static_cast<double*>(workspace);
    rss_partial_kernel<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(x, n, k, ld, y, b, partial);
    BMC_LAUNCH_CHECK();
    sum_partials_kernel<<<1, 32, 0, as_stream(stream)>>>(partial, blocks, 1, rss);
    BMC_LAUNCH_CHECK();
    return BMC_OK;
}

int bmc_nearest_class(const double* points, int64_t n, const double* refs, int64_t r, int dim, double d1,
                      double d2, int32_t* cls, void* stream) {
    BMC_REQUIRE(points && refs && cls, "bmc_nearest_class: null pointer");
    BMC_REQUIRE(n >= 1 && r >= 1 && dim >= 1 && dim <= 1024, "bmc_nearest_class: bad shape n=%lld r=%lld dim=%d",
                (long long)n, (long long)r, dim);
    const int tile_refs = 1024 / dim > 0 ? 1024 / dim : 1;
    const size_t smem = sizeof(double) * tile_refs * dim;
    nearest_class_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, smem, as_stream(stream)>>>(
        points, n, refs, r, dim, d1, d2, cls);
    BMC_LAUNCH_CHECK();
    return BMC_OK;
}

}  // extern "C"
