/* Plain-C client of libbmc_b200.so: no Python, no torch -- only the CUDA runtime for device memory.
 * Built with gcc and run by tests/test_gpu_cabi_from_c.py, which checks the printed numbers against
 * the oracle.  Problem: the 3 x 2 toy regression of the upstream tests (tests/test_inference_utils.py:6-14)
 * with the sufficient statistics and the diagonalising transform handed in by the caller. */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include "../include/bmc_b200.h"

#define CK(x) do { int rc_ = (x); if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, bmc_last_error()); return 2; } } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 3; } } while (0)

static void* to_dev(const void* src, size_t bytes) {
    void* p = NULL;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return NULL;
    if (src) cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice);
    return p;
}

int main(int argc, char** argv) {
    /* constants of the diagonalised problem come from the caller (argv): d0 d1 pull0 pull1 g0 g1 w00 w01 w10 w11 rss_min sigma2_init */
    if (argc != 13) { fprintf(stderr, "usage: cabi_demo <12 numbers>\n"); return 1; }
    double c[12];
    for (int i = 0; i < 12; ++i) c[i] = atof(argv[i + 1]);
    printf("version %d\n", bmc_version());

    /* 1. sufficient statistics of [X | y] on the device */
    const double x[6] = {1, 0, 0, 1, 1, 1}, y[3] = {1.0, 2.0, 3.0};
    double* dx = (double*)to_dev(x, sizeof x);
    double* dy = (double*)to_dev(y, sizeof y);
    double* dgram = (double*)to_dev(NULL, 9 * sizeof(double));
    size_t wsb = bmc_gram_workspace_bytes(3, 3);
    void* ws = to_dev(NULL, wsb);
    CK(bmc_gram(dx, 3, 2, 2, NULL, dy, dgram, ws, wsb, NULL));
    double gram[9];
    CU(cudaMemcpy(gram, dgram, sizeof gram, cudaMemcpyDeviceToHost));
    printf("gram");
    for (int i = 0; i < 9; ++i) printf(" %.17g", gram[i]);
    printf("\n");

    /* 2. four conjugate chains of 5 iterations, fp64, every iterate kept */
    double* dconst = (double*)to_dev(c, 10 * sizeof(double));
    bmc_gibbs_problem p;
    p.k = 2; p.d = dconst; p.pull = dconst + 2; p.g_ols = dconst + 4; p.w = dconst + 6; p.dense_w = 1;
    p.rss_min = c[10]; p.n_obs = 3.0; p.nu0 = 1.0; p.sigma20 = 1.0; p.sigma2_init = c[11];
    p.layout = BMC_LAYOUT_AUTO;
    p.workspace = NULL; p.workspace_bytes = 0;          /* plain launch */
    const int chains = 4, iters = 5;
    double* dsamples = (double*)to_dev(NULL, sizeof(double) * iters * 3 * chains);
    CK(bmc_gibbs_run(BMC_F64, &p, 42, 0, chains, iters, 0, 1, iters, dsamples, NULL, BMC_STATS_NONE, NULL, NULL));
    double samples[5 * 3 * 4];
    CU(cudaMemcpy(samples, dsamples, sizeof samples, cudaMemcpyDeviceToHost));
    for (int ch = 0; ch < chains; ++ch)
        for (int t = 0; t < iters; ++t)
            printf("sample %d %d %.17g %.17g %.17g\n", ch, t, samples[(t * 3 + 0) * chains + ch],
                   samples[(t * 3 + 1) * chains + ch], samples[(t * 3 + 2) * chains + ch]);

    /* 3. order counts of a small matrix against a truth vector */
    const double mat[8] = {1, 5, 2, 5, 3, 7, 4, 5}, truth[2] = {2.5, 5.0};       /* 4 draws x 2 points */
    double* dmat = (double*)to_dev(mat, sizeof mat);
    double* dtruth = (double*)to_dev(truth, sizeof truth);
    int64_t* dlt = (int64_t*)to_dev(NULL, 2 * sizeof(int64_t));
    int64_t* dle = (int64_t*)to_dev(NULL, 2 * sizeof(int64_t));
    CK(bmc_coverage_counts(dmat, 4, 2, 2, dtruth, dlt, dle, NULL));
    int64_t lt[2], le[2];
    CU(cudaMemcpy(lt, dlt, sizeof lt, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(le, dle, sizeof le, cudaMemcpyDeviceToHost));
    printf("counts %lld %lld %lld %lld\n", (long long)lt[0], (long long)le[0], (long long)lt[1], (long long)le[1]);

    /* 4. argument errors come back as codes + messages, not crashes */
    int rc = bmc_gibbs_run(BMC_F64, &p, 42, 0, 0, iters, 0, 1, iters, dsamples, NULL, BMC_STATS_NONE, NULL, NULL);
    printf("bad call -> %d (%s)\n", rc, bmc_last_error());
    CU(cudaDeviceSynchronize());
    return 0;
}
