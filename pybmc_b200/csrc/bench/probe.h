/*
 * probe.h -- private header of libbmc_probe.so (benchmark tooling, not the product ABI).
 *
 * Register-only kernels that measure what this GPU sustains on the pipes the sampler and the fused
 * prediction are bound by (MEASURED_PEAKS.json only holds the HBM and bf16 tensor
 * peaks): FP32 FMA (kind 0), MUFU (1), the Philox integer mix IMAD.WIDE + LOP3 (2), FFMA + LOP3 dual issue
 * (3); single-instruction streams IMAD.WIDE+IADD / IMAD.HI / IMAD / LOP3 (4-7), packed FFMA2 (8), IMAD.WIDE
 * alone (9), and the mixes that show which classes share a pipe: IMAD.WIDE + FFMA (10), + 2 FFMA (11),
 * + FFMA2 (12), MUFU + 4 FFMA (13), MUFU + 2 IMAD.WIDE (14), FFMA2 + FFMA (15); DFMA with two shared operands (16),
 * DFMA with three distinct register operands (17), the same with a LOP3 beside each (18); FFMA / FFMA2 with three distinct register operands (19, 20).
 * `iters` loop iterations per thread, each issuing bmc_probe_ops_per_iteration(kind) thread-level
 * operations; `sink` is a device float[1].  Returns 0, -1 (bad argument) or -2 (launch failure).
 */
#ifndef BMC_PROBE_H
#define BMC_PROBE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
int bmc_probe_ops_per_iteration(int kind);
int bmc_probe(int kind, int64_t iters, int blocks, int threads, float* sink, void* stream);
#ifdef __cplusplus
}
#endif
#endif /* BMC_PROBE_H */
