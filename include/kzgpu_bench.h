/* kzgpu_bench.h -- throughput microbenchmarks (libkzgpu_bench.so): measurement tooling, NOT part of the product ABI.
 *
 * bench.py uses kind 0 as the live denominator of the integer roofline (the reference has no counterpart: it is pure
 * Python); the other kinds back the design measurements quoted in DESIGN.md section 4 / section 7.
 * The library is stand-alone (own CUDA stream on `device`) and is never loaded by the product path. */
#ifndef KZGPU_BENCH_H
#define KZGPU_BENCH_H

#ifdef __cplusplus
extern "C" {
#endif

/* runs `iters` dependent operations per thread on blocks*threads threads of `device` and returns the elapsed device ms
 * (second of two runs).  Returns 0, -1 (bad argument) or -2 (CUDA failure).
 * kind: 0 = IMAD.WIDE.U32 carry chain (raw pipe), 1 = Fp(BN254) Montgomery mul,
 *       2 = Fp(BLS12-381) Montgomery mul, 3 = XYZZ mixed add BN254, 4 = XYZZ mixed add BLS,
 *       5..11 = issue-slot probes (ALU / FP64 beside IMAD.WIDE), 12 / 13 = batched-affine pair additions with one
 *       inversion per thread over `iters` pairs, operands consecutive (12) or gathered from an 8 GiB table (13);
 *       *ops = operations performed */
int kzgpu_microbench(int device, int kind, int blocks, int threads, int iters, float* ms, double* ops);

#ifdef __cplusplus
}
#endif
#endif
