// Throughput microbenchmarks and the batched-affine feasibility probe: MEASUREMENT TOOLING, built into its own library
// (libkzgpu_bench.so, include/kzgpu_bench.h) so that the product library carries no benchmark kernels.  bench.py uses kind 0
// (raw IMAD.WIDE.U32 issue rate) as the live denominator of the integer roofline; the other kinds back the measurements
// quoted in DESIGN.md section 4 and profiles/r1_microbench.txt.  Stand-alone: own stream and events on the device the caller
// names (this library links its own static CUDA runtime, whose current device is independent of libkzgpu.so's).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "../../include/kzgpu_bench.h"
#include "params_gen.cuh"
#include "curve.cuh"

namespace {

struct BenchCtx { int device = -1; cudaStream_t stream = nullptr; cudaEvent_t ev0 = nullptr, ev1 = nullptr; char err[256] = {0}; };
BenchCtx g_bx;
int mb_fail(const char* what, cudaError_t e) { snprintf(g_bx.err, sizeof(g_bx.err), "%s: %s", what, cudaGetErrorString(e)); return -2; }
#define MB_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return mb_fail(#expr, e__); } while (0)
static inline size_t mb_div_up(size_t a, size_t b) { return (a + b - 1) / b; }

// --- microbenchmarks ---------------------------------------------------------------------
// raw IMAD.WIDE.U32(.X) issue rate: 4 independent 8-limb carry chains per thread
__global__ void mb_imad_kernel(uint32_t* sink, int iters, uint32_t seed) {
  uint32_t acc[4][8], a[8];
  for (int i = 0; i < 8; i++) a[i] = seed * (i + 3) + threadIdx.x;
  for (int c = 0; c < 4; c++)
    for (int i = 0; i < 8; i++) acc[c][i] = seed + c * 17 + i;
  uint32_t top = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
      MpPrims<8>::mad_even(acc[c], a, acc[(c + 1) & 3][0], top);       // 4 IMAD.WIDE each
      MpPrims<8>::mad_even(acc[c], a + 1, acc[(c + 2) & 3][1], top);
    }
  }
  uint32_t s = top;
  for (int c = 0; c < 4; c++)
    for (int i = 0; i < 8; i++) s ^= acc[c][i];
  if (s == 0x12345678u) sink[0] = s;
}

// IMAD.WIDE chains as in mb_imad_kernel plus ALU independent 3-input adds per 8 wide multiplies:
// measures whether ALU-pipe work issues in the shadow of a saturated IMAD.WIDE stream.
template <int ALU> __global__ void mb_imad_alu_kernel(uint32_t* sink, int iters, uint32_t seed) {
  uint32_t acc[4][8], a[8], z[8];
  for (int i = 0; i < 8; i++) { a[i] = seed * (i + 3) + threadIdx.x; z[i] = seed + i * 5 + threadIdx.x; }
  for (int c = 0; c < 4; c++)
    for (int i = 0; i < 8; i++) acc[c][i] = seed + c * 17 + i;
  uint32_t top = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
      MpPrims<8>::mad_even(acc[c], a, acc[(c + 1) & 3][0], top);
#pragma unroll
      for (int k = 0; k < ALU / 2; k++) asm volatile("xor.b32 %0, %0, %1; shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(z[k & 7]) : "r"(a[(k + 1) & 7]), "r"(seed));
      MpPrims<8>::mad_even(acc[c], a + 1, acc[(c + 2) & 3][1], top);
#pragma unroll
      for (int k = 0; k < ALU / 2; k++) asm volatile("xor.b32 %0, %0, %1; shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(z[(k + 4) & 7]) : "r"(a[(k + 2) & 7]), "r"(seed));
    }
  }
  uint32_t s = top;
  for (int c = 0; c < 4; c++)
    for (int i = 0; i < 8; i++) s ^= acc[c][i];
  for (int i = 0; i < 8; i++) s ^= z[i];
  if (s == 0x12345678u) sink[0] = s;
}

// narrow multiply-add rate: independent 32-bit mad.lo chains (IMAD)
__global__ void mb_imad32_kernel(uint32_t* sink, int iters, uint32_t seed) {
  uint32_t x[16], a = seed + threadIdx.x, b = seed * 3 + 1;
  for (int i = 0; i < 16; i++) x[i] = seed + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int i = 0; i < 16; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
  }
  uint32_t s = 0;
  for (int i = 0; i < 16; i++) s ^= x[i];
  if (s == 0x12345678u) sink[0] = s;
}

// FP64 pipe: independent DFMA.RZ chains, optionally with ALU (IADD3 pairs = 64-bit adds) and
// IMAD.WIDE work interleaved -- is the FP64 pipe a second multiplier next to the integer one?
template <int ALU, int WIDE> __global__ void mb_dfma_kernel(uint32_t* sink, int iters, uint32_t seed) {
  double x[8], a = 1.0 + 1e-9 * (seed & 7), b = 1e-3 * threadIdx.x;
  unsigned long long z[4];
  uint32_t acc[8], m[8];
  uint32_t top = 0;
  for (int i = 0; i < 8; i++) { x[i] = (double)(seed + i); acc[i] = seed + i; m[i] = seed * (i + 3) + threadIdx.x; }
  for (int i = 0; i < 4; i++) z[i] = seed + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) x[i] = __fma_rz(x[i], a, b);
#pragma unroll
      for (int k = 0; k < ALU; k++) z[k & 3] += __double_as_longlong(x[k & 7]) + z[(k + 1) & 3];
      if (WIDE) MpPrims<8>::mad_even(acc, m, acc[r], top);
    }
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += x[i];
  unsigned long long t = top;
  for (int i = 0; i < 4; i++) t ^= z[i];
  for (int i = 0; i < 8; i++) t ^= acc[i];
  if (s == 0.12345 || t == 0x12345678ull) sink[0] = (uint32_t)t;
}

template <class P> __global__ void mb_mul_kernel(uint32_t* sink, int iters, uint32_t seed) {
  Fe<P> x[4];
  for (int c = 0; c < 4; c++)
    for (int i = 0; i < P::N; i++) x[c].v[i] = (seed * (c + 1) + i * 7 + threadIdx.x) & 0x0fffffffu;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < 4; c++) x[c] = fe_mul<P>(x[c], x[(c + 1) & 3]);
  }
  uint32_t s = 0;
  for (int c = 0; c < 4; c++)
    for (int i = 0; i < P::N; i++) s ^= x[c].v[i];
  if (s == 0x12345678u) sink[0] = s;
}

template <class P> __global__ void mb_madd_kernel(uint32_t* sink, int iters, uint32_t seed) {
  XYZZ<P> acc;
  Affine<P> pt;
  for (int i = 0; i < P::N; i++) {
    uint32_t v = (seed + i * 13 + threadIdx.x) & 0x0fffffffu;
    acc.x.v[i] = v; acc.y.v[i] = v ^ 0x55; acc.zz.v[i] = v + 9; acc.zzz.v[i] = v + 11;
    pt.x.v[i] = v + 3; pt.y.v[i] = v + 5;
  }
  for (int it = 0; it < iters; it++) {
    xyzz_madd<P>(acc, pt);
    pt.x.v[0] ^= acc.x.v[0] & 1;     // keep the operand live and varying
  }
  uint32_t s = 0;
  for (int i = 0; i < P::N; i++) s ^= acc.x.v[i] ^ acc.y.v[i] ^ acc.zz.v[i] ^ acc.zzz.v[i];
  if (s == 0x12345678u) sink[0] = s;
}

}  // namespace

// ---- batched-affine feasibility probe ---------------------------------------------------------
// npairs independent affine additions P[2k] + P[2k+1] with ONE inversion per thread: thread t owns pairs t, t+T, ...
// (coalesced); forward pass stores the running product of the denominators x2 - x1 (32 B per pair), one Fermat inversion,
// backward pass peels the inverses off and finishes the additions (6 modmul per addition + 381 / K for the inversion).
// GATHER: operands are fetched through a random index into a table far larger than L2 (the first pairing round of a bucket
// sum reads the key's window tables this way); otherwise they are consecutive (the later rounds).  Operands are arbitrary
// field elements, which is all the arithmetic cares about.
namespace {
template <class P> __device__ __forceinline__ Fe<P> mb_ld(const uint32_t* p) {
  Fe<P> r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < P::N / 4; i++) { uint4 t = __ldg(q + i); r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w; }
  return r;
}
template <class P> __device__ __forceinline__ void mb_st(uint32_t* p, const Fe<P>& a) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < P::N / 4; i++) q[i] = make_uint4(a.v[4 * i], a.v[4 * i + 1], a.v[4 * i + 2], a.v[4 * i + 3]);
}
__global__ void mb_fill_kernel(uint32_t* buf, size_t words, uint32_t seed) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= words) return;
  uint32_t x = (uint32_t)i * 2654435761u ^ seed ^ (uint32_t)(i >> 32);
  x ^= x << 13; x ^= x >> 17; x ^= x << 5;
  buf[i] = (i & 7) == 7 ? (x & 0x0fffffffu) : x;          // every 8-word element stays below the 254-bit moduli
}
__global__ void mb_index_kernel(uint32_t* idx, size_t n, uint32_t mask, uint32_t seed) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x = (uint32_t)i * 747796405u + seed;
  x ^= x >> 16; x *= 2246822519u; x ^= x >> 13; x *= 3266489917u; x ^= x >> 16;
  idx[i] = x & mask;
}
template <class P, bool GATHER>
__global__ void __launch_bounds__(128) mb_affine_pairs_kernel(const uint32_t* __restrict__ pts, const uint32_t* __restrict__ idx,
                                                             uint32_t npairs, uint32_t* __restrict__ pre, uint32_t* __restrict__ out) {
  const uint32_t T = gridDim.x * blockDim.x, t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npairs) return;
  auto at = [&](uint32_t slot) { return pts + (size_t)(GATHER ? __ldg(idx + slot) : slot) * 2 * P::N; };
  Fe<P> acc = fe_one<P>();
  Fe<P> nx1 = mb_ld<P>(at(2 * t)), nx2 = mb_ld<P>(at(2 * t + 1));
  uint32_t last = t;
  for (uint32_t k = t; k < npairs; k += T) {
    Fe<P> x1 = nx1, x2 = nx2;
    if (k + T < npairs) { nx1 = mb_ld<P>(at(2 * (k + T))); nx2 = mb_ld<P>(at(2 * (k + T) + 1)); }
    mb_st<P>(pre + (size_t)k * P::N, acc);
    acc = fe_mul<P>(acc, fe_sub<P>(x2, x1));
    last = k;
  }
  Fe<P> inv = fe_inv<P>(acc);
  const uint32_t* p1 = at(2 * last); const uint32_t* p2 = at(2 * last + 1);
  Fe<P> a1 = mb_ld<P>(p1), b1 = mb_ld<P>(p1 + P::N), a2 = mb_ld<P>(p2), b2 = mb_ld<P>(p2 + P::N), pr = mb_ld<P>(pre + (size_t)last * P::N);
  for (uint32_t k = last;; k -= T) {
    Fe<P> x1 = a1, y1 = b1, x2 = a2, y2 = b2, pk = pr;
    if (k >= T) {
      p1 = at(2 * (k - T)); p2 = at(2 * (k - T) + 1);
      a1 = mb_ld<P>(p1); b1 = mb_ld<P>(p1 + P::N); a2 = mb_ld<P>(p2); b2 = mb_ld<P>(p2 + P::N); pr = mb_ld<P>(pre + (size_t)(k - T) * P::N);
    }
    Fe<P> d = fe_sub<P>(x2, x1);
    Fe<P> di = fe_mul<P>(inv, pk);
    inv = fe_mul<P>(inv, d);
    Fe<P> lam = fe_mul<P>(fe_sub<P>(y2, y1), di);
    Fe<P> x3 = fe_sub<P>(fe_sub<P>(fe_sqr<P>(lam), x1), x2);
    Fe<P> y3 = fe_sub<P>(fe_mul<P>(lam, fe_sub<P>(x1, x3)), y1);
    mb_st<P>(out + (size_t)k * 2 * P::N, x3);
    mb_st<P>(out + (size_t)k * 2 * P::N + P::N, y3);
    if (k < T) break;
  }
}
}  // namespace

extern "C" {

// kind 12 / 13: `iters` pairs per thread; *ops = additions performed
static int microbench_affine(bool gather, int blocks, int threads, int iters, float* ms, double* ops) {
  BenchCtx& cx = g_bx;
  const size_t npairs = (size_t)blocks * threads * iters;
  if (npairs >= (1ull << 31)) return -1;
  const size_t table_pts = gather ? (1ull << 27) : 2 * npairs;         // 8 GiB table for the gather probe
  uint32_t *pts = nullptr, *idx = nullptr, *pre = nullptr, *out = nullptr;
  MB_CUDA(cudaMalloc(&pts, table_pts * 64));
  MB_CUDA(cudaMalloc(&pre, npairs * 32));
  MB_CUDA(cudaMalloc(&out, npairs * 64));
  mb_fill_kernel<<<(unsigned)mb_div_up(table_pts * 16, 256), 256, 0, cx.stream>>>(pts, table_pts * 16, 99u);
  if (gather) {
    MB_CUDA(cudaMalloc(&idx, 2 * npairs * 4));
    mb_index_kernel<<<(unsigned)mb_div_up(2 * npairs, 256), 256, 0, cx.stream>>>(idx, 2 * npairs, (uint32_t)(table_pts - 1), 7u);
  }
  for (int rep = 0; rep < 2; rep++) {
    MB_CUDA(cudaEventRecord(cx.ev0, cx.stream));
    if (gather) mb_affine_pairs_kernel<FpBN254, true><<<blocks, threads, 0, cx.stream>>>(pts, idx, (uint32_t)npairs, pre, out);
    else mb_affine_pairs_kernel<FpBN254, false><<<blocks, threads, 0, cx.stream>>>(pts, idx, (uint32_t)npairs, pre, out);
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaEventRecord(cx.ev1, cx.stream));
    MB_CUDA(cudaEventSynchronize(cx.ev1));
    MB_CUDA(cudaEventElapsedTime(ms, cx.ev0, cx.ev1));
  }
  if (ops) *ops = (double)npairs;
  cudaFree(pts); cudaFree(idx); cudaFree(pre); cudaFree(out);
  return 0;
}

int kzgpu_microbench(int device, int kind, int blocks, int threads, int iters, float* ms, double* ops) {
  if (g_bx.device != device) {
    MB_CUDA(cudaSetDevice(device));
    if (!g_bx.stream) {
      MB_CUDA(cudaStreamCreateWithFlags(&g_bx.stream, cudaStreamNonBlocking));
      MB_CUDA(cudaEventCreate(&g_bx.ev0));
      MB_CUDA(cudaEventCreate(&g_bx.ev1));
    }
    g_bx.device = device;
  }
  if (kind == 12 || kind == 13) {
    if (blocks <= 0 || threads <= 0 || threads > 128 || iters <= 0 || !ms) return -1;
    return microbench_affine(kind == 13, blocks, threads, iters, ms, ops);
  }
  if (blocks <= 0 || threads <= 0 || threads > 1024 || iters <= 0 || !ms) return -1;
  BenchCtx& cx = g_bx;
  uint32_t* sink = nullptr;
  MB_CUDA(cudaMalloc(&sink, 4));
  double per_thread = 0;
  for (int rep = 0; rep < 2; rep++) {       // rep 0 = warm-up
    MB_CUDA(cudaEventRecord(cx.ev0, cx.stream));
    switch (kind) {
      case 0: mb_imad_kernel<<<blocks, threads, 0, cx.stream>>>(sink, iters, 12345u); per_thread = 32.0 * iters; break;
      case 1: mb_mul_kernel<FpBN254><<<blocks, threads, 0, cx.stream>>>(sink, iters, 12345u); per_thread = 4.0 * iters; break;
      case 2: mb_mul_kernel<FpBLS381><<<blocks, threads, 0, cx.stream>>>(sink, iters, 12345u); per_thread = 4.0 * iters; break;
      case 3: mb_madd_kernel<FpBN254><<<blocks, threads, 0, cx.stream>>>(sink, iters, 12345u); per_thread = 1.0 * iters; break;
      case 4: mb_madd_kernel<FpBLS381><<<blocks, threads, 0, cx.stream>>>(sink, iters, 12345u); per_thread = 1.0 * iters; break;
      case 5: mb_imad_alu_kernel<8><<<blocks, threads, 0, cx.stream>>>(sink, iters, 12345u); per_thread = 32.0 * iters; break;    // 2 ALU ops per wide
      case 6: mb_imad_alu_kernel<16><<<blocks, threads, 0, cx.stream>>>(sink, iters, 12345u); per_thread = 32.0 * iters; break;   // 4 ALU ops per wide
      case 7: mb_imad32_kernel<<<blocks, threads, 0, cx.stream>>>(sink, iters, 12345u); per_thread = 32.0 * iters; break;
      case 8: mb_dfma_kernel<0, 0><<<blocks, threads, 0, cx.stream>>>(sink, iters, 12345u); per_thread = 32.0 * iters; break;    // DFMA only
      case 9: mb_dfma_kernel<8, 0><<<blocks, threads, 0, cx.stream>>>(sink, iters, 12345u); per_thread = 32.0 * iters; break;    // + one 64-bit 3-input add per DFMA
      case 10: mb_dfma_kernel<0, 1><<<blocks, threads, 0, cx.stream>>>(sink, iters, 12345u); per_thread = 32.0 * iters; break;   // + one IMAD.WIDE per 2 DFMA
      case 11: mb_dfma_kernel<8, 1><<<blocks, threads, 0, cx.stream>>>(sink, iters, 12345u); per_thread = 32.0 * iters; break;
      default: cudaFree(sink); return -1;
    }
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaEventRecord(cx.ev1, cx.stream));
    MB_CUDA(cudaEventSynchronize(cx.ev1));
    MB_CUDA(cudaEventElapsedTime(ms, cx.ev0, cx.ev1));
  }
  if (ops) *ops = per_thread * (double)blocks * (double)threads;
  cudaFree(sink);
  return 0;
}

}  // extern "C"
