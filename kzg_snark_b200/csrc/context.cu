// Library context, device-memory plumbing, field self-test and throughput microbenchmarks.
#include "common.cuh"
#include <cstring>
#include <cctype>
#include <unistd.h>
#include <sys/syscall.h>

void kz_ntt_release();
void kz_msm_release();
void kz_poly_release();
void kz_plonk_release();

KzgpuCtx& kz_ctx() {
  static KzgpuCtx ctx;
  return ctx;
}

int kz_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(kz_ctx().err, sizeof(kz_ctx().err), fmt, ap);
  va_end(ap);
  return code;
}

int KzScratch::ensure(size_t bytes) {
  if (bytes <= cap) return 0;
  if (p) { cudaFree(p); p = nullptr; cap = 0; }
  size_t want = bytes + bytes / 8;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    e = cudaMalloc(&p, bytes);
    want = bytes;
  }
  if (e != cudaSuccess) { p = nullptr; return kz_fail(KZGPU_ECUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
  cap = want;
  return 0;
}

void KzScratch::release() {
  if (p) cudaFree(p);
  p = nullptr; cap = 0;
}

namespace {

template <class P>
__global__ void field_op_kernel(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fe<P> x, y, r;
  for (int k = 0; k < P::N; k++) { x.v[k] = a[i * P::N + k]; y.v[k] = b ? b[i * P::N + k] : 0; }
  x = fe_to_mont<P>(x);
  y = fe_to_mont<P>(y);
  if (op == 0) r = fe_mul<P>(x, y);
  else if (op == 1) r = fe_add<P>(x, y);
  else if (op == 2) r = fe_sub<P>(x, y);
  else r = fe_inv<P>(x);
  r = fe_from_mont<P>(r);
  for (int k = 0; k < P::N; k++) out[i * P::N + k] = r.v[k];
}

template <class P>
int field_op_impl(int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  KzgpuCtx& cx = kz_ctx();
  size_t bytes = n * P::N * 4;
  uint32_t *da = nullptr, *db = nullptr, *dout = nullptr;
  KZ_CUDA(cudaMalloc(&da, bytes));
  KZ_CUDA(cudaMalloc(&dout, bytes));
  KZ_CUDA(cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, cx.stream));
  if (b) {
    KZ_CUDA(cudaMalloc(&db, bytes));
    KZ_CUDA(cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, cx.stream));
  }
  field_op_kernel<P><<<(unsigned)kz_div_up(n, 128), 128, 0, cx.stream>>>(op, da, db, dout, n);
  KZ_LAUNCHED();
  KZ_CUDA(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, cx.stream));
  KZ_CUDA(cudaStreamSynchronize(cx.stream));
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return 0;
}

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

extern "C" {

int kzgpu_init(int device) {
  KzgpuCtx& cx = kz_ctx();
  if (cx.inited) {
    if (device == cx.device) return 0;
    return kz_fail(KZGPU_EINVAL, "already initialised on device %d", cx.device);
  }
  int count = 0;
  KZ_CUDA(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return kz_fail(KZGPU_EINVAL, "device %d out of range (%d devices)", device, count);
  KZ_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  KZ_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return kz_fail(KZGPU_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  cx.sm_count = prop.multiProcessorCount;
  KZ_CUDA(cudaStreamCreateWithFlags(&cx.own_stream, cudaStreamNonBlocking));
  cx.stream = cx.own_stream;
  KZ_CUDA(cudaStreamCreateWithFlags(&cx.copy_stream, cudaStreamNonBlocking));
  for (int k = 0; k < 4; k++) KZ_CUDA(cudaEventCreateWithFlags(&cx.copy_ev[k], cudaEventDisableTiming));
  {
    int lo_prio = 0, hi_prio = 0;
    KZ_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
    KZ_CUDA(cudaStreamCreateWithPriority(&cx.sort_stream, cudaStreamNonBlocking, hi_prio));
    for (int k = 0; k < 2; k++) {
      KZ_CUDA(cudaEventCreateWithFlags(&cx.sort_ev[k], cudaEventDisableTiming));
      KZ_CUDA(cudaEventCreateWithFlags(&cx.acc_ev[k], cudaEventDisableTiming));
    }
    KZ_CUDA(cudaEventCreateWithFlags(&cx.start_ev, cudaEventDisableTiming));
  }
  KZ_CUDA(cudaEventCreate(&cx.ev0));
  KZ_CUDA(cudaEventCreate(&cx.ev1));
  cx.device = device;
  cx.inited = true;
  cx.launches = 0;
  return 0;
}

int kzgpu_shutdown(void) {
  KzgpuCtx& cx = kz_ctx();
  if (!cx.inited) return 0;
  cudaStreamSynchronize(cx.stream);
  kz_ntt_release();
  kz_msm_release();
  kz_poly_release();
  kz_plonk_release();
  cudaEventDestroy(cx.ev0);
  cudaEventDestroy(cx.ev1);
  cudaStreamDestroy(cx.own_stream);
  cudaStreamDestroy(cx.copy_stream);
  for (int k = 0; k < 4; k++) cudaEventDestroy(cx.copy_ev[k]);
  cudaStreamDestroy(cx.sort_stream);
  for (int k = 0; k < 2; k++) { cudaEventDestroy(cx.sort_ev[k]); cudaEventDestroy(cx.acc_ev[k]); }
  cudaEventDestroy(cx.start_ev);
  cx.sort_stream = nullptr;
  cx.stream = cx.own_stream = nullptr;
  cx.inited = false;
  cx.device = -1;
  return 0;
}

int kzgpu_last_error(char* buf, size_t cap) {
  if (!buf || cap == 0) return KZGPU_EINVAL;
  strncpy(buf, kz_ctx().err, cap - 1);
  buf[cap - 1] = 0;
  return 0;
}

int kzgpu_device_info(char* name, size_t cap, int* sm_count, size_t* total_mem) {
  KZ_REQUIRE_INIT();
  cudaDeviceProp prop;
  KZ_CUDA(cudaGetDeviceProperties(&prop, kz_ctx().device));
  if (name && cap) { strncpy(name, prop.name, cap - 1); name[cap - 1] = 0; }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (total_mem) *total_mem = prop.totalGlobalMem;
  return 0;
}

int kzgpu_fp_limbs64(int curve) {
  if (curve == KZGPU_BN254) return 4;
  if (curve == KZGPU_BLS12_381) return 6;
  return KZGPU_EINVAL;
}

int kzgpu_alloc(void** d_ptr, size_t bytes) {
  KZ_REQUIRE_INIT();
  if (!d_ptr) return kz_fail(KZGPU_EINVAL, "null pointer");
  KZ_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 1));
  return 0;
}

int kzgpu_free(void* d_ptr) {
  KZ_REQUIRE_INIT();
  KZ_CUDA(cudaFree(d_ptr));
  return 0;
}

int kzgpu_h2d(void* d_dst, const void* src, size_t bytes) {
  KZ_REQUIRE_INIT();
  KZ_CUDA(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, kz_ctx().stream));
  KZ_CUDA(cudaStreamSynchronize(kz_ctx().stream));
  return 0;
}

int kzgpu_d2h(void* dst, const void* d_src, size_t bytes) {
  KZ_REQUIRE_INIT();
  KZ_CUDA(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, kz_ctx().stream));
  KZ_CUDA(cudaStreamSynchronize(kz_ctx().stream));
  return 0;
}

int kzgpu_d2d(void* d_dst, const void* d_src, size_t bytes) {
  KZ_REQUIRE_INIT();
  KZ_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, kz_ctx().stream));
  return 0;
}

int kzgpu_memset(void* d_dst, int byte, size_t bytes) {
  KZ_REQUIRE_INIT();
  KZ_CUDA(cudaMemsetAsync(d_dst, byte, bytes, kz_ctx().stream));
  return 0;
}

int kzgpu_sync(void) {
  KZ_REQUIRE_INIT();
  KZ_CUDA(cudaStreamSynchronize(kz_ctx().stream));
  return 0;
}

// NUMA node of the GPU (from sysfs), -1 if unknown
static int gpu_numa_node(int device) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) return -1;
  for (char* c = bus; *c; c++) *c = (char)tolower(*c);
  char path[128];
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE* f = fopen(path, "r");
  if (!f) return -1;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  return node;
}

// Page-locked host memory, placed on the GPU's own NUMA node when the kernel lets us say so:
// a buffer on the far socket halves the device->host rate of the e2e path.
int kzgpu_host_alloc(void** h_ptr, size_t bytes) {
  KZ_REQUIRE_INIT();
  if (!h_ptr) return kz_fail(KZGPU_EINVAL, "null pointer");
  const int node = gpu_numa_node(kz_ctx().device);
  bool policy_set = false;
  if (node >= 0 && node < 64) {
    unsigned long mask = 1ul << node;
    policy_set = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, &mask, 65ul) == 0;
  }
  cudaError_t e = cudaHostAlloc(h_ptr, bytes ? bytes : 1, cudaHostAllocDefault);
  if (policy_set) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0ul);
  if (e != cudaSuccess) return kz_fail(KZGPU_ECUDA, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  return 0;
}

int kzgpu_host_free(void* h_ptr) {
  KZ_REQUIRE_INIT();
  KZ_CUDA(cudaFreeHost(h_ptr));
  return 0;
}

int kzgpu_set_stream(void* cuda_stream) {
  KZ_REQUIRE_INIT();
  KzgpuCtx& cx = kz_ctx();
  KZ_CUDA(cudaStreamSynchronize(cx.stream));
  cx.stream = cuda_stream ? (cudaStream_t)cuda_stream : cx.own_stream;
  return 0;
}

int kzgpu_profile_enable(int on) {
  kz_ctx().profile = on != 0;
  return 0;
}

int kzgpu_profile_reset(void) {
  KzgpuCtx& cx = kz_ctx();
  for (int i = 0; i < 4; i++) { cx.prof_ms[i] = 0; cx.prof_launches[i] = 0; cx.prof_work[i] = 0; }
  return 0;
}

int kzgpu_profile_get(int which, double* total_ms, uint64_t* launches, double* work_units) {
  if (which < 0 || which > 3) return KZGPU_EINVAL;
  KzgpuCtx& cx = kz_ctx();
  if (total_ms) *total_ms = cx.prof_ms[which];
  if (launches) *launches = cx.prof_launches[which];
  if (work_units) *work_units = cx.prof_work[which];
  return 0;
}

int kzgpu_timer_start(void) {
  KZ_REQUIRE_INIT();
  KZ_CUDA(cudaEventRecord(kz_ctx().ev0, kz_ctx().stream));
  return 0;
}

int kzgpu_timer_stop(float* ms) {
  KZ_REQUIRE_INIT();
  KZ_CUDA(cudaEventRecord(kz_ctx().ev1, kz_ctx().stream));
  KZ_CUDA(cudaEventSynchronize(kz_ctx().ev1));
  KZ_CUDA(cudaEventElapsedTime(ms, kz_ctx().ev0, kz_ctx().ev1));
  return 0;
}

int kzgpu_launch_count(uint64_t* count) {
  if (!count) return KZGPU_EINVAL;
  *count = kz_ctx().launches;
  return 0;
}

int kzgpu_field_op(int curve, int which, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  KZ_REQUIRE_INIT();
  if (!a || !out || (op != 3 && !b) || op < 0 || op > 3) return kz_fail(KZGPU_EINVAL, "bad argument");
  if (n == 0) return 0;
  if (curve == KZGPU_BN254) return which == 0 ? field_op_impl<FpBN254>(op, a, b, out, n) : field_op_impl<FrBN254>(op, a, b, out, n);
  if (curve == KZGPU_BLS12_381) return which == 0 ? field_op_impl<FpBLS381>(op, a, b, out, n) : field_op_impl<FrBLS381>(op, a, b, out, n);
  return kz_fail(KZGPU_EINVAL, "unknown curve id %d", curve);
}

}  // extern "C"

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
  KzgpuCtx& cx = kz_ctx();
  const size_t npairs = (size_t)blocks * threads * iters;
  if (npairs >= (1ull << 31)) return kz_fail(KZGPU_EINVAL, "too many pairs");
  const size_t table_pts = gather ? (1ull << 27) : 2 * npairs;         // 8 GiB table for the gather probe
  uint32_t *pts = nullptr, *idx = nullptr, *pre = nullptr, *out = nullptr;
  KZ_CUDA(cudaMalloc(&pts, table_pts * 64));
  KZ_CUDA(cudaMalloc(&pre, npairs * 32));
  KZ_CUDA(cudaMalloc(&out, npairs * 64));
  mb_fill_kernel<<<(unsigned)kz_div_up(table_pts * 16, 256), 256, 0, cx.stream>>>(pts, table_pts * 16, 99u);
  if (gather) {
    KZ_CUDA(cudaMalloc(&idx, 2 * npairs * 4));
    mb_index_kernel<<<(unsigned)kz_div_up(2 * npairs, 256), 256, 0, cx.stream>>>(idx, 2 * npairs, (uint32_t)(table_pts - 1), 7u);
  }
  for (int rep = 0; rep < 2; rep++) {
    KZ_CUDA(cudaEventRecord(cx.ev0, cx.stream));
    if (gather) mb_affine_pairs_kernel<FpBN254, true><<<blocks, threads, 0, cx.stream>>>(pts, idx, (uint32_t)npairs, pre, out);
    else mb_affine_pairs_kernel<FpBN254, false><<<blocks, threads, 0, cx.stream>>>(pts, idx, (uint32_t)npairs, pre, out);
    KZ_LAUNCHED();
    KZ_CUDA(cudaEventRecord(cx.ev1, cx.stream));
    KZ_CUDA(cudaEventSynchronize(cx.ev1));
    KZ_CUDA(cudaEventElapsedTime(ms, cx.ev0, cx.ev1));
  }
  if (ops) *ops = (double)npairs;
  cudaFree(pts); cudaFree(idx); cudaFree(pre); cudaFree(out);
  return 0;
}

int kzgpu_microbench(int kind, int blocks, int threads, int iters, float* ms, double* ops) {
  KZ_REQUIRE_INIT();
  if (kind == 12 || kind == 13) {
    if (blocks <= 0 || threads <= 0 || threads > 128 || iters <= 0 || !ms) return kz_fail(KZGPU_EINVAL, "bad argument");
    return microbench_affine(kind == 13, blocks, threads, iters, ms, ops);
  }
  if (blocks <= 0 || threads <= 0 || threads > 1024 || iters <= 0 || !ms) return kz_fail(KZGPU_EINVAL, "bad argument");
  KzgpuCtx& cx = kz_ctx();
  uint32_t* sink = nullptr;
  KZ_CUDA(cudaMalloc(&sink, 4));
  double per_thread = 0;
  for (int rep = 0; rep < 2; rep++) {       // rep 0 = warm-up
    KZ_CUDA(cudaEventRecord(cx.ev0, cx.stream));
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
      default: cudaFree(sink); return kz_fail(KZGPU_EINVAL, "unknown microbench kind %d", kind);
    }
    KZ_LAUNCHED();
    KZ_CUDA(cudaEventRecord(cx.ev1, cx.stream));
    KZ_CUDA(cudaEventSynchronize(cx.ev1));
    KZ_CUDA(cudaEventElapsedTime(ms, cx.ev0, cx.ev1));
  }
  if (ops) *ops = per_thread * (double)blocks * (double)threads;
  cudaFree(sink);
  return 0;
}

}  // extern "C"
