// Library context shared by the translation units of libkzgpu.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <functional>
#include "../../include/kzgpu.h"
#include "params_gen.cuh"
#include "curve.cuh"

// Devices: the library is initialised on 1 .. KZ_MAX_DEV devices (kzgpu_init / kzgpu_init_multi).  Every device has a
// SLOT with its own context (streams, events, error text) and its own workspaces in each translation unit.  Slot 0 is the
// primary device and belongs to the caller's thread; slots >= 1 each belong to a worker thread of context.cu that has
// made its device current once and for all.  kz_slot() is the calling thread's slot, so the single-device code paths run
// unchanged on any slot; multi-device entry points fan out with kz_parallel().
constexpr int KZ_MAX_DEV = 16;
int kz_slot();
int kz_ndev();
int kz_device_of(int slot);
// fn(slot) on every slot < ndev concurrently (slot 0 on the calling thread); returns the first non-zero code, whose message
// is copied into slot 0's error text
int kz_parallel(const std::function<int(int)>& fn);

struct KzgpuCtx {
  bool inited = false;
  int device = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t own_stream = nullptr;
  cudaStream_t copy_stream = nullptr;     // chunked host->device uploads overlapped with compute
  cudaStream_t d2h_stream = nullptr;      // batched NTTs: download of vector k while vector k+1 uploads (full duplex)
  cudaEvent_t copy_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaStream_t sort_stream = nullptr;     // MSM: sort of chunk k+1 (HBM/LSU-bound) overlapped with the accumulate of chunk k (multiplier-bound)
  cudaEvent_t sort_ev[2] = {nullptr, nullptr}, acc_ev[2] = {nullptr, nullptr}, start_ev = nullptr;
  uint64_t launches = 0;
  char err[512] = {0};
  // per-kernel profiling (bench.py roofline): enabled -> events around selected launches
  bool profile = false;
  double prof_ms[4] = {0, 0, 0, 0};
  uint64_t prof_launches[4] = {0, 0, 0, 0};
  double prof_work[4] = {0, 0, 0, 0};
};

KzgpuCtx& kz_ctx();                        // context of the calling thread's slot
KzgpuCtx& kz_ctx_of(int slot);
int kz_fail(int code, const char* fmt, ...);

// Host <-> device copies of caller buffers that may be PAGEABLE (the Python drop-in hands over ordinary numpy arrays).
// Pinned memory goes straight to cudaMemcpyAsync on `stream`.  Pageable memory of >= 8 MiB is staged by a few host threads
// through page-locked double buffers (4 MiB pieces) instead of the driver's single bounce buffer: the call returns when
// every piece has been ISSUED on `stream` (upload) / has ARRIVED in dst (download).  context.cu.
int kz_upload(void* d_dst, const void* h_src, size_t bytes, cudaStream_t stream);
int kz_download(void* h_dst, const void* d_src, size_t bytes, cudaStream_t stream);
bool kz_host_is_pinned(const void* p);

// scoped CUDA-event timer around one or more launches of a profiled kernel class
struct KzProf {
  int which;
  bool on;
  cudaEvent_t a = nullptr, b = nullptr;
  KzProf(int which_) : which(which_), on(kz_ctx().profile) {
    if (on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, kz_ctx().stream); }
  }
  void stop(uint64_t launches, double work) {
    if (!on) return;
    cudaEventRecord(b, kz_ctx().stream);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    KzgpuCtx& cx = kz_ctx();
    cx.prof_ms[which] += ms; cx.prof_launches[which] += launches; cx.prof_work[which] += work;
    cudaEventDestroy(a); cudaEventDestroy(b);
    on = false;
  }
  ~KzProf() { if (on) { cudaEventDestroy(a); cudaEventDestroy(b); } }
};

#define KZ_REQUIRE_INIT() \
  do { if (!kz_ctx().inited) return kz_fail(KZGPU_ENOTINIT, "kzgpu_init has not been called"); } while (0)

// per-slot instance of a translation unit's global state
template <class T> struct KzPerSlot {
  T v[KZ_MAX_DEV];
  T& get() { return v[kz_slot()]; }
  T& of(int slot) { return v[slot]; }
};

#define KZ_CUDA(expr) \
  do { cudaError_t e__ = (expr); \
       if (e__ != cudaSuccess) return kz_fail(KZGPU_ECUDA, "%s failed: %s (%s:%d)", #expr, \
                                              cudaGetErrorString(e__), __FILE__, __LINE__); } while (0)

// after a kernel launch: count it and surface launch-configuration errors
#define KZ_LAUNCHED() \
  do { kz_ctx().launches++; KZ_CUDA(cudaGetLastError()); } while (0)

// device scratch that grows on demand and is reused between calls
struct KzScratch {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes);
  void release();
};

static inline size_t kz_div_up(size_t a, size_t b) { return (a + b - 1) / b; }

// 64-bit-limb canonical <-> 32-bit-limb Fe (little endian: identical byte layout)
template <class P> static inline Fe<P> kz_fe_from_u64(const uint64_t* w) {
  Fe<P> r;
  for (int i = 0; i < P::N; i++) r.v[i] = (uint32_t)(w[i / 2] >> (32 * (i & 1)));
  return r;
}
template <class P> static inline void kz_fe_to_u64(const Fe<P>& a, uint64_t* w) {
  for (int i = 0; i < P::N / 2; i++) w[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
}
template <class P> static inline bool kz_fe_reduced(const Fe<P>& a) {
  for (int i = P::N - 1; i >= 0; i--) {
    if (a.v[i] < P::mod(i)) return true;
    if (a.v[i] > P::mod(i)) return false;
  }
  return false;   // equal to the modulus
}
