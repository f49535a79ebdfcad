// Library context (one slot per device, worker threads for the devices beyond the first), device-memory plumbing and the
// field self-test.
#include "common.cuh"
#include <cstring>
#include <cctype>
#include <unistd.h>
#include <sys/syscall.h>
#include <thread>
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <condition_variable>

int kz_msm_pending_check();
void kz_ntt_release();
void kz_msm_release();
void kz_poly_release();
void kz_plonk_release();

// ---------------------------------------------------------------- device slots and their worker threads
namespace {
KzgpuCtx g_ctx[KZ_MAX_DEV];
int g_ndev = 0;
thread_local int tl_slot = 0;

// One worker thread per slot >= 1.  It makes its device current once, then runs the jobs kz_parallel hands it.
struct KzWorker {
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  const std::function<int(int)>* job = nullptr;
  int rc = 0;
  bool done = true, quit = false;
};
KzWorker* g_workers[KZ_MAX_DEV] = {nullptr};

void worker_main(int slot) {
  tl_slot = slot;
  cudaSetDevice(g_ctx[slot].device);
  KzWorker& w = *g_workers[slot];
  std::unique_lock<std::mutex> lk(w.m);
  for (;;) {
    w.cv.wait(lk, [&] { return w.job != nullptr || w.quit; });
    if (w.quit) return;
    const std::function<int(int)>* job = w.job;
    lk.unlock();
    int rc = (*job)(slot);
    lk.lock();
    w.rc = rc; w.job = nullptr; w.done = true;
    w.cv.notify_all();
  }
}
}  // namespace

int kz_slot() { return tl_slot; }
int kz_ndev() { return g_ndev; }
int kz_device_of(int slot) { return g_ctx[slot].device; }
KzgpuCtx& kz_ctx() { return g_ctx[tl_slot]; }
KzgpuCtx& kz_ctx_of(int slot) { return g_ctx[slot]; }

int kz_parallel(const std::function<int(int)>& fn) {
  if (tl_slot != 0) return kz_fail(KZGPU_EINVAL, "kz_parallel from a worker thread");
  for (int s = 1; s < g_ndev; s++) {
    KzWorker& w = *g_workers[s];
    std::lock_guard<std::mutex> lk(w.m);
    w.job = &fn; w.done = false;
    w.cv.notify_all();
  }
  int rc = fn(0);
  for (int s = 1; s < g_ndev; s++) {
    KzWorker& w = *g_workers[s];
    std::unique_lock<std::mutex> lk(w.m);
    w.cv.wait(lk, [&] { return w.done; });
    if (w.rc && !rc) {
      rc = w.rc;
      memcpy(g_ctx[0].err, g_ctx[s].err, sizeof(g_ctx[0].err));
    }
  }
  return rc;
}

// ---------------------------------------------------------------- staged copies of pageable host memory
namespace {
constexpr size_t kPiece = 4u << 20;          // bytes per staged piece
constexpr int kStageThreads = 4;             // host threads per copy (the caller is one of them)
struct KzStage {                             // per slot: two page-locked pieces and their events per thread
  char* buf[kStageThreads][2] = {{nullptr}};
  cudaEvent_t ev[kStageThreads][2] = {{nullptr}};
  bool ready = false;
};
KzStage g_stage[KZ_MAX_DEV];

int stage_ready(KzStage& st) {
  if (st.ready) return 0;
  for (int t = 0; t < kStageThreads; t++)
    for (int k = 0; k < 2; k++) {
      KZ_CUDA(cudaHostAlloc((void**)&st.buf[t][k], kPiece, cudaHostAllocDefault));
      KZ_CUDA(cudaEventCreateWithFlags(&st.ev[t][k], cudaEventDisableTiming));
    }
  st.ready = true;
  return 0;
}

void stage_release(KzStage& st) {
  if (!st.ready) return;
  for (int t = 0; t < kStageThreads; t++)
    for (int k = 0; k < 2; k++) { cudaFreeHost(st.buf[t][k]); cudaEventDestroy(st.ev[t][k]); st.buf[t][k] = nullptr; }
  st.ready = false;
}

bool staging_wanted(const void* h, size_t bytes) {
  static const bool off = getenv("KZGPU_NO_STAGING") != nullptr;
  return !off && bytes >= (8u << 20) && !kz_host_is_pinned(h);
}
}  // namespace

bool kz_host_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

int kz_upload(void* d_dst, const void* h_src, size_t bytes, cudaStream_t stream) {
  if (!bytes) return 0;
  if (!staging_wanted(h_src, bytes)) {
    KZ_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, stream));
    return 0;
  }
  KzStage& st = g_stage[kz_slot()];
  int rc = stage_ready(st);
  if (rc) return rc;
  const size_t pieces = (bytes + kPiece - 1) / kPiece;
  const int device = kz_ctx().device;
  std::atomic<size_t> next{0};
  std::atomic<int> err{0};
  auto work = [&](int t) {
    cudaSetDevice(device);
    int k = 0;
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= pieces || err.load()) break;
      const size_t off = i * kPiece, len = off + kPiece <= bytes ? kPiece : bytes - off;
      cudaEventSynchronize(st.ev[t][k]);                            // this thread's previous use of the buffer has left the host
      memcpy(st.buf[t][k], (const char*)h_src + off, len);
      cudaError_t e = cudaMemcpyAsync((char*)d_dst + off, st.buf[t][k], len, cudaMemcpyHostToDevice, stream);
      if (e == cudaSuccess) e = cudaEventRecord(st.ev[t][k], stream);
      if (e != cudaSuccess) err.store((int)e);
      k ^= 1;
    }
  };
  std::thread th[kStageThreads - 1];
  for (int t = 1; t < kStageThreads; t++) th[t - 1] = std::thread(work, t);
  work(0);
  for (auto& x : th) x.join();
  if (err.load()) return kz_fail(KZGPU_ECUDA, "staged upload failed: %s", cudaGetErrorString((cudaError_t)err.load()));
  return 0;
}

int kz_download(void* h_dst, const void* d_src, size_t bytes, cudaStream_t stream) {
  if (!bytes) return 0;
  if (!staging_wanted(h_dst, bytes)) {
    KZ_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, stream));
    return 0;
  }
  KzStage& st = g_stage[kz_slot()];
  int rc = stage_ready(st);
  if (rc) return rc;
  const size_t pieces = (bytes + kPiece - 1) / kPiece;
  const int device = kz_ctx().device;
  std::atomic<size_t> next{0};
  std::atomic<int> err{0};
  auto work = [&](int t) {
    cudaSetDevice(device);
    // two pieces in flight per thread: piece b is copied out of its buffer while piece b^1 is still arriving
    size_t cur[2] = {(size_t)-1, (size_t)-1};
    int k = 0;
    auto drain = [&](int b) {
      if (cur[b] == (size_t)-1) return;
      cudaEventSynchronize(st.ev[t][b]);
      const size_t off = cur[b] * kPiece, len = off + kPiece <= bytes ? kPiece : bytes - off;
      memcpy((char*)h_dst + off, st.buf[t][b], len);
      cur[b] = (size_t)-1;
    };
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= pieces || err.load()) break;
      drain(k);
      const size_t off = i * kPiece, len = off + kPiece <= bytes ? kPiece : bytes - off;
      cudaError_t e = cudaMemcpyAsync(st.buf[t][k], (const char*)d_src + off, len, cudaMemcpyDeviceToHost, stream);
      if (e == cudaSuccess) e = cudaEventRecord(st.ev[t][k], stream);
      if (e != cudaSuccess) { err.store((int)e); break; }
      cur[k] = i;
      k ^= 1;
    }
    drain(k); drain(k ^ 1);
  };
  std::thread th[kStageThreads - 1];
  for (int t = 1; t < kStageThreads; t++) th[t - 1] = std::thread(work, t);
  work(0);
  for (auto& x : th) x.join();
  if (err.load()) return kz_fail(KZGPU_ECUDA, "staged download failed: %s", cudaGetErrorString((cudaError_t)err.load()));
  return 0;
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

}  // namespace

extern "C" {

// streams and events of one slot; the slot's device must be current
static int slot_create(int slot, int device) {
  KzgpuCtx& cx = g_ctx[slot];
  cudaDeviceProp prop;
  KZ_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return kz_fail(KZGPU_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  cx.sm_count = prop.multiProcessorCount;
  KZ_CUDA(cudaStreamCreateWithFlags(&cx.own_stream, cudaStreamNonBlocking));
  cx.stream = cx.own_stream;
  KZ_CUDA(cudaStreamCreateWithFlags(&cx.copy_stream, cudaStreamNonBlocking));
  KZ_CUDA(cudaStreamCreateWithFlags(&cx.d2h_stream, cudaStreamNonBlocking));
  for (int k = 0; k < 4; k++) KZ_CUDA(cudaEventCreateWithFlags(&cx.copy_ev[k], cudaEventDisableTiming));
  int lo_prio = 0, hi_prio = 0;
  KZ_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
  KZ_CUDA(cudaStreamCreateWithPriority(&cx.sort_stream, cudaStreamNonBlocking, hi_prio));
  for (int k = 0; k < 2; k++) {
    KZ_CUDA(cudaEventCreateWithFlags(&cx.sort_ev[k], cudaEventDisableTiming));
    KZ_CUDA(cudaEventCreateWithFlags(&cx.acc_ev[k], cudaEventDisableTiming));
  }
  KZ_CUDA(cudaEventCreateWithFlags(&cx.start_ev, cudaEventDisableTiming));
  KZ_CUDA(cudaEventCreate(&cx.ev0));
  KZ_CUDA(cudaEventCreate(&cx.ev1));
  cx.device = device;
  cx.inited = true;
  cx.launches = 0;
  return 0;
}

static void slot_destroy(int slot) {
  KzgpuCtx& cx = g_ctx[slot];
  if (!cx.inited) return;
  cudaStreamSynchronize(cx.stream);
  stage_release(g_stage[slot]);
  cudaEventDestroy(cx.ev0);
  cudaEventDestroy(cx.ev1);
  cudaStreamDestroy(cx.own_stream);
  cudaStreamDestroy(cx.copy_stream);
  cudaStreamDestroy(cx.d2h_stream);
  for (int k = 0; k < 4; k++) cudaEventDestroy(cx.copy_ev[k]);
  cudaStreamDestroy(cx.sort_stream);
  for (int k = 0; k < 2; k++) { cudaEventDestroy(cx.sort_ev[k]); cudaEventDestroy(cx.acc_ev[k]); }
  cudaEventDestroy(cx.start_ev);
  cx = KzgpuCtx();
}

int kzgpu_init_multi(int ndev, const int* devs) {
  int count = 0;
  KZ_CUDA(cudaGetDeviceCount(&count));
  if (ndev <= 0) ndev = count;
  if (ndev > KZ_MAX_DEV || ndev > count) return kz_fail(KZGPU_EINVAL, "%d devices requested, %d visible (limit %d)", ndev, count, KZ_MAX_DEV);
  int list[KZ_MAX_DEV];
  for (int i = 0; i < ndev; i++) {
    list[i] = devs ? devs[i] : i;
    if (list[i] < 0 || list[i] >= count) return kz_fail(KZGPU_EINVAL, "device %d out of range (%d devices)", list[i], count);
    for (int j = 0; j < i; j++) if (list[j] == list[i]) return kz_fail(KZGPU_EINVAL, "device %d listed twice", list[i]);
  }
  if (g_ndev) {                                  // idempotent for the same device list
    bool same = g_ndev == ndev;
    for (int i = 0; same && i < ndev; i++) same = g_ctx[i].device == list[i];
    if (same) return 0;
    return kz_fail(KZGPU_EINVAL, "already initialised on %d device(s), first = %d", g_ndev, g_ctx[0].device);
  }
  tl_slot = 0;
  for (int i = 0; i < ndev; i++) {
    KZ_CUDA(cudaSetDevice(list[i]));
    int rc = slot_create(i, list[i]);
    if (rc) {
      if (i) memcpy(g_ctx[0].err, g_ctx[i].err, sizeof(g_ctx[0].err));
      for (int j = 0; j <= i; j++) { cudaSetDevice(list[j]); slot_destroy(j); }
      return rc;
    }
  }
  // peer access in both directions between every pair (NVLink / NVSwitch); without it cudaMemcpyPeer stages through the host
  for (int i = 0; i < ndev; i++) {
    cudaSetDevice(list[i]);
    for (int j = 0; j < ndev; j++) {
      int ok = 0;
      if (i != j && cudaDeviceCanAccessPeer(&ok, list[i], list[j]) == cudaSuccess && ok) cudaDeviceEnablePeerAccess(list[j], 0);
    }
  }
  cudaGetLastError();                            // "peer access already enabled" is not an error
  KZ_CUDA(cudaSetDevice(list[0]));
  g_ndev = ndev;
  for (int i = 1; i < ndev; i++) {
    g_workers[i] = new KzWorker();
    g_workers[i]->th = std::thread(worker_main, i);
  }
  return 0;
}

int kzgpu_init(int device) {
  if (g_ndev) {
    if (device == g_ctx[0].device) return 0;
    return kz_fail(KZGPU_EINVAL, "already initialised on device %d", g_ctx[0].device);
  }
  return kzgpu_init_multi(1, &device);
}

int kzgpu_device_count(int* ndev) {
  if (!ndev) return KZGPU_EINVAL;
  *ndev = g_ndev;
  return 0;
}

int kzgpu_shutdown(void) {
  if (!g_ndev) return 0;
  // every slot releases its own workspaces on its own device (its thread has that device current)
  kz_parallel([](int) {
    cudaStreamSynchronize(kz_ctx().stream);
    kz_ntt_release();
    kz_msm_release();
    kz_poly_release();
    kz_plonk_release();
    return 0;
  });
  for (int i = 1; i < g_ndev; i++) {
    KzWorker* w = g_workers[i];
    { std::lock_guard<std::mutex> lk(w->m); w->quit = true; w->cv.notify_all(); }
    w->th.join();
    delete w;
    g_workers[i] = nullptr;
  }
  for (int i = g_ndev - 1; i >= 0; i--) { cudaSetDevice(g_ctx[i].device); slot_destroy(i); }
  g_ndev = 0;
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
  int rc = kz_upload(d_dst, src, bytes, kz_ctx().stream);
  if (rc) return rc;
  KZ_CUDA(cudaStreamSynchronize(kz_ctx().stream));
  return 0;
}

int kzgpu_d2h(void* dst, const void* d_src, size_t bytes) {
  KZ_REQUIRE_INIT();
  int rc = kz_download(dst, d_src, bytes, kz_ctx().stream);
  if (rc) return rc;
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
  int rc = kz_msm_pending_check();             // a kzgpu_msm_partial* that returned before its flag was read
  if (rc) return rc;
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
  uint64_t sum = 0;
  for (int i = 0; i < (g_ndev ? g_ndev : 1); i++) sum += g_ctx[i].launches;
  *count = sum;
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

