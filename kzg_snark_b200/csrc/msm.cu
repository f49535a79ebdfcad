// G1 multi-scalar multiplication: the device replacement of KZG.commit's inner loop
// (reference kzg.py:108-116:  commitment = sum_i coeff_i * ck[i], zero coefficients skipped),
// plus the SRS store behind `ck` (kzg.py:69-72).
//
// Pippenger bucket method:
//   1. every scalar is recoded into W signed c-bit digits d_w in [-2^(c-1), 2^(c-1)]
//      (zero digits are skipped, as the reference skips zero coefficients, kzg.py:113);
//   2. counting sort of (window, |d|-1) -> per-bucket lists of point indices (sign in bit 31):
//      histogram with L2 atomics, exclusive scan, scatter;
//   3. one thread per bucket accumulates its points with XYZZ mixed additions (8M+2S);
//   4. per window, sum_b (b+1) * B_b by chunked running sums, then a block reduction;
//   5. Horner over the windows (c doublings each), normalisation to affine, canonical limbs.
// Only step 3 is O(n * W); it is integer-pipe (IMAD) bound, see DESIGN.md.
#include "common.cuh"
#include <map>
#include <vector>
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace {

struct BN254Cfg { using Fp = FpBN254; using Fr = FrBN254; static constexpr int id = KZGPU_BN254; };
struct BLS381Cfg { using Fp = FpBLS381; using Fr = FrBLS381; static constexpr int id = KZGPU_BLS12_381; };

// The commitment key is fixed for the life of a prover, so by default it is stored as W window
// tables T_w[i] = 2^(c*w) * P_i (w < W; T_0 is the key itself): every signed digit of every
// scalar then lands in ONE shared set of 2^(c-1) buckets and the MSM needs neither the
// per-window bucket sets nor the final Horner chain of c*W doublings.  Costs W x the key's
// footprint (12 GiB for 2^24 BN254 points at c = 22) -- what the 180 GB of HBM3e is for -- and a
// one-off build.  c_tab == 0: plain key (tables disabled or too large), per-call window choice.
struct SrsPart {          // one device's copy of the index range [first, first + n) of a key
  uint32_t* d_points = nullptr;     // affine, Montgomery form, [x | y] per point; W_tab tables of n points
  size_t first = 0, n = 0;
  uint32_t c_tab = 0, W_tab = 1;
  int slot = 0;
};
struct Srs {
  int curve;
  size_t n;
  SrsPart full;                      // the whole key on the primary device (slot 0)
  // multi-device layouts, built on first use from `full` (peer copy of the plain points, local table build):
  bool has_replicas = false;
  // shard sets by length class E (a power of two, or the key's length): slot d owns [d E / N, (d + 1) E / N) of the key with
  // tables sized for THAT range, so a 2^20-coefficient polynomial committed against a 12 M-point key is spread over all
  // devices (not just the ones owning the head of the key) and each runs the window a 2^17-point MSM wants
  std::map<size_t, std::vector<SrsPart>>* shard_sets = nullptr;
  SrsPart replica[KZ_MAX_DEV];       // the whole key, tables included, on every slot (replica[0] aliases full)
};

std::map<uint64_t, Srs> g_srs;
uint64_t g_next_handle = 1;

struct MsmWs {                         // shared by all chunks of a call
  KzScratch buckets, partials, winsums, result, flag, scal, gather;
};
KzPerSlot<MsmWs> g_ws_slots;
#define g_ws (g_ws_slots.get())
// what the sort + task phase of one chunk produces and its accumulate / merge consumes: two sets, so that the
// sort of chunk k+1 (sort stream) runs while chunk k is accumulated (main stream)
struct MsmSortWs {
  KzScratch counts, offsets, cursor, entries, blocksums;
  KzScratch ntasks, task_off, size_hist, t_start, t_len, t_dest, tparts, multi, sort_tmp, coarse;
};
struct MsmSortPair { MsmSortWs w[2]; };
KzPerSlot<MsmSortPair> g_sw_slots;
#define g_sw (g_sw_slots.get().w)

// ---------------------------------------------------------------- device helpers
template <class P> __device__ __forceinline__ void ld_words(uint32_t* dst, const uint32_t* src) {
  const uint4* q = reinterpret_cast<const uint4*>(src);
#pragma unroll
  for (int i = 0; i < P::N / 4; i++) {
    uint4 t = __ldg(q + i);
    dst[4 * i] = t.x; dst[4 * i + 1] = t.y; dst[4 * i + 2] = t.z; dst[4 * i + 3] = t.w;
  }
}
template <class P> __device__ __forceinline__ Affine<P> ld_affine(const uint32_t* pts, size_t idx) {
  Affine<P> a;
  const uint32_t* p = pts + idx * (2 * P::N);
  ld_words<P>(a.x.v, p);
  ld_words<P>(a.y.v, p + P::N);
  return a;
}
template <class P> __device__ __forceinline__ XYZZ<P> ld_xyzz(const uint32_t* buf, size_t idx) {
  XYZZ<P> a;
  const uint4* q = reinterpret_cast<const uint4*>(buf + idx * (4 * P::N));
  uint32_t* d = a.x.v;     // XYZZ is 4 contiguous Fe
#pragma unroll
  for (int i = 0; i < P::N; i++) {
    uint4 t = q[i];
    d[4 * i] = t.x; d[4 * i + 1] = t.y; d[4 * i + 2] = t.z; d[4 * i + 3] = t.w;
  }
  return a;
}
template <class P> __device__ __forceinline__ void st_xyzz(uint32_t* buf, size_t idx, const XYZZ<P>& a) {
  uint4* q = reinterpret_cast<uint4*>(buf + idx * (4 * P::N));
  const uint32_t* d = a.x.v;
#pragma unroll
  for (int i = 0; i < P::N; i++) q[i] = make_uint4(d[4 * i], d[4 * i + 1], d[4 * i + 2], d[4 * i + 3]);
}

// Signed-window recoding without a serial carry: add 2^(c-1) to every window below the top one
// once (s' = s + OFF), then window w of s' minus 2^(c-1) is a digit in [-2^(c-1), 2^(c-1)) and
// the top window (unsigned, it absorbs the last carry) is at most 2^(c-1) because the scalar has
// at most c*W - 1 bits.  Each digit then depends on s' alone, so windows can be processed in
// any order (the scatter pass below is window-major).
struct DigitOffset { uint32_t w[8]; };

__device__ __forceinline__ void load_scalar_plus_offset(const uint32_t* __restrict__ scalars, size_t i, const DigitOffset& off,
                                                        uint32_t* s) {
  const uint4* q = reinterpret_cast<const uint4*>(scalars + i * 8);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  uint32_t t[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  MpPrims<8>::add_cc(s, t, off.w);
}

__device__ __forceinline__ int signed_digit(const uint32_t* s, uint32_t w, uint32_t c, uint32_t W) {
  uint32_t bit = w * c;
  uint32_t word = bit >> 5, sh = bit & 31;
  uint64_t lo = word < 8 ? s[word] : 0u;
  uint64_t hi = word + 1 < 8 ? s[word + 1] : 0u;
  uint32_t raw = (uint32_t)(((lo | (hi << 32)) >> sh) & ((1ull << c) - 1));
  return w + 1 < W ? (int)raw - (int)(1u << (c - 1)) : (int)raw;
}

// ---------------------------------------------------------------- kernels
// Bucket sort of the n*W (key, value) digit entries, key = bucket (plain key: w * B + bucket),
// value = index of the point to add (tabled: w * n_srs + first + i, in table w) with the digit's
// sign in bit 31.  Two-level MSD counting sort; every atomic that is hit once per entry lives in
// shared memory, global atomics are one per (block, bin):
//   pass 1 (coarse = key >> f, <= 4096 bins): coarse histogram -> scan -> partition into tmp[] as
//          64-bit (key << 32 | value) records, block-aggregated range reservation;
//   pass 2 (fine = key & (2^f - 1)), one block per <= kSortChunk records of ONE coarse bin:
//          fine histogram -> (global scan over all keys) -> scatter of the 32-bit values.
// A block of pass 2 writes into the <= 2^f bucket lists of its coarse bin, a window of a few
// hundred KB, so the 4-byte stores merge into full sectors in L2.  HBM traffic: scalars 2 x 32 B
// per point, records 8 B written + 2 x 8 B read per entry, values 4 B written per entry.
constexpr uint32_t kSortTile = 2048;        // scalars per block in the coarse histogram
constexpr uint32_t kSortStage = 11264;      // records staged per block in the partition (88 KB)
constexpr uint32_t kSortChunk = 16384;      // records per block in pass 2 (64 KB staged)
constexpr uint32_t kMaxCoarse = 4096;

struct SortGeom { uint32_t c, W, B, f, ncoarse, tabled, first, n_srs, top_bits, batch_len; };   // batch_len: scalars per polynomial of a batched pass (0 = one MSM)

__device__ __forceinline__ bool digit_entry(const uint32_t* s, uint32_t w, const SortGeom& g, uint32_t i, uint32_t& key, uint32_t& val) {
  int d = signed_digit(s, w, g.c, g.W);
  if (d == 0) return false;
  uint32_t neg = d < 0 ? 1u : 0u;
  uint32_t mag = neg ? (uint32_t)(-d) : (uint32_t)d;
  // batched pass: scalar i belongs to polynomial j = i / batch_len, whose buckets form their own set(s)
  uint32_t j = 0;
  if (g.batch_len) { j = i / g.batch_len; i -= j * g.batch_len; }
  if (g.tabled) { key = j * g.B + (mag - 1); val = (w * g.n_srs + g.first + i) | (neg << 31); }
  else { key = (j * g.W + w) * g.B + (mag - 1); val = (g.first + i) | (neg << 31); }
  return true;
}

// pass 1a: coarse histogram (block-local in shared memory, one global add per non-empty bin)
__global__ void __launch_bounds__(256) msm_coarse_hist_kernel(const uint32_t* __restrict__ scalars, size_t n, SortGeom g, DigitOffset off,
                                                              uint32_t* __restrict__ coarse_counts, uint32_t* __restrict__ flag) {
  extern __shared__ uint32_t sh[];
  for (uint32_t b = threadIdx.x; b < g.ncoarse; b += blockDim.x) sh[b] = 0;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * kSortTile;
  for (uint32_t k = threadIdx.x; k < kSortTile; k += blockDim.x) {
    size_t i = base + k;
    if (i >= n) break;
    uint32_t s[8];
    load_scalar_plus_offset(scalars, i, off, s);
    // canonical scalars only: s + OFF must stay below 2^(c*W) <=> the raw scalar has <= BITS bits
    {
      const uint4* q = reinterpret_cast<const uint4*>(scalars + i * 8);
      uint32_t top = __ldg(q + 1).w;
      if (g.top_bits < 32 && (top >> g.top_bits)) atomicOr(flag, 1u);
    }
    for (uint32_t w = 0; w < g.W; w++) {
      uint32_t key, val;
      if (digit_entry(s, w, g, (uint32_t)i, key, val)) atomicAdd(&sh[key >> g.f], 1u);
    }
  }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < g.ncoarse; b += blockDim.x)
    if (sh[b]) atomicAdd(&coarse_counts[b], sh[b]);
}

// pass 1b: scan of the coarse counts (single block): coarse_off[0..ncoarse], write cursors, and
// the block table of pass 2 (blk_off[b] = first block of coarse bin b, kSortChunk records each)
__global__ void __launch_bounds__(1024) msm_coarse_scan_kernel(const uint32_t* __restrict__ coarse_counts, uint32_t ncoarse,
                                                               uint32_t* __restrict__ coarse_off, uint32_t* __restrict__ coarse_cur,
                                                               uint32_t* __restrict__ blk_off) {
  __shared__ uint32_t sa[1024], sb[1024];
  // 4 bins per thread (ncoarse <= 4096)
  uint32_t v[4], u[4], sv = 0, su = 0;
  for (int k = 0; k < 4; k++) {
    uint32_t b = threadIdx.x * 4 + k;
    v[k] = b < ncoarse ? coarse_counts[b] : 0;
    u[k] = (v[k] + kSortChunk - 1) / kSortChunk;
    sv += v[k]; su += u[k];
  }
  sa[threadIdx.x] = sv; sb[threadIdx.x] = su;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    uint32_t ta = threadIdx.x >= o ? sa[threadIdx.x - o] : 0, tb = threadIdx.x >= o ? sb[threadIdx.x - o] : 0;
    __syncthreads();
    sa[threadIdx.x] += ta; sb[threadIdx.x] += tb;
    __syncthreads();
  }
  uint32_t ea = sa[threadIdx.x] - sv, eb = sb[threadIdx.x] - su;
  for (int k = 0; k < 4; k++) {
    uint32_t b = threadIdx.x * 4 + k;
    if (b < ncoarse) { coarse_off[b] = ea; coarse_cur[b] = ea; blk_off[b] = eb; }
    ea += v[k]; eb += u[k];
  }
  if (threadIdx.x == 1023) { coarse_off[ncoarse] = sa[1023]; blk_off[ncoarse] = sb[1023]; }
}

// exclusive scan of arr[0..nbins) in shared memory, in place; returns the total.
// scratch: blockDim.x words.  Every thread of the block must call it.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t* arr, uint32_t nbins, uint32_t* scratch) {
  const uint32_t per = (nbins + blockDim.x - 1) / blockDim.x;
  const uint32_t lo = threadIdx.x * per, hi = lo + per < nbins ? lo + per : nbins;
  uint32_t sum = 0;
  for (uint32_t b = lo; b < hi; b++) sum += arr[b];
  scratch[threadIdx.x] = sum;
  __syncthreads();
  for (uint32_t o = 1; o < blockDim.x; o <<= 1) {
    uint32_t t = threadIdx.x >= o ? scratch[threadIdx.x - o] : 0;
    __syncthreads();
    scratch[threadIdx.x] += t;
    __syncthreads();
  }
  uint32_t run = scratch[threadIdx.x] - sum;
  const uint32_t total = scratch[blockDim.x - 1];
  for (uint32_t b = lo; b < hi; b++) { uint32_t v = arr[b]; arr[b] = run; run += v; }
  __syncthreads();
  return total;
}

// pass 1c: partition into coarse bins.  The tile's records are first sorted by bin in shared
// memory (count -> scan -> place), the range of every bin is reserved with ONE global atomic, and
// the staged records are written out in order, so that a bin's run leaves the SM as consecutive
// 8-byte stores within a few instructions (partial sectors are completed while still in L2).
// Shared memory: cur[ncoarse] | delta[ncoarse] | scan scratch[blockDim] | stage[tile * W] (u64)
__global__ void __launch_bounds__(512) msm_partition_kernel(const uint32_t* __restrict__ scalars, size_t n, uint32_t tile, SortGeom g,
                                                            DigitOffset off, uint32_t* __restrict__ coarse_cur,
                                                            uint64_t* __restrict__ tmp) {
  extern __shared__ uint64_t sh64[];
  uint64_t* stage = sh64;
  uint32_t* cur = reinterpret_cast<uint32_t*>(stage + (size_t)tile * g.W);
  uint32_t* delta = cur + g.ncoarse;
  uint32_t* scratch = delta + g.ncoarse;
  for (uint32_t b = threadIdx.x; b < g.ncoarse; b += blockDim.x) cur[b] = 0;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * tile;
  for (uint32_t k = threadIdx.x; k < tile; k += blockDim.x) {
    size_t i = base + k;
    if (i >= n) break;
    uint32_t s[8];
    load_scalar_plus_offset(scalars, i, off, s);
    for (uint32_t w = 0; w < g.W; w++) {
      uint32_t key, val;
      if (digit_entry(s, w, g, (uint32_t)i, key, val)) atomicAdd(&cur[key >> g.f], 1u);
    }
  }
  __syncthreads();
  // reserve the global range of every non-empty bin (count still in cur[]), then local offsets
  for (uint32_t b = threadIdx.x; b < g.ncoarse; b += blockDim.x) {
    uint32_t cnt = cur[b];
    delta[b] = cnt ? atomicAdd(&coarse_cur[b], cnt) : 0u;
  }
  __syncthreads();
  const uint32_t total = block_excl_scan(cur, g.ncoarse, scratch);
  for (uint32_t b = threadIdx.x; b < g.ncoarse; b += blockDim.x) delta[b] -= cur[b];     // global pos = local pos + delta
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < tile; k += blockDim.x) {
    size_t i = base + k;
    if (i >= n) break;
    uint32_t s[8];
    load_scalar_plus_offset(scalars, i, off, s);
    for (uint32_t w = 0; w < g.W; w++) {
      uint32_t key, val;
      if (digit_entry(s, w, g, (uint32_t)i, key, val)) {
        uint32_t pos = atomicAdd(&cur[key >> g.f], 1u);
        stage[pos] = ((uint64_t)key << 32) | val;
      }
    }
  }
  __syncthreads();
  for (uint32_t j = threadIdx.x; j < total; j += blockDim.x) {
    uint64_t rec = stage[j];
    tmp[j + delta[(uint32_t)(rec >> 32) >> g.f]] = rec;
  }
}

// block -> (coarse bin, record range) through the block table
__device__ __forceinline__ bool sort_chunk_range(const uint32_t* __restrict__ coarse_off, const uint32_t* __restrict__ blk_off,
                                                 uint32_t ncoarse, uint32_t& bin, uint32_t& lo, uint32_t& hi) {
  __shared__ uint32_t s_bin;
  if (blockIdx.x >= blk_off[ncoarse]) return false;
  if (threadIdx.x == 0) {
    uint32_t a = 0, b = ncoarse;                 // last bin with blk_off[bin] <= blockIdx.x
    while (b - a > 1) {
      uint32_t m = (a + b) >> 1;
      if (blk_off[m] <= blockIdx.x) a = m; else b = m;
    }
    s_bin = a;
  }
  __syncthreads();
  bin = s_bin;
  uint32_t chunk = blockIdx.x - blk_off[bin];
  lo = coarse_off[bin] + chunk * kSortChunk;
  hi = coarse_off[bin + 1];
  if (hi > lo + kSortChunk) hi = lo + kSortChunk;
  return true;
}

// pass 2a: per-key counts
__global__ void __launch_bounds__(256) msm_fine_hist_kernel(const uint64_t* __restrict__ tmp, const uint32_t* __restrict__ coarse_off,
                                                            const uint32_t* __restrict__ blk_off, uint32_t ncoarse, uint32_t f,
                                                            uint32_t nkeys, uint32_t* __restrict__ counts) {
  extern __shared__ uint32_t sh[];
  uint32_t bin, lo, hi;
  if (!sort_chunk_range(coarse_off, blk_off, ncoarse, bin, lo, hi)) return;
  const uint32_t F = 1u << f;
  for (uint32_t j = threadIdx.x; j < F; j += blockDim.x) sh[j] = 0;
  __syncthreads();
  for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&sh[(uint32_t)(tmp[i] >> 32) & (F - 1)], 1u);
  __syncthreads();
  for (uint32_t j = threadIdx.x; j < F; j += blockDim.x) {
    uint32_t key = (bin << f) + j;
    if (sh[j] && key < nkeys) atomicAdd(&counts[key], sh[j]);
  }
}

// pass 2b: scatter of the values into the per-key lists, staged like pass 1c: the chunk's values
// are sorted by fine key in shared memory and every key's run is written by a group of 8 lanes.
// Shared memory: cur[F] | gpos[F] | scan scratch[blockDim] | stage[kSortChunk] (u32)
__global__ void __launch_bounds__(512) msm_fine_scatter_kernel(const uint64_t* __restrict__ tmp, const uint32_t* __restrict__ coarse_off,
                                                               const uint32_t* __restrict__ blk_off, uint32_t ncoarse, uint32_t f,
                                                               uint32_t nkeys, uint32_t* __restrict__ cursor, uint32_t* __restrict__ entries) {
  extern __shared__ uint32_t sh[];
  uint32_t bin, lo, hi;
  if (!sort_chunk_range(coarse_off, blk_off, ncoarse, bin, lo, hi)) return;
  const uint32_t F = 1u << f;
  uint32_t* cur = sh;
  uint32_t* gpos = cur + F;
  uint32_t* scratch = gpos + F;
  uint32_t* stage = scratch + blockDim.x;
  for (uint32_t j = threadIdx.x; j < F; j += blockDim.x) cur[j] = 0;
  __syncthreads();
  for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&cur[(uint32_t)(tmp[i] >> 32) & (F - 1)], 1u);
  __syncthreads();
  for (uint32_t j = threadIdx.x; j < F; j += blockDim.x) {
    uint32_t key = (bin << f) + j, cnt = cur[j];
    gpos[j] = (cnt && key < nkeys) ? atomicAdd(&cursor[key], cnt) : 0u;
  }
  __syncthreads();
  block_excl_scan(cur, F, scratch);
  for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    uint64_t rec = tmp[i];
    uint32_t pos = atomicAdd(&cur[(uint32_t)(rec >> 32) & (F - 1)], 1u);
    stage[pos] = (uint32_t)rec;
  }
  __syncthreads();
  // cur[j] is now the END of key j's run in stage[]; its start is cur[j-1] (0 for j == 0)
  const uint32_t grp = threadIdx.x >> 3, ln = threadIdx.x & 7, ngrp = blockDim.x >> 3;
  for (uint32_t j = grp; j < F; j += ngrp) {
    uint32_t s0 = j ? cur[j - 1] : 0u, s1 = cur[j], gp = gpos[j];
    for (uint32_t k = s0 + ln; k < s1; k += 8) entries[gp + (k - s0)] = stage[k];
  }
}

// exclusive scan, 3 kernels: (1) per-block scan of 1024 items + block totals, (2) scan of totals,
// (3) add block offsets.  `out` gets len + 1 entries (out[len] = total).
__global__ void scan_block_kernel(const uint32_t* in, uint32_t* out, uint32_t* blocksums, size_t len) {
  __shared__ uint32_t sh[256];
  size_t base = (size_t)blockIdx.x * 1024 + threadIdx.x * 4;
  uint32_t v[4], s = 0;
  for (int k = 0; k < 4; k++) { v[k] = base + k < len ? in[base + k] : 0; s += v[k]; }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int off = 1; off < 256; off <<= 1) {
    uint32_t t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  uint32_t excl = sh[threadIdx.x] - s;
  for (int k = 0; k < 4; k++) { if (base + k < len) out[base + k] = excl; excl += v[k]; }
  if (threadIdx.x == 255) blocksums[blockIdx.x] = sh[255];
}
__global__ void scan_sums_kernel(uint32_t* blocksums, size_t nblocks, uint32_t* total_out) {
  // single block, sequential over chunks of 256
  __shared__ uint32_t sh[256];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (size_t base = 0; base < nblocks; base += 256) {
    size_t i = base + threadIdx.x;
    uint32_t v = i < nblocks ? blocksums[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
      uint32_t t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nblocks) blocksums[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 255) carry += sh[255];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry;
}
__global__ void scan_add_kernel(uint32_t* out, const uint32_t* blocksums, size_t len, uint32_t* cursor) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  uint32_t v = out[i] + blocksums[i >> 10];
  out[i] = v;
  if (cursor) cursor[i] = v;
}

// ---- task construction -------------------------------------------------------------------
// A bucket of `size` points becomes ceil(size / T) tasks of at most T points, so that no thread
// ever walks more than T points (heavy buckets: skewed scalars, short top window).  Tasks are
// then counting-sorted by length, longest first, so that the 32 lanes of a warp run the same
// number of mixed additions (Poisson-distributed bucket sizes would otherwise cost ~25% in
// divergence) and the tail of the grid is made of the shortest tasks.
//
// pass 1: ntasks[b] and the global histogram of task lengths
__global__ void task_count_kernel(const uint32_t* __restrict__ offsets, uint32_t ostride, uint32_t nb, uint32_t T,
                                  uint32_t* __restrict__ ntasks,
                                  uint32_t* __restrict__ size_hist, uint32_t* __restrict__ multi_count,
                                  uint32_t* __restrict__ multi_list) {
  extern __shared__ uint32_t sh_hist[];           // T + 1 bins
  for (uint32_t i = threadIdx.x; i <= T; i += blockDim.x) sh_hist[i] = 0;
  __syncthreads();
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nb) {
    uint32_t size = offsets[(size_t)(b + 1) * ostride] - offsets[(size_t)b * ostride];
    uint32_t full = size / T, rem = size - full * T;
    uint32_t nt = full + (rem ? 1u : 0u);
    ntasks[b] = nt;
    if (nt > 1) multi_list[atomicAdd(multi_count, 1u)] = b;     // buckets whose partial sums need a fold
    if (full) atomicAdd(&sh_hist[T], full);
    if (rem) atomicAdd(&sh_hist[rem], 1u);
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i <= T; i += blockDim.x)
    if (sh_hist[i]) atomicAdd(&size_hist[i], sh_hist[i]);
}

// size_cursor[s] = number of tasks longer than s (descending order), single block
__global__ void task_size_scan_kernel(const uint32_t* __restrict__ size_hist, uint32_t T, uint32_t* __restrict__ size_cursor) {
  if (threadIdx.x == 0) {
    uint32_t acc = 0;
    for (uint32_t s = T; s >= 1; s--) { size_cursor[s] = acc; acc += size_hist[s]; }
    size_cursor[0] = acc;                          // total number of tasks
  }
}

// pass 2: emit the tasks into their length class.  One warp per 32 buckets: each lane emits its
// bucket's first task, the remaining tasks of multi-task buckets are emitted by the whole warp.
__global__ void task_emit_kernel(const uint32_t* __restrict__ offsets, uint32_t ostride, const uint32_t* __restrict__ ntasks,
                                 const uint32_t* __restrict__ task_off, uint32_t nb, uint32_t T,
                                 uint32_t* __restrict__ size_cursor, uint32_t* __restrict__ t_start, uint32_t* __restrict__ t_len,
                                 uint32_t* __restrict__ t_dest) {
  const uint32_t lane = threadIdx.x & 31;
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t my_start = 0, my_size = 0, my_nt = 0, my_off = 0;
  if (b < nb) {
    my_start = offsets[(size_t)b * ostride]; my_size = offsets[(size_t)(b + 1) * ostride] - my_start;
    my_nt = ntasks[b]; my_off = task_off[b];
  }
  auto emit = [&](uint32_t bucket, uint32_t bstart, uint32_t bsize, uint32_t nt, uint32_t toff, uint32_t j) {
    uint32_t s0 = j * T;
    uint32_t len = bsize - s0 < T ? bsize - s0 : T;
    // warp-aggregated slot reservation: lanes emitting the same length share one atomic
    uint32_t am = __activemask();
    uint32_t peers = __match_any_sync(am, len);
    int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(&size_cursor[len], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    uint32_t pos = base + __popc(peers & ((1u << lane) - 1u));
    t_start[pos] = bstart + s0;
    t_len[pos] = len;
    t_dest[pos] = nt == 1 ? (bucket | 0x80000000u) : (toff + j);     // direct to the bucket, or a partial slot
  };
  if (my_nt) emit(b, my_start, my_size, my_nt, my_off, 0);
  uint32_t multi = __ballot_sync(0xffffffffu, my_nt > 1);
  while (multi) {
    int src = __ffs(multi) - 1;
    multi &= multi - 1;
    uint32_t bb = __shfl_sync(0xffffffffu, b, src), bs = __shfl_sync(0xffffffffu, my_start, src);
    uint32_t bz = __shfl_sync(0xffffffffu, my_size, src), bn = __shfl_sync(0xffffffffu, my_nt, src);
    uint32_t bo = __shfl_sync(0xffffffffu, my_off, src);
    for (uint32_t j = 1 + lane; j < bn; j += 32) emit(bb, bs, bz, bn, bo, j);
  }
}

// one thread per task: XYZZ accumulator in registers, points gathered through the sorted index;
// the next point is fetched while the current mixed addition runs
template <class Cfg>
__global__ void __launch_bounds__(128, (Cfg::Fp::N == 8 ? 4 : 1)) msm_accumulate_kernel(const uint32_t* __restrict__ points,
                                                            const uint32_t* __restrict__ entries,
                                                            const uint32_t* __restrict__ t_start, const uint32_t* __restrict__ t_len,
                                                            const uint32_t* __restrict__ t_dest, const uint32_t* __restrict__ n_tasks,
                                                            uint32_t add_existing, uint32_t* __restrict__ buckets,
                                                            uint32_t* __restrict__ partials) {
  using P = typename Cfg::Fp;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *n_tasks) return;
  uint32_t e = t_start[t], len = t_len[t], dest = t_dest[t];
  // later chunks of a chunked MSM continue from the bucket's running sum
  XYZZ<P> acc = (add_existing && (dest >> 31)) ? ld_xyzz<P>(buckets, dest & 0x7fffffffu) : xyzz_inf<P>();
  uint32_t ent = __ldg(entries + e);
  Affine<P> nxt = ld_affine<P>(points, ent & 0x7fffffffu);
  uint32_t nneg = ent >> 31;
  for (uint32_t k = 0; k < len; k++) {
    Affine<P> pt = nxt;
    uint32_t neg = nneg;
    if (k + 1 < len) {
      ent = __ldg(entries + e + k + 1);
      nxt = ld_affine<P>(points, ent & 0x7fffffffu);
      nneg = ent >> 31;
    }
    if (neg) pt.y = fe_neg<P>(pt.y);
    // 254/381-bit base fields: accumulator coordinates semi-reduced in [0, 2p) inside the loop (curve.cuh), folded once below
    if constexpr (FeLz<P>::ok) xyzz_madd_lz<P>(acc, pt); else xyzz_madd<P>(acc, pt);
  }
  if constexpr (FeLz<P>::ok) acc = xyzz_reduce_lz<P>(acc);
  if (dest >> 31) st_xyzz<P>(buckets, dest & 0x7fffffffu, acc);
  else st_xyzz<P>(partials, dest, acc);
}

// Buckets that were split into several tasks: fold their partial sums.  Persistent grid, one
// warp per split bucket (from the compact list built by task_count_kernel): lanes sum strided
// partials serially, then a warp-shuffle tree reduction combines the 32 lane sums.
template <class P> __device__ __forceinline__ XYZZ<P> shfl_xyzz(const XYZZ<P>& a, int delta) {
  XYZZ<P> r;
  const uint32_t* s = a.x.v;
  uint32_t* d = r.x.v;
#pragma unroll
  for (int i = 0; i < 4 * P::N; i++) d[i] = __shfl_down_sync(0xffffffffu, s[i], delta);
  return r;
}
template <class Cfg>
__global__ void __launch_bounds__(128) msm_merge_kernel(const uint32_t* __restrict__ ntasks, const uint32_t* __restrict__ task_off,
                                                       const uint32_t* __restrict__ multi_count,
                                                       const uint32_t* __restrict__ multi_list, uint32_t add_existing,
                                                       const uint32_t* __restrict__ partials, uint32_t* __restrict__ buckets) {
  using P = typename Cfg::Fp;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t nmulti = *multi_count;
  for (uint32_t k = warp; k < nmulti; k += nwarps) {
    uint32_t b = multi_list[k];
    uint32_t nt = ntasks[b], off = task_off[b];
    XYZZ<P> acc = xyzz_inf<P>();
    for (uint32_t j = lane; j < nt; j += 32) acc = xyzz_add<P>(acc, ld_xyzz<P>(partials, off + j));
    if (add_existing && lane == 0) acc = xyzz_add<P>(acc, ld_xyzz<P>(buckets, b));
    for (int d = 16; d > 0; d >>= 1) {
      XYZZ<P> o = shfl_xyzz<P>(acc, d);
      acc = xyzz_add<P>(acc, o);
    }
    if (lane == 0) st_xyzz<P>(buckets, b, acc);
  }
}

// empty buckets hold the identity
template <class Cfg>
__global__ void msm_clear_empty_kernel(const uint32_t* __restrict__ ntasks, uint32_t nb, uint32_t* __restrict__ buckets) {
  using P = typename Cfg::Fp;
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nb && ntasks[b] == 0) st_xyzz<P>(buckets, b, xyzz_inf<P>());
}

// chunked running sum: thread (w, chunk) reduces CH consecutive buckets of window w to
//   sum_b (b + 1) * B_b  over its chunk  =  acc + lo * running
template <class Cfg>
__global__ void __launch_bounds__(128) msm_reduce_kernel(const uint32_t* __restrict__ buckets, uint32_t B, uint32_t CH,
                                                        uint32_t chunks_per_window, uint32_t W, uint32_t* __restrict__ partials) {
  using P = typename Cfg::Fp;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= chunks_per_window * W) return;
  uint32_t w = t / chunks_per_window, ch = t % chunks_per_window;
  uint32_t lo = ch * CH, hi = lo + CH < B ? lo + CH : B;
  XYZZ<P> running = xyzz_inf<P>(), acc = xyzz_inf<P>();
  for (uint32_t b = hi; b-- > lo;) {
    XYZZ<P> bk = ld_xyzz<P>(buckets, (size_t)w * B + b);
    // the chain runs semi-reduced where the field allows it (curve.cuh xyzz_add_lz): ~12 % fewer dependent instructions per addition
    if constexpr (FeLz<P>::ok && FeSq<P>::ok) { running = xyzz_add_lz<P>(running, bk); acc = xyzz_add_lz<P>(acc, running); }
    else { running = xyzz_add<P>(running, bk); acc = xyzz_add<P>(acc, running); }
  }
  if constexpr (FeLz<P>::ok && FeSq<P>::ok) { running = xyzz_reduce_lz<P>(running); acc = xyzz_reduce_lz<P>(acc); }
  if (lo) acc = xyzz_add<P>(acc, xyzz_mul_u32<P>(running, lo));
  st_xyzz<P>(partials, t, acc);
}

// one level of the per-window tree sum: block (bx, w) adds `per_block` consecutive points of
// window w (strided serial sums, then a shared-memory tree) into out[w * gridDim.x + bx]
template <class Cfg>
__global__ void __launch_bounds__(128) msm_window_kernel(const uint32_t* __restrict__ in, uint32_t count, uint32_t per_block,
                                                        uint32_t* __restrict__ out) {
  using P = typename Cfg::Fp;
  extern __shared__ uint32_t shw[];
  const uint32_t w = blockIdx.y;
  const uint32_t lo = blockIdx.x * per_block, hi = lo + per_block < count ? lo + per_block : count;
  XYZZ<P> acc = xyzz_inf<P>();
  for (uint32_t k = lo + threadIdx.x; k < hi; k += blockDim.x) acc = xyzz_add<P>(acc, ld_xyzz<P>(in, (size_t)w * count + k));
  st_xyzz<P>(shw, threadIdx.x, acc);
  __syncthreads();
  for (uint32_t off = blockDim.x / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off) {
      XYZZ<P> a = ld_xyzz<P>(shw, threadIdx.x), b = ld_xyzz<P>(shw, threadIdx.x + off);
      st_xyzz<P>(shw, threadIdx.x, xyzz_add<P>(a, b));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) st_xyzz<P>(out, (size_t)w * gridDim.x + blockIdx.x, ld_xyzz<P>(shw, 0));
}

// Horner over windows (top first), optional normalisation.
// out layout: mode 0 -> XYZZ Montgomery (4N words); mode 1 -> [x | y] canonical (2N words) + inf flag word
template <class Cfg>
__global__ void msm_final_kernel(const uint32_t* __restrict__ winsums, uint32_t W, uint32_t c, int mode, uint32_t* __restrict__ out) {
  using P = typename Cfg::Fp;
  if (threadIdx.x != 0) return;
  winsums += (size_t)blockIdx.x * W * 4 * P::N;        // block j: polynomial j of a batched pass
  XYZZ<P> acc = xyzz_inf<P>();
  for (uint32_t w = W; w-- > 0;) {
    if (!xyzz_is_inf<P>(acc))
      for (uint32_t k = 0; k < c; k++) acc = xyzz_dbl<P>(acc);
    acc = xyzz_add<P>(acc, ld_xyzz<P>(winsums, w));
  }
  if (mode == 0) { st_xyzz<P>(out, blockIdx.x, acc); return; }
  Affine<P> a = xyzz_to_affine<P>(acc);
  Fe<P> x = fe_from_mont<P>(a.x), y = fe_from_mont<P>(a.y);
  for (int i = 0; i < P::N; i++) { out[i] = x.v[i]; out[P::N + i] = y.v[i]; }
  out[2 * P::N] = xyzz_is_inf<P>(acc) ? 1u : 0u;
}

// sum `count` XYZZ points (one block) -> one XYZZ point; the caller normalises it on the host (host_xyzz_to_canonical: one
// inversion is ~20 us there against ~190 us for a lone GPU thread -- at 8 ranks that was 3 % of the sharded step)
template <class Cfg>
__global__ void __launch_bounds__(128) g1_fold_kernel(const uint32_t* __restrict__ pts, uint32_t count, uint32_t* __restrict__ out) {
  using P = typename Cfg::Fp;
  extern __shared__ uint32_t shw[];
  XYZZ<P> acc = xyzz_inf<P>();
  for (uint32_t i = threadIdx.x; i < count; i += blockDim.x) acc = xyzz_add<P>(acc, ld_xyzz<P>(pts, i));
  st_xyzz<P>(shw, threadIdx.x, acc);
  __syncthreads();
  for (uint32_t off = blockDim.x / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off) {
      XYZZ<P> a = ld_xyzz<P>(shw, threadIdx.x), b = ld_xyzz<P>(shw, threadIdx.x + off);
      st_xyzz<P>(shw, threadIdx.x, xyzz_add<P>(a, b));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) st_xyzz<P>(out, 0, ld_xyzz<P>(shw, 0));
}

// Variable-base linear combination sum_i s_i * P_i of a handful of arbitrary points (the verifier-side combinations
// of kzg.py:183-205, 252-281): thread i runs an MSB-first double-and-add over its scalar, then the block folds the
// XYZZ partials exactly like g1_fold_kernel.  points: canonical affine ((0,0) = identity); scalars: canonical limbs.
template <class Cfg>
__global__ void __launch_bounds__(128) g1_lincomb_kernel(const uint32_t* __restrict__ pts, const uint32_t* __restrict__ scalars,
                                                        uint32_t count, uint32_t* __restrict__ out) {
  using P = typename Cfg::Fp;
  extern __shared__ uint32_t shw[];
  XYZZ<P> acc = xyzz_inf<P>();
  for (uint32_t i = threadIdx.x; i < count; i += blockDim.x) {
    Affine<P> a;
    for (int k = 0; k < P::N; k++) { a.x.v[k] = pts[(size_t)i * 2 * P::N + k]; a.y.v[k] = pts[(size_t)i * 2 * P::N + P::N + k]; }
    if (aff_is_inf<P>(a)) continue;
    a.x = fe_to_mont<P>(a.x); a.y = fe_to_mont<P>(a.y);
    XYZZ<P> r = xyzz_inf<P>();
    bool started = false;
    for (int w = 7; w >= 0; w--) {
      uint32_t word = scalars[(size_t)i * 8 + w];
      for (int b = 31; b >= 0; b--) {
        if (started) r = xyzz_dbl<P>(r);
        if ((word >> b) & 1) { xyzz_madd<P>(r, a); started = true; }
      }
    }
    acc = xyzz_add<P>(acc, r);
  }
  st_xyzz<P>(shw, threadIdx.x, acc);
  __syncthreads();
  for (uint32_t off = blockDim.x / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off) {
      XYZZ<P> a = ld_xyzz<P>(shw, threadIdx.x), b = ld_xyzz<P>(shw, threadIdx.x + off);
      st_xyzz<P>(shw, threadIdx.x, xyzz_add<P>(a, b));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) st_xyzz<P>(out, 0, ld_xyzz<P>(shw, 0));
}

// canonical affine -> Montgomery affine (in place); (0,0) stays (0,0)
template <class Cfg> __global__ void srs_to_mont_kernel(uint32_t* pts, size_t n) {
  using P = typename Cfg::Fp;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * n) return;
  Fe<P> v;
  for (int k = 0; k < P::N; k++) v.v[k] = pts[i * P::N + k];
  v = fe_to_mont<P>(v);
  for (int k = 0; k < P::N; k++) pts[i * P::N + k] = v.v[k];
}
template <class Cfg> __global__ void srs_from_mont_kernel(const uint32_t* pts, size_t first, size_t count, uint32_t* out) {
  using P = typename Cfg::Fp;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * count) return;
  Fe<P> v;
  for (int k = 0; k < P::N; k++) v.v[k] = pts[(2 * first + i) * P::N + k];
  v = fe_from_mont<P>(v);
  for (int k = 0; k < P::N; k++) out[i * P::N + k] = v.v[k];
}

// window table w from table w-1: T_w[i] = 2^c * T_{w-1}[i], normalised back to affine
template <class Cfg>
__global__ void __launch_bounds__(128) srs_table_kernel(const uint32_t* __restrict__ prev, uint32_t* __restrict__ next, size_t n,
                                                       uint32_t c) {
  using P = typename Cfg::Fp;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<P> a = ld_affine<P>(prev, i);
  Affine<P> r;
  if (aff_is_inf<P>(a)) {
    r = a;
  } else {
    XYZZ<P> q = xyzz_dbl_affine<P>(a);
    for (uint32_t k = 1; k < c; k++) q = xyzz_dbl<P>(q);
    r = xyzz_to_affine<P>(q);
  }
  uint32_t* o = next + i * (2 * P::N);
  for (int k = 0; k < P::N; k++) { o[k] = r.x.v[k]; o[P::N + k] = r.y.v[k]; }
}

// All window tables of a short key in ONE launch.  Thread i walks the doubling chain of point i -- T_w[i] = 2^(c w) P_i, c (W - 1)
// doublings in XYZZ -- leaving X, Y in the table slots and ZZ, ZZZ plus the running product of the denominators ZZ * ZZZ in
// scratch; ONE Fermat inversion per thread then normalises all W - 1 entries on the way back (Montgomery's trick: 9 products per
// entry instead of a 380-product inversion).  The per-table kernel above pays that inversion W - 1 times per point in W - 1
// dependent launches: 8 ms for a 22-point key (c = 7, 37 tables) against 1.5 ms here, which is what a caller that hands over
// a fresh `ck` list pays before its first commit (kzg.py:80).  scratch: (W - 1) * n * 3N words, laid out [w][i] like the tables.
template <class Cfg>
__global__ void __launch_bounds__(128) srs_tables_fused_kernel(uint32_t* __restrict__ tables, size_t n, uint32_t c, uint32_t W,
                                                              uint32_t* __restrict__ scratch) {
  using P = typename Cfg::Fp;
  constexpr int N = P::N;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t tstride = n * 2 * N, sstride = n * 3 * N;
  auto st_fe = [](uint32_t* p, const Fe<P>& a) { for (int k = 0; k < N; k++) p[k] = a.v[k]; };
  auto ld_fe = [](const uint32_t* p) { Fe<P> a; for (int k = 0; k < N; k++) a.v[k] = p[k]; return a; };
  Affine<P> a = ld_affine<P>(tables, i);
  if (aff_is_inf<P>(a)) {
    for (uint32_t w = 1; w < W; w++) { st_fe(tables + w * tstride + i * 2 * N, a.x); st_fe(tables + w * tstride + i * 2 * N + N, a.y); }
    return;
  }
  XYZZ<P> q = xyzz_from_affine<P>(a);
  Fe<P> run = fe_one<P>();
  for (uint32_t w = 1; w < W; w++) {
    for (uint32_t k = 0; k < c; k++) q = xyzz_dbl<P>(q);
    uint32_t* t = tables + w * tstride + i * 2 * N;
    uint32_t* sc = scratch + (w - 1) * sstride + i * 3 * N;
    st_fe(t, q.x); st_fe(t + N, q.y);
    st_fe(sc, q.zz); st_fe(sc + N, q.zzz);
    st_fe(sc + 2 * N, run);                               // product of the denominators of tables 1 .. w-1
    Fe<P> den = fe_mul<P>(q.zz, q.zzz);
    if (fe_is_zero<P>(den)) den = fe_one<P>();            // 2^(cw) P_i = infinity (a point outside the prime-order group): written as (0, 0) below
    run = fe_mul<P>(run, den);
  }
  Fe<P> inv = fe_inv<P>(run);
  for (uint32_t w = W - 1; w >= 1; w--) {
    uint32_t* t = tables + w * tstride + i * 2 * N;
    const uint32_t* sc = scratch + (w - 1) * sstride + i * 3 * N;
    Fe<P> zz = ld_fe(sc), zzz = ld_fe(sc + N), pre = ld_fe(sc + 2 * N);
    Fe<P> den = fe_mul<P>(zz, zzz);
    if (fe_is_zero<P>(den)) { st_fe(t, fe_zero<P>()); st_fe(t + N, fe_zero<P>()); continue; }
    Fe<P> dinv = fe_mul<P>(inv, pre);                     // 1 / (zz * zzz) = Z^-5
    inv = fe_mul<P>(inv, den);
    Fe<P> x = fe_mul<P>(ld_fe(t), fe_mul<P>(dinv, zzz));  // X / ZZ
    Fe<P> y = fe_mul<P>(ld_fe(t + N), fe_mul<P>(dinv, zz));   // Y / ZZZ
    st_fe(t, x); st_fe(t + N, y);
  }
}

// SRS generation (kzg.py:69-72): point i = tau^i * G1.  dbl_table[j] = 2^j * G1 (affine, Montgomery).
template <class Cfg>
__global__ void __launch_bounds__(128) srs_generate_kernel(uint32_t* pts, size_t start, size_t n, Fe<typename Cfg::Fr> tau_mont,
                                                          const uint32_t* __restrict__ dbl_table) {
  using P = typename Cfg::Fp;
  using R = typename Cfg::Fr;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // tau^i in the scalar field
  Fe<R> base = tau_mont, e = fe_one<R>();
  for (size_t k = start + i; k; k >>= 1) {
    if (k & 1) e = fe_mul<R>(e, base);
    if (k >> 1) base = fe_sqr<R>(base);
  }
  e = fe_from_mont<R>(e);
  XYZZ<P> acc = xyzz_inf<P>();
  for (int bit = 0; bit < R::BITS; bit++) {
    if ((e.v[bit >> 5] >> (bit & 31)) & 1) {
      Affine<P> t = ld_affine<P>(dbl_table, bit);
      xyzz_madd<P>(acc, t);
    }
  }
  Affine<P> a = xyzz_to_affine<P>(acc);
  uint32_t* o = pts + i * (2 * P::N);
  for (int k = 0; k < P::N; k++) { o[k] = a.x.v[k]; o[P::N + k] = a.y.v[k]; }
}

// ---- tiny MSMs (n * W <= kTinyEntries digit entries: the 22-point key of the bundled PLONK instance, the 201-point key of
// the bundled Marlin instance) -- no sort, no buckets, two launches.  With window tables every term of the sum is
// digit * T_w[i] with |digit| <= 2^(c-1) and c <= 10 for such keys: thread (i, w) forms that small multiple by double-and-add
// (<= c doublings, a ~50 us chain), the block tree-sums its 256 terms, and msm_tiny_fold_kernel sums the block partials of
// each polynomial.  The bucket path spends ~0.4 ms of fixed cost (21 launches) on the same work.
constexpr uint32_t kTinyEntries = 98304;      // up to ~4,000 points: above that the bucket method does less work
constexpr uint32_t kTinyThreads = 256;

template <class Cfg>
__global__ void __launch_bounds__(kTinyThreads) msm_tiny_kernel(const uint32_t* __restrict__ points, const uint32_t* __restrict__ scalars,
                                                              uint32_t poly_len, uint32_t first, uint32_t n_srs, uint32_t c, uint32_t W,
                                                              uint32_t top_bits, DigitOffset off, uint32_t* __restrict__ flag,
                                                              uint32_t* __restrict__ partials) {
  using P = typename Cfg::Fp;
  extern __shared__ uint32_t shw[];
  const uint32_t e = blockIdx.x * kTinyThreads + threadIdx.x;      // entry = (point i, window w) of polynomial blockIdx.y
  XYZZ<P> acc = xyzz_inf<P>();
  if (e < poly_len * W) {
    const uint32_t i = e / W, w = e - i * W;
    const size_t gi = (size_t)blockIdx.y * poly_len + i;
    uint32_t s[8];
    load_scalar_plus_offset(scalars, gi, off, s);
    if (w == 0 && top_bits < 32) {
      const uint32_t top = __ldg(reinterpret_cast<const uint4*>(scalars + gi * 8) + 1).w;
      if (top >> top_bits) atomicOr(flag, 1u);
    }
    const int d = signed_digit(s, w, c, W);
    if (d != 0) {
      Affine<P> pt = ld_affine<P>(points, (size_t)w * n_srs + first + i);
      if (!aff_is_inf<P>(pt)) {
        if (d < 0) pt.y = fe_neg<P>(pt.y);
        acc = xyzz_mul_u32<P>(xyzz_from_affine<P>(pt), (uint32_t)(d < 0 ? -d : d));
      }
    }
  }
  st_xyzz<P>(shw, threadIdx.x, acc);
  __syncthreads();
  for (uint32_t o = kTinyThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      XYZZ<P> a = ld_xyzz<P>(shw, threadIdx.x), b = ld_xyzz<P>(shw, threadIdx.x + o);
      st_xyzz<P>(shw, threadIdx.x, xyzz_add<P>(a, b));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) st_xyzz<P>(partials, (size_t)blockIdx.y * gridDim.x + blockIdx.x, ld_xyzz<P>(shw, 0));
}

// block j: sum of the `count` block partials of polynomial j -> out[j] (XYZZ)
template <class Cfg>
__global__ void __launch_bounds__(64) msm_tiny_fold_kernel(const uint32_t* __restrict__ partials, uint32_t count, uint32_t* __restrict__ out) {
  using P = typename Cfg::Fp;
  extern __shared__ uint32_t shw[];
  XYZZ<P> acc = xyzz_inf<P>();
  for (uint32_t k = threadIdx.x; k < count; k += blockDim.x) acc = xyzz_add<P>(acc, ld_xyzz<P>(partials, (size_t)blockIdx.x * count + k));
  st_xyzz<P>(shw, threadIdx.x, acc);
  __syncthreads();
  for (uint32_t o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      XYZZ<P> a = ld_xyzz<P>(shw, threadIdx.x), b = ld_xyzz<P>(shw, threadIdx.x + o);
      st_xyzz<P>(shw, threadIdx.x, xyzz_add<P>(a, b));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) st_xyzz<P>(out, blockIdx.x, ld_xyzz<P>(shw, 0));
}

// ---------------------------------------------------------------- host side
uint32_t choose_c(size_t n, int bits) {
  const char* env = getenv("KZGPU_MSM_C");
  if (env) {
    int v = atoi(env);
    if (v >= 2 && v <= 24) return (uint32_t)v;
  }
  uint32_t logn = 0;
  while ((1ull << (logn + 1)) <= n) logn++;
  int c0 = (int)logn - 4;
  if (c0 < 3) c0 = 3;
  if (c0 > 20) c0 = 20;
  // the top window holds bits - c*(W-1) bits; a nearly empty top window concentrates all points
  // in a handful of buckets, so prefer the nearest c whose top window is at least half full
  for (int d = 0; d <= 4; d++) {
    for (int sgn = 1; sgn >= -1; sgn -= 2) {
      int c = c0 + sgn * d;
      if (c < 3 || c > 22) continue;
      int W = (bits + 1 + c - 1) / c;
      int top = bits - c * (W - 1);
      if (2 * top >= c) return (uint32_t)c;
    }
  }
  return (uint32_t)c0;
}

// window size of the precomputed tables for a key of n points: minimise
//   10 * n * W(c)  (mixed additions, 10 modmul each)  +  28 * 2^(c-1)  (bucket reduction, 2 full adds per bucket)
// (checked against forced-window sweeps, scripts/c_sweep.py / profiles/r1_c_sweep.txt: the model's choice is the
// measured optimum at 2^16, 2^18, 2^20, 2^22 and 2^24; a larger bucket term would pick windows whose few, heavy
// buckets leave too few accumulate tasks to fill the SMs)
// subject to the memory cap and to 31-bit point indices.  Returns 0 when tables are disabled.
uint32_t choose_table_c(size_t n, int bits, size_t point_bytes) {
  const char* env = getenv("KZGPU_SRS_TABLES");
  if (env && env[0] == '0') return 0;
  if (env && env[0] == 'c' && env[1] == '=') {
    int v = atoi(env + 2);
    if (v >= 2 && v <= 24) return (uint32_t)v;
  }
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  double cap = 0.35 * (double)total_b;
  const char* cap_env = getenv("KZGPU_SRS_TABLE_GIB");
  if (cap_env) cap = atof(cap_env) * 1073741824.0;
  if ((double)free_b * 0.6 < cap) cap = (double)free_b * 0.6;
  double best = 1e300;
  uint32_t best_c = 0;
  for (uint32_t c = 2; c <= 24; c++) {
    uint32_t W = (bits + 1 + c - 1) / c;
    double mem = (double)W * (double)n * (double)point_bytes;
    if (mem > cap || (double)W * (double)n >= 2147483648.0) continue;
    double cost = 10.0 * (double)(n ? n : 1) * W + 28.0 * (double)(1ull << (c - 1));
    if (cost < best) { best = cost; best_c = c; }
  }
  return best_c;
}

// The "scalar not canonical" flag of an MSM whose caller did not wait for it (msm_core with defer = true: the partial sum stays
// on the device and the next thing queued behind it is an exchange, so a host synchronisation here would only open a gap):
// collected by whoever synchronises the stream next -- the fold of the partials, kzgpu_sync, or the next MSM.
bool g_flag_pending[KZ_MAX_DEV] = {false};
int msm_flag_collect(uint32_t* host_word) {             // queue the read behind everything on the stream (no-op if nothing is pending)
  *host_word = 0;
  if (!g_flag_pending[kz_slot()]) return 0;
  KZ_CUDA(cudaMemcpyAsync(host_word, g_ws.flag.p, 4, cudaMemcpyDeviceToHost, kz_ctx().stream));
  return 0;
}
int msm_flag_result(uint32_t host_word) {               // after the stream has been synchronised
  if (!g_flag_pending[kz_slot()]) return 0;
  g_flag_pending[kz_slot()] = false;
  if (host_word) return kz_fail(KZGPU_ERANGE, "a scalar of the preceding MSM is not a canonical residue");
  return 0;
}
int msm_flag_wait() {
  if (!g_flag_pending[kz_slot()]) return 0;
  uint32_t h = 0;
  int rc = msm_flag_collect(&h);
  if (rc) return rc;
  KZ_CUDA(cudaStreamSynchronize(kz_ctx().stream));
  return msm_flag_result(h);
}

// `srs`: one device's part of a key (the calling thread's device); `first` is an index INSIDE that part
template <class Cfg>
int msm_core(const SrsPart& srs, size_t first, const uint32_t* d_scalars, size_t n, int mode, uint32_t* d_out,
             const uint64_t* h_scalars = nullptr, uint32_t* h_out_xyzz = nullptr, uint32_t batch = 1, bool defer = false) {
  { int rcp = msm_flag_wait(); if (rcp) return rcp; }
  // batch > 1: d_scalars holds `batch` polynomials of n / batch scalars each (zero padded to equal length); they
  // share one sort / accumulate / reduce pass, every polynomial owning its own bucket set(s), and d_out receives
  // `batch` XYZZ results (mode 0 / 2 only)
  using P = typename Cfg::Fp;
  using R = typename Cfg::Fr;
  KzgpuCtx& cx = kz_ctx();
  cudaStream_t st = cx.stream;
  // Chunks.  Host scalars (the e2e entry point): the upload is split into chunks on the copy stream and the MSM runs
  // chunk by chunk into the same buckets, so that all but the first chunk's PCIe time hides behind the work on the
  // previous chunk.  Chunk sizes grow (1/16, 3/16, 1/4, 1/2 of the scalars): only the first chunk's transfer is exposed.
  // The sort + task construction of chunk k+1 (HBM / shared-memory-atomic bound) runs on a high-priority second stream
  // while chunk k is accumulated (multiplier bound), each chunk owning one of two sort workspaces.  Measured at 2^24
  // (scripts/pipe_ab.py): 38.0 -> 37.3 ms for host scalars.  The same split applied to device-resident scalars
  // (KZGPU_MSM_DEV_SPLIT=d: chunks n/d and the rest) does NOT pay -- 35.7 -> 35.9 ms: a resident sort block takes the
  // registers of one of the four accumulate blocks of its SM, and the accumulate kernel at 3 warps per scheduler
  // loses about what the overlap wins, plus the second chunk's bucket fix-up -- so it stays off by default.
  static const bool pipe_ok = !getenv("KZGPU_MSM_NO_PIPE");
  uint32_t nchunks = 1;
  bool host_two = false;
  if (batch == 1 && h_scalars && n >= (1u << 20)) {
    // every chunk pays the fixed cost of a sort / task / accumulate round (~0.2 ms) and a bucket load + store per task, so short
    // MSMs -- the shards of a point-sharded MSM -- upload in two chunks (1/4, 3/4), long ones in four
    static const int forced = getenv("KZGPU_MSM_NCHUNKS") ? atoi(getenv("KZGPU_MSM_NCHUNKS")) : 0;
    nchunks = forced == 1 || forced == 2 || forced == 4 ? (uint32_t)forced : (n >= (1u << 23) ? 4u : 2u);
    host_two = nchunks == 2;
  }
  else if (batch == 1 && !h_scalars && n >= (1u << 22) && pipe_ok && getenv("KZGPU_MSM_DEV_SPLIT")) nchunks = 2;
  size_t chunk_lo[5] = {0, n, n, n, n};
  if (nchunks == 4) {
    const char* env = getenv("KZGPU_MSM_CHUNKS");      // "uniform": four equal chunks (A/B measurements)
    if (env && env[0] == 'u') { chunk_lo[1] = n / 4; chunk_lo[2] = n / 2; chunk_lo[3] = 3 * (n / 4); }
    else { chunk_lo[1] = n / 16; chunk_lo[2] = n / 4; chunk_lo[3] = n / 2; }
  } else if (host_two) {
    chunk_lo[1] = n / 4;
  } else if (nchunks == 2) {
    static const char* env = getenv("KZGPU_MSM_DEV_SPLIT");     // first chunk = n / split (A/B measurements)
    const size_t split = env && atoi(env) >= 2 ? (size_t)atoi(env) : 8;
    chunk_lo[1] = n / split;
  }
  // profiling (per-kernel CUDA events on the main stream) keeps everything on one stream
  const bool piped = nchunks > 1 && pipe_ok && !cx.profile;
  cudaStream_t sst = piped ? cx.sort_stream : st;
  size_t chunk_n = 0;                                   // largest chunk: sizes the sort scratch
  for (uint32_t k = 0; k < nchunks; k++) if (chunk_lo[k + 1] - chunk_lo[k] > chunk_n) chunk_n = chunk_lo[k + 1] - chunk_lo[k];
  // chunk k of the host scalars -> device, on the copy stream (pinned memory: one asynchronous copy; pageable memory: staged
  // through page-locked buffers by a few host threads, kz_upload).  Chunk 0 goes now, chunk k + 1 right after chunk k's
  // kernels have been queued, so that its transfer (and, for pageable memory, its host-side staging) runs under them.
  const uint32_t* const d_base = d_scalars;              // (d_scalars itself walks through the chunks below)
  auto upload_chunk = [&](uint32_t k) -> int {
    const size_t lo = chunk_lo[k], cnt = chunk_lo[k + 1] - chunk_lo[k];
    if (cnt) { int r = kz_upload((void*)(d_base + lo * 8), h_scalars + lo * 4, cnt * 32, cx.copy_stream); if (r) return r; }
    KZ_CUDA(cudaEventRecord(cx.copy_ev[k], cx.copy_stream));
    return 0;
  };
  if (h_scalars && n) { int r = upload_chunk(0); if (r) return r; }
  const bool tabled = srs.c_tab != 0;
  const uint32_t c = tabled ? srs.c_tab : choose_c(n, R::BITS);
  const uint32_t W = (R::BITS + 1 + c - 1) / c;          // digits per scalar
  const uint32_t Wp = tabled ? 1u : W;                   // bucket sets per polynomial
  const uint32_t Wb = Wp * batch;                        // bucket sets of the pass
  const uint32_t B = 1u << (c - 1);
  const size_t nb = (size_t)Wb * B;                     // buckets = sort keys
  if (first + n > 0x7fffffffull) return kz_fail(KZGPU_EINVAL, "MSM index range exceeds 2^31");
  // sort geometry: f fine bits, <= kMaxCoarse coarse bins
  uint32_t f = 10;
  while ((1ull << f) > nb && f > 0) f--;
  while (((nb + (1ull << f) - 1) >> f) > kMaxCoarse && f < 13) f++;
  const uint32_t ncoarse = (uint32_t)((nb + (1ull << f) - 1) >> f);
  if (ncoarse > kMaxCoarse) return kz_fail(KZGPU_EINVAL, "MSM window c=%u gives too many buckets (%zu)", c, nb);
  const size_t max_entries = (size_t)chunk_n * W;
  if (max_entries >= 0xffffffffull) return kz_fail(KZGPU_EINVAL, "MSM of %zu points x %u digits exceeds 2^32 entries", n, W);
  const size_t nblk = kz_div_up(nb, 1024);
  // tasks: heavy buckets are split into <= T-point tasks, T from the chunk's mean bucket load -- and from the load of the TOP
  // window's buckets: that window holds only BITS - c (W - 1) bits, so its 2^top buckets receive every point (128 per bucket for
  // a 2^21-point shard at c = 20, against a mean of 53).  A T just below that load split half of them in two and left ~8,000
  // two-task buckets for msm_merge_kernel (140 us of a 5.2 ms MSM); T covers it now whenever it fits the 1024 cap.
  const uint32_t top_window_bits = (uint32_t)R::BITS - c * (W - 1);
  auto task_T = [&](size_t cn) {
    uint32_t mean = (uint32_t)((((size_t)cn / batch) * (tabled ? W : 1u)) / B) + 1, t = 32;
    while (t < 2 * mean && t < 1024) t <<= 1;
    // (not for the four-chunk host-scalar pipeline of long MSMs: measured 34.3 -> 36.3 ms at 2^24 with it)
    const size_t top_load = (top_window_bits < 31 && nchunks <= 2) ? (((size_t)cn / batch) >> top_window_bits) : 0;
    while ((size_t)t < top_load + top_load / 4 + 8 && t < 1024) t <<= 1;
    return t;
  };
  size_t max_tasks = 0, max_multi = 0;
  for (uint32_t k = 0; k < nchunks; k++) {
    const size_t cn = chunk_lo[k + 1] - chunk_lo[k];
    const size_t split = (cn * W) / task_T(cn) + 1;               // a split bucket holds more than T points
    if (nb + split > max_tasks) max_tasks = nb + split;
    if (split > max_multi) max_multi = split;
  }
  int rc;
  for (uint32_t s = 0; s < (nchunks > 1 ? 2u : 1u); s++) {
    MsmSortWs& w = g_sw[s];
    if ((rc = w.counts.ensure(nb * 4))) return rc;
    if ((rc = w.offsets.ensure((nb + 1) * 4))) return rc;
    if ((rc = w.cursor.ensure(nb * 4))) return rc;
    if ((rc = w.entries.ensure(max_entries * 4 + 4))) return rc;
    if ((rc = w.sort_tmp.ensure(max_entries * 8 + 8))) return rc;
    if ((rc = w.coarse.ensure((size_t)(4 * kMaxCoarse + 4) * 4))) return rc;
    if ((rc = w.blocksums.ensure(nblk * 4))) return rc;
    if ((rc = w.ntasks.ensure(nb * 4))) return rc;
    if ((rc = w.task_off.ensure((nb + 1) * 4))) return rc;
    if ((rc = w.size_hist.ensure(2 * (1024 + 1) * 4))) return rc;
    if ((rc = w.t_start.ensure(max_tasks * 4))) return rc;
    if ((rc = w.t_len.ensure(max_tasks * 4))) return rc;
    if ((rc = w.t_dest.ensure(max_tasks * 4))) return rc;
    if ((rc = w.tparts.ensure(max_tasks * 4 * P::N * 4))) return rc;
    if ((rc = w.multi.ensure((max_multi + 1) * 4))) return rc;
  }
  if ((rc = g_ws.buckets.ensure(nb * 4 * P::N * 4))) return rc;
  if ((rc = g_ws.flag.ensure(4))) return rc;
  // buckets per reduce thread: each thread walks a chain of 2*CH dependent XYZZ additions (~7 us each when a
  // warp runs alone), so small bucket sets get short chunks -- enough threads to fill the SMs matters more than
  // the ~20-addition fix-up (lo * running) every chunk pays
  uint32_t CH = 64;
  while (CH > 8 && (nb / CH) < 32768) CH >>= 1;        // 2^21 buckets keep CH = 64 (work-bound there), 2^19 get 16
  if (CH > B) CH = B;
  const uint32_t cpw = (B + CH - 1) / CH;
  if ((rc = g_ws.partials.ensure((size_t)cpw * Wb * 4 * P::N * 4))) return rc;
  // tree sum of the per-chunk partials: kTreeSpan partials per block and level (128 threads: span / 128 serial additions, then 7
  // tree levels); 256 keeps the serial part at 2 additions -- the phase is latency-bound
  constexpr uint32_t kTreeSpan = 256;
  const size_t lvl1 = kz_div_up(cpw, kTreeSpan);
  if ((rc = g_ws.winsums.ensure(2 * (size_t)Wb * (lvl1 + 1) * 4 * P::N * 4))) return rc;

  uint32_t* flag = (uint32_t*)g_ws.flag.p;
  KZ_CUDA(cudaMemsetAsync(flag, 0, 4, st));
  if (piped) {                                           // the sort stream starts after everything queued so far
    KZ_CUDA(cudaEventRecord(cx.start_ev, st));
    KZ_CUDA(cudaStreamWaitEvent(sst, cx.start_ev, 0));
  }
  SortGeom geo;
  geo.c = c; geo.W = W; geo.B = B; geo.f = f; geo.ncoarse = ncoarse; geo.tabled = tabled ? 1u : 0u;
  geo.first = (uint32_t)first; geo.n_srs = (uint32_t)srs.n;
  geo.top_bits = R::BITS - 224;                  // bits allowed in the top 32-bit word
  geo.batch_len = batch > 1 ? (uint32_t)(n / batch) : 0u;
  DigitOffset doff;
  for (int k = 0; k < 8; k++) doff.w[k] = 0;
  for (uint32_t w = 0; w + 1 < W; w++) {
    uint32_t bit = w * c + (c - 1);
    if (bit < 256) doff.w[bit >> 5] |= 1u << (bit & 31);
  }
  static const bool tiny_off = getenv("KZGPU_MSM_NO_TINY") != nullptr;
  if (!tiny_off && tabled && nchunks == 1 && n && (n / batch) * (size_t)W <= kTinyEntries && !cx.profile) {
    // tiny MSM: every (point, window) term by double-and-add, two launches (see msm_tiny_kernel)
    const uint32_t poly_len = (uint32_t)(n / batch);
    const uint32_t blocks = (uint32_t)kz_div_up((size_t)poly_len * W, kTinyThreads);
    if ((rc = g_ws.partials.ensure((size_t)blocks * batch * 4 * P::N * 4))) return rc;
    if ((rc = g_ws.winsums.ensure((size_t)batch * 4 * P::N * 4))) return rc;
    if (h_scalars) KZ_CUDA(cudaStreamWaitEvent(st, cx.copy_ev[0], 0));
    msm_tiny_kernel<Cfg><<<dim3(blocks, batch), kTinyThreads, kTinyThreads * 4 * P::N * 4, st>>>(
        srs.d_points, d_scalars, poly_len, (uint32_t)first, (uint32_t)srs.n, c, W, geo.top_bits, doff, flag, (uint32_t*)g_ws.partials.p);
    KZ_LAUNCHED();
    msm_tiny_fold_kernel<Cfg><<<batch, 64, 64 * 4 * P::N * 4, st>>>((uint32_t*)g_ws.partials.p, blocks, (uint32_t*)g_ws.winsums.p);
    KZ_LAUNCHED();
    if (mode == 1) {                                       // affine on the device (rare): the window fold kernel with one window
      msm_final_kernel<Cfg><<<batch, 32, 0, st>>>((uint32_t*)g_ws.winsums.p, 1, c, 1, d_out);
      KZ_LAUNCHED();
    } else {
      KZ_CUDA(cudaMemcpyAsync(d_out, g_ws.winsums.p, (size_t)batch * 4 * P::N * 4, cudaMemcpyDeviceToDevice, st));
    }
    if (defer && mode == 0) { g_flag_pending[kz_slot()] = true; return 0; }
    uint32_t hflag = 0;
    KZ_CUDA(cudaMemcpyAsync(&hflag, flag, 4, cudaMemcpyDeviceToHost, st));
    if (mode == 2 && h_out_xyzz) KZ_CUDA(cudaMemcpyAsync(h_out_xyzz, g_ws.winsums.p, (size_t)batch * 4 * P::N * 4, cudaMemcpyDeviceToHost, st));
    KZ_CUDA(cudaStreamSynchronize(st));
    if (hflag) return kz_fail(KZGPU_ERANGE, "a scalar is not a canonical residue (>= 2^%d)", R::BITS);
    return 0;
  }
  const uint32_t* d_scalars_all = d_scalars;
  const size_t n_all = n, first_all = first;
  (void)n_all;
  for (uint32_t chunk = 0; chunk < nchunks; chunk++) {
  const size_t c_lo = chunk_lo[chunk];
  n = chunk_lo[chunk + 1] - c_lo;
  if (chunk && !n) break;
  d_scalars = d_scalars_all + c_lo * 8;
  first = first_all + c_lo;
  geo.first = (uint32_t)first;
  const uint32_t add_existing = chunk ? 1u : 0u;
  MsmSortWs& w = g_sw[chunk & 1];
  uint32_t* counts = (uint32_t*)w.counts.p;
  uint32_t* offsets = (uint32_t*)w.offsets.p;
  uint32_t* cursor = (uint32_t*)w.cursor.p;
  uint32_t* entries = (uint32_t*)w.entries.p;
  uint64_t* sort_tmp = (uint64_t*)w.sort_tmp.p;
  uint32_t* coarse_counts = (uint32_t*)w.coarse.p;
  uint32_t* coarse_off = coarse_counts + kMaxCoarse;        // ncoarse + 1 entries
  uint32_t* coarse_cur = coarse_off + kMaxCoarse + 1;
  uint32_t* blk_off = coarse_cur + kMaxCoarse;              // ncoarse + 1 entries
  uint32_t* blocksums = (uint32_t*)w.blocksums.p;
  // ---- sort + tasks of this chunk (stream sst)
  if (h_scalars) KZ_CUDA(cudaStreamWaitEvent(sst, cx.copy_ev[chunk], 0));
  if (piped && chunk >= 2) KZ_CUDA(cudaStreamWaitEvent(sst, cx.acc_ev[chunk & 1], 0));   // workspace still read by chunk - 2
  KZ_CUDA(cudaMemsetAsync(counts, 0, nb * 4, sst));
  KZ_CUDA(cudaMemsetAsync(coarse_counts, 0, kMaxCoarse * 4, sst));
  KzProf prof_sort(2);
  if (n) {
    const unsigned tiles = (unsigned)kz_div_up(n, kSortTile);
    msm_coarse_hist_kernel<<<tiles, 256, ncoarse * 4, sst>>>(d_scalars, n, geo, doff, coarse_counts, flag);
    KZ_LAUNCHED();
    msm_coarse_scan_kernel<<<1, 1024, 0, sst>>>(coarse_counts, ncoarse, coarse_off, coarse_cur, blk_off);
    KZ_LAUNCHED();
    const uint32_t ptile = kSortStage / W;                       // scalars per partition block
    const size_t psmem = (size_t)ptile * W * 8 + (2 * (size_t)ncoarse + 512) * 4;
    msm_partition_kernel<<<(unsigned)kz_div_up(n, ptile), 512, psmem, sst>>>(d_scalars, n, ptile, geo, doff, coarse_cur, sort_tmp);
    KZ_LAUNCHED();
    const size_t sort_blocks = (n * W) / kSortChunk + ncoarse + 1;
    msm_fine_hist_kernel<<<(unsigned)sort_blocks, 256, (4u << f), sst>>>(sort_tmp, coarse_off, blk_off, ncoarse, f, (uint32_t)nb, counts);
    KZ_LAUNCHED();
  }
  scan_block_kernel<<<(unsigned)nblk, 256, 0, sst>>>(counts, offsets, blocksums, nb);
  KZ_LAUNCHED();
  scan_sums_kernel<<<1, 256, 0, sst>>>(blocksums, nblk, offsets + nb);
  KZ_LAUNCHED();
  scan_add_kernel<<<(unsigned)kz_div_up(nb, 256), 256, 0, sst>>>(offsets, blocksums, nb, cursor);
  KZ_LAUNCHED();
  if (n) {
    const size_t sort_blocks = (n * W) / kSortChunk + ncoarse + 1;
    msm_fine_scatter_kernel<<<(unsigned)sort_blocks, 512, ((8u << f) + (512 + kSortChunk) * 4), sst>>>(sort_tmp, coarse_off, blk_off, ncoarse, f, (uint32_t)nb, cursor,
                                                                             entries);
    KZ_LAUNCHED();
  }
  const uint32_t ostride = 1u;                          // offsets[] entries per bucket
  const uint32_t T = task_T(n);
  uint32_t* multi_count = (uint32_t*)w.multi.p;
  uint32_t* multi_list = multi_count + 1;
  KZ_CUDA(cudaMemsetAsync(multi_count, 0, 4, sst));
  uint32_t* ntasks = (uint32_t*)w.ntasks.p;
  uint32_t* task_off = (uint32_t*)w.task_off.p;
  uint32_t* size_hist = (uint32_t*)w.size_hist.p;
  uint32_t* size_cursor = size_hist + (1024 + 1);
  KZ_CUDA(cudaMemsetAsync(size_hist, 0, (1024 + 1) * 4, sst));
  task_count_kernel<<<(unsigned)kz_div_up(nb, 256), 256, (T + 1) * 4, sst>>>(offsets, ostride, (uint32_t)nb, T, ntasks, size_hist, multi_count, multi_list);
  KZ_LAUNCHED();
  scan_block_kernel<<<(unsigned)nblk, 256, 0, sst>>>(ntasks, task_off, blocksums, nb);
  KZ_LAUNCHED();
  scan_sums_kernel<<<1, 256, 0, sst>>>(blocksums, nblk, task_off + nb);
  KZ_LAUNCHED();
  scan_add_kernel<<<(unsigned)kz_div_up(nb, 256), 256, 0, sst>>>(task_off, blocksums, nb, nullptr);
  KZ_LAUNCHED();
  task_size_scan_kernel<<<1, 32, 0, sst>>>(size_hist, T, size_cursor);
  KZ_LAUNCHED();
  // size_cursor[0] holds the task total and is not used as a cursor (no task has length 0)
  task_emit_kernel<<<(unsigned)kz_div_up(nb, 256), 256, 0, sst>>>(offsets, ostride, ntasks, task_off, (uint32_t)nb, T, size_cursor,
                                                                (uint32_t*)w.t_start.p, (uint32_t*)w.t_len.p,
                                                                (uint32_t*)w.t_dest.p);
  KZ_LAUNCHED();
  prof_sort.stop(chunk == 0 ? 1 : 0, (double)n);
  if (piped) {
    KZ_CUDA(cudaEventRecord(cx.sort_ev[chunk & 1], sst));
    KZ_CUDA(cudaStreamWaitEvent(st, cx.sort_ev[chunk & 1], 0));
  }
  // ---- accumulate + merge into the shared buckets (main stream)
  const size_t chunk_tasks = nb + ((size_t)n * W) / T + 1;
  KzProf prof_acc(0);
  msm_accumulate_kernel<Cfg><<<(unsigned)kz_div_up(chunk_tasks, 128), 128, 0, st>>>(
      srs.d_points, entries, (uint32_t*)w.t_start.p, (uint32_t*)w.t_len.p, (uint32_t*)w.t_dest.p, size_cursor,
      add_existing, (uint32_t*)g_ws.buckets.p, (uint32_t*)w.tparts.p);
  KZ_LAUNCHED();
  prof_acc.stop(chunk == 0 ? 1 : 0, (double)n * W);
  KzProf prof_merge(3);
  msm_merge_kernel<Cfg><<<cx.sm_count * 4, 128, 0, st>>>(ntasks, task_off, multi_count, multi_list, add_existing,
                                                        (uint32_t*)w.tparts.p, (uint32_t*)g_ws.buckets.p);
  KZ_LAUNCHED();
  if (!add_existing) {
    msm_clear_empty_kernel<Cfg><<<(unsigned)kz_div_up(nb, 256), 256, 0, st>>>(ntasks, (uint32_t)nb, (uint32_t*)g_ws.buckets.p);
    KZ_LAUNCHED();
  }
  prof_merge.stop(0, 0.0);
  if (piped) KZ_CUDA(cudaEventRecord(cx.acc_ev[chunk & 1], st));
  if (h_scalars && chunk + 1 < nchunks) { int r = upload_chunk(chunk + 1); if (r) return r; }
  }  // chunks
  KzProf prof_red(3);
  msm_reduce_kernel<Cfg><<<(unsigned)kz_div_up((size_t)cpw * Wb, 128), 128, 0, st>>>((uint32_t*)g_ws.buckets.p, B, CH, cpw, Wb,
                                                                                    (uint32_t*)g_ws.partials.p);
  KZ_LAUNCHED();
  const uint32_t* lvl_in = (const uint32_t*)g_ws.partials.p;
  uint32_t count = cpw, lvl_launches = 0;
  uint32_t* pong[2] = {(uint32_t*)g_ws.winsums.p, (uint32_t*)g_ws.winsums.p + (size_t)Wb * (lvl1 + 1) * 4 * P::N};
  while (count > 1) {
    uint32_t blocks = (count + kTreeSpan - 1) / kTreeSpan;
    uint32_t* lvl_out = pong[lvl_launches & 1];
    msm_window_kernel<Cfg><<<dim3(blocks, Wb), 128, 128 * 4 * P::N * 4, st>>>(lvl_in, count, kTreeSpan, lvl_out);
    KZ_LAUNCHED();
    lvl_in = lvl_out; count = blocks; lvl_launches++;
  }
  // mode 2: the caller normalises on the host (one 256/384-bit inversion is ~20 us of CPU but ~200 us for a lone
  // GPU thread); only the window fold stays on the device
  msm_final_kernel<Cfg><<<batch, 32, 0, st>>>(lvl_in, Wp, c, mode == 2 ? 0 : mode, d_out);
  KZ_LAUNCHED();
  prof_red.stop(1, (double)nb);
  if (defer && mode == 0 && !cx.profile) { g_flag_pending[kz_slot()] = true; return 0; }
  uint32_t hflag = 0;
  KZ_CUDA(cudaMemcpyAsync(&hflag, flag, 4, cudaMemcpyDeviceToHost, st));
  if (mode == 2 && h_out_xyzz) KZ_CUDA(cudaMemcpyAsync(h_out_xyzz, d_out, (size_t)batch * 4 * P::N * 4, cudaMemcpyDeviceToHost, st));
  KZ_CUDA(cudaStreamSynchronize(st));
  if (hflag) return kz_fail(KZGPU_ERANGE, "a scalar is not a canonical residue (>= 2^%d)", R::BITS);
  return 0;
}

// a^-1 mod p for a Montgomery-form element, on the host: binary extended Euclid on the raw limbs
// (inverse of a*R is a^-1 R^-1), then two Montgomery multiplications by R^2 bring it back to a^-1 R.
template <class P> Fe<P> host_fe_inv(const Fe<P>& a) {
  constexpr int N = P::N;
  auto is_one = [](const uint32_t* x) { for (int i = 1; i <= N; i++) if (x[i]) return false; return x[0] == 1; };
  auto geq = [](const uint32_t* x, const uint32_t* y) { for (int i = N; i >= 0; i--) { if (x[i] != y[i]) return x[i] > y[i]; } return true; };
  auto sub = [](uint32_t* x, const uint32_t* y) { uint64_t br = 0; for (int i = 0; i <= N; i++) { uint64_t d = (uint64_t)x[i] - y[i] - br; x[i] = (uint32_t)d; br = (d >> 63) & 1; } };
  auto add = [](uint32_t* x, const uint32_t* y) { uint64_t c = 0; for (int i = 0; i <= N; i++) { c += (uint64_t)x[i] + y[i]; x[i] = (uint32_t)c; c >>= 32; } };
  auto shr = [](uint32_t* x) { for (int i = 0; i < N; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 31); x[N] >>= 1; };
  uint32_t p[N + 1], u[N + 1], v[N + 1], x1[N + 1] = {0}, x2[N + 1] = {0};
  for (int i = 0; i < N; i++) { p[i] = P::mod(i); u[i] = a.v[i]; v[i] = p[i]; }
  p[N] = u[N] = v[N] = 0;
  x1[0] = 1;
  bool zero = true;
  for (int i = 0; i < N; i++) zero = zero && !u[i];
  if (zero) return a;
  while (!is_one(u) && !is_one(v)) {
    while (!(u[0] & 1)) { shr(u); if (x1[0] & 1) add(x1, p); shr(x1); }
    while (!(v[0] & 1)) { shr(v); if (x2[0] & 1) add(x2, p); shr(x2); }
    if (geq(u, v)) { sub(u, v); if (!geq(x1, x2)) add(x1, p); sub(x1, x2); }
    else { sub(v, u); if (!geq(x2, x1)) add(x2, p); sub(x2, x1); }
  }
  const uint32_t* res = is_one(u) ? x1 : x2;
  Fe<P> r;
  for (int i = 0; i < N; i++) r.v[i] = res[i];
  return fe_to_mont<P>(fe_to_mont<P>(r));
}

// XYZZ (Montgomery) -> canonical affine limbs + infinity flag, on the host
template <class P> void host_xyzz_to_canonical(const uint32_t* h, uint32_t* out_xy, int* is_inf) {
  XYZZ<P> a;
  memcpy(a.x.v, h, 4 * P::N * 4);
  if (xyzz_is_inf<P>(a)) { memset(out_xy, 0, 2 * P::N * 4); if (is_inf) *is_inf = 1; return; }
  Fe<P> t = host_fe_inv<P>(fe_mul<P>(a.zz, a.zzz));
  Fe<P> x = fe_from_mont<P>(fe_mul<P>(a.x, fe_mul<P>(t, a.zzz)));
  Fe<P> y = fe_from_mont<P>(fe_mul<P>(a.y, fe_mul<P>(t, a.zz)));
  memcpy(out_xy, x.v, P::N * 4);
  memcpy(out_xy + P::N, y.v, P::N * 4);
  if (is_inf) *is_inf = 0;
}

template <class Cfg>
int msm_affine(const SrsPart& srs, size_t first, const uint32_t* d_scalars, size_t n, uint64_t* out_xy, int* is_inf,
               const uint64_t* h_scalars = nullptr) {
  using P = typename Cfg::Fp;
  int rc;
  if ((rc = g_ws.result.ensure((4 * P::N + 1) * 4))) return rc;
  uint32_t h[4 * 12];
  if ((rc = msm_core<Cfg>(srs, first, d_scalars, n, 2, (uint32_t*)g_ws.result.p, h_scalars, h))) return rc;
  host_xyzz_to_canonical<P>(h, (uint32_t*)out_xy, is_inf);
  return 0;
}

// k equal-length polynomials in one pass -> k canonical affine points
template <class Cfg>
int msm_affine_batch(const SrsPart& srs, const uint32_t* d_scalars, size_t poly_len, size_t k, uint64_t* out_xy, int* is_inf) {
  using P = typename Cfg::Fp;
  int rc;
  if ((rc = g_ws.result.ensure(k * 4 * P::N * 4 + 4))) return rc;
  std::vector<uint32_t> h(k * 4 * P::N);
  if ((rc = msm_core<Cfg>(srs, 0, d_scalars, poly_len * k, 2, (uint32_t*)g_ws.result.p, nullptr, h.data(), (uint32_t)k))) return rc;
  for (size_t j = 0; j < k; j++)
    host_xyzz_to_canonical<P>(&h[j * 4 * P::N], (uint32_t*)out_xy + j * 2 * P::N, is_inf ? is_inf + j : nullptr);
  return 0;
}

// dynamic shared memory limits are per device: once per slot
int set_smem_attrs() {
  static bool done[KZ_MAX_DEV] = {false};
  if (done[kz_slot()]) return 0;
  KZ_CUDA(cudaFuncSetAttribute(msm_window_kernel<BLS381Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 4 * 12 * 4));
  KZ_CUDA(cudaFuncSetAttribute(g1_fold_kernel<BLS381Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 4 * 12 * 4));
  KZ_CUDA(cudaFuncSetAttribute(msm_tiny_kernel<BLS381Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTinyThreads * 4 * 12 * 4));
  KZ_CUDA(cudaFuncSetAttribute(msm_partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortStage * 8 + (2 * kMaxCoarse + 512) * 4));
  KZ_CUDA(cudaFuncSetAttribute(msm_fine_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (8 << 13) + (512 + kSortChunk) * 4));
  done[kz_slot()] = true;
  return 0;
}

Srs* find_srs(uint64_t handle) {
  auto it = g_srs.find(handle);
  return it == g_srs.end() ? nullptr : &it->second;
}

// allocate one part of n points on the calling thread's device (W tables when enabled); build_tables() fills tables
// 1..W-1 from table 0
template <class Cfg>
int part_alloc(SrsPart& s, size_t first, size_t n) {
  using P = typename Cfg::Fp;
  using R = typename Cfg::Fr;
  s.first = first; s.n = n; s.d_points = nullptr; s.slot = kz_slot();
  const size_t pt_bytes = 2 * P::N * 4;
  s.c_tab = n ? choose_table_c(n, R::BITS, pt_bytes) : 0;
  s.W_tab = s.c_tab ? (R::BITS + 1 + s.c_tab - 1) / s.c_tab : 1;
  size_t bytes = n * pt_bytes * s.W_tab;
  cudaError_t e = cudaMalloc((void**)&s.d_points, bytes ? bytes : 16);
  if (e != cudaSuccess && s.c_tab) {          // not enough memory for the tables: plain key
    cudaGetLastError();
    s.c_tab = 0; s.W_tab = 1;
    bytes = n * pt_bytes;
    e = cudaMalloc((void**)&s.d_points, bytes ? bytes : 16);
  }
  if (e != cudaSuccess) return kz_fail(KZGPU_ECUDA, "cudaMalloc(%zu) for the SRS failed: %s", bytes, cudaGetErrorString(e));
  return 0;
}

template <class Cfg>
int part_build_tables(SrsPart& s) {
  using P = typename Cfg::Fp;
  KzgpuCtx& cx = kz_ctx();
  const size_t stride = s.n * 2 * P::N;
  if (s.c_tab && s.W_tab > 1 && s.n <= ((size_t)1 << 18) && !getenv("KZGPU_SRS_TABLES_PER_WINDOW")) {
    // short keys: every table in one launch, one inversion per point (srs_tables_fused_kernel)
    uint32_t* scratch = nullptr;
    const size_t sbytes = (size_t)(s.W_tab - 1) * s.n * 3 * P::N * 4;
    if (cudaMalloc((void**)&scratch, sbytes) == cudaSuccess) {
      srs_tables_fused_kernel<Cfg><<<(unsigned)kz_div_up(s.n, 128), 128, 0, cx.stream>>>(s.d_points, s.n, s.c_tab, s.W_tab, scratch);
      KZ_LAUNCHED();
      KZ_CUDA(cudaStreamSynchronize(cx.stream));
      cudaFree(scratch);
      return 0;
    }
    cudaGetLastError();                                   // no room for the scratch: table by table
  }
  for (uint32_t w = 1; w < s.W_tab && s.c_tab; w++) {
    srs_table_kernel<Cfg><<<(unsigned)kz_div_up(s.n, 128), 128, 0, cx.stream>>>(s.d_points + (w - 1) * stride, s.d_points + w * stride,
                                                                              s.n, s.c_tab);
    KZ_LAUNCHED();
  }
  return 0;
}

template <class Cfg>
int srs_create_impl(const uint64_t* affine_xy, size_t n, uint64_t* handle) {
  using P = typename Cfg::Fp;
  KzgpuCtx& cx = kz_ctx();
  Srs s;
  s.curve = Cfg::id; s.n = n;
  int rc0 = part_alloc<Cfg>(s.full, 0, n);
  if (rc0) return rc0;
  size_t bytes = n * 2 * P::N * 4;
  if (n) {
    KZ_CUDA(cudaMemcpyAsync(s.full.d_points, affine_xy, bytes, cudaMemcpyHostToDevice, cx.stream));
    srs_to_mont_kernel<Cfg><<<(unsigned)kz_div_up(2 * n, 128), 128, 0, cx.stream>>>(s.full.d_points, n);
    KZ_LAUNCHED();
    int rc1 = part_build_tables<Cfg>(s.full);
    if (rc1) return rc1;
    KZ_CUDA(cudaStreamSynchronize(cx.stream));
  }
  *handle = g_next_handle++;
  g_srs[*handle] = s;
  return 0;
}

template <class Cfg>
int srs_generate_impl(const uint64_t* tau, size_t start, size_t n, uint64_t* handle) {
  using P = typename Cfg::Fp;
  using R = typename Cfg::Fr;
  KzgpuCtx& cx = kz_ctx();
  Fe<R> t = kz_fe_from_u64<R>(tau);
  if (!kz_fe_reduced<R>(t)) return kz_fail(KZGPU_ERANGE, "tau is not a canonical scalar");
  // host: table of 2^j * G1 in affine Montgomery form (O(bits) work, done once per call)
  std::vector<uint32_t> table((size_t)R::BITS * 2 * P::N);
  Affine<P> g;
  for (int i = 0; i < P::N; i++) { g.x.v[i] = P::gx_mont(i); g.y.v[i] = P::gy_mont(i); }
  for (int j = 0; j < R::BITS; j++) {
    memcpy(&table[(size_t)j * 2 * P::N], g.x.v, P::N * 4);
    memcpy(&table[(size_t)j * 2 * P::N + P::N], g.y.v, P::N * 4);
    g = xyzz_to_affine<P>(xyzz_dbl_affine<P>(g));
  }
  uint32_t* d_table = nullptr;
  KZ_CUDA(cudaMalloc((void**)&d_table, table.size() * 4));
  KZ_CUDA(cudaMemcpyAsync(d_table, table.data(), table.size() * 4, cudaMemcpyHostToDevice, cx.stream));
  Srs s;
  s.curve = Cfg::id; s.n = n;
  int rc0 = part_alloc<Cfg>(s.full, 0, n);
  if (rc0) return rc0;
  if (n) {
    srs_generate_kernel<Cfg><<<(unsigned)kz_div_up(n, 128), 128, 0, cx.stream>>>(s.full.d_points, start, n, fe_to_mont<R>(t), d_table);
    KZ_LAUNCHED();
    int rc1 = part_build_tables<Cfg>(s.full);
    if (rc1) return rc1;
  }
  KZ_CUDA(cudaStreamSynchronize(cx.stream));
  cudaFree(d_table);
  *handle = g_next_handle++;
  g_srs[*handle] = s;
  return 0;
}

void srs_free(Srs& s) {
  cudaFree(s.full.d_points);
  if (s.shard_sets) {
    for (auto& kv : *s.shard_sets)
      for (auto& part : kv.second) if (part.d_points) cudaFree(part.d_points);
    delete s.shard_sets;
    s.shard_sets = nullptr;
  }
  for (int d = 1; d < KZ_MAX_DEV; d++)
    if (s.has_replicas && s.replica[d].d_points) cudaFree(s.replica[d].d_points);
}

// ---------------------------------------------------------------- several devices (kzgpu_init_multi)
size_t shard_min() {
  static size_t v = 0;
  if (!v) {
    const char* env = getenv("KZGPU_SHARD_MIN");
    v = env && atoll(env) > 0 ? (size_t)atoll(env) : (size_t)1 << 20;
  }
  return v;
}

// Point shards of the length class E (the index range [0, E) of the key): slot d owns [d E / N, (d + 1) E / N).  Its plain
// points come from the primary device's table 0 by a peer copy (NVLink when peer access is on); the window tables are built
// locally, for the SHARD's size -- a 2^21-point shard of a 2^24-point key gets c = 20 (2^19 buckets to reduce per MSM), not
// the key's c = 22.  Returns the set through *out.
template <class Cfg>
int ensure_shards(Srs& s, size_t end, std::vector<SrsPart>** out) {
  using P = typename Cfg::Fp;
  size_t E = 1;
  while (E < end) E <<= 1;
  if (E > s.n) E = s.n;
  if (!s.shard_sets) s.shard_sets = new std::map<size_t, std::vector<SrsPart>>();
  auto it = s.shard_sets->find(E);
  if (it != s.shard_sets->end()) { *out = &it->second; return 0; }
  const int nd = kz_ndev();
  std::vector<SrsPart> parts(nd);
  KZ_CUDA(cudaStreamSynchronize(kz_ctx_of(0).stream));
  const int dev0 = kz_device_of(0);
  const size_t pt_words = 2 * P::N;
  int rc = kz_parallel([&](int slot) -> int {
    const size_t base = E / nd, rem = E % nd;
    const size_t first = slot * base + ((size_t)slot < rem ? slot : rem), cnt = base + ((size_t)slot < rem ? 1 : 0);
    SrsPart& sh = parts[slot];
    int r = part_alloc<Cfg>(sh, first, cnt);
    if (r) return r;
    if (!cnt) return 0;
    KzgpuCtx& cx = kz_ctx();
    KZ_CUDA(cudaMemcpyPeerAsync(sh.d_points, kz_device_of(slot), s.full.d_points + first * pt_words, dev0, cnt * pt_words * 4, cx.stream));
    if ((r = part_build_tables<Cfg>(sh))) return r;
    KZ_CUDA(cudaStreamSynchronize(cx.stream));
    return 0;
  });
  if (rc) {
    for (auto& part : parts) if (part.d_points) cudaFree(part.d_points);
    return rc;
  }
  *out = &(*s.shard_sets)[E];
  **out = parts;
  return 0;
}

// One device: a polynomial much shorter than the key (Marlin at 2^20 rows commits 1-2 M coefficients against a 12.6 M-point key)
// would inherit the KEY's window -- c = 22, 2^21 buckets to reduce for ~7 entries per bucket.  The same length-class mechanism
// gives it a table set of its own: the head [0, E) of the key with the window an E-point MSM wants (built once per class,
// at most twice the key's tables in total).  Returns the key itself when the polynomial is not at least 4x shorter.
template <class Cfg>
const SrsPart* part_for_length(Srs& s, size_t end, int* rc) {
  *rc = 0;
  static const bool off = getenv("KZGPU_NO_CLASS_TABLES") != nullptr;
  if (off || kz_ndev() != 1 || !s.full.c_tab || end < ((size_t)1 << 14) || end * 4 > s.n) return &s.full;
  std::vector<SrsPart>* set = nullptr;
  if ((*rc = ensure_shards<Cfg>(s, end, &set))) return nullptr;
  return (*set)[0].c_tab ? &(*set)[0] : &s.full;
}

// Replicas: the whole key with its window tables on every device (whole polynomials of a batched commit run where they
// are placed): one peer copy of the primary's tables per device.
template <class Cfg>
int ensure_replicas(Srs& s) {
  using P = typename Cfg::Fp;
  if (s.has_replicas) return 0;
  KZ_CUDA(cudaStreamSynchronize(kz_ctx_of(0).stream));
  const int dev0 = kz_device_of(0);
  const size_t bytes = s.n * 2 * P::N * 4 * s.full.W_tab;
  s.replica[0] = s.full;
  int rc = kz_parallel([&](int slot) -> int {
    if (slot == 0) return 0;
    SrsPart& rp = s.replica[slot];
    rp = s.full;
    rp.slot = slot; rp.d_points = nullptr;
    KZ_CUDA(cudaMalloc((void**)&rp.d_points, bytes ? bytes : 16));
    KzgpuCtx& cx = kz_ctx();
    KZ_CUDA(cudaMemcpyPeerAsync(rp.d_points, kz_device_of(slot), s.full.d_points, dev0, bytes, cx.stream));
    KZ_CUDA(cudaStreamSynchronize(cx.stream));
    return 0;
  });
  if (rc) return rc;
  s.has_replicas = true;
  return 0;
}

// One MSM over the index range [first, first + n) of the key, point-sharded over every device.  Scalars: host memory
// (each device uploads its own slice, chunked and overlapped with its compute) or device memory of the primary device
// (each peer pulls its slice over NVLink).  Every device leaves one XYZZ partial in the primary's gather buffer (peer
// write); the primary folds them and the host normalises the one result.
template <class Cfg>
int msm_sharded(Srs& s, size_t first, size_t n, const uint64_t* h_scalars, const uint32_t* d0_scalars, uint64_t* out_xy, int* is_inf) {
  using P = typename Cfg::Fp;
  const int nd = kz_ndev();
  int rc;
  std::vector<SrsPart>* shards = nullptr;
  if ((rc = ensure_shards<Cfg>(s, first + n, &shards))) return rc;
  const size_t xyzz_bytes = 4 * P::N * 4;
  MsmWs& w0 = g_ws_slots.of(0);
  if ((rc = w0.gather.ensure(nd * xyzz_bytes))) return rc;
  if ((rc = w0.result.ensure((4 * P::N + 1) * 4 + 64))) return rc;
  KzgpuCtx& cx0 = kz_ctx_of(0);
  KZ_CUDA(cudaStreamSynchronize(cx0.stream));            // device-resident scalars may still be in flight on the primary's stream
  const int dev0 = kz_device_of(0);
  uint32_t* gather = (uint32_t*)w0.gather.p;
  rc = kz_parallel([&](int slot) -> int {
    const SrsPart& sh = (*shards)[slot];
    const size_t lo = first > sh.first ? first : sh.first;
    const size_t hi_all = first + n, hi_sh = sh.first + sh.n;
    const size_t hi = hi_all < hi_sh ? hi_all : hi_sh;
    KzgpuCtx& cx = kz_ctx();
    MsmWs& w = g_ws;
    int r;
    if ((r = set_smem_attrs())) return r;
    if ((r = w.result.ensure(xyzz_bytes + 64))) return r;
    uint32_t* d_part = (uint32_t*)w.result.p;
    if (hi <= lo) {
      KZ_CUDA(cudaMemsetAsync(d_part, 0, xyzz_bytes, cx.stream));          // ZZ = 0: the identity
    } else {
      const size_t cnt = hi - lo;
      if ((r = w.scal.ensure(cnt * 32 + 32))) return r;
      const uint32_t* d_sc = (const uint32_t*)w.scal.p;
      const uint64_t* h_sc = nullptr;
      if (h_scalars) h_sc = h_scalars + (lo - first) * 4;
      else if (slot == 0) d_sc = d0_scalars + (lo - first) * 8;
      else KZ_CUDA(cudaMemcpyPeerAsync(w.scal.p, kz_device_of(slot), d0_scalars + (lo - first) * 8, dev0, cnt * 32, cx.stream));
      if ((r = msm_core<Cfg>(sh, lo - sh.first, d_sc, cnt, 0, d_part, h_sc, nullptr, 1, true))) return r;
    }
    KZ_CUDA(cudaMemcpyPeerAsync(gather + (size_t)slot * 4 * P::N, dev0, d_part, kz_device_of(slot), xyzz_bytes, cx.stream));
    uint32_t hflag = 0;
    if ((r = msm_flag_collect(&hflag))) return r;
    KZ_CUDA(cudaStreamSynchronize(cx.stream));
    return msm_flag_result(hflag);
  });
  if (rc) return rc;
  g1_fold_kernel<Cfg><<<1, 128, 128 * 4 * P::N * 4, cx0.stream>>>(gather, (uint32_t)nd, (uint32_t*)w0.result.p);
  KZ_LAUNCHED();
  uint32_t h[4 * 12];
  KZ_CUDA(cudaMemcpyAsync(h, w0.result.p, 4 * P::N * 4, cudaMemcpyDeviceToHost, cx0.stream));
  KZ_CUDA(cudaStreamSynchronize(cx0.stream));
  host_xyzz_to_canonical<P>(h, (uint32_t*)out_xy, is_inf);
  return 0;
}

// can the k polynomials share one pass?  (bucket sets and digit entries must fit the sort's key and index ranges)
bool batch_fits(const SrsPart& s, int curve, size_t poly_len, size_t k) {
  if (k < 2 || k > 64 || poly_len == 0) return false;
  const int bits = curve == KZGPU_BN254 ? FrBN254::BITS : FrBLS381::BITS;
  const uint32_t c = s.c_tab ? s.c_tab : choose_c(poly_len, bits);
  const uint32_t W = (bits + 1 + c - 1) / c;
  const size_t nb = (size_t)(s.c_tab ? 1u : W) * k << (c - 1);
  return nb <= ((size_t)kMaxCoarse << 13) && (size_t)poly_len * k * W < 0xffffffffull && poly_len * k < 0x7fffffffull;
}

// k polynomials of poly_len scalars each, back to back on the calling thread's device, against that device's copy of the key
int msm_batch_on_part(const SrsPart& part, int curve, const uint32_t* d_scalars, size_t poly_len, size_t k, uint64_t* out_xy, int* is_inf) {
  const int L = kzgpu_fp_limbs64(curve);
  int rc = set_smem_attrs();
  if (rc) return rc;
  if (!batch_fits(part, curve, poly_len, k)) {                 // one pass per polynomial
    for (size_t j = 0; j < k; j++) {
      int* fl = is_inf ? is_inf + j : nullptr;
      rc = curve == KZGPU_BN254 ? msm_affine<BN254Cfg>(part, 0, d_scalars + j * poly_len * 8, poly_len, out_xy + j * 2 * L, fl)
                                : msm_affine<BLS381Cfg>(part, 0, d_scalars + j * poly_len * 8, poly_len, out_xy + j * 2 * L, fl);
      if (rc) return rc;
    }
    return 0;
  }
  if (curve == KZGPU_BN254) return msm_affine_batch<BN254Cfg>(part, d_scalars, poly_len, k, out_xy, is_inf);
  return msm_affine_batch<BLS381Cfg>(part, d_scalars, poly_len, k, out_xy, is_inf);
}

// `count` host polynomials (ptrs / lens) of one commit() call on the calling thread's device: uploaded zero padded to a
// common length and committed in one pass (zero scalars emit no digits); results to outs[j] / infs[j]
int msm_batch_host_on_part(const SrsPart& part, int curve, const uint64_t* const* ptrs, const size_t* lens, size_t count,
                           uint64_t* const* outs, int* const* infs) {
  if (!count) return 0;
  KzgpuCtx& cx = kz_ctx();
  const int L = kzgpu_fp_limbs64(curve);
  size_t maxlen = 0, total = 0;
  for (size_t j = 0; j < count; j++) { total += lens[j]; if (lens[j] > maxlen) maxlen = lens[j]; }
  int rc = g_ws.scal.ensure(count * maxlen * 32 + 32);
  if (rc) return rc;
  uint32_t* d = (uint32_t*)g_ws.scal.p;
  if (total != count * maxlen) KZ_CUDA(cudaMemsetAsync(d, 0, count * maxlen * 32, cx.stream));
  for (size_t j = 0; j < count; j++)
    if (lens[j]) { int r = kz_upload(d + j * maxlen * 8, ptrs[j], lens[j] * 32, cx.stream); if (r) return r; }
  std::vector<uint64_t> out(count * 2 * L);
  std::vector<int> inf(count);
  if ((rc = set_smem_attrs())) return rc;
  if (maxlen == 0 || count == 1) {
    for (size_t j = 0; j < count; j++) {
      rc = curve == KZGPU_BN254 ? msm_affine<BN254Cfg>(part, 0, d + j * maxlen * 8, maxlen, &out[j * 2 * L], &inf[j])
                                : msm_affine<BLS381Cfg>(part, 0, d + j * maxlen * 8, maxlen, &out[j * 2 * L], &inf[j]);
      if (rc) return rc;
    }
  } else if ((rc = msm_batch_on_part(part, curve, d, maxlen, count, out.data(), inf.data()))) {
    return rc;
  }
  for (size_t j = 0; j < count; j++) {
    memcpy(outs[j], &out[j * 2 * L], 2 * L * 8);
    if (infs[j]) *infs[j] = inf[j];
  }
  return 0;
}

int upload_scalars(const uint64_t* scalars, size_t n, uint32_t** d) {
  int rc = g_ws.scal.ensure(n * 32 + 32);
  if (rc) return rc;
  if (n) KZ_CUDA(cudaMemcpyAsync(g_ws.scal.p, scalars, n * 32, cudaMemcpyHostToDevice, kz_ctx().stream));
  *d = (uint32_t*)g_ws.scal.p;
  return 0;
}

}  // namespace

// every slot releases its own workspaces (called on the slot's own thread); slot 0 also frees the keys
void kz_msm_release() {
  if (kz_slot() == 0) {
    for (auto& kv : g_srs) srs_free(kv.second);
    g_srs.clear();
  }
  MsmWs& ws = g_ws;
  KzScratch* all[] = {&ws.buckets, &ws.partials, &ws.winsums, &ws.result, &ws.flag, &ws.scal, &ws.gather};
  for (auto* s : all) s->release();
  for (auto& w : g_sw) {
    KzScratch* per[] = {&w.counts, &w.offsets, &w.cursor, &w.entries, &w.blocksums, &w.ntasks, &w.task_off, &w.size_hist,
                        &w.t_start, &w.t_len, &w.t_dest, &w.tparts, &w.multi, &w.sort_tmp, &w.coarse};
    for (auto* s : per) s->release();
  }
}

// used by poly.cu (open) and the MSM entry points: scalars on the primary device (or the host) against a handle.
// With several devices an MSM of >= KZGPU_SHARD_MIN points is point-sharded over all of them.
int kz_msm_dev_internal(uint64_t handle, size_t first, const uint32_t* d_scalars, size_t n, uint64_t* out_xy, int* is_inf,
                        const uint64_t* h_scalars) {
  Srs* s = find_srs(handle);
  if (!s) return kz_fail(KZGPU_EHANDLE, "unknown SRS handle %llu", (unsigned long long)handle);
  if (first + n > s->n)
    return kz_fail(KZGPU_ERANGE, "Polynomial degree %zu exceeds maximum allowed degree %zu", first + n - 1, s->n - 1);
  int rc = set_smem_attrs();
  if (rc) return rc;
  if (kz_ndev() > 1 && n >= shard_min()) {
    if (s->curve == KZGPU_BN254) return msm_sharded<BN254Cfg>(*s, first, n, h_scalars, d_scalars, out_xy, is_inf);
    return msm_sharded<BLS381Cfg>(*s, first, n, h_scalars, d_scalars, out_xy, is_inf);
  }
  if (s->curve == KZGPU_BN254) {
    const SrsPart* part = part_for_length<BN254Cfg>(*s, first + n, &rc);
    return part ? msm_affine<BN254Cfg>(*part, first, d_scalars, n, out_xy, is_inf, h_scalars) : rc;
  }
  const SrsPart* part = part_for_length<BLS381Cfg>(*s, first + n, &rc);
  return part ? msm_affine<BLS381Cfg>(*part, first, d_scalars, n, out_xy, is_inf, h_scalars) : rc;
}

// kzgpu_sync: a deferred MSM flag is collected here too
int kz_msm_pending_check() { return msm_flag_wait(); }

int kz_srs_curve(uint64_t handle) {
  const Srs* s = find_srs(handle);
  return s ? s->curve : -1;
}

extern "C" {

int kzgpu_srs_create(int curve, const uint64_t* affine_xy, size_t n, uint64_t* handle) {
  KZ_REQUIRE_INIT();
  if (!handle || (n && !affine_xy)) return kz_fail(KZGPU_EINVAL, "null pointer");
  if (curve == KZGPU_BN254) return srs_create_impl<BN254Cfg>(affine_xy, n, handle);
  if (curve == KZGPU_BLS12_381) return srs_create_impl<BLS381Cfg>(affine_xy, n, handle);
  return kz_fail(KZGPU_EINVAL, "Unsupported curve type: %d", curve);
}

int kzgpu_srs_generate_range(int curve, const uint64_t* tau, size_t start, size_t n, uint64_t* handle) {
  KZ_REQUIRE_INIT();
  if (!handle || !tau) return kz_fail(KZGPU_EINVAL, "null pointer");
  if (curve == KZGPU_BN254) return srs_generate_impl<BN254Cfg>(tau, start, n, handle);
  if (curve == KZGPU_BLS12_381) return srs_generate_impl<BLS381Cfg>(tau, start, n, handle);
  return kz_fail(KZGPU_EINVAL, "Unsupported curve type: %d", curve);
}

int kzgpu_srs_generate(int curve, const uint64_t* tau, size_t n, uint64_t* handle) {
  return kzgpu_srs_generate_range(curve, tau, 0, n, handle);
}

int kzgpu_srs_destroy(uint64_t handle) {
  KZ_REQUIRE_INIT();
  auto it = g_srs.find(handle);
  if (it == g_srs.end()) return kz_fail(KZGPU_EHANDLE, "unknown SRS handle %llu", (unsigned long long)handle);
  srs_free(it->second);
  g_srs.erase(it);
  return 0;
}

int kzgpu_srs_size(uint64_t handle, size_t* n) {
  KZ_REQUIRE_INIT();
  const Srs* s = find_srs(handle);
  if (!s || !n) return kz_fail(KZGPU_EHANDLE, "unknown SRS handle");
  *n = s->n;
  return 0;
}

int kzgpu_srs_info(uint64_t handle, int* c_tab, int* w_tab, size_t* device_bytes) {
  KZ_REQUIRE_INIT();
  const Srs* s = find_srs(handle);
  if (!s) return kz_fail(KZGPU_EHANDLE, "unknown SRS handle");
  if (c_tab) *c_tab = (int)s->full.c_tab;
  if (w_tab) *w_tab = (int)s->full.W_tab;
  if (device_bytes) *device_bytes = s->n * (s->curve == KZGPU_BN254 ? 64 : 96) * s->full.W_tab;
  return 0;
}

int kzgpu_srs_read(uint64_t handle, size_t first, size_t count, uint64_t* affine_xy) {
  KZ_REQUIRE_INIT();
  const Srs* s = find_srs(handle);
  if (!s) return kz_fail(KZGPU_EHANDLE, "unknown SRS handle");
  if (first + count > s->n || !affine_xy) return kz_fail(KZGPU_EINVAL, "range outside the SRS");
  if (!count) return 0;
  KzgpuCtx& cx = kz_ctx();
  const int N = s->curve == KZGPU_BN254 ? 8 : 12;
  size_t bytes = count * 2 * N * 4;
  uint32_t* tmp = nullptr;
  KZ_CUDA(cudaMalloc((void**)&tmp, bytes));
  if (s->curve == KZGPU_BN254)
    srs_from_mont_kernel<BN254Cfg><<<(unsigned)kz_div_up(2 * count, 128), 128, 0, cx.stream>>>(s->full.d_points, first, count, tmp);
  else
    srs_from_mont_kernel<BLS381Cfg><<<(unsigned)kz_div_up(2 * count, 128), 128, 0, cx.stream>>>(s->full.d_points, first, count, tmp);
  KZ_LAUNCHED();
  KZ_CUDA(cudaMemcpyAsync(affine_xy, tmp, bytes, cudaMemcpyDeviceToHost, cx.stream));
  KZ_CUDA(cudaStreamSynchronize(cx.stream));
  cudaFree(tmp);
  return 0;
}

int kzgpu_msm_dev(uint64_t handle, size_t first, const uint64_t* d_scalars, size_t n, uint64_t* out_affine_xy, int* is_inf) {
  KZ_REQUIRE_INIT();
  if (!out_affine_xy || (n && !d_scalars)) return kz_fail(KZGPU_EINVAL, "null pointer");
  return kz_msm_dev_internal(handle, first, (const uint32_t*)d_scalars, n, out_affine_xy, is_inf, nullptr);
}

int kzgpu_msm(uint64_t handle, size_t first, const uint64_t* scalars, size_t n, uint64_t* out_affine_xy, int* is_inf) {
  KZ_REQUIRE_INIT();
  if (!out_affine_xy || (n && !scalars)) return kz_fail(KZGPU_EINVAL, "null pointer");
  // the upload happens inside the MSM, chunked and overlapped with the compute (per device when the MSM is sharded)
  const uint32_t* d = nullptr;
  if (!(kz_ndev() > 1 && n >= shard_min())) {
    int rc = g_ws.scal.ensure(n * 32 + 32);
    if (rc) return rc;
    d = (const uint32_t*)g_ws.scal.p;
  }
  return kz_msm_dev_internal(handle, first, d, n, out_affine_xy, is_inf, scalars);
}

// k polynomials of poly_len scalars each, back to back on the PRIMARY device.  Several devices: polynomials of
// >= KZGPU_SHARD_MIN coefficients are point-sharded one after the other; shorter ones are dealt round-robin (they are
// equally long) and each device pulls its polynomials from the primary over NVLink and commits them in one pass.
int kz_msm_batch_dev_internal(uint64_t handle, const uint32_t* d_scalars, size_t poly_len, size_t k, uint64_t* out_xy, int* is_inf) {
  Srs* s = find_srs(handle);
  if (!s) return kz_fail(KZGPU_EHANDLE, "unknown SRS handle %llu", (unsigned long long)handle);
  if (poly_len > s->n)
    return kz_fail(KZGPU_ERANGE, "Polynomial degree %zu exceeds maximum allowed degree %zu", poly_len - 1, s->n - 1);
  const int nd = kz_ndev();
  const int L = kzgpu_fp_limbs64(s->curve);
  int rc;
  if (nd == 1) {
    const SrsPart* part = s->curve == KZGPU_BN254 ? part_for_length<BN254Cfg>(*s, poly_len, &rc) : part_for_length<BLS381Cfg>(*s, poly_len, &rc);
    return part ? msm_batch_on_part(*part, s->curve, d_scalars, poly_len, k, out_xy, is_inf) : rc;
  }
  if (poly_len * k < ((size_t)1 << 17)) return msm_batch_on_part(s->full, s->curve, d_scalars, poly_len, k, out_xy, is_inf);
  if (poly_len >= shard_min()) {
    for (size_t j = 0; j < k; j++) {
      int* fl = is_inf ? is_inf + j : nullptr;
      rc = s->curve == KZGPU_BN254 ? msm_sharded<BN254Cfg>(*s, 0, poly_len, nullptr, d_scalars + j * poly_len * 8, out_xy + j * 2 * L, fl)
                                   : msm_sharded<BLS381Cfg>(*s, 0, poly_len, nullptr, d_scalars + j * poly_len * 8, out_xy + j * 2 * L, fl);
      if (rc) return rc;
    }
    return 0;
  }
  if (k > 1) {
    rc = s->curve == KZGPU_BN254 ? ensure_replicas<BN254Cfg>(*s) : ensure_replicas<BLS381Cfg>(*s);
    if (rc) return rc;
  }
  KZ_CUDA(cudaStreamSynchronize(kz_ctx_of(0).stream));            // the polynomials may still be in flight on the primary's stream
  const int dev0 = kz_device_of(0);
  return kz_parallel([&](int slot) -> int {
    // polynomials slot, slot + nd, ... ; the primary works in place on contiguous runs only when nd == 1, so it copies too
    std::vector<size_t> js;
    for (size_t j = slot; j < k; j += nd) js.push_back(j);
    if (js.empty()) return 0;
    KzgpuCtx& cx = kz_ctx();
    int r = g_ws.scal.ensure(js.size() * poly_len * 32 + 32);
    if (r) return r;
    uint32_t* d = (uint32_t*)g_ws.scal.p;
    for (size_t q = 0; q < js.size(); q++)
      KZ_CUDA(cudaMemcpyPeerAsync(d + q * poly_len * 8, kz_device_of(slot), d_scalars + js[q] * poly_len * 8, dev0, poly_len * 32, cx.stream));
    std::vector<uint64_t> o(js.size() * 2 * L);
    std::vector<int> f(js.size());
    if ((r = msm_batch_on_part(slot == 0 ? s->full : s->replica[slot], s->curve, d, poly_len, js.size(), o.data(), f.data()))) return r;
    for (size_t q = 0; q < js.size(); q++) {
      memcpy(out_xy + js[q] * 2 * L, &o[q * 2 * L], 2 * L * 8);
      if (is_inf) is_inf[js[q]] = f[q];
    }
    return 0;
  });
}

int kzgpu_msm_batch_dev(uint64_t handle, const uint64_t* d_scalars, size_t poly_len, size_t k, uint64_t* out_affine_xy, int* is_inf) {
  KZ_REQUIRE_INIT();
  if (k && (!out_affine_xy || (poly_len && !d_scalars))) return kz_fail(KZGPU_EINVAL, "null pointer");
  if (!k) return 0;
  return kz_msm_batch_dev_internal(handle, (const uint32_t*)d_scalars, poly_len, k, out_affine_xy, is_inf);
}

// One commit() call (kzg.py:102): k host polynomials, concatenated.  One device: a single pass over all of them.  Several
// devices: polynomials of >= KZGPU_SHARD_MIN coefficients are point-sharded over every device, the others are placed
// whole, longest first, on the least loaded device (longest-processing-time rule; each device commits its share in one
// pass against its replica of the key).
int kzgpu_msm_batch(uint64_t handle, const uint64_t* scalars, const size_t* lens, size_t k, uint64_t* out_affine_xy, int* is_inf) {
  KZ_REQUIRE_INIT();
  if (k && (!lens || !out_affine_xy)) return kz_fail(KZGPU_EINVAL, "null pointer");
  Srs* s = find_srs(handle);
  if (!s) return kz_fail(KZGPU_EHANDLE, "unknown SRS handle %llu", (unsigned long long)handle);
  if (!k) return 0;
  size_t maxlen = 0;
  for (size_t j = 0; j < k; j++) if (lens[j] > maxlen) maxlen = lens[j];
  if (maxlen > s->n)                                   // the degree check of kzg.py:103-106, before any work
    return kz_fail(KZGPU_ERANGE, "Polynomial degree %zu exceeds maximum allowed degree %zu", maxlen - 1, s->n - 1);
  const int L = kzgpu_fp_limbs64(s->curve);
  std::vector<const uint64_t*> ptr(k);
  std::vector<uint64_t*> outp(k);
  std::vector<int*> infp(k);
  size_t off = 0;
  for (size_t j = 0; j < k; j++) {
    ptr[j] = scalars + off * 4; off += lens[j];
    outp[j] = out_affine_xy + j * 2 * L; infp[j] = is_inf ? is_inf + j : nullptr;
  }
  const int nd = kz_ndev();
  if (nd == 1) {
    int rcp;
    const SrsPart* part = s->curve == KZGPU_BN254 ? part_for_length<BN254Cfg>(*s, maxlen, &rcp) : part_for_length<BLS381Cfg>(*s, maxlen, &rcp);
    return part ? msm_batch_host_on_part(*part, s->curve, ptr.data(), lens, k, outp.data(), infp.data()) : rcp;
  }
  if (k == 1 && lens[0] < shard_min()) return msm_batch_host_on_part(s->full, s->curve, ptr.data(), lens, k, outp.data(), infp.data());
  // several devices
  std::vector<size_t> order, owner(k, 0);
  int rc;
  for (size_t j = 0; j < k; j++) {
    if (lens[j] >= shard_min()) {
      rc = s->curve == KZGPU_BN254 ? msm_sharded<BN254Cfg>(*s, 0, lens[j], ptr[j], nullptr, outp[j], infp[j])
                                   : msm_sharded<BLS381Cfg>(*s, 0, lens[j], ptr[j], nullptr, outp[j], infp[j]);
      if (rc) return rc;
    } else {
      order.push_back(j);
    }
  }
  if (order.empty()) return 0;
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return lens[a] != lens[b] ? lens[a] > lens[b] : a < b; });
  size_t load[KZ_MAX_DEV] = {0};
  std::vector<std::vector<size_t>> mine(nd);
  for (size_t j : order) {
    int best = 0;
    for (int d = 1; d < nd; d++) if (load[d] < load[best]) best = d;
    load[best] += lens[j] + 4096;                      // + a fixed cost per polynomial
    mine[best].push_back(j);
  }
  bool others = false;
  for (int d = 1; d < nd; d++) others = others || !mine[d].empty();
  if (others) {
    rc = s->curve == KZGPU_BN254 ? ensure_replicas<BN254Cfg>(*s) : ensure_replicas<BLS381Cfg>(*s);
    if (rc) return rc;
  } else {
    s->replica[0] = s->full;
  }
  return kz_parallel([&](int slot) -> int {
    const std::vector<size_t>& js = mine[slot];
    if (js.empty()) return 0;
    std::vector<const uint64_t*> p;
    std::vector<size_t> ln;
    std::vector<uint64_t*> o;
    std::vector<int*> f;
    for (size_t j : js) { p.push_back(ptr[j]); ln.push_back(lens[j]); o.push_back(outp[j]); f.push_back(infp[j]); }
    return msm_batch_host_on_part(slot == 0 ? s->full : s->replica[slot], s->curve, p.data(), ln.data(), js.size(), o.data(), f.data());
  });
}

int kzgpu_msm_partial_dev(uint64_t handle, size_t first, const uint64_t* d_scalars, size_t n, uint64_t* d_out_xyzz) {
  KZ_REQUIRE_INIT();
  const Srs* s = find_srs(handle);
  if (!s) return kz_fail(KZGPU_EHANDLE, "unknown SRS handle %llu", (unsigned long long)handle);
  if (first + n > s->n) return kz_fail(KZGPU_ERANGE, "scalar range exceeds the SRS shard");
  if (!d_out_xyzz || (n && !d_scalars)) return kz_fail(KZGPU_EINVAL, "null pointer");
  int rc = set_smem_attrs();
  if (rc) return rc;
  // returns with the work queued: the partial stays on the device and the caller's exchange is queued right behind it; a
  // non-canonical scalar is reported by the kzgpu_g1_fold / kzgpu_sync / MSM call that synchronises next
  if (s->curve == KZGPU_BN254) return msm_core<BN254Cfg>(s->full, first, (const uint32_t*)d_scalars, n, 0, (uint32_t*)d_out_xyzz, nullptr, nullptr, 1, true);
  return msm_core<BLS381Cfg>(s->full, first, (const uint32_t*)d_scalars, n, 0, (uint32_t*)d_out_xyzz, nullptr, nullptr, 1, true);
}

int kzgpu_msm_partial(uint64_t handle, size_t first, const uint64_t* scalars, size_t n, uint64_t* d_out_xyzz) {
  KZ_REQUIRE_INIT();
  const Srs* s = find_srs(handle);
  if (!s) return kz_fail(KZGPU_EHANDLE, "unknown SRS handle %llu", (unsigned long long)handle);
  if (first + n > s->n) return kz_fail(KZGPU_ERANGE, "scalar range exceeds the SRS shard");
  if (!d_out_xyzz || (n && !scalars)) return kz_fail(KZGPU_EINVAL, "null pointer");
  int rc = set_smem_attrs();
  if (rc) return rc;
  // host scalars of this rank's shard: uploaded inside the MSM, chunked and overlapped with the compute like kzgpu_msm
  if ((rc = g_ws.scal.ensure(n * 32 + 32))) return rc;
  const uint32_t* d = (const uint32_t*)g_ws.scal.p;
  if (s->curve == KZGPU_BN254) return msm_core<BN254Cfg>(s->full, first, d, n, 0, (uint32_t*)d_out_xyzz, scalars, nullptr, 1, true);
  return msm_core<BLS381Cfg>(s->full, first, d, n, 0, (uint32_t*)d_out_xyzz, scalars, nullptr, 1, true);
}

int kzgpu_g1_fold(int curve, const uint64_t* d_xyzz, size_t count, uint64_t* out_affine_xy, int* is_inf) {
  KZ_REQUIRE_INIT();
  if (!d_xyzz || !out_affine_xy) return kz_fail(KZGPU_EINVAL, "null pointer");
  if (curve != KZGPU_BN254 && curve != KZGPU_BLS12_381) return kz_fail(KZGPU_EINVAL, "Unsupported curve type: %d", curve);
  KzgpuCtx& cx = kz_ctx();
  int rc = set_smem_attrs();
  if (rc) return rc;
  const int N = curve == KZGPU_BN254 ? 8 : 12;
  if ((rc = g_ws.result.ensure(4 * N * 4 + 4))) return rc;
  if (curve == KZGPU_BN254)
    g1_fold_kernel<BN254Cfg><<<1, 128, 128 * 4 * N * 4, cx.stream>>>((const uint32_t*)d_xyzz, (uint32_t)count, (uint32_t*)g_ws.result.p);
  else
    g1_fold_kernel<BLS381Cfg><<<1, 128, 128 * 4 * N * 4, cx.stream>>>((const uint32_t*)d_xyzz, (uint32_t)count, (uint32_t*)g_ws.result.p);
  KZ_LAUNCHED();
  uint32_t h[4 * 12];
  KZ_CUDA(cudaMemcpyAsync(h, g_ws.result.p, 4 * N * 4, cudaMemcpyDeviceToHost, cx.stream));
  uint32_t hflag = 0;
  if ((rc = msm_flag_collect(&hflag))) return rc;
  KZ_CUDA(cudaStreamSynchronize(cx.stream));
  if ((rc = msm_flag_result(hflag))) return rc;
  if (curve == KZGPU_BN254) host_xyzz_to_canonical<FpBN254>(h, (uint32_t*)out_affine_xy, is_inf);
  else host_xyzz_to_canonical<FpBLS381>(h, (uint32_t*)out_affine_xy, is_inf);
  return 0;
}

int kzgpu_g1_lincomb(int curve, const uint64_t* affine_xy, const uint64_t* scalars, size_t count, uint64_t* out_affine_xy, int* is_inf) {
  KZ_REQUIRE_INIT();
  if (!out_affine_xy || (count && (!affine_xy || !scalars))) return kz_fail(KZGPU_EINVAL, "null pointer");
  if (curve != KZGPU_BN254 && curve != KZGPU_BLS12_381) return kz_fail(KZGPU_EINVAL, "Unsupported curve type: %d", curve);
  if (count > 65536) return kz_fail(KZGPU_EINVAL, "kzgpu_g1_lincomb is for verifier-sized combinations (<= 65536 terms); use an SRS handle and kzgpu_msm");
  KzgpuCtx& cx = kz_ctx();
  int rc = set_smem_attrs();
  if (rc) return rc;
  const int N = curve == KZGPU_BN254 ? 8 : 12;
  const size_t pb = count * 2 * N * 4, sb = count * 32;
  if ((rc = g_ws.scal.ensure(pb + sb + 64)) || (rc = g_ws.result.ensure(4 * N * 4 + 4))) return rc;
  uint32_t* d_pts = (uint32_t*)g_ws.scal.p;
  uint32_t* d_sc = d_pts + count * 2 * N;
  if (count) {
    KZ_CUDA(cudaMemcpyAsync(d_pts, affine_xy, pb, cudaMemcpyHostToDevice, cx.stream));
    KZ_CUDA(cudaMemcpyAsync(d_sc, scalars, sb, cudaMemcpyHostToDevice, cx.stream));
  }
  if (curve == KZGPU_BN254)
    g1_lincomb_kernel<BN254Cfg><<<1, 128, 128 * 4 * N * 4, cx.stream>>>(d_pts, d_sc, (uint32_t)count, (uint32_t*)g_ws.result.p);
  else
    g1_lincomb_kernel<BLS381Cfg><<<1, 128, 128 * 4 * N * 4, cx.stream>>>(d_pts, d_sc, (uint32_t)count, (uint32_t*)g_ws.result.p);
  KZ_LAUNCHED();
  uint32_t h[4 * 12];
  KZ_CUDA(cudaMemcpyAsync(h, g_ws.result.p, 4 * N * 4, cudaMemcpyDeviceToHost, cx.stream));
  KZ_CUDA(cudaStreamSynchronize(cx.stream));
  if (curve == KZGPU_BN254) host_xyzz_to_canonical<FpBN254>(h, (uint32_t*)out_affine_xy, is_inf);
  else host_xyzz_to_canonical<FpBLS381>(h, (uint32_t*)out_affine_xy, is_inf);
  return 0;
}

}  // extern "C"
