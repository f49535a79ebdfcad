// NTT / iNTT / coset NTT over the scalar field: the device replacement of the reference's
// fft_ff / ifft_ff (fft_ff.py:3-58).  Contract (fft_ff.py:32-35 unrolled):
//     out[k] = sum_j in[j] * w^(j k),  natural order in and out, caller-supplied w.
//
// Algorithm: mixed-radix Stockham decomposition n = R_1 * ... * R_m (R_t = 2^b_t, b_t <= 9).
// Pass t (executed t = m .. 1) views its input as a [R_t][Q] matrix (Q = n / R_t), transforms
// each column with an R_t-point DFT held in shared memory, and writes the result so that the
// NEXT pass again reads contiguous runs and the last pass lands in natural order:
//     in [j_t][jlo][kk]  ->  out[jlo][k_t][kk],   jlo < J = R_1..R_{t-1},  kk < K = R_{t+1}..R_m.
// The inter-pass twiddle w^(J * j_t * kk) (and, folded into it, the 1/n scale of the inverse
// and the coset powers) comes from a per-pass 2-D table indexed exactly like the data tile, so
// it is read with the same coalesced pattern.  Inside a block the R_t-point DFT is again
// decomposed into radix-8/4/2 layers done in registers (5 constant multiplications per 8
// points) with one shared-memory exchange per layer.
//
// Data stays in canonical form end to end: every table entry is in Montgomery form, and
// mont_mul(canonical, montgomery) is canonical.  HBM traffic per pass: 32 B read + 32 B
// written per element (+32 B of table for t < m).
#include "common.cuh"
#include <vector>
#include <list>
#include <cstring>

namespace {

constexpr int kLogTile = 10;        // 1024 elements (32 KB of shared memory) per block: four blocks per SM, so the
                                    // global-load / store phases of one block hide behind the butterflies of the others
constexpr int kMaxB = 9;            // R_t <= 512
constexpr int kThreads = 128;

template <class P> struct NttConsts { Fe<P> w8, w4, w8_3; };   // omega_8, omega_8^2, omega_8^3

struct NttPassArgs {
  const uint32_t* in;
  uint32_t* out;
  const uint32_t* bnd;    // [R][K] boundary table or nullptr
  const uint32_t* post;   // [R] multiplier applied at the final store or nullptr
  const uint32_t* wR;     // omega_R^e, e < R
  uint64_t n;
  uint32_t b, logC, logK, logQ;
  uint32_t nd;
  uint32_t dpack;         // layer widths, 2 bits each (kept out of an array: no local-memory indexing)
  uint32_t blk0;          // first column tile of this launch (column-slab launches of the host-buffer entry point)
  uint32_t lazy, last;    // compile-time-shape kernels: values semi-reduced in [0, 2r) between passes; last pass folds them
};

__device__ __forceinline__ uint32_t layer_width(uint32_t dpack, uint32_t s) { return (dpack >> (2 * s)) & 3u; }

template <class P> __device__ __forceinline__ Fe<P> ld_fe(const uint32_t* p) {
  Fe<P> r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < P::N / 4; i++) {
    uint4 t = q[i];
    r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w;
  }
  return r;
}
template <class P> __device__ __forceinline__ Fe<P> ldg_fe(const uint32_t* p) {
  Fe<P> r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < P::N / 4; i++) {
    uint4 t = __ldg(q + i);
    r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w;
  }
  return r;
}
template <class P> __device__ __forceinline__ void st_fe(uint32_t* p, const Fe<P>& a) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < P::N / 4; i++) q[i] = make_uint4(a.v[4 * i], a.v[4 * i + 1], a.v[4 * i + 2], a.v[4 * i + 3]);
}

// shared memory: 8 word-planes of `tile` words; XOR swizzle so that both the "consecutive
// columns" and the "stride 8C" access patterns of the layers fall on distinct banks
__device__ __forceinline__ uint32_t swz(uint32_t pos, uint32_t logC) {
  return pos ^ (((pos >> (logC + 3)) & 7u) << logC);
}
template <class P> __device__ __forceinline__ Fe<P> sm_ld(const uint32_t* sm, uint32_t tile, uint32_t p) {
  Fe<P> r;
#pragma unroll
  for (int w = 0; w < P::N; w++) r.v[w] = sm[w * tile + p];
  return r;
}
template <class P> __device__ __forceinline__ void sm_st(uint32_t* sm, uint32_t tile, uint32_t p, const Fe<P>& a) {
#pragma unroll
  for (int w = 0; w < P::N; w++) sm[w * tile + p] = a.v[w];
}

template <class P> __device__ __forceinline__ void bfly(Fe<P>& a, Fe<P>& b) {
  Fe<P> s = fe_add<P>(a, b);
  b = fe_sub<P>(a, b);
  a = s;
}

// y[v] = sum_u x[u] * root^(u v), root = omega_{2^D}; in place, natural order
template <class P, int D> __device__ __forceinline__ void dft_small(Fe<P>* x, const NttConsts<P>& c) {
  if constexpr (D == 1) {
    bfly<P>(x[0], x[1]);
  } else if constexpr (D == 2) {
    bfly<P>(x[0], x[2]);                 // s0, s1
    bfly<P>(x[1], x[3]);                 // s2, (a1 - a3)
    x[3] = fe_mul<P>(x[3], c.w4);
    Fe<P> y0 = fe_add<P>(x[0], x[1]), y2 = fe_sub<P>(x[0], x[1]);
    Fe<P> y1 = fe_add<P>(x[2], x[3]), y3 = fe_sub<P>(x[2], x[3]);
    x[0] = y0; x[1] = y1; x[2] = y2; x[3] = y3;
  } else {
    Fe<P> e[4] = {x[0], x[2], x[4], x[6]};
    Fe<P> o[4] = {x[1], x[3], x[5], x[7]};
    dft_small<P, 2>(e, c);
    dft_small<P, 2>(o, c);
    o[1] = fe_mul<P>(o[1], c.w8);
    o[2] = fe_mul<P>(o[2], c.w4);
    o[3] = fe_mul<P>(o[3], c.w8_3);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      x[i] = fe_add<P>(e[i], o[i]);
      x[i + 4] = fe_sub<P>(e[i], o[i]);
    }
  }
}

template <class P, int D>
__device__ __forceinline__ void ntt_layer(uint32_t* sm, const uint32_t* smw, const NttPassArgs& a, const NttConsts<P>& c,
                                          uint32_t li, uint32_t off, uint32_t done) {
  const uint32_t tile_log = a.b + a.logC, tile = 1u << tile_log;
  const uint32_t ngroups = tile >> D;
  const uint32_t Cmask = (1u << a.logC) - 1;
#pragma unroll 1
  for (uint32_t gsub = 0; gsub < (8u >> D); gsub++) {
    uint32_t G = threadIdx.x + blockDim.x * gsub;
    if (G >= ngroups) break;
    uint32_t cc = G & Cmask, rest = G >> a.logC;
    uint32_t low = rest & ((1u << off) - 1), high = rest >> off;
    // V = number formed by the output digits of the previous layers (first layer = least significant)
    uint32_t V = 0, hb = high, sh = done;
    for (int s = (int)li - 1; s >= 0; s--) {
      uint32_t dd = layer_width(a.dpack, s);
      sh -= dd;
      V |= (hb & ((1u << dd) - 1)) << sh;
      hb >>= dd;
    }
    Fe<P> x[1 << D];
    uint32_t base_l = (high << (off + D)) | low;
#pragma unroll
    for (int u = 0; u < (1 << D); u++) {
      uint32_t l = base_l | ((uint32_t)u << off);
      x[u] = sm_ld<P>(sm, tile, swz((l << a.logC) | cc, a.logC));
    }
    if (li > 0) {
#pragma unroll
      for (int u = 1; u < (1 << D); u++) {
        uint32_t e = ((uint32_t)u << off) * V;
        x[u] = fe_mul<P>(x[u], sm_ld<P>(smw, 1u << a.b, e));
      }
    }
    dft_small<P, D>(x, c);
#pragma unroll
    for (int u = 0; u < (1 << D); u++) {
      uint32_t l = base_l | ((uint32_t)u << off);
      sm_st<P>(sm, tile, swz((l << a.logC) | cc, a.logC), x[u]);
    }
  }
}

template <class P>
__global__ void __launch_bounds__(kThreads, 4) ntt_pass_kernel(NttPassArgs a, NttConsts<P> c) {
  extern __shared__ uint32_t sm[];
  const uint32_t tile_log = a.b + a.logC, tile = 1u << tile_log;
  // omega_R^e table as 8 word-planes behind the tile (read by the layers li > 0)
  uint32_t* smw = sm + (size_t)P::N * tile;
  if (a.nd > 1) {
    const uint32_t R = 1u << a.b;
    for (uint32_t e = threadIdx.x; e < R; e += blockDim.x) sm_st<P>(smw, R, e, ldg_fe<P>(a.wR + (size_t)e * P::N));
  }
  const uint32_t Cmask = (1u << a.logC) - 1, Kmask = (1u << a.logK) - 1, Rmask = (1u << a.b) - 1;
  const size_t boff = (size_t)blockIdx.y * a.n;
  const size_t q0 = ((size_t)blockIdx.x + a.blk0) << a.logC;

  // load phase, four elements (and their boundary twiddles) in flight per thread
  for (uint32_t o0 = threadIdx.x; o0 < tile; o0 += 4 * blockDim.x) {
    Fe<P> v[4], tw[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint32_t o = o0 + k * blockDim.x;
      if (o < tile) {
        uint32_t l = o >> a.logC, cc = o & Cmask;
        size_t qq = q0 + cc;
        v[k] = ld_fe<P>(a.in + (boff + ((size_t)l << a.logQ) + qq) * P::N);
        if (a.bnd) tw[k] = ldg_fe<P>(a.bnd + (((size_t)l << a.logK) + (qq & Kmask)) * P::N);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint32_t o = o0 + k * blockDim.x;
      if (o < tile) {
        if (a.bnd) v[k] = fe_mul<P>(v[k], tw[k]);
        sm_st<P>(sm, tile, swz(o, a.logC), v[k]);
      }
    }
  }
  __syncthreads();

  uint32_t off = a.b, done = 0;
  for (uint32_t li = 0; li < a.nd; li++) {
    uint32_t d = layer_width(a.dpack, li);
    off -= d;
    if (d == 3) ntt_layer<P, 3>(sm, smw, a, c, li, off, done);
    else if (d == 2) ntt_layer<P, 2>(sm, smw, a, c, li, off, done);
    else ntt_layer<P, 1>(sm, smw, a, c, li, off, done);
    done += d;
    __syncthreads();
  }

  const uint32_t logCm = a.logC < a.logK ? a.logC : a.logK;
  for (uint32_t o = threadIdx.x; o < tile; o += blockDim.x) {
    uint32_t cc_lo = o & ((1u << logCm) - 1);
    uint32_t kt = (o >> logCm) & Rmask;
    uint32_t cc_hi = o >> (logCm + a.b);
    uint32_t cc = (cc_hi << logCm) | cc_lo;
    size_t qq = q0 + cc;
    size_t jlo = qq >> a.logK, kk = qq & Kmask;
    // X[kt] sits at the digit-reversed row
    uint32_t l = 0, ob = a.b, kr = kt;
    for (uint32_t s = 0; s < a.nd; s++) {
      uint32_t dd = layer_width(a.dpack, s);
      ob -= dd;
      l |= (kr & ((1u << dd) - 1)) << ob;
      kr >>= dd;
    }
    Fe<P> v = sm_ld<P>(sm, tile, swz((l << a.logC) | cc, a.logC));
    if (a.post) v = fe_mul<P>(v, ldg_fe<P>(a.post + (size_t)kt * P::N));
    st_fe<P>(a.out + (boff + (((jlo << a.b) + kt) << a.logK) + kk) * P::N, v);
  }
}

// ---- arithmetic policy of the compile-time-shape kernels.  LZ (fields with 4r <= 2^256, i.e. BN254 r): data stays
// semi-reduced in [0, 2r) in shared memory and between passes, products skip their final conditional subtraction and
// sums / differences that only feed a constant multiplication are not folded (field.cuh); the last pass folds once.
template <class P, bool LZ> struct Ar {
  static __device__ __forceinline__ Fe<P> mul(const Fe<P>& a, const Fe<P>& b) {
    if constexpr (LZ) return fe_mul_lz<P>(a, b); else return fe_mul<P>(a, b);
  }
  static __device__ __forceinline__ Fe<P> add(const Fe<P>& a, const Fe<P>& b) {
    if constexpr (LZ) return fe_add_lz<P>(a, b); else return fe_add<P>(a, b);
  }
  static __device__ __forceinline__ Fe<P> sub(const Fe<P>& a, const Fe<P>& b) {
    if constexpr (LZ) return fe_sub_lz<P>(a, b); else return fe_sub<P>(a, b);
  }
  static __device__ __forceinline__ Fe<P> add_m(const Fe<P>& a, const Fe<P>& b) {      // result is multiplied next
    if constexpr (LZ) return fe_add_nr<P>(a, b); else return fe_add<P>(a, b);
  }
  static __device__ __forceinline__ Fe<P> sub_m(const Fe<P>& a, const Fe<P>& b) {
    if constexpr (LZ) return fe_sub_nr<P>(a, b); else return fe_sub<P>(a, b);
  }
};

// 4-point DFT; M: outputs 1..3 are multiplied by constants next (left unfolded under LZ)
template <class P, bool LZ, bool M> __device__ __forceinline__ void dft4_t(Fe<P>* x, const NttConsts<P>& c) {
  using A = Ar<P, LZ>;
  Fe<P> s0 = A::add(x[0], x[2]), d0 = A::sub(x[0], x[2]);
  Fe<P> s1 = A::add(x[1], x[3]);
  Fe<P> t = A::mul(A::sub_m(x[1], x[3]), c.w4);
  x[0] = A::add(s0, s1);
  if constexpr (M) { x[1] = A::add_m(d0, t); x[2] = A::sub_m(s0, s1); x[3] = A::sub_m(d0, t); }
  else { x[1] = A::add(d0, t); x[2] = A::sub(s0, s1); x[3] = A::sub(d0, t); }
}

template <class P, int D, bool LZ> __device__ __forceinline__ void dft_small_t(Fe<P>* x, const NttConsts<P>& c) {
  using A = Ar<P, LZ>;
  if constexpr (D == 1) {
    Fe<P> s = A::add(x[0], x[1]);
    x[1] = A::sub(x[0], x[1]);
    x[0] = s;
  } else if constexpr (D == 2) {
    dft4_t<P, LZ, false>(x, c);
  } else {
    Fe<P> e[4] = {x[0], x[2], x[4], x[6]};
    Fe<P> o[4] = {x[1], x[3], x[5], x[7]};
    dft4_t<P, LZ, false>(e, c);
    dft4_t<P, LZ, true>(o, c);
    o[1] = A::mul(o[1], c.w8);
    o[2] = A::mul(o[2], c.w4);
    o[3] = A::mul(o[3], c.w8_3);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      x[i] = A::add(e[i], o[i]);
      x[i + 4] = A::sub(e[i], o[i]);
    }
  }
}

// ---- compile-time shapes: R = 2^B with a full 1024-element tile (B = 6..9, C = 2^(10-B) columns).  Same algorithm and
// shared-memory layout as ntt_pass_kernel; with B fixed every plane stride, digit width and swizzle shift is an immediate and
// the digit loops fold away (the generic kernel spends ~15 % of its instructions on that address arithmetic).
template <int B> struct NttShape {
  static constexpr int LOGC = kLogTile - B;
  static constexpr int ND = (B + 2) / 3;
  static constexpr int width(int s) { return B - 3 * s >= 3 ? 3 : B - 3 * s; }
  static constexpr int done(int li) { return 3 * li; }                 // all layers before the last are radix 8
  static constexpr int off(int li) { return B - done(li) - width(li); }
};

template <class P, int B, int LI, bool LZ>
__device__ __forceinline__ void ntt_layer_c(uint32_t* sm, const uint32_t* smw, const NttConsts<P>& c) {
  using S = NttShape<B>;
  constexpr int D = S::width(LI), OFF = S::off(LI), DONE = S::done(LI), LOGC = S::LOGC;
  constexpr uint32_t tile = 1u << kLogTile, ngroups = tile >> D, Cmask = (1u << LOGC) - 1;
#pragma unroll
  for (uint32_t gsub = 0; gsub < (8u >> D); gsub++) {
    const uint32_t G = threadIdx.x + kThreads * gsub;
    if (ngroups < kThreads * (8u >> D) && G >= ngroups) break;
    const uint32_t cc = G & Cmask, rest = G >> LOGC;
    const uint32_t low = rest & ((1u << OFF) - 1), high = rest >> OFF;
    uint32_t V = 0;
    {
      uint32_t hb = high;
      int sh = DONE;
#pragma unroll
      for (int s = LI - 1; s >= 0; s--) {
        constexpr int dd = 3;
        sh -= dd;
        V |= (hb & ((1u << dd) - 1)) << sh;
        hb >>= dd;
      }
    }
    Fe<P> x[1 << D];
    const uint32_t base_l = (high << (OFF + D)) | low;
#pragma unroll
    for (int u = 0; u < (1 << D); u++) {
      uint32_t l = base_l | ((uint32_t)u << OFF);
      x[u] = sm_ld<P>(sm, tile, swz((l << LOGC) | cc, LOGC));
    }
    if constexpr (LI > 0) {
#pragma unroll
      for (int u = 1; u < (1 << D); u++) {
        uint32_t e = ((uint32_t)u << OFF) * V;
        x[u] = Ar<P, LZ>::mul(x[u], sm_ld<P>(smw, 1u << B, e));
      }
    }
    dft_small_t<P, D, LZ>(x, c);
#pragma unroll
    for (int u = 0; u < (1 << D); u++) {
      uint32_t l = base_l | ((uint32_t)u << OFF);
      sm_st<P>(sm, tile, swz((l << LOGC) | cc, LOGC), x[u]);
    }
  }
}

template <class P, int B, bool LZ>
__global__ void __launch_bounds__(kThreads, 4) ntt_pass_kernel_c(NttPassArgs a, NttConsts<P> c) {
  using S = NttShape<B>;
  constexpr int LOGC = S::LOGC;
  constexpr uint32_t tile = 1u << kLogTile, R = 1u << B, Cmask = (1u << LOGC) - 1, Rmask = R - 1;
  extern __shared__ uint32_t sm[];
  uint32_t* smw = sm + (size_t)P::N * tile;
  for (uint32_t e = threadIdx.x; e < R; e += kThreads) sm_st<P>(smw, R, e, ldg_fe<P>(a.wR + (size_t)e * P::N));
  const uint32_t Kmask = (1u << a.logK) - 1;
  const size_t boff = (size_t)blockIdx.y * a.n;
  const size_t q0 = ((size_t)blockIdx.x + a.blk0) << LOGC;

#pragma unroll 1
  for (uint32_t o0 = threadIdx.x; o0 < tile; o0 += 4 * kThreads) {
    Fe<P> v[4], tw[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint32_t o = o0 + k * kThreads;
      uint32_t l = o >> LOGC, cc = o & Cmask;
      size_t qq = q0 + cc;
      v[k] = ld_fe<P>(a.in + (boff + ((size_t)l << a.logQ) + qq) * P::N);
      if (a.bnd) tw[k] = ldg_fe<P>(a.bnd + (((size_t)l << a.logK) + (qq & Kmask)) * P::N);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint32_t o = o0 + k * kThreads;
      if (a.bnd) v[k] = Ar<P, LZ>::mul(v[k], tw[k]);
      sm_st<P>(sm, tile, swz(o, LOGC), v[k]);
    }
  }
  __syncthreads();

  ntt_layer_c<P, B, 0, LZ>(sm, smw, c);
  __syncthreads();
  if constexpr (S::ND > 1) { ntt_layer_c<P, B, 1, LZ>(sm, smw, c); __syncthreads(); }
  if constexpr (S::ND > 2) { ntt_layer_c<P, B, 2, LZ>(sm, smw, c); __syncthreads(); }

  const uint32_t logCm = (uint32_t)LOGC < a.logK ? (uint32_t)LOGC : a.logK;
#pragma unroll 1
  for (uint32_t o = threadIdx.x; o < tile; o += kThreads) {
    uint32_t cc_lo = o & ((1u << logCm) - 1);
    uint32_t kt = (o >> logCm) & Rmask;
    uint32_t cc_hi = o >> (logCm + B);
    uint32_t cc = (cc_hi << logCm) | cc_lo;
    size_t qq = q0 + cc;
    size_t jlo = qq >> a.logK, kk = qq & Kmask;
    uint32_t l = 0, kr = kt;
    int ob = B;
#pragma unroll
    for (int s = 0; s < S::ND; s++) {
      constexpr int dummy = 0; (void)dummy;
      const int dd = S::width(s);
      ob -= dd;
      l |= (kr & ((1u << dd) - 1)) << ob;
      kr >>= dd;
    }
    Fe<P> v = sm_ld<P>(sm, tile, swz((l << LOGC) | cc, LOGC));
    if (a.post) v = fe_mul<P>(v, ldg_fe<P>(a.post + (size_t)kt * P::N));      // only on the last pass: canonical result
    else if (LZ && a.last) v = fe_reduce_lz<P>(v);
    st_fe<P>(a.out + (boff + (((jlo << B) + kt) << a.logK) + kk) * P::N, v);
  }
}

// a pass runs on the compile-time-shape kernel when its tile is full (1024 elements) and R = 2^6..2^9
inline bool pass_is_full(uint32_t b, uint32_t logC) { return b + logC == (uint32_t)kLogTile && b >= 6 && b <= 9; }

template <class P, bool LZ>
void launch_pass_c(const NttPassArgs& a, const NttConsts<P>& c, dim3 grid, size_t smem, cudaStream_t st) {
  if (a.b == 9) ntt_pass_kernel_c<P, 9, LZ><<<grid, kThreads, smem, st>>>(a, c);
  else if (a.b == 8) ntt_pass_kernel_c<P, 8, LZ><<<grid, kThreads, smem, st>>>(a, c);
  else if (a.b == 7) ntt_pass_kernel_c<P, 7, LZ><<<grid, kThreads, smem, st>>>(a, c);
  else ntt_pass_kernel_c<P, 6, LZ><<<grid, kThreads, smem, st>>>(a, c);
}

template <class P>
void launch_pass(const NttPassArgs& a, const NttConsts<P>& c, dim3 grid, uint32_t threads, size_t smem, cudaStream_t st) {
  if (pass_is_full(a.b, a.logC)) {
    if constexpr (FeLz<P>::ok) {
      if (a.lazy) { launch_pass_c<P, true>(a, c, grid, smem, st); return; }
    }
    launch_pass_c<P, false>(a, c, grid, smem, st);
    return;
  }
  (void)threads;
  ntt_pass_kernel<P><<<grid, threads, smem, st>>>(a, c);
}

template <class P, int B, bool LZ> int set_pass_smem() {
  KZ_CUDA(cudaFuncSetAttribute(ntt_pass_kernel_c<P, B, LZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (8 * 4 << kLogTile) + (32 << kMaxB)));
  return 0;
}
template <class P> int set_all_pass_smem() {
  int rc;
  if ((rc = set_pass_smem<P, 6, false>()) || (rc = set_pass_smem<P, 7, false>()) || (rc = set_pass_smem<P, 8, false>()) ||
      (rc = set_pass_smem<P, 9, false>())) return rc;
  if constexpr (FeLz<P>::ok) {
    if ((rc = set_pass_smem<P, 6, true>()) || (rc = set_pass_smem<P, 7, true>()) || (rc = set_pass_smem<P, 8, true>()) ||
        (rc = set_pass_smem<P, 9, true>())) return rc;
  }
  return 0;
}

template <class P> __device__ Fe<P> fe_pow_u64(Fe<P> base, uint64_t e) {
  Fe<P> r = fe_one<P>();
  while (e) {
    if (e & 1) r = fe_mul<P>(r, base);
    e >>= 1;
    if (e) base = fe_sqr<P>(base);
  }
  return r;
}

// out[l*K + kk] = a0 * a1^l * (h * g^l)^kk   (all Montgomery form); one thread per 64 kk's
template <class P>
__global__ void ntt_gen_table_kernel(uint32_t* out, uint32_t R, uint64_t K, Fe<P> a0, Fe<P> a1, Fe<P> h, Fe<P> g) {
  constexpr uint64_t CH = 64;
  uint64_t chunks_per_row = (K + CH - 1) / CH;
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= chunks_per_row * R) return;
  uint32_t l = (uint32_t)(t / chunks_per_row);
  uint64_t kk0 = (t % chunks_per_row) * CH;
  Fe<P> A = fe_mul<P>(a0, fe_pow_u64<P>(a1, l));
  Fe<P> B = fe_mul<P>(h, fe_pow_u64<P>(g, l));
  Fe<P> cur = fe_mul<P>(A, fe_pow_u64<P>(B, kk0));
  uint64_t end = kk0 + CH < K ? kk0 + CH : K;
  for (uint64_t kk = kk0; kk < end; kk++) {
    st_fe<P>(out + ((uint64_t)l * K + kk) * P::N, cur);
    cur = fe_mul<P>(cur, B);
  }
}

// ------------------------------------------------------------------ host-side plan
struct PassPlan {
  uint32_t b, logC, logK, logQ, nd, d[3];
  uint32_t* bnd = nullptr;
  uint32_t* post = nullptr;
  uint32_t* wR = nullptr;
};

struct NttPlan {
  int field;
  uint32_t logn;
  int inverse;
  bool has_shift;
  uint32_t w[8], shift[8];
  std::vector<PassPlan> passes;      // in execution order (t = m .. 1)
  uint32_t consts[3 * 8];            // w8, w4, w8^3 (Montgomery)
  size_t bytes = 0;
  void free_tables() {
    for (auto& p : passes) { cudaFree(p.bnd); cudaFree(p.post); cudaFree(p.wR); }
    passes.clear();
  }
};

KzPerSlot<std::list<NttPlan>> g_plans_slots;                 // per device; most recently used first
#define g_plans (g_plans_slots.get())
constexpr size_t kMaxPlans = 8;
struct NttScratchPair { KzScratch s[2]; };
KzPerSlot<NttScratchPair> g_ntt_scratch_slots;
#define g_ntt_scratch (g_ntt_scratch_slots.get().s)

template <class P> Fe<P> host_pow(Fe<P> base, uint64_t e) {
  Fe<P> r = fe_one<P>();
  while (e) {
    if (e & 1) r = fe_mul<P>(r, base);
    e >>= 1;
    if (e) base = fe_sqr<P>(base);
  }
  return r;
}

template <class P>
int gen_table(uint32_t** out, uint32_t R, uint64_t K, const Fe<P>& a0, const Fe<P>& a1, const Fe<P>& h, const Fe<P>& g) {
  KzgpuCtx& cx = kz_ctx();
  size_t bytes = (size_t)R * K * P::N * 4;
  KZ_CUDA(cudaMalloc((void**)out, bytes));
  uint64_t threads = (uint64_t)R * ((K + 63) / 64);
  ntt_gen_table_kernel<P><<<(unsigned)kz_div_up(threads, 128), 128, 0, cx.stream>>>(*out, R, K, a0, a1, h, g);
  KZ_LAUNCHED();
  return 0;
}

template <class P>
int build_plan(NttPlan& pl, uint32_t logn, const Fe<P>& w_canon, int inverse, const Fe<P>* shift_canon) {
  const uint64_t n = 1ull << logn;
  Fe<P> w = fe_to_mont<P>(w_canon);
  if (inverse) w = fe_inv<P>(w);                                   // fft_ff.py:53
  Fe<P> one = fe_one<P>();
  Fe<P> ninv = one;
  if (inverse) {                                                   // fft_ff.py:57
    Fe<P> nn = fe_zero<P>();
    nn.v[0] = (uint32_t)n; nn.v[1] = (uint32_t)(n >> 32);
    ninv = fe_inv<P>(fe_to_mont<P>(nn));
  }
  Fe<P> s = one, sinv = one;
  if (shift_canon) { s = fe_to_mont<P>(*shift_canon); sinv = fe_inv<P>(s); }

  Fe<P> w4 = logn >= 2 ? host_pow<P>(w, n / 4) : one;
  Fe<P> w8 = logn >= 3 ? host_pow<P>(w, n / 8) : one;
  Fe<P> w83 = fe_mul<P>(w8, w4);
  memcpy(pl.consts, w8.v, 32); memcpy(pl.consts + 8, w4.v, 32); memcpy(pl.consts + 16, w83.v, 32);

  uint32_t m = (logn + kMaxB - 1) / kMaxB;
  if (m == 0) m = 1;
  std::vector<uint32_t> bs(m);
  for (uint32_t i = 0; i < m; i++) bs[i] = logn / m + (i < logn % m ? 1 : 0);

  for (int t = (int)m; t >= 1; t--) {
    PassPlan pp;
    pp.b = bs[t - 1];
    uint32_t logJ = 0;
    for (int u = 0; u < t - 1; u++) logJ += bs[u];
    pp.logK = logn - logJ - pp.b;
    pp.logQ = logn - pp.b;
    uint32_t logC = kLogTile - pp.b;
    if (logC > pp.logQ) logC = pp.logQ;
    pp.logC = logC;
    pp.nd = 0;
    for (uint32_t rem = pp.b; rem > 0;) { uint32_t dd = rem >= 3 ? 3 : rem; pp.d[pp.nd++] = dd; rem -= dd; }
    const uint32_t R = 1u << pp.b;
    const uint64_t K = 1ull << pp.logK, J = 1ull << logJ;
    int rc;
    // omega_R^e
    Fe<P> wR = host_pow<P>(w, n >> pp.b);
    if ((rc = gen_table<P>(&pp.wR, 1, R, one, one, wR, one))) return rc;
    pl.bytes += (size_t)R * 32;
    // boundary table T[l][kk] = w^(J l kk) * [fwd coset: s^(J l)] * [inv, t==1, m>=2: 1/n] * [inv coset, t==1: s^-kk]
    bool fwd_coset = shift_canon && !inverse, inv_coset = shift_canon && inverse;
    bool need_bnd = (t < (int)m) || fwd_coset;
    if (need_bnd) {
      Fe<P> a0 = (inverse && t == 1 && m >= 2) ? ninv : one;
      Fe<P> a1 = fwd_coset ? host_pow<P>(s, J) : one;
      Fe<P> h = (inv_coset && t == 1) ? sinv : one;
      Fe<P> g = (t < (int)m) ? host_pow<P>(w, J) : one;
      if ((rc = gen_table<P>(&pp.bnd, R, K, a0, a1, h, g))) return rc;
      pl.bytes += (size_t)R * K * 32;
    }
    // final-store multiplier: [m == 1 inverse: 1/n] * [inverse coset: s^-(kt K)]
    if (t == 1 && ((inverse && m == 1) || inv_coset)) {
      Fe<P> a0 = (inverse && m == 1) ? ninv : one;
      Fe<P> a1 = inv_coset ? host_pow<P>(sinv, K) : one;
      if ((rc = gen_table<P>(&pp.post, R, 1, a0, a1, one, one))) return rc;
    }
    pl.passes.push_back(pp);
  }
  return 0;
}

// Host-buffer entry point (kzgpu_ntt): the first executed pass reads its input as a [R][Q] matrix by column tiles and the
// last one writes a [R][Q] matrix by column tiles, so both can run slab by slab (kSlabs column ranges): slab s of the input
// is uploaded with one 2-D copy and transformed while slab s+1 is in flight, and slab s of the result goes back to the host
// while slab s+1 is computed.  Only one slab of each transfer and the middle passes are not hidden behind PCIe.
struct NttHostIo { const uint64_t* h_in; uint64_t* h_out; };
constexpr uint32_t kSlabs = 4;

template <class P>
int run_plan(const NttPlan& pl, uint32_t* d_data, size_t batch, const NttHostIo* io = nullptr) {
  KzgpuCtx& cx = kz_ctx();
  const uint64_t n = 1ull << pl.logn;
  const size_t bytes = n * batch * 32;
  const size_t m = pl.passes.size();
  // ping-pong so that the last pass writes d_data: data -> s0 [-> s1 -> ...] -> data
  int rc;
  if ((rc = g_ntt_scratch[0].ensure(bytes))) return rc;
  if (m >= 3 && (rc = g_ntt_scratch[1].ensure(bytes))) return rc;
  NttConsts<P> c;
  memcpy(c.w8.v, pl.consts, 32); memcpy(c.w4.v, pl.consts + 8, 32); memcpy(c.w8_3.v, pl.consts + 16, 32);
  static bool attr_set_slot[KZ_MAX_DEV] = {false};       // function attributes are per device
  bool& attr_set = attr_set_slot[kz_slot()];
  if (!attr_set) {
    KZ_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<FrBN254>, cudaFuncAttributeMaxDynamicSharedMemorySize, (8 * 4 << kLogTile) + (32 << kMaxB)));
    KZ_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<FrBLS381>, cudaFuncAttributeMaxDynamicSharedMemorySize, (8 * 4 << kLogTile) + (32 << kMaxB)));
    { int rc2; if ((rc2 = set_all_pass_smem<FrBN254>()) || (rc2 = set_all_pass_smem<FrBLS381>())) return rc2; }
    attr_set = true;
  }
  bool lazy = FeLz<P>::ok && !getenv("KZGPU_NTT_STRICT");
  for (size_t i = 0; i < m; i++) lazy = lazy && pass_is_full(pl.passes[i].b, pl.passes[i].logC);   // every pass must speak [0, 2r)
  const uint32_t* src = d_data;
  for (size_t i = 0; i < m; i++) {
    const PassPlan& pp = pl.passes[i];
    uint32_t* dst;
    if (i == m - 1) dst = (m == 1) ? (uint32_t*)g_ntt_scratch[0].p : d_data;
    else dst = (uint32_t*)g_ntt_scratch[(i & 1)].p;
    if (dst == src) dst = (uint32_t*)g_ntt_scratch[1].p;   // cannot happen with the scheme above; defensive
    NttPassArgs a;
    a.in = src; a.out = dst; a.bnd = pp.bnd; a.post = pp.post; a.wR = pp.wR;
    a.n = n; a.b = pp.b; a.logC = pp.logC; a.logK = pp.logK; a.logQ = pp.logQ; a.nd = pp.nd;
    a.dpack = 0;
    for (uint32_t k = 0; k < pp.nd; k++) a.dpack |= pp.d[k] << (2 * k);
    a.blk0 = 0;
    a.lazy = lazy ? 1u : 0u; a.last = (i == m - 1) ? 1u : 0u;
    uint32_t tile = 1u << (pp.b + pp.logC);
    uint32_t threads = tile / 8 < 32 ? 32 : (tile / 8 > (uint32_t)kThreads ? kThreads : tile / 8);
    const uint32_t tiles = (uint32_t)(1ull << (pp.logQ - pp.logC));
    const size_t smem = (size_t)tile * 32 + (pp.nd > 1 ? (32u << pp.b) : 0u);
    const bool slab_in = io && m >= 2 && i == 0, slab_out = io && m >= 2 && i == m - 1 && pp.logK == pp.logQ;
    if (slab_in || slab_out) {
      // [R][Q] matrix, slab = tiles/kSlabs column tiles = (Q / kSlabs) columns of every row
      static cudaEvent_t done_ev_slot[KZ_MAX_DEV][kSlabs] = {{nullptr}};
      cudaEvent_t* done_ev = done_ev_slot[kz_slot()];
      if (!done_ev[0]) for (uint32_t k = 0; k < kSlabs; k++) KZ_CUDA(cudaEventCreateWithFlags(&done_ev[k], cudaEventDisableTiming));
      const size_t R = 1ull << pp.b, Q = 1ull << pp.logQ, pitch = Q * 32, width = pitch / kSlabs;
      if (slab_in)
        for (uint32_t k = 0; k < kSlabs; k++) {
          KZ_CUDA(cudaMemcpy2DAsync((char*)src + k * width, pitch, (const char*)io->h_in + k * width, pitch, width, R,
                                    cudaMemcpyHostToDevice, cx.copy_stream));
          KZ_CUDA(cudaEventRecord(cx.copy_ev[k], cx.copy_stream));
        }
      for (uint32_t k = 0; k < kSlabs; k++) {
        if (slab_in) KZ_CUDA(cudaStreamWaitEvent(cx.stream, cx.copy_ev[k], 0));
        a.blk0 = k * (tiles / kSlabs);
        KzProf prof(1);
        launch_pass<P>(a, c, dim3(tiles / kSlabs, 1), threads, smem, cx.stream);
        KZ_LAUNCHED();
        prof.stop(k == 0 ? 1 : 0, k == 0 ? (double)n : 0.0);
        if (slab_out) {
          KZ_CUDA(cudaEventRecord(done_ev[k], cx.stream));
          KZ_CUDA(cudaStreamWaitEvent(cx.copy_stream, done_ev[k], 0));
          KZ_CUDA(cudaMemcpy2DAsync((char*)io->h_out + k * width, pitch, (const char*)dst + k * width, pitch, width, R,
                                    cudaMemcpyDeviceToHost, cx.copy_stream));
        }
      }
      if (slab_out) KZ_CUDA(cudaStreamSynchronize(cx.copy_stream));
      src = dst;
      continue;
    }
    dim3 grid((unsigned)tiles, (unsigned)batch);
    KzProf prof(1);
    launch_pass<P>(a, c, grid, threads, smem, cx.stream);
    KZ_LAUNCHED();
    prof.stop(1, (double)n * (double)batch);
    src = dst;
  }
  if (m == 1) KZ_CUDA(cudaMemcpyAsync(d_data, src, bytes, cudaMemcpyDeviceToDevice, cx.stream));
  if (io && !(m >= 2 && pl.passes[m - 1].logK == pl.passes[m - 1].logQ)) {      // no slab-wise download possible: plain copy
    KZ_CUDA(cudaMemcpyAsync(io->h_out, d_data, bytes, cudaMemcpyDeviceToHost, cx.stream));
    KZ_CUDA(cudaStreamSynchronize(cx.stream));
  }
  return 0;
}

template <class P>
int ntt_impl(int field, uint32_t* d_data, size_t n, size_t batch, const uint64_t* w, int inverse, const uint64_t* shift,
             const NttHostIo* io = nullptr) {
  if (n == 0 || (n & (n - 1))) return kz_fail(KZGPU_EINVAL, "NTT length %zu is not a power of two (fft_ff.py:74)", n);
  if (batch == 0) return 0;
  uint32_t logn = 0;
  while ((1ull << logn) < n) logn++;
  if (logn > 32) return kz_fail(KZGPU_EINVAL, "NTT length 2^%u too large", logn);
  Fe<P> wc = kz_fe_from_u64<P>(w), sc;
  if (!kz_fe_reduced<P>(wc)) return kz_fail(KZGPU_ERANGE, "w is not a canonical field element");
  if (shift) {
    sc = kz_fe_from_u64<P>(shift);
    if (!kz_fe_reduced<P>(sc) || fe_is_zero<P>(sc)) return kz_fail(KZGPU_ERANGE, "coset shift must be a non-zero canonical element");
  }
  if (logn == 0) {
    // n == 1: fft_ff returns its input (fft_ff.py:16-17); inverse scales by 1; coset shift^0 = 1
    return 0;
  }
  if (fe_is_zero<P>(wc)) return kz_fail(KZGPU_ERANGE, "w must be non-zero");
  // plan lookup
  for (auto it = g_plans.begin(); it != g_plans.end(); ++it) {
    if (it->field == field && it->logn == logn && it->inverse == (inverse ? 1 : 0) && it->has_shift == (shift != nullptr) &&
        !memcmp(it->w, wc.v, 32) && (!shift || !memcmp(it->shift, sc.v, 32))) {
      g_plans.splice(g_plans.begin(), g_plans, it);
      return run_plan<P>(g_plans.front(), d_data, batch, io);
    }
  }
  NttPlan pl;
  pl.field = field; pl.logn = logn; pl.inverse = inverse ? 1 : 0; pl.has_shift = shift != nullptr;
  memcpy(pl.w, wc.v, 32);
  if (shift) memcpy(pl.shift, sc.v, 32); else memset(pl.shift, 0, 32);
  int rc = build_plan<P>(pl, logn, wc, inverse, shift ? &sc : nullptr);
  if (rc) { pl.free_tables(); return rc; }
  g_plans.push_front(pl);
  while (g_plans.size() > kMaxPlans) { g_plans.back().free_tables(); g_plans.pop_back(); }
  return run_plan<P>(g_plans.front(), d_data, batch, io);
}

int ntt_dispatch(int field, uint32_t* d_data, size_t n, size_t batch, const uint64_t* w, int inverse, const uint64_t* shift,
                 const NttHostIo* io = nullptr) {
  if (field == KZGPU_BN254) return ntt_impl<FrBN254>(field, d_data, n, batch, w, inverse, shift, io);
  if (field == KZGPU_BLS12_381) return ntt_impl<FrBLS381>(field, d_data, n, batch, w, inverse, shift, io);
  return kz_fail(KZGPU_EINVAL, "unknown field id %d", field);
}

KzPerSlot<KzScratch> g_ntt_io_slots;
#define g_ntt_io (g_ntt_io_slots.get())

}  // namespace

void kz_ntt_release() {
  for (auto& p : g_plans) p.free_tables();
  g_plans.clear();
  g_ntt_scratch[0].release(); g_ntt_scratch[1].release(); g_ntt_io.release();
}

extern "C" {

int kzgpu_ntt_batch_dev(int field, uint64_t* d_data, size_t n, size_t batch, const uint64_t* w, int inverse,
                        const uint64_t* coset_shift) {
  KZ_REQUIRE_INIT();
  if (!d_data || !w) return kz_fail(KZGPU_EINVAL, "null pointer");
  return ntt_dispatch(field, (uint32_t*)d_data, n, batch, w, inverse, coset_shift);
}

int kzgpu_ntt_dev(int field, uint64_t* d_data, size_t n, const uint64_t* w, int inverse, const uint64_t* coset_shift) {
  return kzgpu_ntt_batch_dev(field, d_data, n, 1, w, inverse, coset_shift);
}

// `batch` host vectors on the calling thread's device.
//  * one long vector: transfers pipelined with the first and last pass, slab by slab (run_plan);
//  * several long vectors: a per-vector pipeline over three streams -- vector k+1 uploads (copy stream) while vector k is
//    transformed (main stream) and vector k-1 downloads (download stream), so the PCIe link runs full duplex;
//  * otherwise one upload, one batched transform, one download.
static int ntt_batch_host_on_slot(int field, uint64_t* data, size_t n, size_t batch, const uint64_t* w, int inverse,
                                  const uint64_t* coset_shift) {
  KzgpuCtx& cx = kz_ctx();
  size_t bytes = n * batch * 32;
  if (bytes == 0) return 0;
  int rc = g_ntt_io.ensure(bytes);
  if (rc) return rc;
  const bool pinned = kz_host_is_pinned(data);           // the pipelined paths need truly asynchronous copies
  if (batch == 1 && n >= (1u << 20) && pinned && !getenv("KZGPU_NTT_NO_OVERLAP")) {
    // one long vector: transfers pipelined with the first and last pass (>= 2 passes at this size, >= kSlabs tiles each)
    NttHostIo io{data, data};
    return ntt_dispatch(field, (uint32_t*)g_ntt_io.p, n, 1, w, inverse, coset_shift, &io);
  }
  if (batch > 1 && n >= (1u << 18) && pinned && !getenv("KZGPU_NTT_NO_OVERLAP")) {
    static cudaEvent_t ev_slot[KZ_MAX_DEV][2][4] = {{{nullptr}}};          // [slot][up | done][vector & 3]
    cudaEvent_t (*ev)[4] = ev_slot[kz_slot()];
    if (!ev[0][0]) for (int a = 0; a < 2; a++) for (int k = 0; k < 4; k++) KZ_CUDA(cudaEventCreateWithFlags(&ev[a][k], cudaEventDisableTiming));
    KZ_CUDA(cudaEventRecord(cx.start_ev, cx.stream));                      // everything queued so far precedes the first upload
    KZ_CUDA(cudaStreamWaitEvent(cx.copy_stream, cx.start_ev, 0));
    const size_t vbytes = n * 32;
    for (size_t v = 0; v < batch; v++) {
      char* d = (char*)g_ntt_io.p + v * vbytes;
      KZ_CUDA(cudaMemcpyAsync(d, (const char*)data + v * vbytes, vbytes, cudaMemcpyHostToDevice, cx.copy_stream));
      KZ_CUDA(cudaEventRecord(ev[0][v & 3], cx.copy_stream));
      KZ_CUDA(cudaStreamWaitEvent(cx.stream, ev[0][v & 3], 0));
      if ((rc = ntt_dispatch(field, (uint32_t*)d, n, 1, w, inverse, coset_shift))) return rc;
      KZ_CUDA(cudaEventRecord(ev[1][v & 3], cx.stream));
      KZ_CUDA(cudaStreamWaitEvent(cx.d2h_stream, ev[1][v & 3], 0));
      KZ_CUDA(cudaMemcpyAsync((char*)data + v * vbytes, d, vbytes, cudaMemcpyDeviceToHost, cx.d2h_stream));
      if (v >= 3) KZ_CUDA(cudaEventSynchronize(ev[1][(v - 3) & 3]));        // an event slot is re-recorded only once its waiters are past it
    }
    KZ_CUDA(cudaStreamSynchronize(cx.d2h_stream));
    KZ_CUDA(cudaStreamSynchronize(cx.stream));
    return 0;
  }
  // pageable caller memory (what the Python drop-in hands over) or short vectors: staged upload, transform, staged download
  if ((rc = kz_upload(g_ntt_io.p, data, bytes, cx.stream))) return rc;
  rc = ntt_dispatch(field, (uint32_t*)g_ntt_io.p, n, batch, w, inverse, coset_shift);
  if (rc) return rc;
  if ((rc = kz_download(data, g_ntt_io.p, bytes, cx.stream))) return rc;
  KZ_CUDA(cudaStreamSynchronize(cx.stream));
  return 0;
}

// Several devices (kzgpu_init_multi): whole vectors per device -- vector v of the batch runs on device v * ndev / batch
// (contiguous blocks), each device running the single-device path above on its block; no exchange.
int kzgpu_ntt_batch(int field, uint64_t* data, size_t n, size_t batch, const uint64_t* w, int inverse,
                    const uint64_t* coset_shift) {
  KZ_REQUIRE_INIT();
  if (!data || !w) return kz_fail(KZGPU_EINVAL, "null pointer");
  if (n == 0 || (n & (n - 1))) return kz_fail(KZGPU_EINVAL, "NTT length %zu is not a power of two (fft_ff.py:74)", n);
  const int nd = kz_ndev();
  if (nd == 1 || batch < 2 || n * batch < (1u << 16)) return ntt_batch_host_on_slot(field, data, n, batch, w, inverse, coset_shift);
  return kz_parallel([&](int slot) -> int {
    const size_t lo = batch * slot / nd, hi = batch * (slot + 1) / nd;
    if (hi <= lo) return 0;
    return ntt_batch_host_on_slot(field, data + lo * n * 4, n, hi - lo, w, inverse, coset_shift);
  });
}

int kzgpu_ntt(int field, uint64_t* data, size_t n, const uint64_t* w, int inverse, const uint64_t* coset_shift) {
  return kzgpu_ntt_batch(field, data, n, 1, w, inverse, coset_shift);
}

}  // extern "C"
