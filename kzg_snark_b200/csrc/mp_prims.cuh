// Multiprecision carry-chain primitives on 32-bit limbs.
//
// Device: MpPrims<N> from mp_prims_gen.cuh (inline PTX, one statement per carry chain).
// Host:   the same contract in portable C++.  Used by tests/host/host_arith_test.cu to check the
//         Montgomery / curve formulas on the CPU build box, and by the library for MARSHALLING only:
//         msm.cu normalises the ONE XYZZ result of an MSM to canonical affine limbs on the host
//         (host_fe_inv / host_xyzz_to_canonical: a single inversion is ~20 us there against ~200 us
//         for a lone GPU thread) and builds the 255-entry doubling table of the generator for
//         kzgpu_srs_generate.  No O(n) field arithmetic ever runs on the host.
#pragma once
#include <cstdint>
#ifdef __CUDACC__
#include "mp_prims_gen.cuh"
#endif

template <int N> struct MpPrimsHost {
  static inline void mul_even(uint32_t* acc, const uint32_t* a, uint32_t b) {
    for (int j = 0; j < N; j += 2) {
      uint64_t t = (uint64_t)a[j] * b;
      acc[j] = (uint32_t)t; acc[j + 1] = (uint32_t)(t >> 32);
    }
  }
  static inline void mad_even(uint32_t* acc, const uint32_t* a, uint32_t b, uint32_t& top) {
    uint64_t c = 0;
    for (int j = 0; j < N; j += 2) {
      unsigned __int128 t = (unsigned __int128)((uint64_t)a[j] * b) + (((uint64_t)acc[j + 1] << 32) | acc[j]) + c;
      acc[j] = (uint32_t)t; acc[j + 1] = (uint32_t)(t >> 32); c = (uint64_t)(t >> 64);
    }
    top += (uint32_t)c;
  }
  static inline void mad_even_nc(uint32_t* acc, const uint32_t* a, uint32_t b) {
    uint32_t dummy = 0; mad_even(acc, a, b, dummy);
  }
  static inline void shift_mad(uint32_t* x, uint32_t& y0, const uint32_t* a, uint32_t b) {
    uint64_t s = (uint64_t)y0 + x[1];
    y0 = (uint32_t)s;
    uint64_t c = s >> 32;
    for (int j = 0; j < N; j += 2) {
      uint64_t add = (j + 3 < N + 1 && j + 2 < N) ? ((((uint64_t)x[j + 3]) << 32) | x[j + 2]) : 0;
      unsigned __int128 t = (unsigned __int128)((uint64_t)a[j] * b) + add + c;
      x[j] = (uint32_t)t; x[j + 1] = (uint32_t)(t >> 32); c = (uint64_t)(t >> 64);
    }
  }
  // squaring rows (field.cuh fe_sqr_nofinal): the same two chains with the first K products left out
  template <int K> static inline void mad_even_s(uint32_t* acc, const uint32_t* a, uint32_t b, uint32_t& top) {
    uint64_t c = 0;
    for (int j = 2 * K; j < N; j += 2) {
      unsigned __int128 t = (unsigned __int128)((uint64_t)a[j] * b) + (((uint64_t)acc[j + 1] << 32) | acc[j]) + c;
      acc[j] = (uint32_t)t; acc[j + 1] = (uint32_t)(t >> 32); c = (uint64_t)(t >> 64);
    }
    top += (uint32_t)c;
  }
  template <int K> static inline void shift_mad_s(uint32_t* x, uint32_t& y0, const uint32_t* a, uint32_t b) {
    uint64_t s = (uint64_t)y0 + x[1];
    y0 = (uint32_t)s;
    uint64_t c = s >> 32;
    for (int j = 0; j < N; j += 2) {
      uint64_t add = (j + 2 < N) ? ((((uint64_t)x[j + 3]) << 32) | x[j + 2]) : 0;
      uint64_t prod = j >= 2 * K ? (uint64_t)a[j] * b : 0;
      unsigned __int128 t = (unsigned __int128)prod + add + c;
      x[j] = (uint32_t)t; x[j + 1] = (uint32_t)(t >> 32); c = (uint64_t)(t >> 64);
    }
  }
  static inline void merge(uint32_t* r, const uint32_t* o, const uint32_t* e) {
    uint64_t c = 0;
    for (int k = 0; k < N; k++) {
      uint64_t t = (uint64_t)o[k] + (k + 1 < N ? e[k + 1] : 0) + c;
      r[k] = (uint32_t)t; c = t >> 32;
    }
  }
  static inline uint32_t add_cc(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint64_t c = 0;
    for (int k = 0; k < N; k++) { uint64_t t = (uint64_t)a[k] + b[k] + c; r[k] = (uint32_t)t; c = t >> 32; }
    return (uint32_t)c;
  }
  static inline uint32_t sub_cc(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint64_t br = 0;
    for (int k = 0; k < N; k++) { uint64_t t = (uint64_t)a[k] - b[k] - br; r[k] = (uint32_t)t; br = (t >> 32) & 1; }
    return br ? 0xffffffffu : 0u;
  }
};

#ifdef __CUDA_ARCH__
template <int N> using Mp = MpPrims<N>;
#else
template <int N> using Mp = MpPrimsHost<N>;
#endif
