// Montgomery prime-field arithmetic on 32-bit limbs (8 limbs: BN254 p/r, BLS12-381 r;
// 12 limbs: BLS12-381 p).  Replaces, for the hot path, the arithmetic the reference
// delegates to py_ecc's pure-Python FQ (kzg.py:27-35) and Sage's GF(r) (kzg.py:52).
//
// Representation: little-endian uint32 limbs, value < p.  "Montgomery form" of x is
// x * 2^(32N) mod p.  mont_mul(a, b) = a * b * 2^(-32N) mod p for a, b < p.
//
// mont_mul keeps two accumulators E ("even-aligned", limb k at bit 32k) and O
// ("odd-aligned", limb k at bit 32(k+1)) so that every 32x32->64 product of a row lands
// on a (lo,hi) register pair of one accumulator and the whole row is one carry chain of
// IMAD.WIDE.U32.X.  After each row the Montgomery step makes E[0] zero; dividing by 2^32
// then simply swaps the roles of E and O (E>>32 is odd-aligned again after dropping E[0]
// and folding E[1] into the new even accumulator).  Bound: the running value stays < 2p
// after each division and < 2p * 2^32 before it, so it fits positions 0..N as long as
// p <= 2^(32N-1) -- true for all four moduli (254/255/254/381 bits in 256/256/256/384).
#pragma once
#include <cstdint>
#include <utility>
#include "mp_prims.cuh"

#ifdef __CUDACC__
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

template <class P> struct Fe {
  static constexpr int N = P::N;
  uint32_t v[P::N];
};

template <class P> HD void fe_load_mod(uint32_t* m) {
#pragma unroll
  for (int i = 0; i < P::N; i++) m[i] = P::mod(i);
}

template <class P> HD bool fe_is_zero(const Fe<P>& a) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < P::N; i++) o |= a.v[i];
  return o == 0;
}

template <class P> HD bool fe_eq(const Fe<P>& a, const Fe<P>& b) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < P::N; i++) o |= a.v[i] ^ b.v[i];
  return o == 0;
}

template <class P> HD Fe<P> fe_zero() {
  Fe<P> r;
#pragma unroll
  for (int i = 0; i < P::N; i++) r.v[i] = 0;
  return r;
}

// 1 in Montgomery form (2^(32N) mod p)
template <class P> HD Fe<P> fe_one() {
  Fe<P> r;
#pragma unroll
  for (int i = 0; i < P::N; i++) r.v[i] = P::one(i);
  return r;
}

// r = (x >= p) ? x - p : x   for x < 2p
template <class P> HD void fe_final_sub(uint32_t* r, const uint32_t* x) {
  constexpr int N = P::N;
  uint32_t m[N], t[N];
  fe_load_mod<P>(m);
  uint32_t borrow = Mp<N>::sub_cc(t, x, m);
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = borrow ? x[i] : t[i];
}

template <class P> HD Fe<P> fe_add(const Fe<P>& a, const Fe<P>& b) {
  constexpr int N = P::N;
  uint32_t s[N];
  Mp<N>::add_cc(s, a.v, b.v);          // a + b < 2p < 2^(32N): no carry out
  Fe<P> r;
  fe_final_sub<P>(r.v, s);
  return r;
}

template <class P> HD Fe<P> fe_sub(const Fe<P>& a, const Fe<P>& b) {
  constexpr int N = P::N;
  uint32_t d[N], m[N];
  uint32_t borrow = Mp<N>::sub_cc(d, a.v, b.v);
#pragma unroll
  for (int i = 0; i < N; i++) m[i] = P::mod(i) & borrow;
  Fe<P> r;
  Mp<N>::add_cc(r.v, d, m);
  return r;
}

template <class P> HD Fe<P> fe_neg(const Fe<P>& a) {
  constexpr int N = P::N;
  uint32_t m[N];
  fe_load_mod<P>(m);
  Fe<P> r;
  Mp<N>::sub_cc(r.v, m, a.v);
  uint32_t nz = 0;
#pragma unroll
  for (int i = 0; i < N; i++) nz |= a.v[i];
#pragma unroll
  for (int i = 0; i < N; i++) r.v[i] = nz ? r.v[i] : 0u;
  return r;
}

template <class P> HD Fe<P> fe_dbl(const Fe<P>& a) { return fe_add<P>(a, a); }

// One Montgomery row-reduction: make E[0] zero by adding m*p, m = E[0] * (-p^-1 mod 2^32).
template <class P> HD void fe_redc_row(uint32_t* E, uint32_t* O, const uint32_t* mod) {
  constexpr int N = P::N;
  uint32_t m = E[0] * P::INV32;
  Mp<N>::mad_even_nc(O, mod + 1, m);
  Mp<N>::mad_even(E, mod, m, O[N - 1]);
}

// Montgomery product without the final conditional subtraction: t = (a*b + m*p) / 2^(32N) < a*b / 2^(32N) + p
template <class P> HD void fe_mul_nofinal(uint32_t* t, const Fe<P>& a, const Fe<P>& b) {
  constexpr int N = P::N;
  uint32_t mod[N], E[N], O[N];
  fe_load_mod<P>(mod);
  // row 0
  Mp<N>::mul_even(E, a.v, b.v[0]);
  Mp<N>::mul_even(O, a.v + 1, b.v[0]);
  fe_redc_row<P>(E, O, mod);
#pragma unroll
  for (int i = 1; i < N; i += 2) {
    // odd row: roles swapped (O is even-aligned now, E is shifted into odd alignment)
    Mp<N>::shift_mad(E, O[0], a.v + 1, b.v[i]);
    Mp<N>::mad_even(O, a.v, b.v[i], E[N - 1]);
    fe_redc_row<P>(O, E, mod);
    if (i + 1 < N) {
      Mp<N>::shift_mad(O, E[0], a.v + 1, b.v[i + 1]);
      Mp<N>::mad_even(E, a.v, b.v[i + 1], O[N - 1]);
      fe_redc_row<P>(E, O, mod);
    }
  }
  // N is even: after the last (odd) row the even-aligned accumulator is O, with O[0] == 0
  Mp<N>::merge(t, E, O);
}

template <class P> HD Fe<P> fe_mul(const Fe<P>& a, const Fe<P>& b) {
  uint32_t t[P::N];
  fe_mul_nofinal<P>(t, a, b);
  Fe<P> r;
  fe_final_sub<P>(r.v, t);
  return r;
}

// ---- lazy ("semi-reduced") arithmetic: values in [0, 2p).  Usable when 4p <= 2^(32N) (BN254 p and r, BLS12-381 p; not
// BLS12-381 r): for a, b < 2p the Montgomery product is < 4p^2 / 2^(32N) + p <= 2p, so the final conditional subtraction
// (17 ALU instructions of ~200) can be dropped; for a < 4p and a canonical b < p it is < 2p as well.  The running value of
// the row loop stays < (a + p) * 2^32 < 2^(32(N+1)), so no carry is lost.  Additions and subtractions fold back into
// [0, 2p) at the cost of the canonical ones; the *_nr forms skip that when the result only feeds a multiplication.
template <class P> struct FeLz { static constexpr bool ok = (P::mod(P::N - 1) >> 30) == 0; };

template <class P> HD void fe_load_mod2(uint32_t* m) {
#pragma unroll
  for (int i = 0; i < P::N; i++) m[i] = (P::mod(i) << 1) | (i ? (P::mod(i > 0 ? i - 1 : 0) >> 31) : 0u);
}

template <class P> HD Fe<P> fe_mul_lz(const Fe<P>& a, const Fe<P>& b) {
  Fe<P> r;
  fe_mul_nofinal<P>(r.v, a, b);
  return r;
}

// a, b < 2p -> a + b folded into [0, 2p)
template <class P> HD Fe<P> fe_add_lz(const Fe<P>& a, const Fe<P>& b) {
  constexpr int N = P::N;
  uint32_t s[N], m[N], t[N];
  Mp<N>::add_cc(s, a.v, b.v);          // < 4p <= 2^(32N): no carry out
  fe_load_mod2<P>(m);
  uint32_t borrow = Mp<N>::sub_cc(t, s, m);
  Fe<P> r;
#pragma unroll
  for (int i = 0; i < N; i++) r.v[i] = borrow ? s[i] : t[i];
  return r;
}

// a, b < 2p -> a - b folded into [0, 2p)
template <class P> HD Fe<P> fe_sub_lz(const Fe<P>& a, const Fe<P>& b) {
  constexpr int N = P::N;
  uint32_t d[N], m[N];
  uint32_t borrow = Mp<N>::sub_cc(d, a.v, b.v);
  fe_load_mod2<P>(m);
#pragma unroll
  for (int i = 0; i < N; i++) m[i] &= borrow;
  Fe<P> r;
  Mp<N>::add_cc(r.v, d, m);
  return r;
}

// a, b < 2p -> a + b < 4p, not folded (multiplication operand only)
template <class P> HD Fe<P> fe_add_nr(const Fe<P>& a, const Fe<P>& b) {
  Fe<P> r;
  Mp<P::N>::add_cc(r.v, a.v, b.v);
  return r;
}

// a, b < 2p -> a - b + 2p in (0, 4p), not folded (multiplication operand only)
template <class P> HD Fe<P> fe_sub_nr(const Fe<P>& a, const Fe<P>& b) {
  constexpr int N = P::N;
  uint32_t m[N], t[N];
  fe_load_mod2<P>(m);
  Mp<N>::add_cc(t, a.v, m);            // < 4p: no carry out
  Fe<P> r;
  Mp<N>::sub_cc(r.v, t, b.v);
  return r;
}

// x < 2p is a multiple of p
template <class P> HD bool fe_is_zero_lz(const Fe<P>& a) {
  uint32_t o0 = 0, o1 = 0;
#pragma unroll
  for (int i = 0; i < P::N; i++) { o0 |= a.v[i]; o1 |= a.v[i] ^ P::mod(i); }
  return o0 == 0 || o1 == 0;
}

// x < 2p -> canonical
template <class P> HD Fe<P> fe_reduce_lz(const Fe<P>& a) {
  Fe<P> r;
  fe_final_sub<P>(r.v, a.v);
  return r;
}

// Dual product with ONE Montgomery reduction: t = (a*b + c*d + m*p) / 2^(32N) < (a*b + c*d) / 2^(32N) + p.  Same row structure as
// fe_mul_nofinal, every row adding a*b_i and c*d_i before its reduction step, so the second product costs N^2 wide
// multiplies instead of 2N^2 - N wide + 2N narrow.  The running value is < (a + c + p) * 2^32, which fits the N + 1 limb
// positions iff a + c + p <= 2^(32N): for semi-reduced operands a, c <= 2p that is 5p <= 2^(32N) (FeSq<P>::ok, defined below:
// BN254 p and r, BLS12-381 p).  For a, b, c, d <= 2p the result is < 8p^2 / 2^(32N) + p, i.e. < 2.52p for BN254 p and
// < 1.82p for BLS12-381 p: fold with fe_fold2_lz before using it as a semi-reduced value.
template <class P> HD void fe_mul2_nofinal(uint32_t* t, const Fe<P>& a, const Fe<P>& b, const Fe<P>& c, const Fe<P>& d) {
  constexpr int N = P::N;
  uint32_t mod[N], E[N], O[N];
  fe_load_mod<P>(mod);
  Mp<N>::mul_even(E, a.v, b.v[0]);
  Mp<N>::mul_even(O, a.v + 1, b.v[0]);
  Mp<N>::mad_even_nc(O, c.v + 1, d.v[0]);
  Mp<N>::mad_even(E, c.v, d.v[0], O[N - 1]);
  fe_redc_row<P>(E, O, mod);
#pragma unroll
  for (int i = 1; i < N; i += 2) {
    Mp<N>::shift_mad(E, O[0], a.v + 1, b.v[i]);
    Mp<N>::mad_even_nc(E, c.v + 1, d.v[i]);
    Mp<N>::mad_even(O, a.v, b.v[i], E[N - 1]);
    Mp<N>::mad_even(O, c.v, d.v[i], E[N - 1]);
    fe_redc_row<P>(O, E, mod);
    if (i + 1 < N) {
      Mp<N>::shift_mad(O, E[0], a.v + 1, b.v[i + 1]);
      Mp<N>::mad_even_nc(O, c.v + 1, d.v[i + 1]);
      Mp<N>::mad_even(E, a.v, b.v[i + 1], O[N - 1]);
      Mp<N>::mad_even(E, c.v, d.v[i + 1], O[N - 1]);
      fe_redc_row<P>(E, O, mod);
    }
  }
  Mp<N>::merge(t, E, O);
}

// x < 4p -> x or x - 2p, in [0, 2p)
template <class P> HD Fe<P> fe_fold2_lz(const uint32_t* x) {
  constexpr int N = P::N;
  uint32_t m[N], t[N];
  fe_load_mod2<P>(m);
  uint32_t borrow = Mp<N>::sub_cc(t, x, m);
  Fe<P> r;
#pragma unroll
  for (int i = 0; i < N; i++) r.v[i] = borrow ? x[i] : t[i];
  return r;
}

// a <= 2p -> 2p - a in [0, 2p], not folded (multiplication operand only: -a up to a multiple of p)
template <class P> HD Fe<P> fe_neg_nr(const Fe<P>& a) {
  constexpr int N = P::N;
  uint32_t m[N];
  fe_load_mod2<P>(m);
  Fe<P> r;
  Mp<N>::sub_cc(r.v, m, a.v);
  return r;
}

// Montgomery square, same row structure as fe_mul_nofinal with row i reduced to the products a_i * a_j, j >= i
// (N(N+1)/2 instead of N^2 wide multiplies): the multiplicand of row i is  [a_i | 2 * (a >> 32(i+1))], i.e. limb i is
// a_i, limb i+1 is a_(i+1) << 1 and the limbs above are those of 2a (funnel shifts, computed once); limbs below i are
// never read.  Needs a < 2^(32N-1) (true for canonical and for semi-reduced values of every modulus here).  Result
// < a^2 / 2^(32N) + p like the product; the running value stays < (2a + p) * 2^32 < 2^(32(N+1)) for a < 2p <= 2^(32N)/2.
template <class P, int I> HD void fe_sqr_row(uint32_t* E, uint32_t* O, uint32_t* Mv, const uint32_t* a, const uint32_t* mod) {
  constexpr int N = P::N;
  Mv[I] = a[I];
  if constexpr (I + 1 < N) Mv[I + 1] = a[I + 1] << 1;
  if constexpr (I & 1) {
    Mp<N>::template shift_mad_s<I / 2>(E, O[0], Mv + 1, a[I]);
    Mp<N>::template mad_even_s<(I + 1) / 2>(O, Mv, a[I], E[N - 1]);
    fe_redc_row<P>(O, E, mod);
  } else {
    Mp<N>::template shift_mad_s<I / 2>(O, E[0], Mv + 1, a[I]);
    Mp<N>::template mad_even_s<(I + 1) / 2>(E, Mv, a[I], O[N - 1]);
    fe_redc_row<P>(E, O, mod);
  }
}
template <class P, int... I> HD void fe_sqr_rows(uint32_t* E, uint32_t* O, uint32_t* Mv, const uint32_t* a, const uint32_t* mod,
                                                  std::integer_sequence<int, I...>) {
  (fe_sqr_row<P, I + 1>(E, O, Mv, a, mod), ...);
}
// the doubled multiplicand makes the running value < (2a + p) * 2^32: it fits the N+1 positions for a < 2p iff 5p <= 2^(32N)
// (BN254 p and r, BLS12-381 p); BLS12-381 r squares through the general product
template <class P> struct FeSq { static constexpr bool ok = P::mod(P::N - 1) < 0x33333333u; };

template <class P> HD void fe_sqr_nofinal(uint32_t* t, const Fe<P>& a) {
  constexpr int N = P::N;
  if constexpr (!FeSq<P>::ok) { fe_mul_nofinal<P>(t, a, a); return; }
  uint32_t mod[N], E[N], O[N], Mv[N];
  fe_load_mod<P>(mod);
  Mv[0] = a.v[0];
  Mv[1] = a.v[1] << 1;
#pragma unroll
  for (int j = 2; j < N; j++) Mv[j] = (a.v[j] << 1) | (a.v[j - 1] >> 31);
  Mp<N>::mul_even(E, Mv, a.v[0]);
  Mp<N>::mul_even(O, Mv + 1, a.v[0]);
  fe_redc_row<P>(E, O, mod);
  fe_sqr_rows<P>(E, O, Mv, a.v, mod, std::make_integer_sequence<int, N - 1>{});
  Mp<N>::merge(t, E, O);
}

template <class P> HD Fe<P> fe_sqr(const Fe<P>& a) {
  uint32_t t[P::N];
  fe_sqr_nofinal<P>(t, a);
  Fe<P> r;
  fe_final_sub<P>(r.v, t);
  return r;
}


// a*b - c*d for semi-reduced operands, semi-reduced result, one Montgomery reduction (dual product with c replaced by 2p - c)
// where the modulus leaves the headroom (5p <= 2^(32N)), two products otherwise
template <class P> HD Fe<P> fe_mulsub_lz(const Fe<P>& a, const Fe<P>& b, const Fe<P>& c, const Fe<P>& d) {
  if constexpr (FeSq<P>::ok) {
    uint32_t t[P::N];
    fe_mul2_nofinal<P>(t, a, b, fe_neg_nr<P>(c), d);
    return fe_fold2_lz<P>(t);
  } else {
    return fe_sub_lz<P>(fe_mul_lz<P>(a, b), fe_mul_lz<P>(c, d));
  }
}

// semi-reduced square: a < 2p -> a^2 / 2^(32N) mod p in [0, 2p)
template <class P> HD Fe<P> fe_sqr_lz(const Fe<P>& a) {
  Fe<P> r;
  fe_sqr_nofinal<P>(r.v, a);
  return r;
}

template <class P> HD Fe<P> fe_to_mont(const Fe<P>& a) {
  Fe<P> r2;
#pragma unroll
  for (int i = 0; i < P::N; i++) r2.v[i] = P::r2(i);
  return fe_mul<P>(a, r2);
}

template <class P> HD Fe<P> fe_from_mont(const Fe<P>& a) {
  Fe<P> one = fe_zero<P>();
  one.v[0] = 1;
  return fe_mul<P>(a, one);
}

// a^e for a in Montgomery form, e given as limbs (not secret; square-and-multiply MSB first)
template <class P> HD Fe<P> fe_pow(const Fe<P>& a, const uint32_t* e, int nlimbs) {
  Fe<P> r = fe_one<P>();
  bool started = false;
  for (int i = nlimbs - 1; i >= 0; i--) {
    for (int b = 31; b >= 0; b--) {
      if (started) r = fe_sqr<P>(r);
      if ((e[i] >> b) & 1) {
        r = started ? fe_mul<P>(r, a) : a;
        started = true;
      }
    }
  }
  return r;
}

// a^(p-2): inverse for a != 0 (returns 0 for 0)
template <class P> HD Fe<P> fe_inv(const Fe<P>& a) {
  uint32_t e[P::N];
#pragma unroll
  for (int i = 0; i < P::N; i++) e[i] = P::mod(i);
  uint32_t borrow = 2;                 // e = p - 2 (r_bls has low limb 1: the borrow ripples)
  for (int i = 0; i < P::N && borrow; i++) {
    uint32_t old = e[i];
    e[i] = old - borrow;
    borrow = old < borrow ? 1u : 0u;
  }
  return fe_pow<P>(a, e, P::N);
}
