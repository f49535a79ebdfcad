// G1 point arithmetic for the bucket method: short-Weierstrass y^2 = x^3 + b (a = 0) over
// Fp, BN254 (b=3) and BLS12-381 (b=4).  Replaces py_ecc's projective add/double/multiply
// (bound at reference kzg.py:27-35, called at kzg.py:115-116) with extended-Jacobian
// "XYZZ" coordinates: x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; infinity <=> ZZ == 0.
// Only the normalised affine result is comparable with the reference.
//
// Affine points use (0, 0) for infinity (not on either curve since b != 0); py_ecc's
// Z1 = (1, 1, 0) is mapped to it at the boundary.
#pragma once
#include "field.cuh"

template <class P> struct Affine { Fe<P> x, y; };
template <class P> struct XYZZ { Fe<P> x, y, zz, zzz; };

template <class P> HD bool aff_is_inf(const Affine<P>& a) { return fe_is_zero<P>(a.x) && fe_is_zero<P>(a.y); }
template <class P> HD bool xyzz_is_inf(const XYZZ<P>& a) { return fe_is_zero<P>(a.zz); }

template <class P> HD XYZZ<P> xyzz_inf() {
  XYZZ<P> r;
  r.x = fe_zero<P>(); r.y = fe_zero<P>(); r.zz = fe_zero<P>(); r.zzz = fe_zero<P>();
  return r;
}

template <class P> HD XYZZ<P> xyzz_from_affine(const Affine<P>& a) {
  XYZZ<P> r;
  if (aff_is_inf<P>(a)) return xyzz_inf<P>();
  r.x = a.x; r.y = a.y; r.zz = fe_one<P>(); r.zzz = fe_one<P>();
  return r;
}

template <class P> HD Affine<P> aff_neg(const Affine<P>& a) {
  Affine<P> r; r.x = a.x; r.y = fe_neg<P>(a.y); return r;
}

template <class P> HD XYZZ<P> xyzz_neg(const XYZZ<P>& a) {
  XYZZ<P> r = a; r.y = fe_neg<P>(a.y); return r;
}

// 2 * (affine point), a = 0 curve.  mdbl-2008-s-1: 2M + 4S... (U=2y, V=U^2, W=UV, S=xV, M=3x^2)
template <class P> HD XYZZ<P> xyzz_dbl_affine(const Affine<P>& a) {
  XYZZ<P> r;
  Fe<P> U = fe_dbl<P>(a.y);
  Fe<P> V = fe_sqr<P>(U);
  Fe<P> W = fe_mul<P>(U, V);
  Fe<P> S = fe_mul<P>(a.x, V);
  Fe<P> X2 = fe_sqr<P>(a.x);
  Fe<P> M = fe_add<P>(fe_dbl<P>(X2), X2);
  r.x = fe_sub<P>(fe_sub<P>(fe_sqr<P>(M), S), S);
  r.y = fe_sub<P>(fe_mul<P>(M, fe_sub<P>(S, r.x)), fe_mul<P>(W, a.y));
  r.zz = V; r.zzz = W;
  return r;
}

// 2 * (XYZZ point).  dbl-2008-s-1 with a = 0.  Infinity (ZZ=0) maps to infinity.
template <class P> HD XYZZ<P> xyzz_dbl(const XYZZ<P>& a) {
  XYZZ<P> r;
  Fe<P> U = fe_dbl<P>(a.y);
  Fe<P> V = fe_sqr<P>(U);
  Fe<P> W = fe_mul<P>(U, V);
  Fe<P> S = fe_mul<P>(a.x, V);
  Fe<P> X2 = fe_sqr<P>(a.x);
  Fe<P> M = fe_add<P>(fe_dbl<P>(X2), X2);
  r.x = fe_sub<P>(fe_sub<P>(fe_sqr<P>(M), S), S);
  r.y = fe_sub<P>(fe_mul<P>(M, fe_sub<P>(S, r.x)), fe_mul<P>(W, a.y));
  r.zz = fe_mul<P>(V, a.zz);
  r.zzz = fe_mul<P>(W, a.zzz);
  return r;
}

// acc += b (affine).  madd-2008-s (8M + 2S) with every special case handled:
// b infinity, acc infinity, acc == b (doubling), acc == -b (infinity).
template <class P> HD void xyzz_madd(XYZZ<P>& acc, const Affine<P>& b) {
  if (aff_is_inf<P>(b)) return;
  if (xyzz_is_inf<P>(acc)) { acc = xyzz_from_affine<P>(b); return; }
  Fe<P> U2 = fe_mul<P>(b.x, acc.zz);
  Fe<P> S2 = fe_mul<P>(b.y, acc.zzz);
  Fe<P> Pd = fe_sub<P>(U2, acc.x);
  Fe<P> Rd = fe_sub<P>(S2, acc.y);
  if (fe_is_zero<P>(Pd)) {
    if (fe_is_zero<P>(Rd)) acc = xyzz_dbl_affine<P>(b);
    else acc = xyzz_inf<P>();
    return;
  }
  Fe<P> PP = fe_sqr<P>(Pd);
  Fe<P> PPP = fe_mul<P>(Pd, PP);
  Fe<P> Q = fe_mul<P>(acc.x, PP);
  Fe<P> X3 = fe_sub<P>(fe_sub<P>(fe_sub<P>(fe_sqr<P>(Rd), PPP), Q), Q);
  Fe<P> Y3 = fe_sub<P>(fe_mul<P>(Rd, fe_sub<P>(Q, X3)), fe_mul<P>(acc.y, PPP));
  acc.zz = fe_mul<P>(acc.zz, PP);
  acc.zzz = fe_mul<P>(acc.zzz, PPP);
  acc.x = X3; acc.y = Y3;
}

// acc += b (canonical affine) with the accumulator's coordinates kept semi-reduced in [0, 2p) (field.cuh, FeLz<P>::ok):
// same formulas and special cases as xyzz_madd, every product without its final conditional subtraction.  acc.zz is a
// product of non-zero factors unless it was set to exactly 0 (infinity), so the exact-zero test still identifies infinity;
// the difference Pd can be 0 or p when the x-coordinates agree.  The caller folds the result with xyzz_reduce_lz.
template <class P> HD void xyzz_madd_lz(XYZZ<P>& acc, const Affine<P>& b) {
  if (aff_is_inf<P>(b)) return;
  if (xyzz_is_inf<P>(acc)) { acc = xyzz_from_affine<P>(b); return; }
  Fe<P> U2 = fe_mul_lz<P>(b.x, acc.zz);
  Fe<P> S2 = fe_mul_lz<P>(b.y, acc.zzz);
  Fe<P> Pd = fe_sub_lz<P>(U2, acc.x);
  Fe<P> Rd = fe_sub_lz<P>(S2, acc.y);
  if (fe_is_zero_lz<P>(Pd)) {
    if (fe_is_zero_lz<P>(Rd)) acc = xyzz_dbl_affine<P>(b);
    else acc = xyzz_inf<P>();
    return;
  }
  Fe<P> PP = fe_sqr_lz<P>(Pd);
  Fe<P> PPP = fe_mul_lz<P>(Pd, PP);
  Fe<P> Q = fe_mul_lz<P>(acc.x, PP);
  Fe<P> X3 = fe_sub_lz<P>(fe_sub_lz<P>(fe_sub_lz<P>(fe_sqr_lz<P>(Rd), PPP), Q), Q);
  Fe<P> Y3 = fe_mulsub_lz<P>(Rd, fe_sub_lz<P>(Q, X3), acc.y, PPP);      // Rd (Q - X3) - Y1 PPP, one reduction for both products
  acc.zz = fe_mul_lz<P>(acc.zz, PP);
  acc.zzz = fe_mul_lz<P>(acc.zzz, PPP);
  acc.x = X3; acc.y = Y3;
}

template <class P> HD XYZZ<P> xyzz_reduce_lz(const XYZZ<P>& a) {
  XYZZ<P> r;
  r.x = fe_reduce_lz<P>(a.x); r.y = fe_reduce_lz<P>(a.y); r.zz = fe_reduce_lz<P>(a.zz); r.zzz = fe_reduce_lz<P>(a.zzz);
  return r;
}

// a + b, both XYZZ.  add-2008-s (12M + 2S) with special cases.
template <class P> HD XYZZ<P> xyzz_add(const XYZZ<P>& a, const XYZZ<P>& b) {
  if (xyzz_is_inf<P>(a)) return b;
  if (xyzz_is_inf<P>(b)) return a;
  Fe<P> U1 = fe_mul<P>(a.x, b.zz);
  Fe<P> U2 = fe_mul<P>(b.x, a.zz);
  Fe<P> S1 = fe_mul<P>(a.y, b.zzz);
  Fe<P> S2 = fe_mul<P>(b.y, a.zzz);
  Fe<P> Pd = fe_sub<P>(U2, U1);
  Fe<P> Rd = fe_sub<P>(S2, S1);
  if (fe_is_zero<P>(Pd)) {
    if (fe_is_zero<P>(Rd)) return xyzz_dbl<P>(a);
    return xyzz_inf<P>();
  }
  XYZZ<P> r;
  Fe<P> PP = fe_sqr<P>(Pd);
  Fe<P> PPP = fe_mul<P>(Pd, PP);
  Fe<P> Q = fe_mul<P>(U1, PP);
  r.x = fe_sub<P>(fe_sub<P>(fe_sub<P>(fe_sqr<P>(Rd), PPP), Q), Q);
  r.y = fe_sub<P>(fe_mul<P>(Rd, fe_sub<P>(Q, r.x)), fe_mul<P>(S1, PPP));
  r.zz = fe_mul<P>(fe_mul<P>(a.zz, b.zz), PP);
  r.zzz = fe_mul<P>(fe_mul<P>(a.zzz, b.zzz), PPP);
  return r;
}

// a + b with every coordinate semi-reduced in [0, 2p) (FeLz<P>::ok fields), result semi-reduced: the formulas of xyzz_add with
// products without their final subtraction and Y3 as one dual product.  Used by the bucket reduction, a chain of dependent
// additions where every instruction saved is latency saved.  Canonical inputs are semi-reduced inputs; infinity is ZZ == 0
// exactly (a product of non-zero residues is never a multiple of p); fold the end result with xyzz_reduce_lz.
template <class P> HD XYZZ<P> xyzz_add_lz(const XYZZ<P>& a, const XYZZ<P>& b) {
  if (xyzz_is_inf<P>(a)) return b;
  if (xyzz_is_inf<P>(b)) return a;
  Fe<P> U1 = fe_mul_lz<P>(a.x, b.zz);
  Fe<P> U2 = fe_mul_lz<P>(b.x, a.zz);
  Fe<P> S1 = fe_mul_lz<P>(a.y, b.zzz);
  Fe<P> S2 = fe_mul_lz<P>(b.y, a.zzz);
  Fe<P> Pd = fe_sub_lz<P>(U2, U1);
  Fe<P> Rd = fe_sub_lz<P>(S2, S1);
  if (fe_is_zero_lz<P>(Pd)) {
    if (fe_is_zero_lz<P>(Rd)) return xyzz_dbl<P>(xyzz_reduce_lz<P>(a));
    return xyzz_inf<P>();
  }
  XYZZ<P> r;
  Fe<P> PP = fe_sqr_lz<P>(Pd);
  Fe<P> PPP = fe_mul_lz<P>(Pd, PP);
  Fe<P> Q = fe_mul_lz<P>(U1, PP);
  r.x = fe_sub_lz<P>(fe_sub_lz<P>(fe_sub_lz<P>(fe_sqr_lz<P>(Rd), PPP), Q), Q);
  r.y = fe_mulsub_lz<P>(Rd, fe_sub_lz<P>(Q, r.x), S1, PPP);
  r.zz = fe_mul_lz<P>(fe_mul_lz<P>(a.zz, b.zz), PP);
  r.zzz = fe_mul_lz<P>(fe_mul_lz<P>(a.zzz, b.zzz), PPP);
  return r;
}

// XYZZ (Montgomery) -> affine (Montgomery); infinity -> (0, 0).  One inversion.
template <class P> HD Affine<P> xyzz_to_affine(const XYZZ<P>& a) {
  Affine<P> r;
  if (xyzz_is_inf<P>(a)) { r.x = fe_zero<P>(); r.y = fe_zero<P>(); return r; }
  Fe<P> t = fe_inv<P>(fe_mul<P>(a.zz, a.zzz));     // Z^-5
  Fe<P> izz = fe_mul<P>(t, a.zzz);                 // Z^-2
  Fe<P> izzz = fe_mul<P>(t, a.zz);                 // Z^-3
  r.x = fe_mul<P>(a.x, izz);
  r.y = fe_mul<P>(a.y, izzz);
  return r;
}

// k * pt for a small/any scalar given as limbs (LSB-first double-and-add); used only for
// O(#windows) fix-ups, never per point.
template <class P> HD XYZZ<P> xyzz_mul_u32(const XYZZ<P>& pt, uint32_t k) {
  XYZZ<P> acc = xyzz_inf<P>(), base = pt;
  while (k) {
    if (k & 1) acc = xyzz_add<P>(acc, base);
    k >>= 1;
    if (k) base = xyzz_dbl<P>(base);
  }
  return acc;
}
