// Host-side check of the Montgomery / XYZZ formulas in kzg_snark_b200/csrc/{field,curve}.cuh
// (compiled with g++; the device build swaps in the PTX primitives with the same contract).
// Protocol on stdin, one op per line, hex operands (canonical, non-Montgomery):
//   <field> mul a b | add a b | sub a b | inv a
//   <curve> madd x1 y1 x2 y2 | dbl x y | smul x y k     (affine in, affine out; inf = 0 0)
// fields: fp_bn fr_bn fp_bls fr_bls ; curves: bn bls
#include <cstdio>
#include <cstring>
#include <string>
#include <iostream>
#include <sstream>
#include "../../kzg_snark_b200/csrc/params_gen.cuh"
#include "../../kzg_snark_b200/csrc/curve.cuh"

template <class P> Fe<P> parse(const std::string& h) {
  Fe<P> r = fe_zero<P>();
  int n = (int)h.size();
  for (int i = 0; i < n; i++) {
    char c = h[n - 1 - i];
    uint32_t d = (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : c - 'A' + 10;
    if (i / 8 < P::N) r.v[i / 8] |= d << (4 * (i % 8));
  }
  return r;
}
template <class P> std::string hex(const Fe<P>& a) {
  char buf[16]; std::string s;
  for (int i = P::N - 1; i >= 0; i--) { snprintf(buf, sizeof buf, "%08x", a.v[i]); s += buf; }
  size_t p = s.find_first_not_of('0');
  return p == std::string::npos ? "0" : s.substr(p);
}
template <class P> Fe<P> M(const std::string& h) { return fe_to_mont<P>(parse<P>(h)); }
template <class P> std::string U(const Fe<P>& a) { return hex<P>(fe_from_mont<P>(a)); }

template <class P> void field_op(const std::string& op, std::istringstream& in) {
  std::string a, b; in >> a;
  if (op == "inv") { std::cout << U<P>(fe_inv<P>(M<P>(a))) << "\n"; return; }
  in >> b;
  if (op == "lzmulsub") {        // a*b - c*d on raw semi-reduced limbs: dual product with one reduction (field.cuh fe_mulsub_lz)
    std::string c, d; in >> c >> d;
    Fe<P> w = fe_mulsub_lz<P>(parse<P>(a), parse<P>(b), parse<P>(c), parse<P>(d));
    std::cout << hex<P>(w) << " " << (fe_is_zero_lz<P>(w) ? 1 : 0) << " " << hex<P>(fe_reduce_lz<P>(w)) << "\n";
    return;
  }
  Fe<P> x = M<P>(a), y = M<P>(b), r;
  if (op == "mul") r = fe_mul<P>(x, y);
  else if (op == "sqr") r = fe_sqr<P>(x);
  else if (op == "add") r = fe_add<P>(x, y);
  else if (op == "sub") r = fe_sub<P>(x, y);
  else if (op == "rawmul") { std::cout << hex<P>(fe_mul<P>(parse<P>(a), parse<P>(b))) << "\n"; return; }
  else if (op == "lzmul" || op == "lzadd" || op == "lzsub" || op == "nradd" || op == "nrsub" || op == "lzsqr") {
    // raw limbs in, raw limbs out (no Montgomery conversion): the semi-reduced primitives of field.cuh
    Fe<P> u = parse<P>(a), v = parse<P>(b), w;
    if (op == "lzmul") w = fe_mul_lz<P>(u, v);
    else if (op == "lzsqr") w = fe_sqr_lz<P>(u);
    else if (op == "lzadd") w = fe_add_lz<P>(u, v);
    else if (op == "lzsub") w = fe_sub_lz<P>(u, v);
    else if (op == "nradd") w = fe_add_nr<P>(u, v);
    else w = fe_sub_nr<P>(u, v);
    std::cout << hex<P>(w) << " " << (fe_is_zero_lz<P>(w) ? 1 : 0) << " " << hex<P>(fe_reduce_lz<P>(w)) << "\n";
    return;
  }
  std::cout << U<P>(r) << "\n";
}
template <class P> void curve_op(const std::string& op, std::istringstream& in) {
  std::string x1, y1, x2, y2;
  in >> x1 >> y1;
  Affine<P> a; a.x = M<P>(x1); a.y = M<P>(y1);
  XYZZ<P> r;
  if (op == "madd") {
    in >> x2 >> y2;
    Affine<P> b; b.x = M<P>(x2); b.y = M<P>(y2);
    r = xyzz_from_affine<P>(a);
    // push acc off Z=1 so the general formulas are exercised: acc = 2a - a when a finite
    xyzz_madd<P>(r, b);
  } else if (op == "lzchain") {   // ((2a + b) + b) - b ... with the semi-reduced mixed addition, folded at the end
    in >> x2 >> y2;
    Affine<P> b; b.x = M<P>(x2); b.y = M<P>(y2);
    r = xyzz_dbl_affine<P>(a);
    xyzz_madd_lz<P>(r, b);
    xyzz_madd_lz<P>(r, b);
    xyzz_madd_lz<P>(r, aff_neg<P>(a));
    xyzz_madd_lz<P>(r, aff_neg<P>(b));
    r = xyzz_reduce_lz<P>(r);       // = a + b
  } else if (op == "lzmadd") {
    in >> x2 >> y2;
    Affine<P> b; b.x = M<P>(x2); b.y = M<P>(y2);
    r = xyzz_from_affine<P>(a);
    xyzz_madd_lz<P>(r, b);
    r = xyzz_reduce_lz<P>(r);
  } else if (op == "add3") {       // (a + b) + b via full add of two non-trivial-Z points
    in >> x2 >> y2;
    Affine<P> b; b.x = M<P>(x2); b.y = M<P>(y2);
    XYZZ<P> t = xyzz_dbl_affine<P>(a);   // 2a
    XYZZ<P> u = xyzz_dbl_affine<P>(b);   // 2b
    r = xyzz_add<P>(t, u);               // 2a + 2b
    xyzz_madd<P>(r, aff_neg<P>(a));      // a + 2b
    xyzz_madd<P>(r, aff_neg<P>(b));      // a + b
  } else if (op == "lzadd3") {     // the same through the semi-reduced full addition, incl. the doubling and cancellation branches
    in >> x2 >> y2;
    Affine<P> b; b.x = M<P>(x2); b.y = M<P>(y2);
    XYZZ<P> t = xyzz_dbl_affine<P>(a);   // 2a
    XYZZ<P> u = xyzz_dbl_affine<P>(b);   // 2b
    r = xyzz_add_lz<P>(t, u);            // 2a + 2b   (a == b: doubling branch; a == -b: infinity)
    r = xyzz_add_lz<P>(r, xyzz_from_affine<P>(aff_neg<P>(a)));   // a + 2b
    r = xyzz_add_lz<P>(r, xyzz_add_lz<P>(xyzz_from_affine<P>(aff_neg<P>(b)), xyzz_inf<P>()));   // a + b
    r = xyzz_add_lz<P>(xyzz_inf<P>(), r);
    r = xyzz_reduce_lz<P>(r);
  } else if (op == "dbl") {
    r = xyzz_dbl<P>(xyzz_dbl_affine<P>(a));   // 4a
  } else if (op == "smul") {
    uint32_t k; in >> k;
    r = xyzz_mul_u32<P>(xyzz_from_affine<P>(a), k);
  }
  Affine<P> o = xyzz_to_affine<P>(r);
  std::cout << U<P>(o.x) << " " << U<P>(o.y) << "\n";
}
int main() {
  std::string line;
  while (std::getline(std::cin, line)) {
    std::istringstream in(line);
    std::string f, op; in >> f >> op;
    if (f == "fp_bn") field_op<FpBN254>(op, in);
    else if (f == "fr_bn") field_op<FrBN254>(op, in);
    else if (f == "fp_bls") field_op<FpBLS381>(op, in);
    else if (f == "fr_bls") field_op<FrBLS381>(op, in);
    else if (f == "bn") curve_op<FpBN254>(op, in);
    else if (f == "bls") curve_op<FpBLS381>(op, in);
  }
  return 0;
}
