// Prover polynomial kernels for the callers of commit / open / fft_ff: the places where
// the reference's PLONK and Marlin provers spend their time once the hot path itself is on the GPU.
//
//   Marlin evaluation loops     marlin/prover.py:248-301 (_compute_t_polynomial), :404-470 (_compute_f2_polynomial)
//        both sum, over the non-zero entries kappa of the three index matrices, eta_M val_M(kappa) divided by
//        (x - row_M(kappa)) (alpha - col_M(kappa)); the reference loops over K with one rational-function division per
//        entry.  Here: numerators / denominators per entry, ONE batched inversion over the 3m denominators, then a sum per
//        kappa (f_2, x = beta_1) or per row of H (t: v_H(X)/(X - h) vanishes on H except at h, where it is n/h).
//
//   permutation grand product   plonk/prover.py:245-258   z(w^0) = 1,
//        z(w^(i+1)) = z(w^i) * num_i / den_i,
//        num_i = (a_i + beta H_i + gamma)(b_i + beta k1 H_i + gamma)(c_i + beta k2 H_i + gamma)
//        den_i = (a_i + beta s*_i + gamma)(b_i + beta s*_(n+i) + gamma)(c_i + beta s*_(2n+i) + gamma)
//     -> ratio kernel, strided Montgomery batch inversion (one Fermat inversion per thread),
//        chunked exclusive prefix product (same two-sweep shape as the synthetic division in
//        poly.cu).  Scratch values are kept in Montgomery form; inputs / outputs are canonical.
//
//   quotient t(X)               plonk/prover.py:297-316   the reference divides four polynomial
//        terms by v_H = X^n - 1 with Sage's long division.  Here every operand is evaluated on
//        the coset s*<w_4n> (coset NTTs, ntt.cu), the numerator is formed point-wise and
//        multiplied by 1/v_H(x) -- v_H takes only 4 values on that coset -- and one inverse
//        coset NTT returns t's coefficients.  deg t = 3n+5 < 4n, so the result is the same
//        polynomial (exact division).
#include "common.cuh"
#include <vector>
#include <cstring>

namespace {

constexpr size_t CH = 64;          // elements per thread in the scan sweeps
constexpr size_t INV_ROUNDS = 64;  // elements per thread in the batch inversion

struct PlonkWs { KzScratch num, den, pre, levels, flag, ptrs; };
PlonkWs g_ws;

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
template <class P> __device__ __forceinline__ void st_fe(uint32_t* p, const Fe<P>& a) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < P::N / 4; i++) q[i] = make_uint4(a.v[4 * i], a.v[4 * i + 1], a.v[4 * i + 2], a.v[4 * i + 3]);
}
// canonical in memory -> Montgomery in registers
template <class P> __device__ __forceinline__ Fe<P> ld_mont(const uint32_t* p) { return fe_to_mont<P>(ld_fe<P>(p)); }

template <class P> struct PermParams { Fe<P> beta, gamma, k1, k2; };   // Montgomery form

// num / den of step i (i < n-1); entry n-1 is padded with 1 so the scan has n inputs
template <class P>
__global__ void perm_ratio_kernel(size_t n, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* sigma,
                                  const uint32_t* H, PermParams<P> pp, uint32_t* num, uint32_t* den) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == n - 1) { st_fe<P>(num + i * P::N, fe_one<P>()); st_fe<P>(den + i * P::N, fe_one<P>()); return; }
  Fe<P> av = fe_add<P>(ld_mont<P>(a + i * P::N), pp.gamma);
  Fe<P> bv = fe_add<P>(ld_mont<P>(b + i * P::N), pp.gamma);
  Fe<P> cv = fe_add<P>(ld_mont<P>(c + i * P::N), pp.gamma);
  Fe<P> bh = fe_mul<P>(ld_mont<P>(H + i * P::N), pp.beta);
  Fe<P> nu = fe_mul<P>(fe_mul<P>(fe_add<P>(av, bh), fe_add<P>(bv, fe_mul<P>(bh, pp.k1))), fe_add<P>(cv, fe_mul<P>(bh, pp.k2)));
  Fe<P> s1 = fe_mul<P>(ld_mont<P>(sigma + i * P::N), pp.beta);
  Fe<P> s2 = fe_mul<P>(ld_mont<P>(sigma + (n + i) * P::N), pp.beta);
  Fe<P> s3 = fe_mul<P>(ld_mont<P>(sigma + (2 * n + i) * P::N), pp.beta);
  Fe<P> de = fe_mul<P>(fe_mul<P>(fe_add<P>(av, s1), fe_add<P>(bv, s2)), fe_add<P>(cv, s3));
  st_fe<P>(num + i * P::N, nu);
  st_fe<P>(den + i * P::N, de);
}

// ratio[i] = num[i] / den[i] in place of num.  Thread t owns elements t, t+T, t+2T, ... (coalesced):
// forward pass stores running products, one inversion, backward pass peels them off.
template <class P>
__global__ void batch_ratio_kernel(size_t n, size_t T, uint32_t* num, const uint32_t* den, uint32_t* pre, int* zero_flag) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  Fe<P> acc = fe_one<P>();
  size_t last = t;
  for (size_t i = t; i < n; i += T) {
    Fe<P> d = ld_fe<P>(den + i * P::N);
    if (fe_is_zero<P>(d)) { *zero_flag = 1; d = fe_one<P>(); }
    st_fe<P>(pre + i * P::N, acc);
    acc = fe_mul<P>(acc, d);
    last = i;
  }
  Fe<P> inv = fe_inv<P>(acc);
  for (size_t i = last;; i -= T) {
    Fe<P> d = ld_fe<P>(den + i * P::N);
    if (fe_is_zero<P>(d)) d = fe_one<P>();
    Fe<P> di = fe_mul<P>(inv, ld_fe<P>(pre + i * P::N));
    inv = fe_mul<P>(inv, d);
    st_fe<P>(num + i * P::N, fe_mul<P>(ld_fe<P>(num + i * P::N), di));
    if (i < T) break;
  }
}

// S[t] = product of chunk t
template <class P>
__global__ void chunk_product_kernel(const uint32_t* in, size_t len, uint32_t* S, size_t nchunks) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nchunks) return;
  size_t lo = t * CH, hi = lo + CH < len ? lo + CH : len;
  Fe<P> acc = ld_fe<P>(in + lo * P::N);
  for (size_t i = lo + 1; i < hi; i++) acc = fe_mul<P>(acc, ld_fe<P>(in + i * P::N));
  st_fe<P>(S + t * P::N, acc);
}

// in place: x[i] <- carry_t * prod_{lo <= j < i} x[j]; carry_t = up[t] (already exclusive) or 1
template <class P>
__global__ void chunk_exclusive_kernel(uint32_t* x, size_t len, const uint32_t* up, size_t nchunks, uint32_t* canonical_out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nchunks) return;
  size_t lo = t * CH, hi = lo + CH < len ? lo + CH : len;
  Fe<P> carry = up ? ld_fe<P>(up + t * P::N) : fe_one<P>();
  for (size_t i = lo; i < hi; i++) {
    Fe<P> v = ld_fe<P>(x + i * P::N);
    if (canonical_out) st_fe<P>(canonical_out + i * P::N, fe_from_mont<P>(carry));
    else st_fe<P>(x + i * P::N, carry);
    carry = fe_mul<P>(carry, v);
  }
}

template <class P>
int permutation_impl(size_t n, const uint32_t* d_a, const uint32_t* d_b, const uint32_t* d_c, const uint32_t* d_sigma, const uint32_t* d_H,
                     const uint64_t* k1, const uint64_t* k2, const uint64_t* beta, const uint64_t* gamma, uint32_t* d_z, int* zero_den) {
  KzgpuCtx& cx = kz_ctx();
  PermParams<P> pp;
  const uint64_t* src[4] = {beta, gamma, k1, k2};
  Fe<P>* dst[4] = {&pp.beta, &pp.gamma, &pp.k1, &pp.k2};
  for (int i = 0; i < 4; i++) {
    Fe<P> v = kz_fe_from_u64<P>(src[i]);
    if (!kz_fe_reduced<P>(v)) return kz_fail(KZGPU_ERANGE, "beta / gamma / k1 / k2 must be canonical field elements");
    *dst[i] = fe_to_mont<P>(v);
  }
  int rc;
  const size_t eb = P::N * 4;
  if ((rc = g_ws.num.ensure(n * eb)) || (rc = g_ws.den.ensure(n * eb)) || (rc = g_ws.pre.ensure(n * eb)) || (rc = g_ws.flag.ensure(64))) return rc;
  uint32_t *num = (uint32_t*)g_ws.num.p, *den = (uint32_t*)g_ws.den.p, *pre = (uint32_t*)g_ws.pre.p;
  int* d_flag = (int*)g_ws.flag.p;
  KZ_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), cx.stream));
  perm_ratio_kernel<P><<<(unsigned)kz_div_up(n, 128), 128, 0, cx.stream>>>(n, d_a, d_b, d_c, d_sigma, d_H, pp, num, den);
  KZ_LAUNCHED();
  size_t T = kz_div_up(n, INV_ROUNDS);
  batch_ratio_kernel<P><<<(unsigned)kz_div_up(T, 128), 128, 0, cx.stream>>>(n, T, num, den, pre, d_flag);
  KZ_LAUNCHED();
  // exclusive prefix product of the n ratios -> z values
  std::vector<size_t> lens{n};
  while (lens.back() > 1) lens.push_back(kz_div_up(lens.back(), CH));
  size_t total = 0;
  for (size_t l = 1; l < lens.size(); l++) total += lens[l];
  if ((rc = g_ws.levels.ensure((total + 1) * eb))) return rc;
  std::vector<uint32_t*> S(lens.size(), nullptr);
  S[0] = num;
  uint32_t* p = (uint32_t*)g_ws.levels.p;
  for (size_t l = 1; l < lens.size(); l++) { S[l] = p; p += lens[l] * P::N; }
  for (size_t l = 1; l < lens.size(); l++) {
    chunk_product_kernel<P><<<(unsigned)kz_div_up(lens[l], 128), 128, 0, cx.stream>>>(S[l - 1], lens[l - 1], S[l], lens[l]);
    KZ_LAUNCHED();
  }
  for (size_t l = lens.size(); l-- > 0;) {
    const uint32_t* up = l + 1 < lens.size() ? S[l + 1] : nullptr;   // top level (one element) has no carry-in
    size_t nch = kz_div_up(lens[l], CH);
    chunk_exclusive_kernel<P><<<(unsigned)kz_div_up(nch, 128), 128, 0, cx.stream>>>(S[l], lens[l], up, nch, l == 0 ? d_z : nullptr);
    KZ_LAUNCHED();
  }
  int flag = 0;
  KZ_CUDA(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, cx.stream));
  KZ_CUDA(cudaStreamSynchronize(cx.stream));
  if (zero_den) *zero_den = flag;
  return 0;
}

// ---------------------------------------------------------------------------- quotient
// operand order of the evaluation-pointer table
enum { Q_A, Q_B, Q_C, Q_Z, Q_QM, Q_QL, Q_QR, Q_QO, Q_QC, Q_S1, Q_S2, Q_S3, Q_PI, Q_L1, Q_X, Q_COUNT };

struct QuotPtrs { const uint32_t* p[Q_COUNT]; };
template <class P> struct QuotParams { Fe<P> alpha, beta, gamma, k1, k2, zh_inv[4]; };   // Montgomery form

// MONT: the operand vectors already hold Montgomery-form values (their coefficient vectors were scaled by R before the
// coset NTTs, which are linear), so the 16 per-point conversions disappear; the result is stored canonical either way.
template <class P, bool MONT>
__global__ void quotient_kernel(size_t n4, QuotPtrs q, QuotParams<P> pp, uint32_t* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const size_t o = i * P::N;
  auto ld_mont = [](const uint32_t* p) { return MONT ? ld_fe<P>(p) : fe_to_mont<P>(ld_fe<P>(p)); };
  Fe<P> a = ld_mont(q.p[Q_A] + o), b = ld_mont(q.p[Q_B] + o), c = ld_mont(q.p[Q_C] + o);
  Fe<P> z = ld_mont(q.p[Q_Z] + o);
  // gate constraint  a b qM + a qL + b qR + c qO + PI + qC           plonk/prover.py:297
  Fe<P> t = fe_mul<P>(fe_mul<P>(a, b), ld_mont(q.p[Q_QM] + o));
  t = fe_add<P>(t, fe_mul<P>(a, ld_mont(q.p[Q_QL] + o)));
  t = fe_add<P>(t, fe_mul<P>(b, ld_mont(q.p[Q_QR] + o)));
  t = fe_add<P>(t, fe_mul<P>(c, ld_mont(q.p[Q_QO] + o)));
  t = fe_add<P>(t, fe_add<P>(ld_mont(q.p[Q_PI] + o), ld_mont(q.p[Q_QC] + o)));
  // permutation numerator  alpha z (a + beta X + gamma)(b + beta k1 X + gamma)(c + beta k2 X + gamma)   :298-300
  Fe<P> ag = fe_add<P>(a, pp.gamma), bg = fe_add<P>(b, pp.gamma), cg = fe_add<P>(c, pp.gamma);
  Fe<P> bx = fe_mul<P>(ld_mont(q.p[Q_X] + o), pp.beta);
  Fe<P> u = fe_mul<P>(fe_mul<P>(fe_add<P>(ag, bx), fe_add<P>(bg, fe_mul<P>(bx, pp.k1))), fe_add<P>(cg, fe_mul<P>(bx, pp.k2)));
  u = fe_mul<P>(u, z);
  // permutation denominator  alpha (a + beta S1 + gamma)(b + beta S2 + gamma)(c + beta S3 + gamma) z(wX)   :301-305
  size_t is = i + 4 < n4 ? i + 4 : i + 4 - n4;            // w = w_4n^4: z(w x_i) = z(x_{i+4})
  Fe<P> zw = ld_mont(q.p[Q_Z] + is * P::N);
  Fe<P> v = fe_mul<P>(fe_mul<P>(fe_add<P>(ag, fe_mul<P>(ld_mont(q.p[Q_S1] + o), pp.beta)),
                                fe_add<P>(bg, fe_mul<P>(ld_mont(q.p[Q_S2] + o), pp.beta))),
                      fe_add<P>(cg, fe_mul<P>(ld_mont(q.p[Q_S3] + o), pp.beta)));
  v = fe_mul<P>(v, zw);
  // alpha^2 (z - 1) L1                                                                           :306-307
  Fe<P> w4 = fe_mul<P>(fe_sub<P>(z, fe_one<P>()), ld_mont(q.p[Q_L1] + o));
  Fe<P> perm = fe_add<P>(fe_sub<P>(u, v), fe_mul<P>(w4, pp.alpha));
  t = fe_add<P>(t, fe_mul<P>(perm, pp.alpha));
  t = fe_mul<P>(t, pp.zh_inv[i & 3]);
  st_fe<P>(out + o, fe_from_mont<P>(t));
}

template <class P>
int quotient_impl(size_t n4, const uint64_t* const* d_evals, const uint64_t* params, int mont_in, uint32_t* d_t) {
  KzgpuCtx& cx = kz_ctx();
  QuotPtrs q;
  for (int i = 0; i < Q_COUNT; i++) {
    if (!d_evals[i]) return kz_fail(KZGPU_EINVAL, "null evaluation vector %d", i);
    q.p[i] = (const uint32_t*)d_evals[i];
  }
  QuotParams<P> pp;
  Fe<P>* dst[9] = {&pp.alpha, &pp.beta, &pp.gamma, &pp.k1, &pp.k2, &pp.zh_inv[0], &pp.zh_inv[1], &pp.zh_inv[2], &pp.zh_inv[3]};
  for (int i = 0; i < 9; i++) {
    Fe<P> v = kz_fe_from_u64<P>(params + 4 * i);
    if (!kz_fe_reduced<P>(v)) return kz_fail(KZGPU_ERANGE, "quotient parameter %d is not a canonical field element", i);
    *dst[i] = fe_to_mont<P>(v);
  }
  if (mont_in) quotient_kernel<P, true><<<(unsigned)kz_div_up(n4, 128), 128, 0, cx.stream>>>(n4, q, pp, d_t);
  else quotient_kernel<P, false><<<(unsigned)kz_div_up(n4, 128), 128, 0, cx.stream>>>(n4, q, pp, d_t);
  KZ_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------- Marlin loops (N4)
template <class P> struct MarlinParams { Fe<P> eta[3], alpha, beta1, scale; };   // Montgomery form

// entry e = M * m + kappa: num = eta_M * val, den = (alpha - col) [* (beta1 - row) when WITH_ROW]; den == 0 -> term skipped
template <class P, bool WITH_ROW>
__global__ void marlin_terms_kernel(size_t m, const uint32_t* row, const uint32_t* col, const uint32_t* val, MarlinParams<P> pp,
                                    uint32_t* num, uint32_t* den) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 3 * m) return;
  const uint32_t M = (uint32_t)(e / m);
  Fe<P> d = fe_sub<P>(pp.alpha, ld_mont<P>(col + e * P::N));
  if (WITH_ROW) d = fe_mul<P>(d, fe_sub<P>(pp.beta1, ld_mont<P>(row + e * P::N)));
  Fe<P> nu = fe_mul<P>(ld_mont<P>(val + e * P::N), pp.eta[M]);
  if (fe_is_zero<P>(d)) { nu = fe_zero<P>(); d = fe_one<P>(); }     // `if denom != 0` of the reference
  st_fe<P>(num + e * P::N, nu);
  st_fe<P>(den + e * P::N, d);
}

// f2(kappa) = scale * sum_M ratio[M * m + kappa]
template <class P>
__global__ void marlin_f2_sum_kernel(size_t m, const uint32_t* ratio, Fe<P> scale, uint32_t* out) {
  size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  Fe<P> s = fe_add<P>(fe_add<P>(ld_fe<P>(ratio + k * P::N), ld_fe<P>(ratio + (m + k) * P::N)), ld_fe<P>(ratio + (2 * m + k) * P::N));
  st_fe<P>(out + k * P::N, fe_from_mont<P>(fe_mul<P>(s, scale)));
}

// t(h_i) = scale * h_i^-1 * sum over the entries whose row is h_i (row_index sorted ascending per matrix, 0xffffffff = padding)
template <class P>
__global__ void marlin_t_rows_kernel(size_t n, size_t m, const uint32_t* row_index, const uint32_t* ratio, const uint32_t* H, Fe<P> scale,
                                     uint32_t* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fe<P> s = fe_zero<P>();
  for (uint32_t M = 0; M < 3; M++) {
    const uint32_t* ri = row_index + (size_t)M * m;
    size_t lo = 0, hi = m;                       // first kappa with ri[kappa] >= i
    while (lo < hi) { size_t mid = (lo + hi) >> 1; if (ri[mid] < (uint32_t)i) lo = mid + 1; else hi = mid; }
    for (size_t k = lo; k < m && ri[k] == (uint32_t)i; k++) s = fe_add<P>(s, ld_fe<P>(ratio + ((size_t)M * m + k) * P::N));
  }
  Fe<P> hinv = ld_fe<P>(H + ((n - i) & (n - 1)) * P::N);            // h_i^-1 = h_(n-i), canonical: the product comes out canonical
  st_fe<P>(out + i * P::N, fe_mul<P>(fe_mul<P>(s, scale), hinv));
}

// out[i] = a[i] * b[i] (canonical in, canonical out): the point-wise step of an NTT-based polynomial product
template <class P>
__global__ void pointwise_mul_kernel(size_t n, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  st_fe<P>(out + i * P::N, fe_mul<P>(ld_fe<P>(a + i * P::N), ld_mont<P>(b + i * P::N)));
}

// out[i] = sum_k val[k] * z[col[k]] over row i of a CSR matrix (z_M = M z, marlin/encoder.py:205-207): one thread per row
template <class P>
__global__ void spmv_kernel(size_t n_rows, const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col,
                            const uint32_t* __restrict__ val, const uint32_t* __restrict__ z, uint32_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  Fe<P> acc = fe_zero<P>();
  for (uint32_t k = row_ptr[i]; k < row_ptr[i + 1]; k++)
    acc = fe_add<P>(acc, fe_mul<P>(ld_fe<P>(val + (size_t)k * P::N), ld_mont<P>(z + (size_t)col[k] * P::N)));
  st_fe<P>(out + i * P::N, acc);
}

// Marlin third round (marlin/prover.py:166-171, 303-353): on the coset {s w_8m^i} form
//   b = prod_M (beta1 - row_M)(alpha - col_M),  a = sum_M eta_M vv val_M prod_{O != M} (beta1 - row_O)(alpha - col_O),
//   out = (a - b f_2) / v_K   (v_K takes 8 values on the coset),  whose inverse coset NTT is h_2.
template <class P> struct H2Params { Fe<P> eta[3], alpha, beta1, vv, vk_inv[8]; };
template <class P>
__global__ void marlin_h2_kernel(size_t m8, const uint32_t* row, const uint32_t* col, const uint32_t* val, const uint32_t* f2,
                                 H2Params<P> pp, uint32_t* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m8) return;
  Fe<P> fac[3], v[3];
#pragma unroll
  for (int M = 0; M < 3; M++) {
    fac[M] = fe_mul<P>(fe_sub<P>(pp.beta1, ld_mont<P>(row + ((size_t)M * m8 + i) * P::N)),
                       fe_sub<P>(pp.alpha, ld_mont<P>(col + ((size_t)M * m8 + i) * P::N)));
    v[M] = fe_mul<P>(fe_mul<P>(ld_mont<P>(val + ((size_t)M * m8 + i) * P::N), pp.eta[M]), pp.vv);
  }
  Fe<P> a = fe_add<P>(fe_add<P>(fe_mul<P>(v[0], fe_mul<P>(fac[1], fac[2])), fe_mul<P>(v[1], fe_mul<P>(fac[0], fac[2]))),
                      fe_mul<P>(v[2], fe_mul<P>(fac[0], fac[1])));
  Fe<P> b = fe_mul<P>(fe_mul<P>(fac[0], fac[1]), fac[2]);
  Fe<P> t = fe_mul<P>(fe_sub<P>(a, fe_mul<P>(b, ld_mont<P>(f2 + i * P::N))), pp.vk_inv[i & 7]);
  st_fe<P>(out + i * P::N, fe_from_mont<P>(t));
}

template <class P>
int h2_impl(size_t m8, const uint32_t* row, const uint32_t* col, const uint32_t* val, const uint32_t* f2, const uint64_t* params, uint32_t* out) {
  KzgpuCtx& cx = kz_ctx();
  H2Params<P> pp;
  Fe<P>* dst[14] = {&pp.eta[0], &pp.eta[1], &pp.eta[2], &pp.alpha, &pp.beta1, &pp.vv, &pp.vk_inv[0], &pp.vk_inv[1], &pp.vk_inv[2],
                    &pp.vk_inv[3], &pp.vk_inv[4], &pp.vk_inv[5], &pp.vk_inv[6], &pp.vk_inv[7]};
  for (int i = 0; i < 14; i++) {
    Fe<P> v = kz_fe_from_u64<P>(params + 4 * i);
    if (!kz_fe_reduced<P>(v)) return kz_fail(KZGPU_ERANGE, "h_2 parameter %d is not a canonical field element", i);
    *dst[i] = fe_to_mont<P>(v);
  }
  marlin_h2_kernel<P><<<(unsigned)kz_div_up(m8, 128), 128, 0, cx.stream>>>(m8, row, col, val, f2, pp, out);
  KZ_LAUNCHED();
  return 0;
}

template <class P>
int marlin_impl(bool is_t, size_t n, size_t m, const uint32_t* d_row_or_index, const uint32_t* d_col, const uint32_t* d_val,
                const uint32_t* d_H, const uint64_t* eta, const uint64_t* alpha, const uint64_t* beta1, const uint64_t* scale,
                uint32_t* d_out) {
  KzgpuCtx& cx = kz_ctx();
  MarlinParams<P> pp;
  const uint64_t* src[6] = {eta, eta + 4, eta + 8, alpha, beta1 ? beta1 : alpha, scale};
  Fe<P>* dst[6] = {&pp.eta[0], &pp.eta[1], &pp.eta[2], &pp.alpha, &pp.beta1, &pp.scale};
  for (int i = 0; i < 6; i++) {
    Fe<P> v = kz_fe_from_u64<P>(src[i]);
    if (!kz_fe_reduced<P>(v)) return kz_fail(KZGPU_ERANGE, "eta / alpha / beta_1 / scale must be canonical field elements");
    *dst[i] = fe_to_mont<P>(v);
  }
  int rc;
  const size_t eb = P::N * 4, cnt = 3 * m;
  if ((rc = g_ws.num.ensure(cnt * eb)) || (rc = g_ws.den.ensure(cnt * eb)) || (rc = g_ws.pre.ensure(cnt * eb)) || (rc = g_ws.flag.ensure(64))) return rc;
  uint32_t *num = (uint32_t*)g_ws.num.p, *den = (uint32_t*)g_ws.den.p, *pre = (uint32_t*)g_ws.pre.p;
  if (is_t) marlin_terms_kernel<P, false><<<(unsigned)kz_div_up(cnt, 128), 128, 0, cx.stream>>>(m, nullptr, d_col, d_val, pp, num, den);
  else marlin_terms_kernel<P, true><<<(unsigned)kz_div_up(cnt, 128), 128, 0, cx.stream>>>(m, d_row_or_index, d_col, d_val, pp, num, den);
  KZ_LAUNCHED();
  size_t T = kz_div_up(cnt, INV_ROUNDS);
  batch_ratio_kernel<P><<<(unsigned)kz_div_up(T, 128), 128, 0, cx.stream>>>(cnt, T, num, den, pre, (int*)g_ws.flag.p);
  KZ_LAUNCHED();
  if (is_t) marlin_t_rows_kernel<P><<<(unsigned)kz_div_up(n, 128), 128, 0, cx.stream>>>(n, m, d_row_or_index, num, d_H, pp.scale, d_out);
  else marlin_f2_sum_kernel<P><<<(unsigned)kz_div_up(m, 128), 128, 0, cx.stream>>>(m, num, pp.scale, d_out);
  KZ_LAUNCHED();
  return 0;
}

}  // namespace

void kz_plonk_release() {
  if (kz_slot() != 0) return;          // polynomial / prover kernels run on the primary device only
  KzScratch* all[] = {&g_ws.num, &g_ws.den, &g_ws.pre, &g_ws.levels, &g_ws.flag, &g_ws.ptrs};
  for (auto* s : all) s->release();
}

extern "C" {

int kzgpu_plonk_permutation_dev(int field, size_t n, const uint64_t* d_a, const uint64_t* d_b, const uint64_t* d_c,
                                const uint64_t* d_sigma_star, const uint64_t* d_H, const uint64_t* k1, const uint64_t* k2,
                                const uint64_t* beta, const uint64_t* gamma, uint64_t* d_z, int* zero_den) {
  KZ_REQUIRE_INIT();
  if (!d_a || !d_b || !d_c || !d_sigma_star || !d_H || !k1 || !k2 || !beta || !gamma || !d_z) return kz_fail(KZGPU_EINVAL, "null pointer");
  if (n < 2 || (n & (n - 1))) return kz_fail(KZGPU_EINVAL, "n must be a power of two >= 2");
  if (field == KZGPU_BN254)
    return permutation_impl<FrBN254>(n, (const uint32_t*)d_a, (const uint32_t*)d_b, (const uint32_t*)d_c, (const uint32_t*)d_sigma_star,
                                     (const uint32_t*)d_H, k1, k2, beta, gamma, (uint32_t*)d_z, zero_den);
  if (field == KZGPU_BLS12_381)
    return permutation_impl<FrBLS381>(n, (const uint32_t*)d_a, (const uint32_t*)d_b, (const uint32_t*)d_c, (const uint32_t*)d_sigma_star,
                                      (const uint32_t*)d_H, k1, k2, beta, gamma, (uint32_t*)d_z, zero_den);
  return kz_fail(KZGPU_EINVAL, "unknown field id %d", field);
}

int kzgpu_plonk_quotient_dev(int field, size_t n4, const uint64_t* const* d_evals, const uint64_t* params, int mont_in,
                             uint64_t* d_t) {
  KZ_REQUIRE_INIT();
  if (!d_evals || !params || !d_t) return kz_fail(KZGPU_EINVAL, "null pointer");
  if (n4 < 8 || (n4 & (n4 - 1))) return kz_fail(KZGPU_EINVAL, "the coset size must be a power of two >= 8");
  if (field == KZGPU_BN254) return quotient_impl<FrBN254>(n4, d_evals, params, mont_in, (uint32_t*)d_t);
  if (field == KZGPU_BLS12_381) return quotient_impl<FrBLS381>(n4, d_evals, params, mont_in, (uint32_t*)d_t);
  return kz_fail(KZGPU_EINVAL, "unknown field id %d", field);
}

int kzgpu_poly_mul_pointwise_dev(int field, uint64_t* d_out, const uint64_t* d_a, const uint64_t* d_b, size_t n) {
  KZ_REQUIRE_INIT();
  if (n && (!d_out || !d_a || !d_b)) return kz_fail(KZGPU_EINVAL, "null pointer");
  if (!n) return 0;
  KzgpuCtx& cx = kz_ctx();
  if (field == KZGPU_BN254)
    pointwise_mul_kernel<FrBN254><<<(unsigned)kz_div_up(n, 128), 128, 0, cx.stream>>>(n, (const uint32_t*)d_a, (const uint32_t*)d_b, (uint32_t*)d_out);
  else if (field == KZGPU_BLS12_381)
    pointwise_mul_kernel<FrBLS381><<<(unsigned)kz_div_up(n, 128), 128, 0, cx.stream>>>(n, (const uint32_t*)d_a, (const uint32_t*)d_b, (uint32_t*)d_out);
  else return kz_fail(KZGPU_EINVAL, "unknown field id %d", field);
  KZ_LAUNCHED();
  return 0;
}

int kzgpu_spmv_dev(int field, size_t n_rows, const uint32_t* d_row_ptr, const uint32_t* d_col, const uint64_t* d_val,
                   const uint64_t* d_z, uint64_t* d_out) {
  KZ_REQUIRE_INIT();
  if (n_rows && (!d_row_ptr || !d_col || !d_val || !d_z || !d_out)) return kz_fail(KZGPU_EINVAL, "null pointer");
  if (!n_rows) return 0;
  KzgpuCtx& cx = kz_ctx();
  if (field == KZGPU_BN254)
    spmv_kernel<FrBN254><<<(unsigned)kz_div_up(n_rows, 128), 128, 0, cx.stream>>>(n_rows, d_row_ptr, d_col, (const uint32_t*)d_val, (const uint32_t*)d_z, (uint32_t*)d_out);
  else if (field == KZGPU_BLS12_381)
    spmv_kernel<FrBLS381><<<(unsigned)kz_div_up(n_rows, 128), 128, 0, cx.stream>>>(n_rows, d_row_ptr, d_col, (const uint32_t*)d_val, (const uint32_t*)d_z, (uint32_t*)d_out);
  else return kz_fail(KZGPU_EINVAL, "unknown field id %d", field);
  KZ_LAUNCHED();
  return 0;
}

int kzgpu_marlin_h2_evals_dev(int field, size_t m8, const uint64_t* d_row, const uint64_t* d_col, const uint64_t* d_val,
                              const uint64_t* d_f2, const uint64_t* params, uint64_t* d_out) {
  KZ_REQUIRE_INIT();
  if (!d_row || !d_col || !d_val || !d_f2 || !params || !d_out) return kz_fail(KZGPU_EINVAL, "null pointer");
  if (m8 < 8 || (m8 & (m8 - 1))) return kz_fail(KZGPU_EINVAL, "the coset size must be a power of two >= 8");
  if (field == KZGPU_BN254)
    return h2_impl<FrBN254>(m8, (const uint32_t*)d_row, (const uint32_t*)d_col, (const uint32_t*)d_val, (const uint32_t*)d_f2, params, (uint32_t*)d_out);
  if (field == KZGPU_BLS12_381)
    return h2_impl<FrBLS381>(m8, (const uint32_t*)d_row, (const uint32_t*)d_col, (const uint32_t*)d_val, (const uint32_t*)d_f2, params, (uint32_t*)d_out);
  return kz_fail(KZGPU_EINVAL, "unknown field id %d", field);
}

int kzgpu_marlin_f2_evals_dev(int field, size_t m, const uint64_t* d_row, const uint64_t* d_col, const uint64_t* d_val,
                              const uint64_t* eta, const uint64_t* alpha, const uint64_t* beta1, const uint64_t* scale, uint64_t* d_out) {
  KZ_REQUIRE_INIT();
  if (!d_row || !d_col || !d_val || !eta || !alpha || !beta1 || !scale || !d_out) return kz_fail(KZGPU_EINVAL, "null pointer");
  if (m == 0) return 0;
  if (field == KZGPU_BN254)
    return marlin_impl<FrBN254>(false, 0, m, (const uint32_t*)d_row, (const uint32_t*)d_col, (const uint32_t*)d_val, nullptr, eta, alpha, beta1, scale, (uint32_t*)d_out);
  if (field == KZGPU_BLS12_381)
    return marlin_impl<FrBLS381>(false, 0, m, (const uint32_t*)d_row, (const uint32_t*)d_col, (const uint32_t*)d_val, nullptr, eta, alpha, beta1, scale, (uint32_t*)d_out);
  return kz_fail(KZGPU_EINVAL, "unknown field id %d", field);
}

int kzgpu_marlin_t_evals_dev(int field, size_t n, size_t m, const uint32_t* d_row_index, const uint64_t* d_col, const uint64_t* d_val,
                             const uint64_t* d_H, const uint64_t* eta, const uint64_t* alpha, const uint64_t* scale, uint64_t* d_out) {
  KZ_REQUIRE_INIT();
  if (!d_row_index || !d_col || !d_val || !d_H || !eta || !alpha || !scale || !d_out) return kz_fail(KZGPU_EINVAL, "null pointer");
  if (n == 0 || (n & (n - 1))) return kz_fail(KZGPU_EINVAL, "|H| must be a power of two");
  if (field == KZGPU_BN254)
    return marlin_impl<FrBN254>(true, n, m, d_row_index, (const uint32_t*)d_col, (const uint32_t*)d_val, (const uint32_t*)d_H, eta, alpha, nullptr, scale, (uint32_t*)d_out);
  if (field == KZGPU_BLS12_381)
    return marlin_impl<FrBLS381>(true, n, m, d_row_index, (const uint32_t*)d_col, (const uint32_t*)d_val, (const uint32_t*)d_H, eta, alpha, nullptr, scale, (uint32_t*)d_out);
  return kz_fail(KZGPU_EINVAL, "unknown field id %d", field);
}

}  // extern "C"
