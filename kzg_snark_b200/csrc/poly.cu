// KZG.open on the device (reference kzg.py:122-159):
//   P(X) = sum_j xi^(j+1) * p_j(X)            kzg.py:147-150   (combine kernel)
//   W(X) = (P(X) - P(z)) // (X - z)           kzg.py:153-154   (parallel synthetic division)
//   proof = commit(ck, [W])[0]                kzg.py:157       (MSM, msm.cu)
//
// Synthetic division is the suffix Horner recurrence T_i = c_i + z * T_{i+1} (T_{d+1} = 0):
// quotient coefficient q_{i-1} = T_i for i >= 1 and P(z) = T_0.  It is parallelised as a
// recursion on chunk sums: level l reduces chunks of 64 entries to S^(l+1) with z^(64^(l+1)),
// the (tiny) top level is solved serially, and a second sweep re-runs each chunk from its
// now-known carry-in.  2 modular multiplications per coefficient in total.
// Coefficients stay canonical; z and the xi powers are in Montgomery form.
#include "common.cuh"
#include <vector>
#include <cstring>

int kz_msm_dev_internal(uint64_t handle, size_t first, const uint32_t* d_scalars, size_t n, uint64_t* out_xy, int* is_inf,
                        const uint64_t* h_scalars);
int kz_srs_curve(uint64_t handle);

namespace {

constexpr size_t CHL = 64;

struct PolyWs { KzScratch polys, meta, comb, levels, t; };
PolyWs g_pw;

template <class P> __device__ __forceinline__ Fe<P> ldc_fe(const uint32_t* p) {
  Fe<P> r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < P::N / 4; i++) {
    uint4 t = q[i];
    r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w;
  }
  return r;
}
template <class P> __device__ __forceinline__ void stc_fe(uint32_t* p, const Fe<P>& a) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < P::N / 4; i++) q[i] = make_uint4(a.v[4 * i], a.v[4 * i + 1], a.v[4 * i + 2], a.v[4 * i + 3]);
}

// out[i] = addend*[i==0] + sum_j mult[j] * poly_j[i]; meta = {device address of poly_j, len_j} pairs;
// mult in Montgomery form, coefficients and addend canonical
template <class P>
__global__ void poly_combine_kernel(const uint64_t* meta, const uint32_t* mult, uint32_t k, size_t outlen, const uint32_t* addend,
                                    uint32_t* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= outlen) return;
  Fe<P> acc = (addend && i == 0) ? ldc_fe<P>(addend) : fe_zero<P>();
  for (uint32_t j = 0; j < k; j++) {
    const uint32_t* base = reinterpret_cast<const uint32_t*>(meta[2 * j]);
    uint64_t len = meta[2 * j + 1];
    if (i < len) acc = fe_add<P>(acc, fe_mul<P>(ldc_fe<P>(base + i * P::N), ldc_fe<P>(mult + (size_t)j * P::N)));
  }
  stc_fe<P>(out + i * P::N, acc);
}

// out[i] = scale * base^i: thread t owns PCH consecutive exponents, seeded by square-and-multiply
constexpr size_t PCH = 16;
template <class P>
__global__ void powers_kernel(uint32_t* out, size_t n, Fe<P> base_m, Fe<P> scale) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t lo = t * PCH;
  if (lo >= n) return;
  Fe<P> acc = scale, b = base_m;            // scale canonical, base Montgomery -> products stay canonical
  for (size_t e = lo; e; e >>= 1) {
    if (e & 1) acc = fe_mul<P>(acc, b);
    b = fe_sqr<P>(b);
  }
  size_t hi = lo + PCH < n ? lo + PCH : n;
  for (size_t i = lo; i < hi; i++) {
    stc_fe<P>(out + i * P::N, acc);
    acc = fe_mul<P>(acc, base_m);
  }
}

// S[t] = sum_{i in chunk t} in[i] * z^(i - 64 t)
template <class P>
__global__ void horner_chunks_kernel(const uint32_t* in, size_t len, Fe<P> z, uint32_t* S, size_t nchunks) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nchunks) return;
  size_t lo = t * CHL, hi = lo + CHL < len ? lo + CHL : len;
  Fe<P> acc = fe_zero<P>();
  for (size_t i = hi; i-- > lo;) acc = fe_add<P>(fe_mul<P>(acc, z), ldc_fe<P>(in + i * P::N));
  stc_fe<P>(S + t * P::N, acc);
}

// out[i] = in[i] + z * out[i+1] inside chunk t, seeded with carry = Tup[t+1] (0 for the last chunk)
template <class P>
__global__ void suffix_fill_kernel(const uint32_t* in, size_t len, Fe<P> z, const uint32_t* Tup, size_t nchunks, uint32_t* out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nchunks) return;
  size_t lo = t * CHL, hi = lo + CHL < len ? lo + CHL : len;
  Fe<P> carry = (Tup && t + 1 < nchunks) ? ldc_fe<P>(Tup + (t + 1) * P::N) : fe_zero<P>();
  for (size_t i = hi; i-- > lo;) {
    carry = fe_add<P>(fe_mul<P>(carry, z), ldc_fe<P>(in + i * P::N));
    stc_fe<P>(out + i * P::N, carry);
  }
}

template <class P> Fe<P> hpow(Fe<P> b, uint64_t e) {
  Fe<P> r = fe_one<P>();
  while (e) {
    if (e & 1) r = fe_mul<P>(r, b);
    e >>= 1;
    if (e) b = fe_sqr<P>(b);
  }
  return r;
}

// d_T receives T_0 .. T_{len-1}; d_c is the combined polynomial (len coefficients)
template <class P>
int suffix_horner(const uint32_t* d_c, size_t len, const Fe<P>& z_mont, uint32_t* d_T) {
  KzgpuCtx& cx = kz_ctx();
  // level sizes
  std::vector<size_t> lens{len};
  while (lens.back() > CHL) lens.push_back(kz_div_up(lens.back(), CHL));
  size_t total = 0;
  for (size_t l = 1; l < lens.size(); l++) total += 2 * lens[l];     // S and T per upper level
  int rc = g_pw.levels.ensure((total + 1) * P::N * 4);
  if (rc) return rc;
  std::vector<uint32_t*> S(lens.size(), nullptr), T(lens.size(), nullptr);
  uint32_t* p = (uint32_t*)g_pw.levels.p;
  for (size_t l = 1; l < lens.size(); l++) { S[l] = p; p += lens[l] * P::N; T[l] = p; p += lens[l] * P::N; }
  std::vector<Fe<P>> zl(lens.size());
  zl[0] = z_mont;
  for (size_t l = 1; l < lens.size(); l++) zl[l] = hpow<P>(zl[l - 1], CHL);
  // upward sweep: chunk sums
  const uint32_t* cur = d_c;
  for (size_t l = 1; l < lens.size(); l++) {
    horner_chunks_kernel<P><<<(unsigned)kz_div_up(lens[l], 128), 128, 0, cx.stream>>>(cur, lens[l - 1], zl[l - 1], S[l], lens[l]);
    KZ_LAUNCHED();
    cur = S[l];
  }
  // downward sweep
  for (size_t l = lens.size(); l-- > 0;) {
    const uint32_t* in = l == 0 ? d_c : S[l];
    uint32_t* out = l == 0 ? d_T : T[l];
    const uint32_t* up = l + 1 < lens.size() ? T[l + 1] : nullptr;
    size_t nch = kz_div_up(lens[l], CHL);
    suffix_fill_kernel<P><<<(unsigned)kz_div_up(nch, 128), 128, 0, cx.stream>>>(in, lens[l], zl[l], up, nch, out);
    KZ_LAUNCHED();
  }
  return 0;
}

template <class P>
int open_impl(uint64_t handle, bool do_msm, const uint64_t* polys, const uint64_t* const* d_polys, const size_t* lens, size_t k,
              const uint64_t* z, const uint64_t* xi, uint64_t* out_xy, int* is_inf, uint64_t* quotient, size_t* quot_len,
              uint64_t* eval_out) {
  KzgpuCtx& cx = kz_ctx();
  Fe<P> zc = kz_fe_from_u64<P>(z), xc = kz_fe_from_u64<P>(xi);
  if (!kz_fe_reduced<P>(zc) || !kz_fe_reduced<P>(xc)) return kz_fail(KZGPU_ERANGE, "z / xi must be canonical field elements");
  size_t total = 0, maxlen = 0;
  std::vector<uint64_t> meta(2 * k + 2);
  for (size_t j = 0; j < k; j++) {
    meta[2 * j] = total; meta[2 * j + 1] = lens[j];
    total += lens[j];
    if (lens[j] > maxlen) maxlen = lens[j];
  }
  if (maxlen == 0) {           // all-zero combination: witness is the zero polynomial -> Z1 (kzg.py:109)
    if (do_msm) {
      int rc = kz_msm_dev_internal(handle, 0, nullptr, 0, out_xy, is_inf, nullptr);
      if (rc) return rc;
    }
    if (quot_len) *quot_len = 0;
    if (eval_out) memset(eval_out, 0, 32);
    return 0;
  }
  // xi^(j+1), Montgomery form (kzg.py:149: exponent starts at 1)
  std::vector<uint32_t> xip(k * P::N);
  Fe<P> xm = fe_to_mont<P>(xc), cur = xm;
  for (size_t j = 0; j < k; j++) { memcpy(&xip[j * P::N], cur.v, P::N * 4); cur = fe_mul<P>(cur, xm); }
  int rc;
  if (!d_polys && (rc = g_pw.polys.ensure(total * 32))) return rc;
  for (size_t j = 0; j < k; j++)       // element offsets -> device addresses
    meta[2 * j] = d_polys ? (uint64_t)(uintptr_t)d_polys[j] : (uint64_t)(uintptr_t)g_pw.polys.p + meta[2 * j] * 32;
  if ((rc = g_pw.meta.ensure(meta.size() * 8 + xip.size() * 4 + 64))) return rc;
  if ((rc = g_pw.comb.ensure(maxlen * 32))) return rc;
  if ((rc = g_pw.t.ensure(maxlen * 32 + 32))) return rc;
  uint64_t* d_meta = (uint64_t*)g_pw.meta.p;
  uint32_t* d_xip = (uint32_t*)(d_meta + meta.size());
  if (!d_polys) KZ_CUDA(cudaMemcpyAsync(g_pw.polys.p, polys, total * 32, cudaMemcpyHostToDevice, cx.stream));
  KZ_CUDA(cudaMemcpyAsync(d_meta, meta.data(), meta.size() * 8, cudaMemcpyHostToDevice, cx.stream));
  KZ_CUDA(cudaMemcpyAsync(d_xip, xip.data(), xip.size() * 4, cudaMemcpyHostToDevice, cx.stream));
  poly_combine_kernel<P><<<(unsigned)kz_div_up(maxlen, 128), 128, 0, cx.stream>>>(d_meta, d_xip, (uint32_t)k, maxlen, nullptr,
                                                                               (uint32_t*)g_pw.comb.p);
  KZ_LAUNCHED();
  if ((rc = suffix_horner<P>((uint32_t*)g_pw.comb.p, maxlen, fe_to_mont<P>(zc), (uint32_t*)g_pw.t.p))) return rc;
  uint32_t* d_T = (uint32_t*)g_pw.t.p;
  if (eval_out) KZ_CUDA(cudaMemcpyAsync(eval_out, d_T, 32, cudaMemcpyDeviceToHost, cx.stream));
  if (quotient && maxlen > 1) KZ_CUDA(cudaMemcpyAsync(quotient, d_T + P::N, (maxlen - 1) * 32, cudaMemcpyDefault, cx.stream));   // host or device destination
  if (quot_len) *quot_len = maxlen - 1;
  KZ_CUDA(cudaStreamSynchronize(cx.stream));
  if (do_msm) return kz_msm_dev_internal(handle, 0, d_T + P::N, maxlen - 1, out_xy, is_inf, nullptr);
  return 0;
}

// p(x) = T_0 of the suffix Horner recurrence: upward sweep only
template <class P>
int eval_impl(const uint32_t* d_c, size_t len, const uint64_t* x, uint64_t* out) {
  KzgpuCtx& cx = kz_ctx();
  Fe<P> xc = kz_fe_from_u64<P>(x);
  if (!kz_fe_reduced<P>(xc)) return kz_fail(KZGPU_ERANGE, "evaluation point must be a canonical field element");
  if (len == 0) { memset(out, 0, 32); return 0; }
  std::vector<size_t> lens{len};
  while (lens.back() > 1) lens.push_back(kz_div_up(lens.back(), CHL));
  size_t total = 0;
  for (size_t l = 1; l < lens.size(); l++) total += lens[l];
  int rc = g_pw.levels.ensure((total + 1) * P::N * 4);
  if (rc) return rc;
  uint32_t* p = (uint32_t*)g_pw.levels.p;
  Fe<P> zl = fe_to_mont<P>(xc);
  const uint32_t* cur = d_c;
  for (size_t l = 1; l < lens.size(); l++) {
    horner_chunks_kernel<P><<<(unsigned)kz_div_up(lens[l], 128), 128, 0, cx.stream>>>(cur, lens[l - 1], zl, p, lens[l]);
    KZ_LAUNCHED();
    cur = p; p += lens[l] * P::N;
    zl = hpow<P>(zl, CHL);
  }
  KZ_CUDA(cudaMemcpyAsync(out, cur, 32, cudaMemcpyDeviceToHost, cx.stream));
  KZ_CUDA(cudaStreamSynchronize(cx.stream));
  return 0;
}

template <class P>
int lincomb_impl(uint32_t* d_out, size_t out_len, const uint64_t* const* d_polys, const size_t* lens, const uint64_t* scalars, size_t k,
                 const uint64_t* constant) {
  KzgpuCtx& cx = kz_ctx();
  if (out_len == 0) return 0;
  std::vector<uint64_t> meta(2 * k + 2);
  std::vector<uint32_t> mult((k + 1) * P::N);
  for (size_t j = 0; j < k; j++) {
    meta[2 * j] = (uint64_t)(uintptr_t)d_polys[j]; meta[2 * j + 1] = lens[j];
    Fe<P> sc = kz_fe_from_u64<P>(scalars + 4 * j);
    if (!kz_fe_reduced<P>(sc)) return kz_fail(KZGPU_ERANGE, "scalar %zu is not a canonical field element", j);
    Fe<P> sm = fe_to_mont<P>(sc);
    memcpy(&mult[j * P::N], sm.v, P::N * 4);
  }
  if (constant) {
    Fe<P> cc = kz_fe_from_u64<P>(constant);
    if (!kz_fe_reduced<P>(cc)) return kz_fail(KZGPU_ERANGE, "constant is not a canonical field element");
    memcpy(&mult[k * P::N], cc.v, P::N * 4);
  }
  int rc;
  if ((rc = g_pw.meta.ensure(meta.size() * 8 + mult.size() * 4 + 64))) return rc;
  uint64_t* d_meta = (uint64_t*)g_pw.meta.p;
  uint32_t* d_mult = (uint32_t*)(d_meta + meta.size());
  KZ_CUDA(cudaMemcpyAsync(d_meta, meta.data(), meta.size() * 8, cudaMemcpyHostToDevice, cx.stream));
  KZ_CUDA(cudaMemcpyAsync(d_mult, mult.data(), mult.size() * 4, cudaMemcpyHostToDevice, cx.stream));
  poly_combine_kernel<P><<<(unsigned)kz_div_up(out_len, 128), 128, 0, cx.stream>>>(d_meta, d_mult, (uint32_t)k, out_len,
                                                                                constant ? d_mult + k * P::N : nullptr, d_out);
  KZ_LAUNCHED();
  KZ_CUDA(cudaStreamSynchronize(cx.stream));     // meta/mult are host vectors: keep them alive until consumed
  return 0;
}

template <class P>
int powers_impl(uint32_t* d_out, size_t n, const uint64_t* base, const uint64_t* scale) {
  KzgpuCtx& cx = kz_ctx();
  Fe<P> b = kz_fe_from_u64<P>(base), sc = fe_zero<P>();
  sc.v[0] = 1;
  if (scale) sc = kz_fe_from_u64<P>(scale);
  if (!kz_fe_reduced<P>(b) || !kz_fe_reduced<P>(sc)) return kz_fail(KZGPU_ERANGE, "base / scale must be canonical field elements");
  if (n == 0) return 0;
  size_t threads = kz_div_up(n, PCH);
  powers_kernel<P><<<(unsigned)kz_div_up(threads, 128), 128, 0, cx.stream>>>(d_out, n, fe_to_mont<P>(b), sc);
  KZ_LAUNCHED();
  return 0;
}

}  // namespace

void kz_poly_release() {
  if (kz_slot() != 0) return;          // polynomial / prover kernels run on the primary device only
  KzScratch* all[] = {&g_pw.polys, &g_pw.meta, &g_pw.comb, &g_pw.levels, &g_pw.t};
  for (auto* s : all) s->release();
}

extern "C" {

int kzgpu_open(uint64_t handle, const uint64_t* polys, const size_t* lens, size_t k, const uint64_t* z, const uint64_t* xi,
               uint64_t* out_affine_xy, int* is_inf, uint64_t* eval_out) {
  KZ_REQUIRE_INIT();
  if ((k && (!polys || !lens)) || !z || !xi || !out_affine_xy) return kz_fail(KZGPU_EINVAL, "null pointer");
  int curve = kz_srs_curve(handle);
  if (curve < 0) return kz_fail(KZGPU_EHANDLE, "unknown SRS handle %llu", (unsigned long long)handle);
  if (curve == KZGPU_BN254) return open_impl<FrBN254>(handle, true, polys, nullptr, lens, k, z, xi, out_affine_xy, is_inf, nullptr, nullptr, eval_out);
  return open_impl<FrBLS381>(handle, true, polys, nullptr, lens, k, z, xi, out_affine_xy, is_inf, nullptr, nullptr, eval_out);
}

int kzgpu_open_quotient(int field, const uint64_t* polys, const size_t* lens, size_t k, const uint64_t* z, const uint64_t* xi,
                        uint64_t* quotient, size_t* quot_len, uint64_t* eval_out) {
  KZ_REQUIRE_INIT();
  if ((k && (!polys || !lens)) || !z || !xi) return kz_fail(KZGPU_EINVAL, "null pointer");
  if (field == KZGPU_BN254) return open_impl<FrBN254>(0, false, polys, nullptr, lens, k, z, xi, nullptr, nullptr, quotient, quot_len, eval_out);
  if (field == KZGPU_BLS12_381) return open_impl<FrBLS381>(0, false, polys, nullptr, lens, k, z, xi, nullptr, nullptr, quotient, quot_len, eval_out);
  return kz_fail(KZGPU_EINVAL, "unknown field id %d", field);
}

int kzgpu_open_dev(uint64_t handle, const uint64_t* const* d_polys, const size_t* lens, size_t k, const uint64_t* z, const uint64_t* xi,
                   uint64_t* out_affine_xy, int* is_inf, uint64_t* eval_out) {
  KZ_REQUIRE_INIT();
  if ((k && (!d_polys || !lens)) || !z || !xi || !out_affine_xy) return kz_fail(KZGPU_EINVAL, "null pointer");
  int curve = kz_srs_curve(handle);
  if (curve < 0) return kz_fail(KZGPU_EHANDLE, "unknown SRS handle %llu", (unsigned long long)handle);
  if (curve == KZGPU_BN254)
    return open_impl<FrBN254>(handle, true, nullptr, d_polys, lens, k, z, xi, out_affine_xy, is_inf, nullptr, nullptr, eval_out);
  return open_impl<FrBLS381>(handle, true, nullptr, d_polys, lens, k, z, xi, out_affine_xy, is_inf, nullptr, nullptr, eval_out);
}

int kzgpu_open_quotient_dev(int field, const uint64_t* const* d_polys, const size_t* lens, size_t k, const uint64_t* z, const uint64_t* xi,
                            uint64_t* d_quotient, size_t* quot_len, uint64_t* eval_out) {
  KZ_REQUIRE_INIT();
  if ((k && (!d_polys || !lens)) || !z || !xi || !d_quotient) return kz_fail(KZGPU_EINVAL, "null pointer");
  if (field == KZGPU_BN254)
    return open_impl<FrBN254>(0, false, nullptr, d_polys, lens, k, z, xi, nullptr, nullptr, d_quotient, quot_len, eval_out);
  if (field == KZGPU_BLS12_381)
    return open_impl<FrBLS381>(0, false, nullptr, d_polys, lens, k, z, xi, nullptr, nullptr, d_quotient, quot_len, eval_out);
  return kz_fail(KZGPU_EINVAL, "unknown field id %d", field);
}

int kzgpu_poly_eval_dev(int field, const uint64_t* d_poly, size_t len, const uint64_t* x, uint64_t* out) {
  KZ_REQUIRE_INIT();
  if ((len && !d_poly) || !x || !out) return kz_fail(KZGPU_EINVAL, "null pointer");
  if (field == KZGPU_BN254) return eval_impl<FrBN254>((const uint32_t*)d_poly, len, x, out);
  if (field == KZGPU_BLS12_381) return eval_impl<FrBLS381>((const uint32_t*)d_poly, len, x, out);
  return kz_fail(KZGPU_EINVAL, "unknown field id %d", field);
}

int kzgpu_poly_lincomb_dev(int field, uint64_t* d_out, size_t out_len, const uint64_t* const* d_polys, const size_t* lens,
                           const uint64_t* scalars, size_t k, const uint64_t* constant) {
  KZ_REQUIRE_INIT();
  if ((out_len && !d_out) || (k && (!d_polys || !lens || !scalars))) return kz_fail(KZGPU_EINVAL, "null pointer");
  if (field == KZGPU_BN254) return lincomb_impl<FrBN254>((uint32_t*)d_out, out_len, d_polys, lens, scalars, k, constant);
  if (field == KZGPU_BLS12_381) return lincomb_impl<FrBLS381>((uint32_t*)d_out, out_len, d_polys, lens, scalars, k, constant);
  return kz_fail(KZGPU_EINVAL, "unknown field id %d", field);
}

int kzgpu_powers_dev(int field, uint64_t* d_out, size_t n, const uint64_t* base, const uint64_t* scale) {
  KZ_REQUIRE_INIT();
  if ((n && !d_out) || !base) return kz_fail(KZGPU_EINVAL, "null pointer");
  if (field == KZGPU_BN254) return powers_impl<FrBN254>((uint32_t*)d_out, n, base, scale);
  if (field == KZGPU_BLS12_381) return powers_impl<FrBLS381>((uint32_t*)d_out, n, base, scale);
  return kz_fail(KZGPU_EINVAL, "unknown field id %d", field);
}

}  // extern "C"
