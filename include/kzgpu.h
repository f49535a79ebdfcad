/* kzgpu.h -- C ABI of the B200 (sm_100a) KZG-MSM / NTT library (libkzgpu.so).
 *
 * The reference (swusjask/kzg-snark) is pure Python and has no FFI of its own; its
 * drop-in boundary is the Python surface of kzg.py and fft_ff.py.
 * The functions below are what a ctypes/cffi binding under those modules calls; each
 * cites the reference code it replaces.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - every function returns 0 on success or a negative KZGPU_E* code; it never throws and
 *    never calls back into Python.  kzgpu_last_error() gives the message of the last
 *    failure on the calling thread's library context.
 *  - field elements cross the boundary as little-endian uint64 limbs of the CANONICAL
 *    residue in [0, modulus): 4 limbs for both scalar fields and the BN254 base field,
 *    6 limbs for the BLS12-381 base field.  Montgomery form never leaves the device.
 *  - G1 points cross as affine (x, y) = 2 * fp_limbs64 limbs; the point at infinity is
 *    (0, 0) (py_ecc's Z1 = (1, 1, 0), kzg.py:42, is mapped to it by the host shim).
 *  - pointers are caller-owned host buffers unless the name starts with d_ (device
 *    pointers obtained from kzgpu_alloc).
 *  - one caller thread at a time.  The library drives one device (kzgpu_init; also the layout of the one-process-per-GPU
 *    torchrun harness) or several (kzgpu_init_multi): then the entry points that shard naturally -- kzgpu_msm,
 *    kzgpu_msm_dev, kzgpu_msm_batch, kzgpu_ntt_batch, kzgpu_open* -- spread their work over the devices from this one
 *    process, with no PyTorch and no launcher.  Device pointers (d_*) and
 *    everything else always refer to the FIRST device of the list.
 */
#ifndef KZGPU_H
#define KZGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KZGPU_OK            0
#define KZGPU_EINVAL       -1   /* bad argument (size, curve id, non power of two ...) */
#define KZGPU_ECUDA        -2   /* CUDA runtime failure (no device, OOM, launch error) */
#define KZGPU_ENOTINIT     -3   /* kzgpu_init not called */
#define KZGPU_ERANGE       -4   /* operand not reduced / degree exceeds the SRS (kzg.py:103-106) */
#define KZGPU_EHANDLE      -5   /* unknown SRS handle */

/* curve ids: the two curve_type strings of KZG.__init__ (kzg.py:26-37) */
#define KZGPU_BN254         0
#define KZGPU_BLS12_381     1

/* ---- context ----------------------------------------------------------------------- */
int kzgpu_init(int device);                  /* cudaSetDevice + streams; idempotent */
/* SURVEY.md 8(b)'s kzgpu_init(ndev, devs): initialise on `ndev` devices (devs == NULL: devices 0 .. ndev-1; ndev <= 0: every
 * visible device).  devs[0] is the primary device.  One worker thread per further device; peer access is enabled where
 * the hardware offers it (NVLink / NVSwitch), otherwise peer copies stage through the host.
 *   kzgpu_msm / kzgpu_msm_dev / kzgpu_open*: an MSM of >= KZGPU_SHARD_MIN points (env, default 2^20) is point-sharded:
 *     device d owns the contiguous range d of the key (with window tables sized for the SHARD), reduces it to one XYZZ
 *     partial sum, writes it peer-to-peer into the primary device's gather buffer, and the primary folds the ndev partials
 *     (kzg.py:112-116 is a plain sum, so any partition of the index range is exact).
 *   kzgpu_msm_batch: the k polynomials of one commit() (kzg.py:102) are placed whole, longest first, on the least
 *     loaded device (each holds a replica of the key); polynomials of >= KZGPU_SHARD_MIN coefficients are point-sharded.
 *   kzgpu_ntt_batch: whole vectors per device.
 * Shards and replicas of a key are built on first use from the primary device's copy (peer copies + local table build). */
int kzgpu_init_multi(int ndev, const int* devs);
int kzgpu_device_count(int* ndev);           /* devices the library is initialised on */
int kzgpu_shutdown(void);                    /* frees SRS handles, NTT plans, workspaces */
int kzgpu_last_error(char* buf, size_t cap);
int kzgpu_device_info(char* name, size_t cap, int* sm_count, size_t* total_mem);
int kzgpu_fp_limbs64(int curve);             /* 4 (BN254) or 6 (BLS12-381); <0 on bad id */

/* ---- device memory (bench / multi-GPU plumbing; lets callers keep data resident) ---- */
int kzgpu_alloc(void** d_ptr, size_t bytes);
int kzgpu_free(void* d_ptr);
int kzgpu_h2d(void* d_dst, const void* src, size_t bytes);
int kzgpu_d2h(void* dst, const void* d_src, size_t bytes);
/* stream-ordered device->device copy and fill (zero padding of coefficient vectors before an NTT) */
int kzgpu_d2d(void* d_dst, const void* d_src, size_t bytes);
int kzgpu_memset(void* d_dst, int byte, size_t bytes);
int kzgpu_sync(void);
/* page-locked host memory (cudaHostAlloc) so that host<->device copies of the timed e2e
 * path run at PCIe rate */
int kzgpu_host_alloc(void** h_ptr, size_t bytes);
int kzgpu_host_free(void* h_ptr);
/* run every subsequent launch/copy on the caller's stream (a cudaStream_t passed as void*);
 * NULL restores the library's own stream.  Lets a host that already owns a stream (e.g. the
 * torch.distributed/NCCL plumbing of the multi-GPU bench) order and time the kernels. */
int kzgpu_set_stream(void* cuda_stream);
/* CUDA-event timing on the library's stream (bench.py: device time of a bracketed region) */
int kzgpu_timer_start(void);
int kzgpu_timer_stop(float* ms);

/* ---- SRS: the commitment key ck = [tau^i * G1] (kzg.py:69-72, consumed at :99,:115) -- */
/* Upload n affine points (canonical limbs); converted to Montgomery form on the device and
 * kept resident (ck is passed on every commit/open call, SURVEY.md section 7 hard part 4). */
int kzgpu_srs_create(int curve, const uint64_t* affine_xy, size_t n, uint64_t* handle);
/* Device-side replacement of the setup loop kzg.py:69-72 for a given secret tau
 * (canonical, 4 limbs): points[i] = tau^i * G1, i < n. */
int kzgpu_srs_generate(int curve, const uint64_t* tau, size_t n, uint64_t* handle);
/* same for the index range [start, start + n): points[i] = tau^(start + i) * G1 -- one GPU's
 * shard of a point-sharded SRS */
int kzgpu_srs_generate_range(int curve, const uint64_t* tau, size_t start, size_t n, uint64_t* handle);
int kzgpu_srs_destroy(uint64_t handle);
int kzgpu_srs_size(uint64_t handle, size_t* n);
/* storage layout of a key: window bits c and table count W of the precomputed fixed-base
 * tables T_w[i] = 2^(c w) * P_i (c = 0, W = 1: plain key), and the device bytes held.
 * Environment: KZGPU_SRS_TABLES=0 disables the tables, KZGPU_SRS_TABLES=c=NN forces c,
 * KZGPU_SRS_TABLE_GIB caps their size (default 35% of device memory). */
int kzgpu_srs_info(uint64_t handle, int* c_tab, int* w_tab, size_t* device_bytes);
/* read back `count` points starting at `first` as canonical affine limbs */
int kzgpu_srs_read(uint64_t handle, size_t first, size_t count, uint64_t* affine_xy);

/* ---- MSM: KZG.commit's inner loop, sum_i scalars[i] * ck[first + i]  (kzg.py:108-116) -- */
/* scalars: n * 4 limbs, canonical mod r.  out_affine_xy: 2 * fp_limbs64 limbs; *is_inf = 1
 * and out = (0,0) when the sum is the identity (zero polynomial -> Z1, kzg.py:109).
 * KZGPU_ERANGE if first + n exceeds the SRS (the degree check kzg.py:103-106). */
int kzgpu_msm(uint64_t handle, size_t first, const uint64_t* scalars, size_t n,
              uint64_t* out_affine_xy, int* is_inf);
/* same with scalars already on the device */
int kzgpu_msm_dev(uint64_t handle, size_t first, const uint64_t* d_scalars, size_t n,
                  uint64_t* out_affine_xy, int* is_inf);
/* k polynomials of one commit() call (kzg.py:102): scalars concatenated, lens[j] limbs-of-4 counts */
int kzgpu_msm_batch(uint64_t handle, const uint64_t* scalars, const size_t* lens, size_t k,
                    uint64_t* out_affine_xy, int* is_inf);
/* same with the polynomials on the device: k vectors of poly_len scalars each, back to back (pad shorter ones with
 * zeros: zero coefficients are skipped, kzg.py:113).  All k MSMs share one sort / accumulate / reduce pass. */
int kzgpu_msm_batch_dev(uint64_t handle, const uint64_t* d_scalars, size_t poly_len, size_t k,
                        uint64_t* out_affine_xy, int* is_inf);
/* Partial sum for point-sharded multi-GPU MSM: result left un-normalised as XYZZ in
 * Montgomery form (4 * fp_limbs64 limbs), to be all-gathered and folded by kzgpu_g1_fold.
 * Both forms return with the work QUEUED on the library's stream (no host synchronisation: the caller's exchange is
 * queued right behind the partial); a non-canonical scalar (KZGPU_ERANGE) is then reported by the call that
 * synchronises next -- kzgpu_g1_fold, kzgpu_sync, or the next MSM. */
int kzgpu_msm_partial_dev(uint64_t handle, size_t first, const uint64_t* d_scalars, size_t n,
                          uint64_t* d_out_xyzz);
/* same with this rank's slice of the scalars in host memory (the multi-GPU form of kzgpu_msm: kzg.py:112-116 over an
 * index range, upload chunked and overlapped with the compute); the partial stays on the device for the all-gather */
int kzgpu_msm_partial(uint64_t handle, size_t first, const uint64_t* scalars, size_t n,
                      uint64_t* d_out_xyzz);
/* sum of `count` XYZZ partials (device) -> canonical affine on the host */
int kzgpu_g1_fold(int curve, const uint64_t* d_xyzz, size_t count,
                  uint64_t* out_affine_xy, int* is_inf);

/* Verifier-side combinations: sum_i scalars[i] * P_i over `count` ARBITRARY affine points given on
 * the host (commitments, proofs, G1) -- the loops of KZG.check / batch_check, kzg.py:183-205 and :252-281, and of the
 * PLONK / Marlin verifiers (plonk/verifier.py:117-157).  Small counts only (<= 65536); the pairing stays with the caller. */
int kzgpu_g1_lincomb(int curve, const uint64_t* affine_xy, const uint64_t* scalars, size_t count,
                     uint64_t* out_affine_xy, int* is_inf);

/* ---- NTT: fft_ff / ifft_ff (fft_ff.py:3-58) and the coset variant ---------------------- */
/* In place, natural order in and out:  out[k] = sum_j data[j] * (shift^j) * w^(j k)
 * inverse != 0: uses w^-1 and scales by n^-1 (fft_ff.py:53-58); with a shift, output j is
 * additionally multiplied by shift^-j (inverse of the forward coset transform).
 * n must be a power of two (fft_ff.py:74) and w a canonical element (4 limbs) whose order
 * is caller-guaranteed (fft_ff.py:77-78 is checked by the host shim).  field = curve id
 * (the scalar field of that curve).  coset_shift may be NULL. */
int kzgpu_ntt(int field, uint64_t* data, size_t n, const uint64_t* w, int inverse,
              const uint64_t* coset_shift);
int kzgpu_ntt_dev(int field, uint64_t* d_data, size_t n, const uint64_t* w, int inverse,
                  const uint64_t* coset_shift);
/* `batch` independent vectors of the same length n, contiguous */
int kzgpu_ntt_batch(int field, uint64_t* data, size_t n, size_t batch, const uint64_t* w,
                    int inverse, const uint64_t* coset_shift);
int kzgpu_ntt_batch_dev(int field, uint64_t* d_data, size_t n, size_t batch, const uint64_t* w,
                        int inverse, const uint64_t* coset_shift);

/* ---- open: KZG.open (kzg.py:122-159) ---------------------------------------------------- */
/* polys: k coefficient vectors concatenated (low -> high, 4 limbs each), lens[j] coefficients.
 * Computes P = sum_j xi^(j+1) * poly_j (kzg.py:148-150), W = (P - P(z)) / (X - z)
 * (kzg.py:153-154) and returns commit(W) (kzg.py:157).  eval_out (4 limbs, may be NULL)
 * receives P(z). */
int kzgpu_open(uint64_t handle, const uint64_t* polys, const size_t* lens, size_t k,
               const uint64_t* z, const uint64_t* xi, uint64_t* out_affine_xy, int* is_inf,
               uint64_t* eval_out);
/* the polynomial half alone: combination + synthetic division, quotient returned to the host
 * (max(lens) - 1 coefficients; *quot_len receives the count) */
int kzgpu_open_quotient(int field, const uint64_t* polys, const size_t* lens, size_t k,
                        const uint64_t* z, const uint64_t* xi, uint64_t* quotient,
                        size_t* quot_len, uint64_t* eval_out);

/* same with the k polynomials already on the device: d_polys[j] -> lens[j] canonical coefficients */
int kzgpu_open_dev(uint64_t handle, const uint64_t* const* d_polys, const size_t* lens, size_t k,
                   const uint64_t* z, const uint64_t* xi, uint64_t* out_affine_xy, int* is_inf,
                   uint64_t* eval_out);

/* the polynomial half with device-resident inputs, quotient left on the device (max(lens) - 1 coefficients): lets a
 * multi-GPU host MSM index ranges of the quotient with kzgpu_msm_partial_dev (point-sharded open) */
int kzgpu_open_quotient_dev(int field, const uint64_t* const* d_polys, const size_t* lens, size_t k,
                            const uint64_t* z, const uint64_t* xi, uint64_t* d_quotient, size_t* quot_len,
                            uint64_t* eval_out);

/* ---- polynomial kernels for the callers of commit / open ----------------- */
/* The reference's PLONK prover does this work with Sage polynomial arithmetic between its
 * kzg.commit / kzg.open / fft_ff_interpolation calls; these keep it on the device so that the
 * polynomials never cross PCIe.  All vectors are canonical residues, 4 limbs per element. */
/* out = p(x): poly(zeta) in round 4 (plonk/prover.py:161-166) */
int kzgpu_poly_eval_dev(int field, const uint64_t* d_poly, size_t len, const uint64_t* x, uint64_t* out);
/* d_out[i] = constant*[i == 0] + sum_j scalars[j] * d_polys[j][i] for i < out_len (coefficients
 * beyond lens[j] read as 0): blinding terms (plonk/prover.py:84-86), the split of t(X) (:336-351)
 * and the linearisation polynomial r(X) (:383-407).  constant may be NULL. */
int kzgpu_poly_lincomb_dev(int field, uint64_t* d_out, size_t out_len, const uint64_t* const* d_polys,
                           const size_t* lens, const uint64_t* scalars, size_t k, const uint64_t* constant);
/* d_out[i] = scale * base^i, i < n (the subgroup H of plonk/encoder.py:45 and the evaluation coset);
 * scale may be NULL (= 1) */
int kzgpu_powers_dev(int field, uint64_t* d_out, size_t n, const uint64_t* base, const uint64_t* scale);
/* Permutation grand product (plonk/prover.py:245-258): d_z[0] = 1,
 * d_z[i+1] = d_z[i] * num_i / den_i over the wire values d_a/d_b/d_c (n each), the 3n permutation
 * images d_sigma_star (plonk/encoder.py:137) and d_H[i] = g^i.  *zero_den = 1 if some den_i was 0
 * (the reference raises ValueError at :254-255). */
int kzgpu_plonk_permutation_dev(int field, size_t n, const uint64_t* d_a, const uint64_t* d_b, const uint64_t* d_c,
                                const uint64_t* d_sigma_star, const uint64_t* d_H, const uint64_t* k1,
                                const uint64_t* k2, const uint64_t* beta, const uint64_t* gamma,
                                uint64_t* d_z, int* zero_den);
/* Quotient numerator over v_H on the evaluation coset {s * w_4n^i} (plonk/prover.py:297-310):
 * d_evals = 15 device vectors of n4 = 4n evaluations in the order a, b, c, z, qM, qL, qR, qO, qC,
 * S_sigma1, S_sigma2, S_sigma3, PI, L1, X (the coset points themselves); params = 9 elements
 * alpha, beta, gamma, k1, k2, 1/v_H(x_0..3) (v_H has period 4 on the coset).  d_t receives
 * t(x_i), canonical; one inverse coset NTT of it gives t's coefficients.  mont_in != 0: the 15 vectors hold
 * Montgomery-form values v * 2^256 mod r (scale the coefficient vectors by 2^256 mod r before their coset NTTs, e.g.
 * with kzgpu_poly_lincomb_dev -- the transform is linear), which saves the 16 per-point conversions. */
int kzgpu_plonk_quotient_dev(int field, size_t n4, const uint64_t* const* d_evals, const uint64_t* params,
                             int mont_in, uint64_t* d_t);

/* d_out[i] = d_a[i] * d_b[i]: the point-wise step of NTT-based polynomial products (the Sage `*` of marlin/prover.py:96,
 * 131) */
int kzgpu_poly_mul_pointwise_dev(int field, uint64_t* d_out, const uint64_t* d_a, const uint64_t* d_b, size_t n);
/* d_out = M d_z for a CSR matrix over the scalar field (row_ptr: n_rows + 1 offsets, col: column indices, d_val: values),
 * the z_A = A z, z_B = B z, z_C = C z of marlin/encoder.py:205-207 */
int kzgpu_spmv_dev(int field, size_t n_rows, const uint32_t* d_row_ptr, const uint32_t* d_col, const uint64_t* d_val,
                   const uint64_t* d_z, uint64_t* d_out);
/* Marlin third round on the coset {s w_8m^i} (marlin/prover.py:166-171, 303-353): d_row / d_col / d_val hold the
 * coset evaluations of row_M, col_M, val_M for M = A, B, C back to back (3 * m8 each), d_f2 those of f_2;
 * params = eta_A, eta_B, eta_C, alpha, beta_1, v_H(beta_1) v_H(alpha), 1 / v_K(x_0..7) (14 elements).
 * d_out[i] = (a - b f_2)(x_i) / v_K(x_i) with a, b as in _compute_a_b_polynomials; its inverse coset NTT is h_2. */
int kzgpu_marlin_h2_evals_dev(int field, size_t m8, const uint64_t* d_row, const uint64_t* d_col, const uint64_t* d_val,
                              const uint64_t* d_f2, const uint64_t* params, uint64_t* d_out);
/* Marlin prover evaluation loops.  d_row / d_col / d_val: the K-domain evaluations of the index
 * polynomials row_M, col_M, val_M for M = A, B, C back to back (3 * m elements each; marlin/encoder.py:98-125).
 * f2 evaluations (marlin/prover.py:404-466):  d_out[kappa] = scale * sum_M eta_M val_M(kappa) /
 * ((beta1 - row_M(kappa)) (alpha - col_M(kappa))), terms with a zero denominator skipped as in the reference;
 * scale = v_H(beta1) v_H(alpha); one inverse NTT over K gives f_2 (:469). */
int kzgpu_marlin_f2_evals_dev(int field, size_t m, const uint64_t* d_row, const uint64_t* d_col, const uint64_t* d_val,
                              const uint64_t* eta, const uint64_t* alpha, const uint64_t* beta1, const uint64_t* scale,
                              uint64_t* d_out);
/* t(X) on H (marlin/prover.py:248-301): d_out[i] = t(h_i) = scale * h_i^-1 * sum over the entries (M, kappa) whose row is
 * h_i of eta_M val_M(kappa) / (alpha - col_M(kappa)), scale = n v_H(alpha) (v_H(X)/(X - h) vanishes on H except at h, where
 * it is n/h).  d_row_index: 3 * m indices i of row_M(kappa) = h_i, ascending per matrix (the index is built in row-major
 * order), 0xffffffff for padding entries; d_H[i] = h_i.  One inverse NTT over H gives t's coefficients. */
int kzgpu_marlin_t_evals_dev(int field, size_t n, size_t m, const uint32_t* d_row_index, const uint64_t* d_col,
                             const uint64_t* d_val, const uint64_t* d_H, const uint64_t* eta, const uint64_t* alpha,
                             const uint64_t* scale, uint64_t* d_out);

/* ---- diagnostics used by the parity tests and bench.py ---------------------------------- */
/* elementwise Montgomery-core check: out[i] = a[i] op b[i] in the chosen field.
 * which: 0 = Fp(curve), 1 = Fr(curve); op: 0 mul, 1 add, 2 sub, 3 inverse(a). */
int kzgpu_field_op(int curve, int which, int op, const uint64_t* a, const uint64_t* b,
                   uint64_t* out, size_t n);
/* (the throughput microbenchmarks live in their own library: include/kzgpu_bench.h, libkzgpu_bench.so) */
/* per-kernel device timing (CUDA events on the launching stream) for bench.py's roofline:
 * which: 0 = MSM bucket-accumulate kernel, 1 = NTT pass kernel, 2 = MSM sort (histogram+scatter),
 *        3 = MSM bucket reduction + window fold.  Accumulates while enabled. */
int kzgpu_profile_enable(int on);
int kzgpu_profile_reset(void);
int kzgpu_profile_get(int which, double* total_ms, uint64_t* launches, double* work_units);
/* number of kernel launches issued by the library since init (bench.py "gpu_launches") */
int kzgpu_launch_count(uint64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* KZGPU_H */
