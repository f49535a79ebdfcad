#!/usr/bin/env python3
"""Generate mp_prims_gen.cuh: carry-chain multiprecision primitives for N = 8 and
N = 12 32-bit limbs, each as ONE inline-PTX statement so the CC flag never
crosses an asm boundary (the compiler does not model CC between statements).

ptxas fuses every (mad.lo.cc, madc.hi.cc) pair below into a single
IMAD.WIDE.U32[.X] with the carry in a predicate register (checked with
cuobjdump -sass, see DESIGN.md "Field core").

Run:  python gen_prims.py > mp_prims_gen.cuh
"""

import sys


def emit(N):
    H = N // 2
    out = []
    w = out.append
    w(f"// ---------------------------------------------------------------- N = {N}")
    w(f"template <> struct MpPrims<{N}> {{")

    # mul_even: acc[2j], acc[2j+1] = a[2j] * b
    w("  // acc[2j..2j+1] = a[2j] * b  for j < N/2 (no carries: the products do not overlap)")
    w("  static __device__ __forceinline__ void mul_even(uint32_t* acc, const uint32_t* a, uint32_t b) {")
    s = []
    for j in range(H):
        s.append(f"mul.lo.u32 %{2*j}, %{N+j}, %{N+H}; mul.hi.u32 %{2*j+1}, %{N+j}, %{N+H};")
    outs = ", ".join(f'"=r"(acc[{k}])' for k in range(N))
    ins = ", ".join(f'"r"(a[{2*j}])' for j in range(H)) + ', "r"(b)'
    w('    asm("' + '"\n        "'.join(s) + '"')
    w(f"        : {outs}\n        : {ins});")
    w("  }")

    # mad_even: acc += sum a[2j]*b << 64j ; top += carry
    w("  // acc[0..N) += sum_j a[2j]*b * 2^(64j); the carry out of acc[N-1] is added to `top`")
    w("  static __device__ __forceinline__ void mad_even(uint32_t* acc, const uint32_t* a, uint32_t b, uint32_t& top) {")
    s = []
    for j in range(H):
        lo = "mad.lo.cc.u32" if j == 0 else "madc.lo.cc.u32"
        s.append(f"{lo} %{2*j}, %{N+1+j}, %{N+1+H}, %{2*j}; madc.hi.cc.u32 %{2*j+1}, %{N+1+j}, %{N+1+H}, %{2*j+1};")
    s.append(f"addc.u32 %{N}, %{N}, 0;")
    outs = ", ".join(f'"+r"(acc[{k}])' for k in range(N)) + ', "+r"(top)'
    ins = ", ".join(f'"r"(a[{2*j}])' for j in range(H)) + ', "r"(b)'
    w('    asm("' + '"\n        "'.join(s) + '"')
    w(f"        : {outs}\n        : {ins});")
    w("  }")

    # mad_even_nc: same without top (carry dropped; caller guarantees none)
    w("  // same, carry out of acc[N-1] known to be zero (value bound, see field.cuh)")
    w("  static __device__ __forceinline__ void mad_even_nc(uint32_t* acc, const uint32_t* a, uint32_t b) {")
    s = []
    for j in range(H):
        lo = "mad.lo.cc.u32" if j == 0 else "madc.lo.cc.u32"
        hi = "madc.hi.cc.u32" if j < H - 1 else "madc.hi.u32"
        s.append(f"{lo} %{2*j}, %{N+j}, %{N+H}, %{2*j}; {hi} %{2*j+1}, %{N+j}, %{N+H}, %{2*j+1};")
    outs = ", ".join(f'"+r"(acc[{k}])' for k in range(N))
    ins = ", ".join(f'"r"(a[{2*j}])' for j in range(H)) + ', "r"(b)'
    w('    asm("' + '"\n        "'.join(s) + '"')
    w(f"        : {outs}\n        : {ins});")
    w("  }")

    # shift_mad: y0 += x[1] (carry into chain); x[j..j+1] = a[2j']*b + x[j+2..j+3]
    w("  // y0 += x[1]; then x := (x >> 64) + sum_j a[2j]*b * 2^(64j), the carry of the first add")
    w("  // entering the chain at x[0].  (x[0] is dropped: the caller has made it zero.)")
    w("  static __device__ __forceinline__ void shift_mad(uint32_t* x, uint32_t& y0, const uint32_t* a, uint32_t b) {")
    s = [f"add.cc.u32 %{N}, %{N}, %1;"]
    for j in range(H - 1):
        s.append(f"madc.lo.cc.u32 %{2*j}, %{N+1+j}, %{N+1+H}, %{2*j+2}; madc.hi.cc.u32 %{2*j+1}, %{N+1+j}, %{N+1+H}, %{2*j+3};")
    j = H - 1
    s.append(f"madc.lo.cc.u32 %{2*j}, %{N+1+j}, %{N+1+H}, 0; madc.hi.u32 %{2*j+1}, %{N+1+j}, %{N+1+H}, 0;")
    outs = ", ".join(f'"+r"(x[{k}])' for k in range(N)) + ', "+r"(y0)'
    ins = ", ".join(f'"r"(a[{2*j}])' for j in range(H)) + ', "r"(b)'
    w('    asm("' + '"\n        "'.join(s) + '"')
    w(f"        : {outs}\n        : {ins});")
    w("  }")

    # squaring rows: the two chains with the first K products left out (field.cuh fe_sqr_nofinal)
    w("  // mad_even without its first K products (chain starts at acc[2K]; K == N/2: nothing to add)")
    w("  template <int K> static __device__ __forceinline__ void mad_even_s(uint32_t* acc, const uint32_t* a, uint32_t b, uint32_t& top) {")
    for K in range(H):
        cnt = H - K                       # products
        nacc = N - 2 * K                  # accumulator limbs touched
        s = []
        for t in range(cnt):
            lo = "mad.lo.cc.u32" if t == 0 else "madc.lo.cc.u32"
            s.append(f"{lo} %{2*t}, %{nacc+1+t}, %{nacc+1+cnt}, %{2*t}; madc.hi.cc.u32 %{2*t+1}, %{nacc+1+t}, %{nacc+1+cnt}, %{2*t+1};")
        s.append(f"addc.u32 %{nacc}, %{nacc}, 0;")
        outs = ", ".join(f'"+r"(acc[{2*K+k}])' for k in range(nacc)) + ', "+r"(top)'
        ins = ", ".join(f'"r"(a[{2*(K+t)}])' for t in range(cnt)) + ', "r"(b)'
        kw = "if" if K == 0 else "else if"
        w(f"    {kw} constexpr (K == {K}) {{")
        w('      asm("' + '"\n          "'.join(s) + '"')
        w(f"          : {outs}\n          : {ins});")
        w("    }")
    w("  }")
    w("  // shift_mad without its first K products: those slots only pass the carry on")
    w("  template <int K> static __device__ __forceinline__ void shift_mad_s(uint32_t* x, uint32_t& y0, const uint32_t* a, uint32_t b) {")
    for K in range(H):
        cnt = H - K
        s = [f"add.cc.u32 %{N}, %{N}, %1;"]
        for j in range(H):
            if j < K:
                s.append(f"addc.cc.u32 %{2*j}, %{2*j+2}, 0; addc.cc.u32 %{2*j+1}, %{2*j+3}, 0;")
            elif j < H - 1:
                s.append(f"madc.lo.cc.u32 %{2*j}, %{N+1+(j-K)}, %{N+1+cnt}, %{2*j+2}; madc.hi.cc.u32 %{2*j+1}, %{N+1+(j-K)}, %{N+1+cnt}, %{2*j+3};")
            else:
                s.append(f"madc.lo.cc.u32 %{2*j}, %{N+1+(j-K)}, %{N+1+cnt}, 0; madc.hi.u32 %{2*j+1}, %{N+1+(j-K)}, %{N+1+cnt}, 0;")
        outs = ", ".join(f'"+r"(x[{k}])' for k in range(N)) + ', "+r"(y0)'
        ins = ", ".join(f'"r"(a[{2*(K+t)}])' for t in range(cnt)) + ', "r"(b)'
        kw = "if" if K == 0 else "else if"
        w(f"    {kw} constexpr (K == {K}) {{")
        w('      asm("' + '"\n          "'.join(s) + '"')
        w(f"          : {outs}\n          : {ins});")
        w("    }")
    w("  }")

    # merge: r[k] = o[k] + e[k+1], r[N-1] = o[N-1] + carry
    w("  // r[k] = o[k] + e[k+1] (k < N-1), r[N-1] = o[N-1] + carry")
    w("  static __device__ __forceinline__ void merge(uint32_t* r, const uint32_t* o, const uint32_t* e) {")
    s = []
    for k in range(N - 1):
        op = "add.cc.u32" if k == 0 else "addc.cc.u32"
        s.append(f"{op} %{k}, %{N+k}, %{2*N+k};")
    s.append(f"addc.u32 %{N-1}, %{2*N-1}, 0;")
    outs = ", ".join(f'"=r"(r[{k}])' for k in range(N))
    ins = ", ".join(f'"r"(o[{k}])' for k in range(N)) + ", " + ", ".join(f'"r"(e[{k+1}])' for k in range(N - 1))
    w('    asm("' + '"\n        "'.join(s) + '"')
    w(f"        : {outs}\n        : {ins});")
    w("  }")

    # add / sub with carry/borrow out
    for name, op0, opc, last, doc in (
        ("add_cc", "add.cc.u32", "addc.cc.u32", "addc.u32 %{c}, 0, 0;", "r = a + b, returns the carry (0/1)"),
        ("sub_cc", "sub.cc.u32", "subc.cc.u32", "subc.u32 %{c}, 0, 0;", "r = a - b, returns 0 or 0xffffffff (borrow mask)"),
    ):
        w(f"  // {doc}")
        w(f"  static __device__ __forceinline__ uint32_t {name}(uint32_t* r, const uint32_t* a, const uint32_t* b) {{")
        w("    uint32_t c;")
        s = []
        for k in range(N):
            op = op0 if k == 0 else opc
            s.append(f"{op} %{k}, %{N+1+k}, %{2*N+1+k};")
        s.append(last.replace("{c}", str(N)))
        outs = ", ".join(f'"=r"(r[{k}])' for k in range(N)) + ', "=r"(c)'
        ins = ", ".join(f'"r"(a[{k}])' for k in range(N)) + ", " + ", ".join(f'"r"(b[{k}])' for k in range(N))
        w('    asm("' + '"\n        "'.join(s) + '"')
        w(f"        : {outs}\n        : {ins});")
        w("    return c;")
        w("  }")
    w("};")
    return "\n".join(out)


def main():
    print("// GENERATED by gen_prims.py -- do not edit; regenerate with `python gen_prims.py > mp_prims_gen.cuh`.")
    print("// Device-only carry-chain primitives; the host mirror of the same contract is in mp_prims.cuh.")
    print("#pragma once")
    print("#include <cstdint>")
    print("template <int N> struct MpPrims;")
    for N in (8, 12):
        print(emit(N))


if __name__ == "__main__":
    main()
