#!/usr/bin/env python3
"""Static instruction mix of the hot kernels from the built objects (`cuobjdump -sass`, no GPU needed): per kernel the
number of SASS instructions by class -- IMAD.WIDE (the 32x32->64 multiplier, 4 clk per warp instruction), other IMAD, ALU
(IADD3 / LOP3 / SEL / SHF / MOV ...), shared / global memory, control -- plus registers and spills from the ELF.
The NTT and accumulate kernels are straight-line per element / per mixed addition, so the mix explains the issue model of
DESIGN.md 4.3 (wide x 4 + narrow x 2 + other x 1.25 clk).

    python scripts/sass_mix.py [regex ...]          # default: accumulate, ntt_pass_kernel_c<BN254, 8>, reduce, quotient
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "kzg_snark_b200", "build")
WANT = sys.argv[1:] or [r"msm_accumulate_kernel.*BN254", r"ntt_pass_kernel_cI7FrBN254Li8ELb1", r"ntt_pass_kernel_cI7FrBN254Li8ELb0",
                        r"ntt_pass_kernelI7FrBN254", r"msm_reduce_kernel.*BN254", r"quotient_kernelI7FrBN254"]
CLASSES = [("IMAD.WIDE", r"^IMAD\.WIDE"), ("IMAD other", r"^IMAD"), ("IADD3", r"^IADD3"), ("LOP3/SEL/SHF/MOV/LEA", r"^(LOP3|SEL|SHF|MOV|LEA|CS2R|PRMT|ISETP)"),
           ("LDS/STS", r"^(LDS|STS)"), ("LDG/STG/LDC/LDL/STL", r"^(LDG|STG|LDC|LDCU|LDL|STL|ST|LD)\b"), ("BAR/BRA/other", r".")]


def main():
    for obj in sorted(os.listdir(OBJ)):
        if not obj.endswith(".o"):
            continue
        path = os.path.join(OBJ, obj)
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
        regs = {}
        for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+).*?LOCAL:(\d+)", res):
            regs[m.group(1)] = (int(m.group(2)), int(m.group(3)))
        for f in re.split(r"\n\s+Function : ", sass)[1:]:
            name = f.split("\n")[0].strip()
            if not any(re.search(w, name) for w in WANT):
                continue
            ops = collections.Counter()
            for line in f.split("\n"):
                m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
                if not m:
                    continue
                for cls, pat in CLASSES:
                    if re.match(pat, m.group(1)):
                        ops[cls] += 1
                        break
            tot = sum(ops.values())
            demangled = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
            cut = demangled.rfind(">(")                       # template kernels: keep <...>, drop the parameter list
            short = demangled[:cut + 1] if cut >= 0 else demangled.split("(")[0]
            short = short.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
            r = regs.get(name, ("?", "?"))
            print(f"{short}  [{obj}]  registers {r[0]}, local bytes {r[1]}, {tot} instructions")
            for cls, _ in CLASSES:
                print(f"    {cls:24s} {ops[cls]:6d}  {100.0 * ops[cls] / tot:5.1f} %")


if __name__ == "__main__":
    main()
