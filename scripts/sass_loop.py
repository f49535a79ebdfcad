#!/usr/bin/env python3
"""Instruction mix of the INNER LOOP of a kernel (the span of its longest backward branch) from `cuobjdump -sass` of the
built object -- the executed work per mixed addition that bench.py's `roofline.frac` is computed from.
    python scripts/sass_loop.py [object.o] [kernel-name regex]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "kzg_snark_b200", "build", "msm.o")
want = sys.argv[2] if len(sys.argv) > 2 else r"msm_accumulate_kernel.*BN254"
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
for f in re.split(r"\n\s+Function : ", sass)[1:]:
    name = f.split("\n")[0].strip()
    if not re.search(want, name):
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+((?:@!?U?P\w+\s+)?)([A-Z0-9_.]+)(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(3), m.group(4), m.group(2).strip()))
    loops = []
    for addr, op, rest, pred in ins:
        if op.startswith("BRA"):
            t = re.search(r"0x([0-9a-f]+)", rest)
            if t and int(t.group(1), 16) < addr:
                loops.append((int(t.group(1), 16), addr))
    print(name, len(ins), "instructions; backward branches:", [(hex(a), hex(b), sum(1 for i in ins if a <= i[0] <= b)) for a, b in loops])
    for a, b in sorted(loops, key=lambda ab: ab[0] - ab[1])[:2]:
        body = [i for i in ins if a <= i[0] <= b]
        c = collections.Counter()
        for _, op, rest, pred in body:
            if op.startswith("IMAD.WIDE"): c["IMAD.WIDE"] += 1
            elif op.startswith("IMAD.HI"): c["IMAD.HI"] += 1
            elif op.startswith("IMAD.MOV") or op.startswith("IMAD.SHL") or op.startswith("IMAD.IADD"): c["IMAD.MOV/SHL/IADD (no multiply)"] += 1
            elif op.startswith("IMAD"): c["IMAD (32-bit)"] += 1
            elif op.startswith("IADD3"): c["IADD3"] += 1
            elif re.match(r"(LOP3|SEL|SHF|MOV|LEA|CS2R|PRMT|ISETP|UIADD|UMOV|ULOP|USHF|UISETP|ULEA|PLOP3)", op): c["other ALU"] += 1
            elif re.match(r"(LDG|STG|LDC|LDCU|LDL|STL|LD|ST)", op): c["memory"] += 1
            else: c["control/other:" + op.split(".")[0]] += 1
        print(f"  loop {hex(a)}..{hex(b)}: {len(body)} instructions")
        for k, v in c.most_common():
            print(f"    {k:36s} {v:6d}")
