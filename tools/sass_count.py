#!/usr/bin/env python
"""Static instruction mix of the step kernels from the built library's SASS (no GPU needed):
    python tools/sass_count.py [regex-on-mangled-name]
Prints, per matching kernel: registers (from build.log), instructions, FP ops, integer/address
ops, loads/stores.  The f32 vector kernels update 4 cells per thread (f64: 2)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("CHEMSIM_LBM_LIB") or os.path.join(ROOT, "chemsim_b200", "libchemsim_lbm.so")
pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else r"step_vec_kernelI[fd]Lb1ELb0ELi\dELb0E")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
name, ops = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1) if pat.search(m.group(1)) else None
        if name:
            ops[name] = collections.Counter()
        continue
    if name:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            ops[name][m.group(1)] += 1
FP = ("FADD", "FMUL", "FFMA", "FADD2", "FMUL2", "FFMA2", "DADD", "DMUL", "DFMA", "MUFU", "FSEL", "FSETP", "DSETP", "FMNMX", "FCHK")
INT = ("IADD3", "IMAD", "LEA", "SHF", "LOP3", "ISETP", "SEL", "MOV", "IADD", "UIADD3", "UIMAD", "ULEA", "USHF", "UMOV",
       "UISETP", "ULOP3", "PRMT", "CS2R", "HFMA2", "I2F", "F2I", "PLOP3", "USEL", "IABS", "UFLO", "R2UR", "S2R", "S2UR")
MEM = ("LDG", "STG", "LDC", "LDCU", "ULDC", "SHFL", "LDS", "STS")
for n, c in ops.items():
    tot = sum(c.values())
    fp = sum(c[k] for k in FP); it = sum(c[k] for k in INT); mem = sum(c[k] for k in MEM)
    short = re.sub(r"^.*?(step_\w+?_kernel)", r"\1", n)
    print(f"{short[:70]:70s} total {tot:5d}  fp {fp:4d}  int {it:4d}  mem {mem:3d}  other {tot-fp-it-mem:3d}")
