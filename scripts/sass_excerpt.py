#!/usr/bin/env python
"""profiles/r02_sass_excerpt.md: instruction mix of the hot kernels in the built libfe_b200.so (cuobjdump -sass),
with the mnemonics that prove TMA bulk copies (UBLKCP), mbarrier transaction counting (SYNCS), cp.async (LDGSTS)
and the FP64 pipe (DFMA / DMUL / DADD).  No tensor-core mnemonics are expected: nothing on the path is a dense
contraction."""
import collections
import re
import subprocess
import sys

LIB = "finite_elements_b200/libfe_b200.so"
KERNELS = [("k_assemble_fanILi0ELb1", "k_assemble_fan<0,1> (plane stress, 4-byte records)"),
           ("k_assemble_fanILi2ELb1", "k_assemble_fan<2,1> (magnetic)"),
           ("k_spmv_streamILb1ELb0", "k_spmv_stream<1,0> (PCG SpMV, 2 DOF per node)"),
           ("k_pcg_persistILb1", "k_pcg_persist<true> (whole-solve cooperative kernel, multi-GPU)"),
           ("k_tet_assemble_stagedILb0", "k_tet_assemble_staged<false> (tetrahedra)"),
           ("k_spmmI", "k_spmm (modal block product, first instance)")]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", sass)
out = [__doc__.strip().replace("\n", " "), ""]
for key, title in KERNELS:
    body = next((f for f in funcs if f.split("\n", 1)[0].find(key) >= 0), None)
    if body is None:
        out.append(f"## {title}\n\nnot found in {LIB}\n")
        continue
    ops = collections.Counter()
    for line in body.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(1).split(".")[0]] += 1
    total = sum(ops.values())
    out.append(f"## {title}\n\n`{body.splitlines()[0].strip()}` — {total} SASS instructions\n")
    out.append(" ".join(f"{k}:{v}" for k, v in ops.most_common(28)) + "\n")
    proof = [l.strip() for l in body.splitlines() if re.search(r"UBLKCP|SYNCS\.(ARRIVE|PHASECHK)|LDGSTS|UTMA|CCTL", l)]
    seen, keep = set(), []
    for l in proof:
        mn = re.sub(r"/\*[0-9a-f]+\*/", "", l).strip().split(" ")[0:2]
        k = " ".join(mn)
        if k not in seen:
            seen.add(k)
            keep.append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/", "", l))
    if keep:
        out.append("```\n" + "\n".join(keep[:10]) + "\n```\n")
tc = len(re.findall(r"UTCMMA|UTCHMMA|LDTM|HMMA|UTMALDG", sass))
out.append(f"Tensor-core / tensor-map mnemonics in the whole library (UTCMMA, LDTM, HMMA, UTMALDG): {tc}.")
open(sys.argv[1] if len(sys.argv) > 1 else "profiles/r02_sass_excerpt.md", "w").write("\n".join(out) + "\n")
print("\n".join(out)[:3000])
