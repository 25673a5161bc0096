#!/bin/bash
# developer check: local-memory instructions of reproj_eval_kernel<0,1,54> and where they sit relative to the Gram (DMMA) loop
d=$(mktemp -d); cd $d
cuobjdump -xelf vg_eval_eucm ${1:-/root/repo/visgeom_b200/libvisgeom_b200.so} >/dev/null 2>&1
nvdisasm -c vg_eval_eucm.sm_100a.cubin > a.sass
python3 - <<'PY'
import re
s=open("a.sass").read()
for p in re.split(r"\n\s*\.text\.", s):
    name=p.split(":")[0][:120]
    if "reproj_eval_kernelILi0ELi1ELi54" in name:
        lines=p.splitlines()
        idx=[i for i,l in enumerate(lines) if re.search(r"\b(STL|LDL)\b",l)]
        d=[i for i,l in enumerate(lines) if "DMMA" in l]
        b=[i for i,l in enumerate(lines) if "BAR.SYNC" in l]
        print("lines", len(lines), "STL/LDL", len(idx), idx)
        print("DMMA", d[0], d[-1], "BAR", b)
PY
rm -rf $d
