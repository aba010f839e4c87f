"""writes profiles/r2_spmv_packed_sass.txt: opcode counts and the TMA / mbarrier sites of the shipped bulk-copy kernels, from
`cuobjdump -sass` of the built library (no GPU needed)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "del-fem_b200", "lib", "libdelfem_b200.so")
KERNELS = [
    ("_Z13k_spmv_packedILi3ELb0ELi8EE", "k_spmv_packed<3,false,8> (single domain, 8-row tiles: the headline kernel)"),
    ("_Z13k_spmv_packedILi3ELb0ELi16EE", "k_spmv_packed<3,false,16> (single domain, 16-row tiles / 2 lanes per row: sparse rows, e.g. the cloth)"),
    ("_Z13k_spmv_packedILi3ELb1ELi8EE", "k_spmv_packed<3,true,8> (peer halo: boundary tiles wait for the neighbours, ghost gathers are L2-coherent loads)"),
    ("_Z11k_spmv_ringILi3EE", "k_spmv_ring<3> (variant 4: canonical arrays)"),
    ("_Z11k_ilu_sweepILi3ELb1ELi512EE", "k_ilu_sweep<3,forward,512> (ILU forward sweep: one bulk copy per 8-row record)"),
]
KEEP = ("UBLKCP", "SYNCS", "UTMA", "UTCMMA")
WANT = ("ATOM", "CCTL", "DFMA", "DADD", "DMUL", "LDG", "LDS", "MEMBAR", "SHFL", "SYNCS", "UBLKCP", "RED", "BAR")

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
funcs = {}
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur is not None:
        funcs[cur].append(line)
out = ["# cuobjdump -sass del-fem_b200/lib/libdelfem_b200.so — excerpt for the shipped bulk-copy kernels (sm_100a); tools/sass_excerpt.py",
       "# The matrix is streamed by TMA bulk copies (UBLKCP = cp.async.bulk global -> shared) whose completion is tracked by mbarriers (SYNCS.*);",
       "# no UTMALDG (1-D bulk stream, no tensor map) and no UTCMMA/tcgen05 (no dense contraction: 0.24 flop/B).", ""]
for prefix, title in KERNELS:
    name = next((k for k in funcs if k.startswith(prefix)), None)
    out.append(f"## {title}")
    if name is None:
        out.append("(not found)")
        continue
    ops = collections.Counter()
    sites = []
    for line in funcs[name]:
        m = re.match(r"\s*/\*([0-9a-f]{4})\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(2).strip()
        op = next((t for t in ins.split() if not t.startswith("@")), "")
        if op.startswith(WANT):
            ops[op] += 1
        if any(k in ins for k in KEEP):
            sites.append(f"  /*{m.group(1)}*/ {ins} ;")
    out.append("opcode counts: " + ", ".join(f"{k} {v}" for k, v in sorted(ops.items())))
    out.extend(sites[:12])
    out.append("")
open(os.path.join(ROOT, "profiles", "r2_spmv_packed_sass.txt"), "w").write("\n".join(out))
print("\n".join(out[:14]))
