#!/usr/bin/env python
"""Opcode histogram per kernel of the in-tree library (cuobjdump -sass; runs without a GPU):
    python profiles/sass_histogram.py > profiles/r2_sass.txt
Shows which sm_100a features each kernel's machine code really uses: UBLKCP (1-D bulk TMA, cp.async.bulk), SYNCS
(mbarrier), FFMA2 / FMUL2 / FADD2 (packed fp32 pairs), UCGABAR / CCTL-free cluster barriers and remote shared-memory
stores (ST.E with mapa'd addresses show as plain ST / ATOM after MAPA), REDUX / MATCH / VOTE / SHFL (warp collectives)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "control-gic_b200", "libcgic_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
kern, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur).split("(")[0].replace("cgic::", "").replace("void ", "")
        kern[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)(\.[A-Z0-9_.]+)?", line)
    if m and cur:
        kern[cur][m.group(1)] += 1
KEYS = ["UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "MAPA", "UCGABAR_ARV", "UCGABAR_WAIT", "ATOMS", "ATOMG", "RED", "REDUX", "MATCH", "VOTE", "SHFL",
        "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR", "MUFU", "DADD", "DFMA"]
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: architectures {arch}")
W = {k: max(6, len(k)) for k in KEYS}
print(f"# {'kernel':58s} {'instr':>6s} " + " ".join(f"{k:>{W[k]}s}" for k in KEYS))
for name, c in kern.items():
    tot = sum(c.values())
    print(f"{name[:60]:60s} {tot:6d} " + " ".join(f"{c.get(k, 0):{W[k]}d}" for k in KEYS))
print("# UTMALDG (tensor-map TMA), UTCxMMA / TCGEN05 (tensor cores), HMMA:",
      {k: sum(c.get(k, 0) for c in kern.values()) for k in ("UTMALDG", "UTCHMMA", "UTCQMMA", "HMMA", "IMMA")},
      "-- none: the path is D = 4 fp32 search + integer / bit work; tables are staged with 1-D bulk copies")
