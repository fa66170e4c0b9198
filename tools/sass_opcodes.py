#!/usr/bin/env python
"""Per-kernel SASS opcode counts of libtgx.so (static), for profiles/rNN_sass_opcodes.txt.

    python tools/sass_opcodes.py [path/to/libtgx.so] > profiles/r02_sass_opcodes.txt

The built library is git-ignored; this file is the tracked evidence that the store path is TMA / 256-bit vector stores,
that the reductions are REDUX, and what the FP64 share of each kernel is."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "trajectory_generator_ros2_b200", "libtgx.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda names: subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()

kernels, cur, arch = collections.OrderedDict(), None, set()
for line in sass.splitlines():
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1

names = list(kernels)
pretty = dict(zip(names, demangle(names)))
KEYS = [("UTMASTG", lambda o: o.startswith("UTMASTG")), ("UTMALDG", lambda o: o.startswith("UTMALDG")),
        ("STG.256", lambda o: o.startswith("STG") and ".256" in o), ("STG.128", lambda o: o.startswith("STG") and ".128" in o),
        ("STG other", lambda o: o.startswith("STG") and ".256" not in o and ".128" not in o),
        ("LDG", lambda o: o.startswith("LDG")), ("LDS", lambda o: o.startswith("LDS")), ("STS", lambda o: o.startswith("STS")),
        ("REDUX", lambda o: o.startswith("REDUX") or o.startswith("CREDUX")),
        ("ATOM/RED", lambda o: o.startswith("ATOM") or o.startswith("RED.") or o.startswith("REDG")),
        ("DFMA", lambda o: o.startswith("DFMA")), ("DMUL", lambda o: o.startswith("DMUL")), ("DADD", lambda o: o.startswith("DADD")),
        ("DSETP", lambda o: o.startswith("DSETP")), ("MUFU", lambda o: o.startswith("MUFU")),
        ("BAR", lambda o: o.startswith("BAR")), ("SYNCS/mbar", lambda o: o.startswith("SYNCS")), ("total", lambda o: True)]
print("libtgx.so: %d kernels, cubins for %s; static SASS opcode counts per kernel (cuobjdump -sass)" % (len(kernels), ", ".join(sorted(arch))))
print("%-118s" % "kernel" + "".join("%10s" % k for k, _ in KEYS))
tot = collections.Counter()
for n, c in kernels.items():
    p = re.sub(r"\(anonymous namespace\)::|tgx::", "", pretty.get(n, n))
    p = p.replace("(int)", "").replace("(bool)", "").replace("void ", "")
    p = re.sub(r"\(.*$", "", p)
    row = [sum(v for o, v in c.items() if f(o)) for _, f in KEYS]
    for (k, _), v in zip(KEYS, row):
        tot[k] += v
    print("%-118s" % p[:117] + "".join("%10d" % v for v in row))
print("%-118s" % "ALL KERNELS" + "".join("%10d" % tot[k] for k, _ in KEYS))
