"""Opcode histogram of the library's main kernels from `cuobjdump -sass` (evidence that the cubins are sm_100a code and of
what the hot loops are made of).   python tools/sass_histogram.py > profiles/<name>.txt"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "redback_b200/_lib/libredback_b200.so"
WANT = ("sn_kernel", "phot_kernel", "ag_cell_kernel", "kn_kernel", "spectra_kernel", "ext_apply_kernel")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
print(f"# {LIB}: cuobjdump -sass, architectures {arch}")
name, hist = None, None
funcs = {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        hist = funcs.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and hist is not None:
        hist[m.group(1).split(".")[0] + ("." + m.group(1).split(".")[1] if m.group(1).startswith(("LDS", "STS", "LDG", "STG")) and "." in m.group(1) else "")] += 1
for fn, h in funcs.items():
    if not any(w in fn for w in WANT):
        continue
    total = sum(h.values())
    top = ", ".join(f"{k} {v}" for k, v in h.most_common(14))
    print(f"{fn}: {total} instructions; {top}")
