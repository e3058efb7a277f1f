"""Per-kernel SASS evidence of the built library: tcgen05 / TMA / TMEM mnemonics and spills.
    python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit-lens_b200", "vitlens_b200", "libvitlens_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "SYNCS", "USETMAXREG", "MUFU", "FFMA2", "HMMA", "STL", "LDL"]
cur = None
counts = collections.OrderedDict()
arch = ""
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["instructions"] += 1
        for k in keys:
            if op.startswith(k):
                counts[cur][k] += 1
print(f"# cuobjdump -sass {os.path.basename(so)}  (arch {arch}); counts of SASS mnemonics per kernel")
print(f"# {'kernel':70s} {'instr':>6s} " + " ".join(f"{k:>9s}" for k in keys))
for name, c in counts.items():
    print(f"{name[:72]:72s} {c['instructions']:6d} " + " ".join(f"{c[k]:9d}" for k in keys))
