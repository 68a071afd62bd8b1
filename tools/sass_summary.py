"""Per-kernel SASS mnemonic counts of libidf_b200.so (cuobjdump -sass): the evidence that the contraction kernels are
tcgen05 (UTCHMMA) / TMEM (LDTM, STTM) / TMA (UTMALDG, UTMASTG) code and contain no legacy HMMA. Writes
profiles/r02_sass_summary.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "image-diffusion_b200", "idf_b200", "libidf_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "MUFU.EX2", "MUFU.TANH",
        "HMMA", "IMMA", "FFMA", "F2FP", "REDG", "ATOMG"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void idf::", "").replace("idf::", "")
        cur = counts.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        cur["_total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                cur[k] += 1
lines = [f"SASS mnemonic counts per kernel, {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a)", ""]
hdr = f"{'kernel':70s} {'instrs':>7s} " + " ".join(f"{k:>12s}" for k in KEYS)
lines.append(hdr)
tot = collections.Counter()
for name, c in counts.items():
    lines.append(f"{name[:70]:70s} {c['_total']:7d} " + " ".join(f"{c[k]:12d}" for k in KEYS))
    tot.update(c)
lines.append(f"{'TOTAL':70s} {tot['_total']:7d} " + " ".join(f"{tot[k]:12d}" for k in KEYS))
text = "\n".join(lines) + "\n"
dst = os.path.join(ROOT, "profiles", sys.argv[1] if len(sys.argv) > 1 else "r02_sass_summary.txt")
open(dst, "w").write(text)
print(text)
