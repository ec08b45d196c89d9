"""Per-kernel counts of the Blackwell-native SASS opcodes in libspine_b200.so -> profiles/sass_opcodes.txt

    python scripts/sass_opcodes.py

UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / .st (TMEM), UTMALDG / UTMASTG = TMA tensor
load / store, UTCBAR = tcgen05.commit, FFMA2 = packed fp32 FMA, SYNCS = mbarrier ops, ACQBULK = griddepcontrol.wait.
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
so = ROOT / "spine_vision_b200" / "libspine_b200.so"
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "FFMA2", "FHFMA", "HFMA2", "SHFL", "SYNCS", "ACQBULK", "UCGABAR", "LDSM", "STSM"]
sass = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
counts: dict = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
    if not m:
        continue
    op = m.group(1)
    counts[cur]["_total"] += 1
    base = op.split(".")[0]
    if base in OPS:
        counts[cur][base] += 1
    if base == "UTCHMMA" and ".2CTA" in op:
        counts[cur]["UTCHMMA.2CTA"] += 1
rows = []
tot = collections.Counter()
for fn, c in counts.items():
    name = demangle(fn)
    name = name.replace("(int)", "").replace("(svb::GemmMode)", "mode ")
    name = (name.split(">(")[0] + ">") if ">(" in name else re.sub(r"\(.*", "", name)
    name = name.replace("void svb::", "").replace("svb::", "").replace("(anonymous namespace)::", "")
    rows.append((name, c))
    tot.update(c)
out = [f"# cuobjdump -sass {so.relative_to(ROOT)}  (sm_100a; {len(rows)} kernels, {tot['_total']} instructions)",
       "# " + " ".join(f"{o:>12s}" for o in OPS) + "  kernel"]
for name, c in sorted(rows, key=lambda r: r[0]):
    if not any(c[o] for o in OPS):
        continue
    out.append("  " + " ".join(f"{c[o]:12d}" for o in OPS) + "  " + name)
out.append("  " + " ".join(f"{tot[o]:12d}" for o in OPS) + "  TOTAL")
text = "\n".join(out) + "\n"
dst = ROOT / "profiles" / (sys.argv[1] if len(sys.argv) > 1 else "sass_opcodes.txt")
dst.write_text(text)
print(text[-1500:])
print("wrote", dst)
