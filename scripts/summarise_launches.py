"""Per-kernel shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list of the bench command.

    python scripts/summarise_launches.py gpurun_out/launches_r01f.csv r01f "<command that was profiled>"

writes profiles/<tag>_launches.csv.gz (the raw list) and profiles/<tag>_launches_summary.txt.
Per-launch times under ncu are cold-cache and serialised: compare SHARES with the CUDA-event shares of the plain run.
"""
import csv
import gzip
import re
import shutil
import sys
from collections import defaultdict
from pathlib import Path

src, tag = Path(sys.argv[1]), sys.argv[2]
cmd = sys.argv[3] if len(sys.argv) > 3 else ""
ROOT = Path(__file__).resolve().parent.parent
lines = [l for l in src.read_text(errors="replace").splitlines() if not l.startswith("==")]
rows = list(csv.reader(lines))
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0.0, 0, 1e30, 0.0])
for r in rows[1:]:
    if len(r) <= iv:
        continue
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("svb::", "")
    v = float(r[iv].replace(",", ""))
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(r[iu], 1e-6)
    a = agg[name]
    a[0] += ms; a[1] += 1; a[2] = min(a[2], ms); a[3] = max(a[3], ms)
total = sum(a[0] for a in agg.values())
out = [f"# {cmd}", f"# {sum(a[1] for a in agg.values())} launches captured, {total:.1f} ms total (cold-cache, serialised: compare SHARES)",
       f"{'total_ms':>10} {'share':>6} {'n':>6} {'min_ms':>8} {'max_ms':>8}  kernel"]
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    out.append(f"{a[0]:10.3f} {100 * a[0] / total:5.1f}% {a[1]:6d} {a[2]:8.4f} {a[3]:8.4f}  {name}")
groups = defaultdict(float)
for name, a in agg.items():
    g = ("gemm" if name.startswith("gemm_kernel") or name.startswith("mlp_fused_kernel")
         else "dwconv_ln" if name.startswith("dwconv_") or name.startswith("ln_stat_finalize") else "k1" if name.startswith("k1_")
         else "k3" if name.startswith("k3_") else "k4" if name.startswith("k4_") else name.split("<")[0])
    groups[g] += a[0]
out.append("# by class: " + ", ".join(f"{g} {100 * v / total:.1f} %" for g, v in sorted(groups.items(), key=lambda kv: -kv[1])))
(ROOT / "profiles" / f"{tag}_launches_summary.txt").write_text("\n".join(out) + "\n")
with open(src, "rb") as fi, gzip.open(ROOT / "profiles" / f"{tag}_launches.csv.gz", "wb") as fo:
    shutil.copyfileobj(fi, fo)
print("\n".join(out))
