"""Summarise an `ncu --set full` capture of scripts/prof_layers.py (one launch per layer kernel) into the two files the
bench and DESIGN.md cite:  profiles/<tag>_ncu_full_layers.csv  and  profiles/<tag>_gemm_traffic.json.

    python scripts/summarise_ncu_layers.py gpurun_out/prof_layers_r01e.ncu-rep r01e
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

rep, tag = sys.argv[1], sys.argv[2]
MB = int(sys.argv[3]) if len(sys.argv) > 3 else 64  # micro-batch the capture was taken at (scripts/prof_layers.py MB 0)
ROOT = Path(__file__).resolve().parent.parent
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
idx = [hdr.index(k) for k in KEEP if k in hdr]
out = ROOT / "profiles" / f"{tag}_ncu_full_layers.csv"
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in data:
        w.writerow([r[i] for i in idx])
print("wrote", out)


def col(r, k):
    return float(r[hdr.index(k)].replace(",", ""))


def mb(r, k):  # ncu prints Mbyte / Gbyte depending on the run: normalise through the unit row
    u = units[hdr.index(k)].lower()
    scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    return col(r, k) * scale


gemms = [r for r in data if "gemm_kernel" in r[hdr.index("Kernel Name")] or "mlp_fused_kernel" in r[hdr.index("Kernel Name")]]
# scripts/prof_layers.py <MB> 0 launches the default forward's MLP kernels in this order
names = ["stage0 C=128 fused MLP (fc1+LN+gelu+fc2+resid)", "stage1 C=256 fused MLP (fc1+LN+gelu+fc2+resid)",
         "stage2 C=512 fc1+LN+gelu", "stage2 C=512 fc2+resid", "stage3 C=1024 fc1+LN+gelu", "stage3 C=1024 fc2+resid"]
launches = [3, 3, 27, 27, 3, 3]
assert len(gemms) == len(names), f"expected {len(names)} MLP launches in the capture, found {len(gemms)}"
per = []
total = 0.0
for r, n, k in zip(gemms, names, launches):
    b = mb(r, "dram__bytes_read.sum") + mb(r, "dram__bytes_write.sum")
    per.append({"layer": n, "dram_bytes": b, "time_us": col(r, "gpu__time_duration.sum"),
                "tensor_pipe_active_pct": col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                "launches_per_micro_batch": k, "dram_read_bytes": mb(r, "dram__bytes_read.sum"), "dram_write_bytes": mb(r, "dram__bytes_write.sum")})
    total += b * k
js = {"source": f"ncu --set full --clock-control none, scripts/prof_layers.py {MB} 0 (micro-batch {MB} shapes), profiles/{tag}_ncu_full_layers.csv",
      "per_launch": per, "micro_batch": MB, "gemm_dram_bytes_per_micro_batch": total,
      "note": "the 3 downsample GEMMs of a forward (0.6 % of the flops) are not in the capture"}
(ROOT / "profiles" / f"{tag}_gemm_traffic.json").write_text(json.dumps(js, indent=1))
print("wrote", ROOT / "profiles" / f"{tag}_gemm_traffic.json", f"{total / 1e9:.2f} GB per micro-batch of {MB}")
