"""BASELINE.json configs[2]: 2,000 patients x {T1, T2} = 4,000 series with variable in-plane size (320-1024 px) and spacing
(0.30-0.95 mm) -> 0.3 mm isotropic middle slices of very different sizes (device-generated), localize + crop, sharded by series
over the ranks (size-balanced, no data-path collective), one gather of coordinates and crops at the end.

    python scripts/bench_config3.py [n_series=4000]                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_config3.py
"""
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from spine_vision_b200 import ops, pipeline, synthetic  # noqa: E402
from spine_vision_b200.cropping import LocalizationModel  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
shapes = synthetic.ragged_shapes(n, seed=0)
mine = pipeline.shard_series([h * w for h, w in shapes], world)[rank]
offs, total = ops.SlicePool.layout([shapes[i] for i in mine])
data = torch.empty(total, dtype=torch.float32, device=dev)
g = torch.Generator(device=dev).manual_seed(rank)
for k, i in enumerate(mine):  # smooth field + noise per slice
    h, w = shapes[i]
    low = torch.rand((1, 1, 10, 10), generator=g, device=dev)
    sl = torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=False)[0, 0] * 900
    sl += torch.rand((h, w), generator=g, device=dev) * 300
    data[offs[k] : offs[k] + h * w] = sl.reshape(-1)
model = LocalizationModel(synthetic.random_state_dict("base", seed=0), dev, dtype="bf16")
B = 256


def run_all():
    coords, crops = [], []
    for b0 in range(0, len(mine), B):
        sel = list(range(b0, min(b0 + B, len(mine))))
        sub_offs = torch.tensor([offs[k] for k in sel], dtype=torch.int64, device=dev)
        sub_hw = torch.tensor([shapes[mine[k]] for k in sel], dtype=torch.int32, device=dev).reshape(-1, 2)
        pool = ops.SlicePool(data, sub_offs, sub_hw, [shapes[mine[k]] for k in sel])
        out = pipeline.localize_and_crop(pool, model, (50, 20, 30, 30), (128, 128), (512, 512), None)
        coords.append(out.coords)
        crops.append(out.crops)
    c, k = torch.cat(coords), torch.cat(crops)
    return pipeline.gather_results(mine, c, k, n) if world > 1 else (c, k)


run_all()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
c, k = run_all()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    mb = sum(h * w for h, w in shapes) * 4 / 1e9
    print(f"config 3: {n} ragged series ({mb:.1f} GB of 0.3 mm slices, {min(shapes)}..{max(shapes)} px) on {world} GPU(s): {ms.item():.1f} ms = "
          f"{n / ms.item() * 1e3:.0f} series/s, {5 * n / ms.item() * 1e3:.0f} crops/s; gathered coords {tuple(c.shape)}, crops {tuple(k.shape)}, "
          f"finite {bool(torch.isfinite(c).all())}")
if world > 1:
    dist.destroy_process_group()
