"""BASELINE.json configs[3], the crop-only microbench: 50,000 precomputed IVD coordinates over 10,000 isotropic slices
(1195 x 1195 float32, generated on the device: 57 GB resident), crop_delta_mm 50/20/30/30, 128 x 128 crops + the 256 x 256
classifier-size resample -- K3 alone, one launch, CUDA events.

    python scripts/bench_crops.py [n_slices]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from spine_vision_b200 import ops, pipeline, synthetic  # noqa: E402

dev = "cuda:0"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
H = W = 1195
per = (H * W + 3) // 4 * 4
g = torch.Generator(device=dev).manual_seed(0)
data = torch.empty(n * per, dtype=torch.float32, device=dev)
low = torch.rand((n, 1, 12, 12), generator=g, device=dev)
for i0 in range(0, n, 64):  # smooth field + noise, 64 slices at a time (keeps the temporary small)
    i1 = min(n, i0 + 64)
    field = torch.nn.functional.interpolate(low[i0:i1], size=(H, W), mode="bilinear", align_corners=False)[:, 0]
    sl = (field * 900 + torch.rand((i1 - i0, H, W), generator=g, device=dev) * 300).clamp_(min=0)
    data[i0 * per : i1 * per].view(i1 - i0, per)[:, : H * W] = sl.view(i1 - i0, -1)
offs = (torch.arange(n, dtype=torch.int64, device=dev) * per).contiguous()
hw = torch.tensor([[H, W]], dtype=torch.int32, device=dev).repeat(n, 1).contiguous()
pool = ops.SlicePool(data, offs, hw, [(H, W)] * n)
xy = torch.from_numpy(synthetic.make_coords(n, seed=0)).to(dev).reshape(n * 5, 2).contiguous()
dpx = pipeline.mm_to_pixels((50, 20, 30, 30), (0.3, 0.3))
idx = torch.arange(n, dtype=torch.int32, device=dev).repeat_interleave(5).contiguous()
delta = torch.tensor([dpx], dtype=torch.int32, device=dev).repeat(n * 5, 1).contiguous()
out = torch.empty((n * 5, 128, 128), dtype=torch.uint8, device=dev)
out2 = torch.empty((n * 5, 256, 256), dtype=torch.uint8, device=dev)


def run():
    ops.crop_resample(pool, idx, xy, delta, (dpx[2] + dpx[3], dpx[0] + dpx[1]), (128, 128), (256, 256), out=out, out2=out2)


run(); torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
crops = n * 5
print(f"config 4: {crops} crops from {n} resident slices ({data.numel() * 4 / 1e9:.1f} GB) in {ms:.2f} ms = {crops / ms / 1e3:.2f} M crops/s, "
      f"{crops * 269120 / ms / 1e6:.0f} GB/s algorithmic; non-zero crops: {int((out.view(crops, -1).max(dim=1).values > 0).sum())}")
