"""K1 / K3 / K4 on the bench shapes (256 slices of 1195^2, 1280 crops): CUDA-event timings, or one launch each for ncu (REPS=0)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from spine_vision_b200 import ops, pipeline, synthetic, volumes  # noqa: E402

dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
base = [synthetic.make_iso_slice(s, 1195, 1195) for s in range(8)]
pool = ops.SlicePool.from_numpy([base[i % 8] for i in range(B)], dev)
xy = torch.from_numpy(synthetic.make_coords(B, seed=1)).to(dev).reshape(B * 5, 2).contiguous()
dpx = pipeline.mm_to_pixels((50, 20, 30, 30), (0.3, 0.3))
idx = torch.arange(B, dtype=torch.int32).repeat_interleave(5).contiguous().to(dev)
delta = torch.tensor([dpx] * (B * 5), dtype=torch.int32).to(dev)
out = torch.empty((B * 5, 128, 128), dtype=torch.uint8, device=dev)
out2 = torch.empty((B * 5, 256, 256), dtype=torch.uint8, device=dev)
planes = torch.empty((B, 512, 512), dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def k1():
    ops.normalize_resize(pool, (512, 512), out=planes)


def k3():
    ops.crop_resample(pool, idx, xy, delta, (dpx[2] + dpx[3], dpx[0] + dpx[1]), (128, 128), (256, 256), out=out, out2=out2)


P = B * 5 // 2
k4_t2 = torch.arange(0, P, dtype=torch.int32, device=dev)
k4_t1 = torch.arange(P, 2 * P, dtype=torch.int32, device=dev)
k4_out = torch.empty((P, 3, 256, 256), dtype=torch.float32, device=dev)


def k4():
    ops.classifier_input(out2, k4_t2, k4_t1, out=k4_out)


norm_out = torch.empty(pool.data.numel(), dtype=torch.uint8, device=dev)


def norm():
    ops.normalize_u8(pool, out=norm_out)


# K0 and the fused K0 + K1 call of the end-to-end path, on the source planes of config-1 volumes
plans = [volumes.plan_midplane(*synthetic.make_volume(s)) for s in range(8)]
pv = volumes.PinnedVolumes.from_plans([plans[i % 8] for i in range(B)])
k0_vols, k0_desc = pv.host.to(dev), pv.chunk_descs(0, B).to(dev)
k0_pool = ops.SlicePool(torch.empty(max(pv.out_total, 4), dtype=torch.float32, device=dev), torch.tensor(pv.out_offs, dtype=torch.int64).to(dev),
                        torch.tensor(pv.shapes, dtype=torch.int32).reshape(-1, 2).to(dev), list(pv.shapes))


def k0():
    ops.midplane_resample_into(k0_vols, k0_desc, k0_pool)


def k01():
    ops.midplane_normalize_resize(k0_vols, k0_desc, k0_pool, (512, 512), out=planes)


for name, fn, nbytes in (("K0 midplane resample", k0, B * (2 * 512 * 512 * 4 + 1195 * 1195 * 4)),
                         ("K0+K1 fused", k01, B * (2 * 512 * 512 * 4 + 2 * 1195 * 1195 * 4 + 512 * 512)),
                         ("normalize_u8 (no resize)", norm, B * 1195 * 1195 * 5), ("K1 normalize+resize", k1, B * (1195 * 1195 * 4 + 512 * 512)), ("K3 crop+resample", k3, B * 5 * 269120),
                         ("K4 classifier input", k4, P * 65536 * 14)):
    if REPS == 0:
        fn(); torch.cuda.synchronize(); continue
    fn(); fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(REPS):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name:22s} B={B}: {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:8.1f} GB/s algorithmic (L2 flushed between reps)", flush=True)
