"""One small pass over EVERY kernel of libspine_b200.so, meant to run under compute-sanitizer (closed on the round-1 GPU pool:
"runs under it have left GPUs needing a reset"; the script also runs plain as a walk over every guarded edge):

    compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_gpu.py          # out-of-bounds / misaligned
    compute-sanitizer --tool racecheck --error-exitcode 9 python tests/sanitize_gpu.py k1k3    # shared-memory hazards (K1/K3)

Ragged shapes on purpose (odd widths, a slice narrower than the crop box, a tail micro-batch), so that every guarded edge is
walked: K0 -> K1 -> ConvNeXt-base localizer (bf16, micro-batch 2 + 1) -> K3 (horizontal and rotated) -> K4, normalize_u8,
the centre-crop fallback, and a ConvNeXt-V2 forward for the GRN kernels.  Prints one line per stage; results are only checked
for being finite -- parity is the test suite's job."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from spine_vision_b200 import cropping, ops, pipeline, synthetic  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = "cuda:0"
slices = [synthetic.make_iso_slice(1, 333, 517), synthetic.make_iso_slice(2, 640, 650), synthetic.make_iso_slice(3, 150, 121)]
pool = ops.SlicePool.from_numpy(slices, dev)

planes = ops.normalize_resize(pool, (512, 512))
u8 = ops.normalize_u8(pool)
torch.cuda.synchronize()
print("K1 normalize+resize, normalize_u8: ok", tuple(planes.shape))

xy = torch.from_numpy(synthetic.make_coords(3, seed=1, border_frac=0.4, hw=(333, 517))).to(dev)
for mode in ("horizontal", "rotated"):
    crops, crops2, _ = pipeline.crop_levels(pool, xy, (50, 20, 30, 30), None, (128, 128), (256, 256), mode=mode)
    torch.cuda.synchronize()
    print(f"K3 {mode}: ok", tuple(crops.shape), tuple(crops2.shape))
fb = pipeline.localize_and_crop(pool, None, (55, 15, 17.5, 20), (256, 256), (512, 512), None)
torch.cuda.synchronize()
print("K3 centre fallback (float64 centres): ok", tuple(fb.crops.shape))

flat = crops2.reshape(-1, 256, 256).contiguous()
i2 = torch.tensor([0, 1, -1], dtype=torch.int32, device=dev)
i1 = torch.tensor([5, -1, 6], dtype=torch.int32, device=dev)
for dt in (torch.float32, torch.bfloat16):
    k4 = ops.classifier_input(flat, i2, i1, dtype=dt)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(k4.float()).all())
print("K4 classifier input: ok")
if what == "k1k3":
    sys.exit(0)

vols, sps, dirs = [], [], []
for k, (h, w, sp) in enumerate(((331, 347, 0.61), (320, 400, 0.9))):
    v, _, d = synthetic.make_volume(70 + k, 11 + 2 * k, h, w, (sp, sp, 4.0))
    vols.append(v); sps.append((sp, sp, 4.0)); dirs.append(d)
model = cropping.LocalizationModel(synthetic.random_state_dict("base", seed=0), dev, dtype="bf16", micro_batch=2)
out = pipeline.localize_and_crop_volumes(vols + vols[:1], sps + sps[:1], dirs + dirs[:1], model, dev, crop_delta_mm=(50, 20, 30, 30),
                                         crop_size=(128, 128))
torch.cuda.synchronize()
c = out.coords.cpu().numpy()
assert np.isfinite(c).all() and c.shape == (3, 5, 2)
print("K0 -> K1 -> ConvNeXt-base (bf16, micro-batch 2 + 1) -> K3: ok")

m16 = cropping.LocalizationModel(synthetic.random_state_dict("base", seed=0), dev, dtype="fp16", micro_batch=64)
c16 = m16.predict_u8(planes)
torch.cuda.synchronize()
assert bool(torch.isfinite(c16).all())
print("ConvNeXt-base fp16: ok")
if what == "all":
    from oracle.convnext import make_model  # test infrastructure: only the V2 random weights come from here

    v2 = cropping.LocalizationModel(make_model("v2_tiny", seed=0).state_dict(), dev, dtype="bf16", micro_batch=2)
    cv2_ = v2.predict_u8(planes)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(cv2_).all())
    print("ConvNeXt-V2 tiny (GRN kernels): ok")
print("sanitize_gpu: done")
