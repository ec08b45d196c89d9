"""Freeze outputs of the reference's OWN functions into ``tests/golden/``.

TEST INFRASTRUCTURE ONLY; runs in the build container only (needs
``/root/reference`` through ``oracle/ref_shim.py``):

    python -m oracle.make_golden

Inputs are named by seed (``spine_vision_b200.synthetic``), outputs are what the
unmodified reference returned here with Pillow 12.2.0 / OpenCV 4.13.0 /
NumPy 2.3.5 / torch 2.11.0 (the container's versions define parity).
"""

from __future__ import annotations

import tempfile
import warnings
from pathlib import Path

import numpy as np
import torch

from oracle import ref_shim
from oracle.convnext import make_model
from spine_vision_b200 import synthetic

GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"

# (seed, H, W): config-1 size, a SPIDER size (test_middle_slice.ipynb:237), a mixed up/down case,
# and a strongly anisotropic one
K1_CASES = [(0, 1195, 1195), (1, 640, 650), (2, 350, 420), (3, 1700, 560)]
CROP_SERIES = [(10, 1195, 1195), (11, 1040, 1040), (12, 640, 650), (13, 400, 380)]
CROP_DELTAS_MM = [(50, 20, 30, 30), (55, 15, 17.5, 20)]
MODEL_SLICES = [(20, 1195, 1195), (21, 640, 650)]


def golden_rotated(ref) -> None:
    """Rotated crop mode (cropping.py:172-313) through the reference's own CropContext / get_rotation_angles,
    plus the notebook's recorded angles (notebooks/compare_crop_modes.ipynb:87-161)."""
    out = {}
    for seed, h, w in CROP_SERIES:
        img = synthetic.make_iso_slice(seed, h, w)
        xy = synthetic.make_coords(2, seed=100 + seed, border_frac=0.3, hw=(h, w))
        xy[1, 0] = (0.02, 0.03)      # corner: replicated border inside the rotated box
        xy[1, 4] = (0.97, 0.985)
        out[f"xy_{seed}_{h}_{w}"] = xy
        for di, dmm in enumerate(CROP_DELTAS_MM):
            dpx = ref.mm_to_pixels(dmm, (0.3, 0.3))
            for boost in (1.0, 2.0):
                res = np.zeros((2, 5, 128, 128), dtype=np.uint8)
                ang = np.zeros((2, 5), dtype=np.float64)
                for s in range(2):
                    locs = {i: (float(xy[s, i, 0]), float(xy[s, i, 1])) for i in range(5)}
                    ctx = ref.CropContext(image=img, ivd_locations=locs, crop_size=(128, 128), crop_delta_px=dpx, mode="rotated",
                                          last_disc_angle_boost=boost)
                    for i in range(5):
                        res[s, i] = ctx.crop(i)
                        ang[s, i] = ctx.rotation_angles[i]
                out[f"crops_{seed}_{h}_{w}_d{di}_b{int(boost)}"] = res
                out[f"angles_{seed}_{h}_{w}_d{di}_b{int(boost)}"] = ang
    # the notebook's printed coordinates (4 dp) and angles for a 1040x1040 slice, boost 2
    nb_locs = {0: (0.5102, 0.2838), 1: (0.4773, 0.6622)}
    out["notebook_angles_2pt"] = np.array([ref.get_rotation_angles(nb_locs, (1040, 1040), 2.0)[i] for i in (0, 1)])
    np.savez_compressed(GOLDEN / "k3_rotated.npz", **out)


INT_SERIES = [(12, 640, 650), (13, 400, 380)]
INT_TYPES = {"int16": np.int16, "uint16": np.uint16, "uint8": np.uint8}


def int_slice(seed: int, h: int, w: int, name: str) -> np.ndarray:
    """The synthetic slice in an integer pixel type, as ``sitk.GetArrayFromImage`` hands an MR volume to the crop step:
    int16 shifted below zero (CT-like offsets occur), uint16 as is, uint8 scaled into 0..255."""
    img = synthetic.make_iso_slice(seed, h, w)
    if name == "int16":
        return (np.rint(img) - 300).astype(np.int16)
    if name == "uint16":
        return np.rint(img * 20).astype(np.uint16)
    return np.rint(img * (255.0 / max(float(img.max()), 1.0))).astype(np.uint8)


def golden_rotated_int(ref) -> None:
    """Rotated crop mode on INTEGER-typed slices through the reference's own CropContext: cv2.warpAffine then works in the
    slice's pixel type (rounds every warped value back to int16 / uint16; fixed-point weights for uint8) before
    normalize_to_uint8 (cropping.py:292-311).  ADVICE r01 (svb_resize.cu:649)."""
    out = {}
    dpx = ref.mm_to_pixels(CROP_DELTAS_MM[0], (0.3, 0.3))
    for seed, h, w in INT_SERIES:
        xy = synthetic.make_coords(2, seed=300 + seed, border_frac=0.3, hw=(h, w))
        xy[1, 0] = (0.02, 0.03)
        xy[1, 4] = (0.97, 0.985)
        out[f"xy_{seed}_{h}_{w}"] = xy
        for name in INT_TYPES:
            img = int_slice(seed, h, w, name)
            res = np.zeros((2, 5, 128, 128), dtype=np.uint8)
            for s in range(2):
                locs = {i: (float(xy[s, i, 0]), float(xy[s, i, 1])) for i in range(5)}
                ctx = ref.CropContext(image=img, ivd_locations=locs, crop_size=(128, 128), crop_delta_px=dpx, mode="rotated")
                for i in range(5):
                    res[s, i] = ctx.crop(i)
            out[f"crops_{seed}_{h}_{w}_{name}"] = res
    np.savez_compressed(GOLDEN / "k3_rotated_int.npz", **out)


SMALL_SERIES = [(30, 150, 121), (31, 90, 300), (32, 200, 234), (33, 260, 180), (34, 64, 64), (35, 7, 500)]


def golden_small(ref) -> None:
    """Slices SMALLER than the crop box (the box is clipped to the whole slice on one or both axes, cropping.py:335-341) and
    degenerate aspect ratios (a 7 x 500 strip letterboxes to 2 rows), both crop modes, crop sizes 128 and the config default
    256 (config.py:47-48), through the reference's own CropContext."""
    out = {}
    for seed, h, w in SMALL_SERIES:
        img = synthetic.make_iso_slice(seed, h, w)
        xy = synthetic.make_coords(2, seed=200 + seed, border_frac=0.3, hw=(h, w))
        xy[1, 0] = (0.0, 0.0)
        xy[1, 4] = (0.99999, 0.99999)
        out[f"xy_{seed}_{h}_{w}"] = xy
        for di, dmm in enumerate(CROP_DELTAS_MM):
            dpx = ref.mm_to_pixels(dmm, (0.3, 0.3))
            for cs in (128, 256):
                for mode in ("horizontal", "rotated"):
                    if cs == 256 and (mode == "rotated" or di == 0):
                        continue  # 256 x 256 (the config default): horizontal mode with the default deltas only (fixture size)
                    res = np.zeros((2, 5, cs, cs), dtype=np.uint8)
                    for s in range(2):
                        locs = {i: (float(xy[s, i, 0]), float(xy[s, i, 1])) for i in range(5)}
                        ctx = ref.CropContext(image=img, ivd_locations=locs, crop_size=(cs, cs), crop_delta_px=dpx, mode=mode)
                        for i in range(5):
                            res[s, i] = ctx.crop(i)
                    out[f"crops_{seed}_{h}_{w}_d{di}_c{cs}_{mode}"] = res
    np.savez_compressed(GOLDEN / "k3_small.npz", **out)


def main() -> None:
    import sys

    ref = ref_shim.load()
    if "--rotated-int-only" in sys.argv:
        golden_rotated_int(ref)
        for f in sorted(GOLDEN.glob("k3_rotated_int.npz")):
            print(f.name, f.stat().st_size)
        return
    golden_rotated_int(ref)
    if "--small-only" in sys.argv:
        golden_small(ref)
        for f in sorted(GOLDEN.glob("k3_small.npz")):
            print(f.name, f.stat().st_size)
        return
    golden_small(ref)
    if "--rotated-only" in sys.argv:
        golden_rotated(ref)
        for f in sorted(GOLDEN.glob("k3_rotated.npz")):
            print(f.name, f.stat().st_size)
        return
    golden_rotated(ref)
    from PIL import Image
    from torchvision import transforms

    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)

    # ---- a4/a5 preprocessing: normalize_to_uint8 -> PIL RGB -> Resize((512,512)) ----
    out = {}
    for seed, h, w in K1_CASES:
        img = synthetic.make_iso_slice(seed, h, w)
        u8 = ref.normalize_to_uint8(img)
        pil = Image.fromarray(u8).convert("RGB")
        res = np.asarray(transforms.Resize((512, 512))(pil))
        assert (res[..., 0] == res[..., 1]).all() and (res[..., 0] == res[..., 2]).all()
        out[f"plane_{seed}_{h}_{w}"] = res[..., 0].copy()
        out[f"u8sum_{seed}_{h}_{w}"] = np.array([int(u8.astype(np.int64).sum()), int(u8[::7, ::5].astype(np.int64).sum())])
    np.savez_compressed(GOLDEN / "k1_normalize_resize.npz", **out)

    # ---- normalize_to_uint8 edge cases (io/__init__.py:15-30) ----
    edge_in = {
        "const300": np.full((4, 5), 300.0, dtype=np.float32),
        "constneg": np.full((3, 3), -1.0, dtype=np.float32),
        "int16": (np.arange(-50, 50, dtype=np.int16).reshape(10, 10) * 37),
        "tiny_range": (np.float32(1000.0) + np.arange(12, dtype=np.float32).reshape(3, 4) * np.float32(1e-4)),
        "neg_pos": np.linspace(-7.5, 9.25, 64, dtype=np.float32).reshape(8, 8),
        "f64": np.linspace(0, 1e6, 35, dtype=np.float64).reshape(5, 7),
    }
    edge = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k, v in edge_in.items():
            edge["in_" + k] = v
            edge["out_" + k] = ref.normalize_to_uint8(v)
    np.savez_compressed(GOLDEN / "normalize_edge.npz", **edge)

    # ---- a8 mm_to_pixels known answers (notebooks/compare_crop_modes.ipynb:65,255 + more) ----
    mm_cases = [((35, 5, 20, 20), (0.3, 0.3)), ((50, 20, 30, 30), (0.3, 0.3)), ((55, 15, 17.5, 20), (0.3, 0.3)),
                ((55, 15, 17.5, 20), (0.5, 0.4)), ((0.75, 0.45, 1.05, 0.15), (0.3, 0.3))]
    mm = np.array([list(ref.mm_to_pixels(d, s)) for d, s in mm_cases], dtype=np.int64)
    assert tuple(mm[0]) == (117, 17, 67, 67), mm[0]  # the notebook's recorded answer
    np.savez_compressed(GOLDEN / "mm_to_pixels.npz", delta_mm=np.array([d for d, _ in mm_cases], dtype=np.float64),
                        spacing=np.array([s for _, s in mm_cases], dtype=np.float64), out=mm)

    # ---- a9-a11 crops via the reference CropContext (horizontal) ----
    crops = {}
    for seed, h, w in CROP_SERIES:
        img = synthetic.make_iso_slice(seed, h, w)
        xy = synthetic.make_coords(3, seed=seed, border_frac=0.25, hw=(h, w))
        # add hand-made extreme points (corners / edges)
        xy[2, 0] = (0.0, 0.0)
        xy[2, 1] = (0.99999, 0.99999)
        xy[2, 2] = (0.5, 0.001)
        for di, dmm in enumerate(CROP_DELTAS_MM):
            dpx = ref.mm_to_pixels(dmm, (0.3, 0.3))
            for cs in (128, 256):
                if cs == 256 and not (seed == CROP_SERIES[0][0] and di == 1):
                    continue  # the code-default 256x256 size on one series only (fixture size)
                res = np.zeros((3, 5, cs, cs), dtype=np.uint8)
                for s in (0, 2):
                    locs = {i: (float(xy[s, i, 0]), float(xy[s, i, 1])) for i in range(5)}
                    ctx = ref.CropContext(image=img, ivd_locations=locs, crop_size=(cs, cs), crop_delta_px=dpx, mode="horizontal")
                    for i in range(5):
                        res[s, i] = ctx.crop(i)
                crops[f"crops_{seed}_{h}_{w}_d{di}_c{cs}"] = res
        crops[f"xy_{seed}_{h}_{w}"] = xy
    np.savez_compressed(GOLDEN / "k3_crops.npz", **crops)

    # ---- a13 classifier input: [T2,T1,T2] -> Resize((256,256)) ----
    k = f"crops_{CROP_SERIES[0][0]}_{CROP_SERIES[0][1]}_{CROP_SERIES[0][2]}_d0_c128"
    t2, t1 = crops[k][0, 2], crops[k][0, 3]
    rgb = ref.construct_3channel(t2, t1)
    up = np.asarray(transforms.Resize((256, 256))(Image.fromarray(rgb)))
    np.savez_compressed(GOLDEN / "classifier_input.npz", t2=t2, t1=t1, up=up)

    # ---- a5-a7 predict_ivd_locations through the reference model classes + checkpoint format ----
    model_out = {}
    for tag, trained_like in (("init", False), ("trained", True)):
        oracle_model = make_model("base", seed=0, trained_like=trained_like)
        sd = oracle_model.state_dict()
        if not trained_like:
            torch.manual_seed(0)
            ref_model = ref.CoordinateRegressor(backbone="convnext_base", pretrained=False, num_levels=5)
            for kk, vv in ref_model.state_dict().items():
                assert torch.equal(vv, sd[kk]), f"RNG/construct-order drift at {kk}"
        with tempfile.TemporaryDirectory() as td:
            ck = Path(td) / "best_model.pt"
            # trainers/base.py:695-706 checkpoint layout
            torch.save({"epoch": 0, "model_state_dict": sd, "optimizer_state_dict": {}, "scheduler_state_dict": None,
                        "best_metric": 0.0, "best_epoch": 0, "history": {}, "config": {}}, ck)
            model = ref.load_localization_model(ck, "base", "cpu")
        for seed, h, w in MODEL_SLICES:
            img = synthetic.make_iso_slice(seed, h, w)
            locs = ref.predict_ivd_locations(model, img, "cpu", (512, 512))
            model_out[f"coords_{tag}_{seed}_{h}_{w}"] = np.array([locs[i] for i in range(5)], dtype=np.float64)
        # weight fingerprints so RNG drift on another box is detected before a parity claim
        model_out[f"wsum_{tag}"] = np.array([float(sd["backbone.stem.0.weight"].double().sum()),
                                             float(sd["backbone.stages.2.blocks.13.mlp.fc1.weight"].double().sum()),
                                             float(sd["head.5.weight"].double().sum())])
    np.savez_compressed(GOLDEN / "model_coords.npz", **model_out)

    for f in sorted(GOLDEN.glob("*.npz")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
