"""GPU parity: K3 (crop + per-crop normalise + OpenCV letterbox + Pillow second output) through the C ABI."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN
from gpu_util import dev
from oracle import fixedpoint as fx
from oracle import reference_path as ref
from spine_vision_b200 import cropping, ops, pipeline, synthetic

pytestmark = pytest.mark.gpu
DELTAS = [(50, 20, 30, 30), (55, 15, 17.5, 20)]


def test_k3_matches_reference_golden_bit_exact():
    g = np.load(GOLDEN / "k3_crops.npz")
    checked = 0
    for k in g.files:
        if not k.startswith("crops_"):
            continue
        seed, h, w, d, c = k[len("crops_"):].split("_")
        seed, h, w, di, cs = int(seed), int(h), int(w), int(d[1:]), int(c[1:])
        img = synthetic.make_iso_slice(seed, h, w)
        xy = g[f"xy_{seed}_{h}_{w}"][[0, 2]]
        pool = ops.SlicePool.from_numpy([img, img], dev())
        crops, crops2, geom = pipeline.crop_levels(pool, torch.from_numpy(xy).to(dev()), DELTAS[di], None, (cs, cs), (256, 256), True)
        got = crops.cpu().numpy()
        for j, s in enumerate((0, 2)):
            for lvl in range(5):
                want = g[k][s, lvl]
                assert np.array_equal(got[j, lvl], want), f"{k} set {s} level {lvl}: {(got[j, lvl] != want).sum()} px differ"
                up = fx.pillow_resize_u8(want, (256, 256))
                assert np.array_equal(crops2[j, lvl].cpu().numpy(), up), f"{k} second output"
                checked += 1
        # letterbox geometry must match exactly (SURVEY 8d parity gates)
        dpx = fx.mm_to_pixels(DELTAS[di], (0.3, 0.3))
        gm = geom.cpu().numpy()
        for j in range(2):
            for lvl in range(5):
                x1, x2, y1, y2 = fx.crop_box(h, w, float(xy[j, lvl, 0]), float(xy[j, lvl, 1]), dpx)
                nh, nw, yo, xo = fx.letterbox_geometry(y2 - y1, x2 - x1, (cs, cs))
                assert tuple(gm[j, lvl]) == (x1, x2, y1, y2, nh, nw, yo, xo)
    assert checked >= 80


def test_k3_random_coords_vs_oracle_ragged():
    shapes = [(1195, 1195), (640, 650), (400, 380), (1040, 1040), (350, 900)]
    slices = [synthetic.make_iso_slice(40 + i, h, w) for i, (h, w) in enumerate(shapes)]
    pool = ops.SlicePool.from_numpy(slices, dev())
    xy = synthetic.make_coords(len(shapes), seed=9, border_frac=0.3)
    crops, crops2, _ = pipeline.crop_levels(pool, torch.from_numpy(xy).to(dev()), (50, 20, 30, 30), None, (128, 128), (256, 256))
    got, got2 = crops.cpu().numpy(), crops2.cpu().numpy()
    dpx = fx.mm_to_pixels((50, 20, 30, 30), (0.3, 0.3))
    for i, sl in enumerate(slices):
        for lvl in range(5):
            want = ref.crop_region_horizontal(sl, float(xy[i, lvl, 0]), float(xy[i, lvl, 1]), (128, 128), dpx)
            assert np.array_equal(got[i, lvl], want), (i, lvl, int((got[i, lvl] != want).sum()))
            want2, _ = ref.classifier_input(want, None)
            assert np.array_equal(got2[i, lvl], want2[..., 0]), (i, lvl)


def test_k3_reference_shaped_api():
    img = synthetic.make_iso_slice(77, 640, 650)
    locs = {i: (0.45 + 0.02 * i, 0.2 + 0.12 * i) for i in range(5)}
    dpx = cropping.mm_to_pixels((55, 15, 17.5, 20), (0.3, 0.3))
    ctx = cropping.CropContext(image=img, ivd_locations=locs, crop_size=(256, 256), crop_delta_px=dpx, mode="horizontal", device=dev())
    for i in range(5):
        want = ref.crop_region_horizontal(img, locs[i][0], locs[i][1], (256, 256), dpx)
        assert np.array_equal(ctx.crop(i), want)
    assert ctx.crop(7) is None
    rot = cropping.CropContext(image=img, ivd_locations=locs, crop_size=(256, 256), crop_delta_px=dpx, mode="rotated", device=dev())
    rref = ref.CropContext(image=img, ivd_locations=locs, crop_size=(256, 256), crop_delta_px=dpx, mode="rotated")
    assert rot.rotation_angles == rref.rotation_angles
    for i in range(5):
        assert np.array_equal(rot.crop(i), rref.crop(i))
    with pytest.raises(ValueError):
        cropping.CropContext(image=img, ivd_locations=locs, crop_size=(256, 256), crop_delta_px=dpx, mode="sideways")
    u8 = ref.normalize_to_uint8(img[100:300, 200:434])
    assert np.array_equal(cropping.resize_with_padding(u8, (128, 128), dev()), ref.resize_with_padding(u8, (128, 128)))
    # constant crop (max == min): values are cast, not scaled
    flat = np.full((300, 300), 300.0, dtype=np.float32)
    got = cropping.crop_region_horizontal(flat, 0.5, 0.5, (128, 128), (40, 40, 30, 30), dev())
    assert np.array_equal(got, ref.crop_region_horizontal(flat, 0.5, 0.5, (128, 128), (40, 40, 30, 30)))


def test_k3_config4_properties():
    """Config-4 scale (5,000 crops here): every crop depends only on its own (slice, xy): permutation invariance,
    and the letterbox bars are exactly zero."""
    n_slices = 8
    slices = [synthetic.make_iso_slice(60 + i) for i in range(n_slices)]
    pool = ops.SlicePool.from_numpy(slices, dev())
    n = 1000
    xy = synthetic.make_coords(n, seed=3)
    idx = torch.arange(n * 5, dtype=torch.int32) % n_slices
    delta = torch.tensor([[167, 67, 100, 100]] * (n * 5), dtype=torch.int32)
    xyf = torch.from_numpy(xy.reshape(-1, 2).copy())
    a, _, geom = ops.crop_resample(pool, idx.to(dev()), xyf.to(dev()), delta.to(dev()), (200, 234), (128, 128), None, True)
    perm = torch.randperm(n * 5, generator=torch.Generator().manual_seed(0))
    b, _, _ = ops.crop_resample(pool, idx[perm].contiguous().to(dev()), xyf[perm].contiguous().to(dev()), delta.to(dev()),
                                (200, 234), (128, 128), None)
    assert torch.equal(a[perm.to(dev())], b)
    gm = geom.cpu().numpy()
    full = (gm[:, 1] - gm[:, 0] == 234) & (gm[:, 3] - gm[:, 2] == 200)
    assert full.mean() > 0.9
    sel = np.nonzero(full)[0][:50]
    ah = a.cpu().numpy()
    for i in sel:
        assert gm[i, 4] == 109 and gm[i, 5] == 128 and gm[i, 6] == 9  # 234x200 -> 128x109, y_off 9
        assert ah[i, :9].max() == 0 and ah[i, 9 + 109:].max() == 0


def test_k3_rotated_matches_reference_golden_bit_exact():
    """Rotated crop mode (CropContext mode="rotated", cropping.py:172-313): crops frozen from the reference's own
    CropContext (cv2.warpAffine + normalise + letterbox) against K3's on-the-fly rotation, through the drop-in
    CropContext and through the batched pipeline entry; angles from the host mirror must equal the reference's."""
    g = np.load(GOLDEN / "k3_rotated.npz")
    total = bad = 0
    for seed, h, w in [(10, 1195, 1195), (11, 1040, 1040), (12, 640, 650), (13, 400, 380)]:
        img = synthetic.make_iso_slice(seed, h, w)
        xy = g[f"xy_{seed}_{h}_{w}"]
        for di, dmm in enumerate([(50, 20, 30, 30), (55, 15, 17.5, 20)]):
            dpx = cropping.mm_to_pixels(dmm, (0.3, 0.3))
            for boost in (1.0, 2.0):
                want = g[f"crops_{seed}_{h}_{w}_d{di}_b{int(boost)}"]
                wang = g[f"angles_{seed}_{h}_{w}_d{di}_b{int(boost)}"]
                for s in range(2):
                    locs = {i: (float(xy[s, i, 0]), float(xy[s, i, 1])) for i in range(5)}
                    ctx = cropping.CropContext(img, locs, (128, 128), dpx, "rotated", last_disc_angle_boost=boost, device=dev())
                    assert [ctx.rotation_angles[i] for i in range(5)] == list(wang[s])
                    got = ctx.crop_all(range(5))
                    for i in range(5):
                        total += 1
                        bad += int((got[i] != want[s, i]).sum())
                # batched entry: both coordinate sets of the series in one launch
                pool = ops.SlicePool.from_numpy([img, img], dev())
                crops, _, _ = pipeline.crop_levels(pool, torch.from_numpy(xy[:2]).to(dev()), dmm, None, (128, 128), None,
                                                   mode="rotated", last_disc_angle_boost=boost)
                assert np.array_equal(crops.cpu().numpy(), want)
    assert total == 160 and bad == 0, f"{bad} differing pixels over {total} rotated crops"
    assert np.allclose(g["notebook_angles_2pt"], [4.96908734, 9.93817468])


def test_k3_rotated_integer_slices_match_reference_golden_bit_exact():
    """ADVICE r01 (svb_resize.cu:649): rotated crops of int16 / uint16 / uint8 slices.  The reference's cv2.warpAffine works in
    the slice's pixel type (rounds back to int16 / uint16, fixed-point weights for uint8); K3 gets the type per slice
    (SlicePool.pixel_kind) -- through the drop-in CropContext (dtype of the array it is given), the batched entry, and from a
    volume on (K0 keeps the type)."""
    from oracle.make_golden import INT_SERIES, INT_TYPES, int_slice

    g = np.load(GOLDEN / "k3_rotated_int.npz")
    dpx = cropping.mm_to_pixels((50, 20, 30, 30), (0.3, 0.3))
    n = 0
    for seed, h, w in INT_SERIES:
        xy = g[f"xy_{seed}_{h}_{w}"]
        imgs = {name: int_slice(seed, h, w, name) for name in INT_TYPES}
        for name, img in imgs.items():
            want = g[f"crops_{seed}_{h}_{w}_{name}"]
            for s in range(2):
                locs = {i: (float(xy[s, i, 0]), float(xy[s, i, 1])) for i in range(5)}
                got = cropping.CropContext(img, locs, (128, 128), dpx, "rotated", device=dev()).crop_all(range(5))
                for i in range(5):
                    assert np.array_equal(got[i], want[s, i]), (seed, name, s, i)
                    n += 1
        # one mixed batch: the three types (and a float32 copy) side by side, per-slice kinds
        names = list(INT_TYPES)
        pool = ops.SlicePool.from_numpy([imgs[k] for k in names] + [imgs["int16"].astype(np.float32)], dev())
        assert pool.pixel_kind is not None and pool.pixel_kind.cpu().tolist() == [1, 2, 3, 0]
        coords = torch.from_numpy(np.stack([xy[0]] * 4)).to(dev())
        crops, _, _ = pipeline.crop_levels(pool, coords, (50, 20, 30, 30), None, (128, 128), None, mode="rotated")
        got = crops.cpu().numpy()
        for k, name in enumerate(names):
            assert np.array_equal(got[k], g[f"crops_{seed}_{h}_{w}_{name}"][0]), name
        assert (got[3] != got[0]).any()  # the float32 copy of the int16 slice is NOT rounded back: a different crop
    assert n == 60


def test_k3_small_slices_match_reference_golden_bit_exact():
    """Slices SMALLER than the crop box (150 x 121, 90 x 300, exactly 200 x 234, ...) and a 7 x 500 strip, both crop modes,
    128 and 256 crops, corner centres (0, 0) and (0.99999, 0.99999): K3 against crops frozen from the reference's own
    CropContext (tests/golden/k3_small.npz)."""
    g = np.load(GOLDEN / "k3_small.npz")
    n = bad = 0
    for k in g.files:
        if not k.startswith("crops_"):
            continue
        seed, h, w, d, c, mode = k[len("crops_"):].split("_")
        seed, h, w, di, cs = int(seed), int(h), int(w), int(d[1:]), int(c[1:])
        img = synthetic.make_iso_slice(seed, h, w)
        xy = g[f"xy_{seed}_{h}_{w}"]
        pool = ops.SlicePool.from_numpy([img, img], dev())
        crops, _, _ = pipeline.crop_levels(pool, torch.from_numpy(xy[:2]).to(dev()), DELTAS[di], None, (cs, cs), None, mode=mode)
        got = crops.cpu().numpy()
        n += 10
        bad += int((got != g[k]).sum())
        dpx = cropping.mm_to_pixels(DELTAS[di], (0.3, 0.3))
        locs = {i: (float(xy[1, i, 0]), float(xy[1, i, 1])) for i in range(5)}
        ctx = cropping.CropContext(img, locs, (cs, cs), dpx, mode, device=dev())  # the one-series drop-in entry
        for i in (0, 4):
            assert np.array_equal(ctx.crop(i), g[k][1, i]), (k, i)
    assert n == 6 * (2 * 2 + 1) * 10 and bad == 0, f"{bad} differing pixels over {n} crops"
