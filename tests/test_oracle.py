"""CPU: the oracle against the golden vectors frozen from the reference's own functions
(oracle/make_golden.py) and against the libraries the reference calls."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import fixedpoint as fx
from oracle import reference_path as ref
from oracle.convnext import make_model
from spine_vision_b200 import synthetic


def _cases(npz, prefix):
    for k in npz.files:
        if k.startswith(prefix):
            yield k, [int(v) for v in k[len(prefix):].split("_")[:3]]


def test_mm_to_pixels_known_answers():
    g = np.load(GOLDEN / "mm_to_pixels.npz")
    # notebooks/compare_crop_modes.ipynb:65,255 -> left=117, right=17, top=67, bottom=67
    assert fx.mm_to_pixels((35, 5, 20, 20), (0.3, 0.3)) == (117, 17, 67, 67)
    for d, s, o in zip(g["delta_mm"], g["spacing"], g["out"]):
        assert fx.mm_to_pixels(tuple(d), tuple(s)) == tuple(int(v) for v in o)
        assert ref.mm_to_pixels(tuple(d), tuple(s)) == tuple(int(v) for v in o)


def test_normalize_edge_cases_match_reference():
    g = np.load(GOLDEN / "normalize_edge.npz")
    for k in g.files:
        if k.startswith("in_"):
            want = g["out_" + k[3:]]
            assert np.array_equal(fx.normalize_to_uint8(g[k]), want), k
            assert np.array_equal(ref.normalize_to_uint8(g[k]), want), k
    assert fx.normalize_to_uint8(np.full((2, 2), 300.0))[0, 0] == 44  # wraps mod 256


def test_k1_restatement_matches_reference_golden():
    g = np.load(GOLDEN / "k1_normalize_resize.npz")
    n = 0
    for k, (seed, h, w) in _cases(g, "plane_"):
        img = synthetic.make_iso_slice(seed, h, w)
        u8 = fx.normalize_to_uint8(img)
        s = g[f"u8sum_{seed}_{h}_{w}"]
        assert int(u8.astype(np.int64).sum()) == int(s[0]) and int(u8[::7, ::5].astype(np.int64).sum()) == int(s[1])
        assert np.array_equal(fx.pillow_resize_u8(u8, (512, 512)), g[k]), k
        plane, _ = ref.preprocess_slice(img, (512, 512))
        assert np.array_equal(plane, g[k]), k
        n += 1
    assert n >= 4


@pytest.mark.parametrize("shape", [(1195, 1195, 512, 512), (3400 // 4, 1100 // 4, 128, 128), (350, 420, 512, 512),
                                   (128, 128, 256, 256), (109, 128, 256, 256), (64, 64, 64, 96)])
def test_pillow_restatement_vs_pillow(shape):
    Image = pytest.importorskip("PIL.Image")
    h, w, oh, ow = shape
    img = np.random.default_rng(h * 7 + w).integers(0, 256, (h, w), dtype=np.uint8)
    want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
    assert np.array_equal(fx.pillow_resize_u8(img, (oh, ow)), want)


@pytest.mark.parametrize("shape", [(200, 234, 109, 128), (234, 200, 128, 109), (125, 233, 69, 128), (50, 60, 128, 107),
                                   (256, 200, 128, 100), (1, 7, 18, 128), (128, 128, 128, 128), (3, 5, 77, 128)])
def test_opencv_restatement_vs_opencv(shape):
    cv2 = pytest.importorskip("cv2")
    h, w, oh, ow = shape
    img = np.random.default_rng(h * 13 + w).integers(0, 256, (h, w), dtype=np.uint8)
    want = cv2.resize(img, (ow, oh), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(fx.cv_resize_u8(img, (oh, ow)), want)


def test_crops_match_reference_golden():
    g = np.load(GOLDEN / "k3_crops.npz")
    n = 0
    deltas = [(50, 20, 30, 30), (55, 15, 17.5, 20)]
    for k in g.files:
        if not k.startswith("crops_"):
            continue
        seed, h, w, d, c = k[len("crops_"):].split("_")
        seed, h, w, di, cs = int(seed), int(h), int(w), int(d[1:]), int(c[1:])
        img = synthetic.make_iso_slice(seed, h, w)
        xy = g[f"xy_{seed}_{h}_{w}"]
        dpx = fx.mm_to_pixels(deltas[di], (0.3, 0.3))
        for s in (0, 2):
            for lvl in range(5):
                x, y = float(xy[s, lvl, 0]), float(xy[s, lvl, 1])
                assert np.array_equal(fx.crop_region_horizontal(img, x, y, (cs, cs), dpx), g[k][s, lvl]), (k, s, lvl)
                assert np.array_equal(ref.crop_region_horizontal(img, x, y, (cs, cs), dpx), g[k][s, lvl]), (k, s, lvl)
                n += 1
    assert n >= 80


def test_classifier_input_matches_reference_golden():
    g = np.load(GOLDEN / "classifier_input.npz")
    u8, t = ref.classifier_input(g["t2"], g["t1"])
    assert np.array_equal(u8, g["up"])
    assert t.shape == (3, 256, 256)
    # per-channel Pillow resize == restated integer resize of each plane
    assert np.array_equal(fx.pillow_resize_u8(g["t2"], (256, 256)), g["up"][..., 0])
    assert np.array_equal(fx.pillow_resize_u8(g["t1"], (256, 256)), g["up"][..., 1])


def test_convnext_restatement_equals_torchvision():
    """timm is absent; the restated backbone must be the same function as torchvision's
    convnext_base once the weights are mapped across."""
    tv = pytest.importorskip("torchvision.models")
    m = make_model("base", seed=3, trained_like=True)
    t = tv.convnext_base(weights=None).eval()
    sd = m.backbone.state_dict()
    tsd = t.state_dict()
    # torchvision: features.0 = stem, features.{1,3,5,7} = stages, features.{2,4,6} = downsample
    tsd["features.0.0.weight"], tsd["features.0.0.bias"] = sd["stem.0.weight"], sd["stem.0.bias"]
    tsd["features.0.1.weight"], tsd["features.0.1.bias"] = sd["stem.1.weight"], sd["stem.1.bias"]
    for s in range(4):
        if s > 0:
            f = 2 * s
            tsd[f"features.{f}.0.weight"], tsd[f"features.{f}.0.bias"] = sd[f"stages.{s}.downsample.0.weight"], sd[f"stages.{s}.downsample.0.bias"]
            tsd[f"features.{f}.1.weight"], tsd[f"features.{f}.1.bias"] = sd[f"stages.{s}.downsample.1.weight"], sd[f"stages.{s}.downsample.1.bias"]
        f = 2 * s + 1
        j = 0
        while f"stages.{s}.blocks.{j}.gamma" in sd:
            p = f"stages.{s}.blocks.{j}."
            q = f"features.{f}.{j}."
            tsd[q + "layer_scale"] = sd[p + "gamma"].reshape(-1, 1, 1)
            tsd[q + "block.0.weight"], tsd[q + "block.0.bias"] = sd[p + "conv_dw.weight"], sd[p + "conv_dw.bias"]
            tsd[q + "block.2.weight"], tsd[q + "block.2.bias"] = sd[p + "norm.weight"], sd[p + "norm.bias"]
            tsd[q + "block.3.weight"], tsd[q + "block.3.bias"] = sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]
            tsd[q + "block.5.weight"], tsd[q + "block.5.bias"] = sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"]
            j += 1
    tsd["classifier.0.weight"], tsd["classifier.0.bias"] = sd["head.norm.weight"], sd["head.norm.bias"]
    t.load_state_dict(tsd)
    t.classifier[2] = torch.nn.Identity()
    x = torch.randn(1, 3, 128, 128, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        a, b = m.backbone(x), t(x)
    assert a.shape == b.shape == (1, 1024)
    assert torch.allclose(a, b, atol=2e-5, rtol=1e-5), (a - b).abs().max()
    assert len(m.state_dict()) == 348  # 342 backbone + 6 head tensors


def test_model_coords_match_reference_golden():
    g = np.load(GOLDEN / "model_coords.npz")
    torch.set_num_threads(8)
    for tag, tl in (("init", False), ("trained", True)):
        m = make_model("base", seed=0, trained_like=tl)
        sd = m.state_dict()
        fp = np.array([float(sd["backbone.stem.0.weight"].double().sum()),
                       float(sd["backbone.stages.2.blocks.13.mlp.fc1.weight"].double().sum()),
                       float(sd["head.5.weight"].double().sum())])
        assert np.allclose(fp, g[f"wsum_{tag}"], rtol=0, atol=1e-9), "torch RNG drift: seeded weights differ from the golden run"
        k = f"coords_{tag}_21_640_650"
        locs = ref.predict_ivd_locations(m, synthetic.make_iso_slice(21, 640, 650), "cpu", (512, 512))
        got = np.array([locs[i] for i in range(5)])
        assert np.abs(got - g[k]).max() < 2e-6, np.abs(got - g[k]).max()


def test_gelu_fast_formula():
    """NumPy replica of gelu_fast (csrc/svb_convnext_kernels.cuh) vs exact erf GELU."""
    from scipy.special import erf

    x = np.linspace(-12, 12, 200001).astype(np.float32)
    u = np.minimum(np.abs(x), np.float32(5.65685425))
    r = np.float32(-5.20460508e-04)
    for c in (7.39751849e-03, -5.25612477e-02, -4.59254682e-01, -1.15109138e+00):
        r = r * u + np.float32(c)
    r = r * u
    e = np.exp2(r.astype(np.float32))
    got = np.maximum(x, 0) - np.abs(np.float32(0.5) * x * e)
    want = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
    assert np.abs(got - want).max() < 3e-6


def test_gelu_pack2_formula():
    """NumPy replica of gelu_pack2 (the GEMM epilogue's GELU: cubic exponent fit in v = -min(|x|, 12)) vs exact erf GELU,
    over the whole range incl. the clamp and huge inputs."""
    from scipy.special import erf

    x = np.concatenate([np.linspace(-40, 40, 400001), [-1e6, 1e6, -65504.0, 65504.0, 0.0]]).astype(np.float32)
    v = np.maximum(-np.abs(x), np.float32(-12.0))
    r = np.float32(4.16166e-03)
    for c in (4.573539e-02, -4.6493057e-01, 1.14956693e+00, -1.0):
        r = r * v + np.float32(c)
    e = np.exp2(r.astype(np.float32))
    got = v * e + np.maximum(x, 0)
    xd = x.astype(np.float64)
    want = 0.5 * xd * (1 + erf(xd / np.sqrt(2)))
    assert np.isfinite(got).all() and np.abs(got - want).max() < 1.2e-5


def test_rotated_crop_restatement_matches_reference_golden_and_opencv():
    """oracle/fixedpoint.py's warpAffine / rotation-angle restatements against (a) crops and angles frozen from the
    reference's own CropContext(mode="rotated") and (b) cv2.warpAffine directly, bit for bit."""
    import cv2

    from oracle import fixedpoint as fp

    g = np.load(GOLDEN / "k3_rotated.npz")
    seed, h, w = 13, 400, 380
    img = synthetic.make_iso_slice(seed, h, w)
    xy = g[f"xy_{seed}_{h}_{w}"]
    for di, dmm in enumerate([(50, 20, 30, 30), (55, 15, 17.5, 20)]):
        dpx = fp.mm_to_pixels(dmm, (0.3, 0.3))
        for boost in (1.0, 2.0):
            want = g[f"crops_{seed}_{h}_{w}_d{di}_b{int(boost)}"]
            wang = g[f"angles_{seed}_{h}_{w}_d{di}_b{int(boost)}"]
            for s in range(2):
                locs = {i: (float(xy[s, i, 0]), float(xy[s, i, 1])) for i in range(5)}
                ang = fp.rotation_angles(locs, (h, w), boost)
                for i in range(5):
                    assert ang[i] == wang[s, i]
                    got = fp.crop_region_rotated(img, locs[i][0], locs[i][1], (128, 128), dpx, ang[i])
                    assert np.array_equal(got, want[s, i])
                    assert np.array_equal(ref.CropContext(img, locs, (128, 128), dpx, "rotated", boost).crop(i), want[s, i])
    rng = np.random.default_rng(3)
    a = (rng.random((211, 173), dtype=np.float32) * 900).astype(np.float32)
    for ang in (0.0, 6.98, -26.68, 45.0, 90.0):
        M = cv2.getRotationMatrix2D((80, 120), ang, 1.0)
        assert np.array_equal(M, fp.rotation_matrix_2d((80, 120), ang))
        want = cv2.warpAffine(a, M, (173, 211), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
        assert np.array_equal(want.view(np.uint32), fp.warp_affine_f32(a, M).view(np.uint32))


def test_itk_restatement_plane_shortcut_equals_whole_volume_path():
    """oracle/itk_resample.py (parity UNPINNED -- SimpleITK absent): the one-plane shortcut K0 is tested against equals
    the reference's order of operations (resample the whole volume, orient to LPI, take arr[:, :, n // 2])."""
    from oracle import itk_resample as itk
    from spine_vision_b200 import volumes

    rng = np.random.default_rng(0)
    dirs = [None, (0, 0, 1, 1, 0, 0, 0, -1, 0), (-1, 0, 0, 0, -1, 0, 0, 0, 1), (0.05, -0.02, 0.998, 0.996, 0.08, -0.04, -0.07, -0.996, -0.03)]
    for d in dirs:
        for shape, sp in [((5, 12, 14), (0.7, 0.9, 4.0)), ((7, 9, 11), (0.45, 0.31, 1.3))]:
            v = (rng.random(shape) * 1000).astype(np.float32)
            full = itk.extract_middle_slice_full(v, sp, d)
            fast, spc = itk.resample_middle_sagittal(v, sp, d)
            assert np.array_equal(full, fast) and spc == (0.3, 0.3)
            plan = volumes.plan_midplane(v, sp, d)  # host planning of the product agrees on geometry
            assert plan.out_hw == fast.shape and plan.spacing == spc
    assert itk.new_size((512, 512, 15), (0.7, 0.7, 4.0)) == [1195, 1195, 200]


def test_convnext_v2_restatement_grn():
    """oracle/convnext.py's ConvNeXt-V2 (timm convnextv2; timm absent -> restated, unpinned): GlobalResponseNorm is the identity
    at its zero init, follows x + b + w * x * g / (mean_c g + eps) with g the per-image, per-channel L2 norm over the tokens,
    the blocks carry no layer scale, and the state dict uses timm's key names."""
    from oracle.convnext import GlobalResponseNorm, make_model

    x = torch.randn(2, 5, 7, 12)
    grn = GlobalResponseNorm(12)
    assert torch.equal(grn(x), x)
    with torch.no_grad():
        grn.weight.copy_(torch.randn(12))
        grn.bias.copy_(torch.randn(12))
    g = x.pow(2).sum(dim=(1, 2), keepdim=True).sqrt()
    want = x * (1 + grn.weight * g / (g.mean(dim=-1, keepdim=True) + 1e-6)) + grn.bias
    assert torch.allclose(grn(x), want, atol=1e-6)
    sd = make_model("v2_tiny", seed=0).state_dict()
    assert "backbone.stages.2.blocks.8.mlp.grn.weight" in sd and sd["backbone.stages.0.blocks.0.mlp.grn.bias"].shape == (384,)
    assert not any(k.endswith(".gamma") for k in sd) and len([k for k in sd if k.endswith("conv_dw.weight")]) == 18


def test_small_slice_crops_match_reference_golden():
    """Slices smaller than the crop box and degenerate strips (tests/golden/k3_small.npz, frozen from the reference's own
    CropContext in both modes): the port and the fixed-point restatements reproduce them bit for bit."""
    from oracle import fixedpoint as fp

    g = np.load(GOLDEN / "k3_small.npz")
    deltas = [(50, 20, 30, 30), (55, 15, 17.5, 20)]
    n = 0
    for k in g.files:
        if not k.startswith("crops_"):
            continue
        seed, h, w, d, c, mode = k[len("crops_"):].split("_")
        seed, h, w, di, cs = int(seed), int(h), int(w), int(d[1:]), int(c[1:])
        img = synthetic.make_iso_slice(seed, h, w)
        xy = g[f"xy_{seed}_{h}_{w}"]
        dpx = fx.mm_to_pixels(deltas[di], (0.3, 0.3))
        for s in range(2):
            locs = {i: (float(xy[s, i, 0]), float(xy[s, i, 1])) for i in range(5)}
            ang = fp.rotation_angles(locs, (h, w), 1.0) if mode == "rotated" else None
            for i in range(5):
                if mode == "horizontal":
                    got = fx.crop_region_horizontal(img, locs[i][0], locs[i][1], (cs, cs), dpx)
                    assert np.array_equal(ref.crop_region_horizontal(img, locs[i][0], locs[i][1], (cs, cs), dpx), g[k][s, i]), (k, s, i)
                else:
                    got = fp.crop_region_rotated(img, locs[i][0], locs[i][1], (cs, cs), dpx, ang[i])
                assert np.array_equal(got, g[k][s, i]), (k, s, i)
                n += 1
    assert n == 6 * (2 * 2 + 1) * 10


def test_pillow_pass_order_for_slivers_is_pinned():
    """The installed Pillow (12.2.0) resizes an image more than 100 times taller than wide VERTICALLY first when the height
    shrinks (``Image.resize``: ``self.size[1] > self.size[0] * 100 and size[1] < self.size[1]``); every other shape -- and every
    shape under the reference's pinned Pillow 10.2.0 (uv.lock:2563-2564) -- goes horizontally first.  The restatement (and K1)
    keep the single horizontal-first order; this test pins where the two part so that the exception is a statement, not a
    surprise: outside the rule the restatement equals Pillow, inside it Pillow equals the transposed order."""
    from PIL import Image

    rng = np.random.default_rng(5)

    def pil(u8, hw):
        return np.asarray(Image.fromarray(u8).resize((hw[1], hw[0]), Image.BILINEAR))

    def vertical_first(u8, hw):
        return fx.pillow_resize_u8(fx.pillow_resize_u8(u8, (hw[0], u8.shape[1])), hw)

    for shape, out_hw in [((873, 9), (512, 512)), ((600, 6), (512, 512)), ((400, 2), (512, 512)), ((1195, 12), (512, 512)),
                          ((2, 793), (512, 512)), ((889, 36), (64, 96))]:
        u8 = rng.integers(0, 256, size=shape).astype(np.uint8)
        assert not (shape[0] > 100 * shape[1] and out_hw[0] < shape[0])
        assert np.array_equal(fx.pillow_resize_u8(u8, out_hw), pil(u8, out_hw)), shape
    import inspect

    if "self.size[0] * 100" in inspect.getsource(Image.Image.resize):  # the rule is in the installed Pillow's Python layer
        for shape, out_hw in [((873, 5), (512, 512)), ((873, 8), (512, 512)), ((601, 6), (512, 512)), ((500, 3), (64, 96))]:
            u8 = rng.integers(0, 256, size=shape).astype(np.uint8)
            assert np.array_equal(vertical_first(u8, out_hw), pil(u8, out_hw)), shape
            assert not np.array_equal(fx.pillow_resize_u8(u8, out_hw), pil(u8, out_hw)), shape


def test_rotated_crop_on_integer_slices_matches_reference_golden_and_opencv():
    """the reference warps the slice in the FILE's pixel type (int16 for SPIDER .mha / MR DICOM), so cv2.warpAffine
    rounds every warped value back to that type before normalize_to_uint8.  The restatement (fixedpoint.warp_affine) equals
    cv2.warpAffine bit for bit for float32 / int16 / uint16 / uint8, and crop_region_rotated equals crops frozen from the
    reference's own CropContext on integer-typed slices (tests/golden/k3_rotated_int.npz)."""
    import cv2

    from oracle.make_golden import INT_SERIES, INT_TYPES, int_slice

    rng = np.random.default_rng(7)
    for dt in (np.float32, np.int16, np.uint16, np.uint8):
        for _ in range(6):
            h, w = (int(v) for v in rng.integers(40, 160, 2))
            if dt == np.float32:
                img = (rng.random((h, w)) * 1500).astype(dt)
            elif dt == np.int16:
                img = rng.integers(-32768, 32768, (h, w)).astype(dt)
            else:
                img = rng.integers(0, np.iinfo(dt).max + 1, (h, w)).astype(dt)
            m = cv2.getRotationMatrix2D((float(rng.integers(0, w)), float(rng.integers(0, h))), float(rng.uniform(-40, 40)), 1.0)
            want = cv2.warpAffine(img, m, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
            got = fx.warp_affine(img, m)
            assert got.dtype == want.dtype and np.array_equal(got, want), dt
    g = np.load(GOLDEN / "k3_rotated_int.npz")
    dpx = fx.mm_to_pixels((50, 20, 30, 30), (0.3, 0.3))
    differs_from_float = 0
    for seed, h, w in INT_SERIES:
        xy = g[f"xy_{seed}_{h}_{w}"]
        for name in INT_TYPES:
            img = int_slice(seed, h, w, name)
            want = g[f"crops_{seed}_{h}_{w}_{name}"]
            for s in range(2):
                locs = {i: (float(xy[s, i, 0]), float(xy[s, i, 1])) for i in range(5)}
                ang = fx.rotation_angles(locs, (h, w), 1.0)
                for i in range(5):
                    got = fx.crop_region_rotated(img, locs[i][0], locs[i][1], (128, 128), dpx, ang[i])
                    assert np.array_equal(got, want[s, i]), (seed, name, s, i)
                    as_float = fx.crop_region_rotated(img.astype(np.float32), locs[i][0], locs[i][1], (128, 128), dpx, ang[i])
                    differs_from_float += int((as_float != want[s, i]).sum())
    assert differs_from_float > 0  # the float32 assumption of round 1 was observably wrong on integer sources
