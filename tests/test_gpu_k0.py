"""GPU parity: K0 (middle isotropic sagittal plane from the source volume) vs oracle/itk_resample.py, through the C ABI.
SimpleITK is absent from the image, so the oracle is a restatement of ITK's documented behaviour: parity UNPINNED."""
import numpy as np
import pytest
import torch

from gpu_util import dev
from oracle import itk_resample as itk
from oracle.convnext import make_model
from spine_vision_b200 import cropping, pipeline, synthetic, volumes

pytestmark = pytest.mark.gpu

DIRS = {
    "lps": None,
    "sagittal": (0.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, -1.0, 0.0),
    "ras": (-1.0, 0.0, 0.0, 0.0, -1.0, 0.0, 0.0, 0.0, 1.0),
    "oblique": (0.05, -0.02, 0.998, 0.996, 0.08, -0.04, -0.07, -0.996, -0.03),
}


def test_k0_vs_oracle_orientations_bit_exact():
    rng = np.random.default_rng(5)
    vols, sps, dirs, want = [], [], [], []
    for name, d in DIRS.items():
        for shape, sp in [((5, 60, 70), (0.7, 0.9, 4.0)), ((9, 33, 47), (0.45, 0.31, 1.3)), ((4, 40, 40), (0.3, 0.3, 0.3)),
                          ((6, 21, 19), (1.7, 2.2, 0.8))]:
            v = (rng.random(shape) * 1500).astype(np.float32)
            vols.append(v); sps.append(sp); dirs.append(d)
            want.append(itk.resample_middle_sagittal(v, sp, d))
    pool, spacings = volumes.midplane_resample(vols, sps, dirs, dev())
    torch.cuda.synchronize()
    flat = pool.data.cpu().numpy()
    offs = pool.offs.cpu().numpy()
    for i, (w, wsp) in enumerate(want):
        h_, w_ = pool.shapes[i]
        assert (h_, w_) == w.shape and spacings[i] == wsp
        got = flat[offs[i] : offs[i] + h_ * w_].reshape(h_, w_)
        assert np.array_equal(got.view(np.uint32), w.view(np.uint32)), f"series {i}: {np.abs(got - w).max()}"


def test_k0_config1_series_and_pipeline():
    """Config-1 geometry (15 x 512 x 512 @ 0.7/0.7/4.0 mm -> 1195 x 1195, last row/column outside the buffer = 0) and
    the volume entry of the pipeline against K1 / model / K3 run on the oracle's plane."""
    vol, sp, d = synthetic.make_volume(3)
    want, wsp = itk.resample_middle_sagittal(vol, sp, d)
    assert want.shape == (1195, 1195) and wsp == (0.3, 0.3)
    assert not want[-1].any() and not want[:, -1].any() and want[:-1, :-1].any()
    pool, spacings = volumes.midplane_resample([vol], [sp], [d], dev())
    got = pool.data[: 1195 * 1195].reshape(1195, 1195).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    om = make_model("base", seed=0)
    model = cropping.LocalizationModel(om.state_dict(), dev(), dtype="bf16")
    a = pipeline.localize_and_crop_volumes([vol], [sp], [d], model, dev(), crop_delta_mm=(50, 20, 30, 30), crop_size=(128, 128))
    from spine_vision_b200 import ops

    b = pipeline.localize_and_crop(ops.SlicePool.from_numpy([want], dev()), model, (50, 20, 30, 30), (128, 128))
    assert torch.equal(a.coords, b.coords) and torch.equal(a.crops, b.crops) and torch.equal(a.crops2, b.crops2)


def test_k0_integer_pixel_types_truncate_like_itk_cast():
    """Integer-typed volumes (what SPIDER ships): ITK casts every interpolated value back to the pixel type
    (static_cast = truncation toward zero); negative values included."""
    rng = np.random.default_rng(9)
    vols, sps, dirs, want = [], [], [], []
    for dt, lo, hi in (("int16", -900, 2500), ("uint16", 0, 4000), ("uint8", 0, 256)):
        for name in ("sagittal", "oblique"):
            v = rng.integers(lo, hi, size=(7, 41, 37)).astype(dt)
            vols.append(v); sps.append((0.62, 0.71, 3.9)); dirs.append(DIRS[name])
            w, _ = itk.resample_middle_sagittal(v, sps[-1], dirs[-1])
            assert w.dtype == np.dtype(dt)
            want.append(w)
    pool, _ = volumes.midplane_resample(vols, sps, dirs, dev())  # integer_pixels inferred from the dtype
    flat, offs = pool.data.cpu().numpy(), pool.offs.cpu().numpy()
    for i, w in enumerate(want):
        h_, w_ = pool.shapes[i]
        got = flat[offs[i] : offs[i] + h_ * w_].reshape(h_, w_)
        assert np.array_equal(got, w.astype(np.float32)), f"series {i}"
    # the same data declared float keeps the fractions
    pool_f, _ = volumes.midplane_resample([vols[0].astype(np.float32)], sps[:1], dirs[:1], dev())
    assert (pool_f.data[: pool.shapes[0][0] * pool.shapes[0][1]].cpu().numpy() % 1 != 0).any()
