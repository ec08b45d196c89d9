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


def test_k01_fused_equals_k0_then_k1_and_streams_from_source_planes():
    """SURVEY 8f row 1 / VERDICT r01 #4: K0 fused into K1 (``svb_k01_midplane_normalize_resize``: K0 accumulates each plane's
    min / max while writing it, K1 makes no min/max pass) is the same function as K0 followed by K1 -- isotropic planes, min / max
    and uint8 planes identical -- on a ragged batch walked in several L2-sized groups; and the streamed driver fed with the two
    SOURCE planes per series (``volumes.PinnedVolumes``) returns the bytes of the resident volume path (coords, crops, 256^2)."""
    from spine_vision_b200 import ops

    rng = np.random.default_rng(11)
    vols, sps, dirs = [], [], []
    for k in range(7):
        h, w = int(rng.integers(200, 520)), int(rng.integers(200, 520))
        sp = float(rng.uniform(0.35, 0.9))
        v, _, d = synthetic.make_volume(70 + k, int(rng.integers(9, 17)), h, w, (sp, sp, 4.0))
        if k == 3:
            v = np.full_like(v, 7.0)  # a constant series: max == min, the un-scaled values are cast (io/__init__.py:28)
        if k == 5:
            v = np.rint(v).astype(np.int16)
        vols.append(v); sps.append((sp, sp, 4.0)); dirs.append(d)
    pool, _ = volumes.midplane_resample(vols, sps, dirs, dev())
    planes, mm = ops.normalize_resize(pool, (512, 512), return_minmax=True)
    pv = volumes.PinnedVolumes(vols, sps, dirs)
    assert pv.shapes == pool.shapes and pv.nbytes < sum(h * w for h, w in pool.shapes) * 4
    d = torch.device(dev())
    pool2 = ops.SlicePool(torch.empty_like(pool.data), pool.offs, pool.hw, list(pool.shapes))
    descs = pv.chunk_descs(0, pv.n).to(d)
    import os
    for group in ("0", "2", "3", ""):  # whole batch in one pair (the default), groups of two / three planes (the L2-sized schedule)
        os.environ["SVB_K1_GROUP"] = group
        planes2, mm2 = ops.midplane_normalize_resize(pv.host.to(d), descs, pool2, (512, 512), return_minmax=True)
        torch.cuda.synchronize()
        for i, (h, w) in enumerate(pool.shapes):
            o = int(pool.offs[i])
            assert torch.equal(pool2.data[o : o + h * w], pool.data[o : o + h * w]), i
        assert torch.equal(mm2, mm) and torch.equal(planes2, planes)
    os.environ.pop("SVB_K1_GROUP", None)
    om = make_model("base", seed=0)
    model = cropping.LocalizationModel(om.state_dict(), dev(), micro_batch=3)
    want = pipeline.localize_and_crop_volumes(vols, sps, dirs, model, dev(), crop_delta_mm=(50, 20, 30, 30), crop_size=(128, 128))
    wc, wk, wk2 = want.to_host()
    streamer = pipeline.StreamedLocalizer(model, dev(), (50, 20, 30, 30), (128, 128), (512, 512), (256, 256), chunk=3)
    for _ in range(2):
        gc, gk, gk2 = streamer.run(pv)
        assert np.array_equal(gc.numpy(), wc) and np.array_equal(gk.numpy(), wk) and np.array_equal(gk2.numpy(), wk2)
    fb = pipeline.StreamedLocalizer(None, dev(), (50, 20, 30, 30), (128, 128), (512, 512), None, chunk=4)
    fc, fk, _ = fb.run(pv)
    ref_fb = pipeline.localize_and_crop_volumes(vols, sps, dirs, None, dev(), crop_delta_mm=(50, 20, 30, 30), crop_size=(128, 128), second_size=None)
    assert np.array_equal(fk.numpy(), ref_fb.crops.cpu().numpy())
