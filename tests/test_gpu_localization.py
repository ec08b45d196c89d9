"""GPU parity of the second drop-in entry point: ``spine_vision_b200.localization_dataset.create_localization_dataset`` on a
synthetic raw tree vs what the reference's OWN builder wrote for the same tree (tests/golden/localization_dataset.npz, frozen
by oracle/make_golden_localization.py): file set, PNG pixels, copied / Pillow-encoded JPG bytes and CSV text; plus the
batched normalise kernel (svb_normalize_u8) against the reference's normalize_to_uint8 edge cases."""
import numpy as np
import torch
from PIL import Image

from conftest import GOLDEN
from gpu_util import dev, requires_gpu
from oracle import reference_path as ref
from spine_vision_b200 import localization_dataset as loc
from spine_vision_b200 import ops, synthetic


@requires_gpu
def test_create_localization_dataset_matches_reference_builder(tmp_path):
    g = np.load(GOLDEN / "localization_dataset.npz")
    synthetic.make_localization_tree(tmp_path, seed=0)
    cfg = loc.LocalizationDatasetConfig(base_path=tmp_path, output_name="loc", device=dev(), chunk_images=2)  # several GPU batches
    res = loc.create_localization_dataset(cfg)
    images = cfg.output_path / "images"
    names = sorted(p.name for p in images.iterdir())
    assert names == [str(n) for n in g["names"]] and res.num_samples == int(g["num_samples"])
    assert (cfg.output_path / "annotations.csv").read_text() == g["csv"].item()
    for n in names:
        if n.endswith(".png"):
            pil = Image.open(images / n)
            assert pil.mode == "L" and np.array_equal(np.asarray(pil), g["png_" + n]), n
        else:  # copied JPGs byte for byte; arrays saved under a .jpg name go through the same Pillow encoder
            assert np.array_equal(np.frombuffer((images / n).read_bytes(), dtype=np.uint8), g["raw_" + n]), n
    cfg2 = loc.LocalizationDatasetConfig(base_path=tmp_path, output_name="loc2", device=dev(), include_neural_foraminal=False,
                                         skip_invalid_instances=False)
    loc.create_localization_dataset(cfg2)
    assert (cfg2.output_path / "annotations.csv").read_text() == g["csv_no_foraminal"].item()


@requires_gpu
def test_normalize_u8_ragged_batch_bit_exact():
    rng = np.random.default_rng(11)
    arrays = [synthetic.make_iso_slice(70, 300, 257), (rng.random((33, 5)) * 1e-3).astype(np.float32), np.full((17, 9), 777.0, np.float32),
              rng.integers(-2000, 3000, size=(64, 64)).astype(np.float32), np.zeros((1, 1), np.float32),
              (rng.normal(0, 1, size=(129, 130)) * 1e6).astype(np.float32)]
    pool = ops.SlicePool.from_numpy(arrays, dev())
    out, mm = ops.normalize_u8(pool, return_minmax=True)
    out, mm, offs = out.cpu().numpy(), mm.cpu().numpy(), pool.offs.cpu().numpy()
    for i, a in enumerate(arrays):
        want = ref.normalize_to_uint8(a)
        got = out[offs[i] : offs[i] + a.size].reshape(a.shape)
        assert np.array_equal(got, want), f"array {i}: {(got != want).sum()} pixels differ"
        assert mm[i, 0] == a.min() and mm[i, 1] == a.max()
