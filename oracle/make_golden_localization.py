"""Freeze what the reference's OWN ``create_localization_dataset`` writes for a synthetic raw tree into ``tests/golden/``.

TEST INFRASTRUCTURE ONLY; build container only (needs ``/root/reference`` through ``oracle/ref_shim.py``):

    python -m oracle.make_golden_localization

The unmodified ``spine_vision/datasets/localization.py`` runs (CSV walks, skip rules, ``normalize_to_uint8``, PIL PNG save,
``write_records_csv``).  Only ``sitk.ReadImage`` / ``sitk.GetArrayFromImage`` -- SimpleITK is not installed -- are
substituted by ``oracle.dicom.read_slice`` (parity of the DICOM decode is UNPINNED, see its header).
Frozen: the CSV text, the names of the files written, the decoded PNGs and the bytes of the copied JPGs.
"""

from __future__ import annotations

import tempfile
from pathlib import Path

import numpy as np
from PIL import Image

from oracle import dicom, ref_shim
from spine_vision_b200 import synthetic

GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"


class _Img:
    def __init__(self, arr):
        self.arr = arr


def _read_image(path):
    s = dicom.read_slice(Path(path))
    px = s["px"]
    if not (s["slope"] == 1.0 and s["inter"] == 0.0):
        px = px.astype(np.float64) * s["slope"] + s["inter"]
        if float(s["slope"]).is_integer() and float(s["inter"]).is_integer():
            px = px.astype(np.int32)
    return _Img(px[None])  # sitk.GetArrayFromImage of a single slice is [1, rows, cols]


def main() -> None:
    ref_shim.install()
    from spine_vision.datasets import localization as ref_loc

    ref_loc.sitk.ReadImage = _read_image
    ref_loc.sitk.GetArrayFromImage = lambda im: im.arr
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        base = Path(tmp)
        synthetic.make_localization_tree(base, seed=0)
        cfg = ref_loc.LocalizationDatasetConfig(base_path=base, output_name="loc")
        res = ref_loc.create_localization_dataset(cfg)
        images = cfg.output_path / "images"
        names = sorted(p.name for p in images.iterdir())
        out["names"] = np.array(names)
        out["csv"] = np.array((cfg.output_path / "annotations.csv").read_text())
        out["num_samples"] = np.array(res.num_samples)
        for n in names:
            if n.endswith(".png"):
                out["png_" + n] = np.asarray(Image.open(images / n))
            else:
                out["raw_" + n] = np.frombuffer((images / n).read_bytes(), dtype=np.uint8)
        cfg2 = ref_loc.LocalizationDatasetConfig(base_path=base, output_name="loc2", include_neural_foraminal=False, skip_invalid_instances=False)
        ref_loc.create_localization_dataset(cfg2)
        out["csv_no_foraminal"] = np.array((cfg2.output_path / "annotations.csv").read_text())
    np.savez_compressed(GOLDEN / "localization_dataset.npz", **out)
    print(f"wrote {GOLDEN / 'localization_dataset.npz'}: {len(names)} files, {res.num_samples} records")


if __name__ == "__main__":
    main()
