"""GPU parity of the drop-in entry point: ``spine_vision_b200.dataset.create_classification_dataset`` on a synthetic
SPIDER tree (MetaImage volumes) + Phenikaa tree (DICOM series folders) vs what the reference's OWN ``create_classification_dataset`` wrote for the same tree
(tests/golden/host_dataset.npz, frozen by oracle/make_golden_host.py: PNG pixels and CSV text), both crop modes, plus the
resume run and the checkpoint path.  Every pixel comes through the C ABI (native MetaImage decode -> K0 -> K3 -> native PNG)."""
import numpy as np
import torch
from PIL import Image

from conftest import GOLDEN
from gpu_util import dev, requires_gpu
from spine_vision_b200 import dataset, synthetic


def _config(base, mode, g, **kw):
    return dataset.ClassificationDatasetConfig(base_path=base, output_name="cls", localization_model_path=None,
                                               crop_size=tuple(int(v) for v in g["crop_size"]),
                                               crop_delta_mm=tuple(float(v) for v in g["delta_mm"]), crop_mode=mode,
                                               last_disc_angle_boost=1.5, device=dev(), **kw)


def _check_tree(out_dir, names, images):
    found = sorted(p.name for p in (out_dir / "images").glob("*.png"))
    assert found == names
    for n, want in zip(names, images):
        pil = Image.open(out_dir / "images" / n)
        assert pil.mode == "L"
        got = np.asarray(pil)
        assert np.array_equal(got, want), f"{n}: {(got != want).sum()} pixels differ"


@requires_gpu
def test_create_classification_dataset_matches_reference_driver(tmp_path):
    g = np.load(GOLDEN / "host_dataset.npz")
    for mode in ("horizontal", "rotated"):
        base = tmp_path / mode
        synthetic.make_spider_tree(base, seed=0)
        synthetic.make_phenikaa_tree(base, seed=0)  # DICOM series folders: the second source of the reference's driver
        cfg = _config(base, mode, g, chunk_series=3 if mode == "horizontal" else 64)  # several GPU batches / one batch
        res = dataset.create_classification_dataset(cfg)
        names = [str(n) for n in g[f"{mode}_names"]]
        assert res.num_samples == len(names) and res.output_path == cfg.output_path
        _check_tree(cfg.output_path, names, g[f"{mode}_images"])
        assert (cfg.output_path / "annotations.csv").read_text() == g[f"{mode}_csv"].item()
        if mode == "horizontal":
            # resume: delete two crops, run again -> only those are recomputed; CSV = recovered (any order) + new (job order)
            for n in g["delete_for_resume"]:
                (cfg.output_path / "images" / str(n)).unlink()
            res2 = dataset.create_classification_dataset(cfg)
            assert res2.num_samples == len(names)
            _check_tree(cfg.output_path, names, g["horizontal_images"])
            got = (cfg.output_path / "annotations.csv").read_text().strip().split("\n")
            want = g["resume_csv"].item().strip().split("\n")
            nd = len(g["delete_for_resume"])
            assert got[0] == want[0] and got[-nd:] == want[-nd:] and sorted(got) == sorted(want)
            # nothing missing -> nothing to do, CSV rebuilt from the label file alone
            res3 = dataset.create_classification_dataset(cfg)
            assert res3.num_samples == len(names) and "0 new" in res3.summary


@requires_gpu
def test_dataset_with_checkpoint_and_unreadable_series(tmp_path):
    """The model path: a LocalizationTrainer-format checkpoint (trainers/base.py:695-706) drives the crops; a corrupt
    volume is skipped like the reference skips a series whose reader raises (spider.py:139-141)."""
    from spine_vision_b200 import cropping, hostio, ops, pipeline, volumes

    base = tmp_path
    pids = synthetic.make_spider_tree(base, n_patients=2, seed=5, in_plane=(160, 150), missing_t1=())
    bad = base / "raw" / "SPIDER" / "images" / f"{pids[1]}_t1.mha"
    bad.write_bytes(bad.read_bytes()[:-100])
    sd = synthetic.random_state_dict("base", seed=3)
    ckpt = base / "model.pt"
    torch.save({"model_state_dict": sd, "epoch": 1}, ckpt)
    cfg = dataset.ClassificationDatasetConfig(base_path=base, output_name="m", localization_model_path=ckpt, model_variant="base",
                                              crop_size=(128, 128), crop_delta_mm=(20.0, 8.0, 9.0, 10.0), device=dev())
    res = dataset.create_classification_dataset(cfg)
    assert res.num_samples == 14  # patient 0: 2 series x 5 levels; patient 1: T1 unreadable, T2 has 4 labelled levels
    assert not list((cfg.output_path / "images").glob(f"spider_{pids[1]}_sag_t1_*"))
    # the same series through the module-level mirrors, one at a time (the reference's call sequence)
    model = cropping.load_localization_model(ckpt, "base", dev())
    vol = hostio.read_medical_image(base / "raw" / "SPIDER" / "images" / f"{pids[0]}_t2.mha")
    pool, sp = volumes.midplane_resample([vol.array], [vol.spacing], [vol.direction], dev(), integer_pixels=[True])
    sl = pool.data[: pool.shapes[0][0] * pool.shapes[0][1]].reshape(pool.shapes[0]).cpu().numpy()
    locs = cropping.predict_ivd_locations(model, sl, dev(), (512, 512))
    ctx = cropping.CropContext(sl, locs, (128, 128), cropping.mm_to_pixels(cfg.crop_delta_mm, sp[0]), device=dev())
    for lvl in range(1, 6):
        got = np.asarray(Image.open(cfg.output_path / "images" / f"spider_{pids[0]}_sag_t2_L{lvl}.png"))
        assert np.array_equal(got, ctx.crop(lvl - 1))


@requires_gpu
def test_cli_entry_point(tmp_path, capsys):
    """``python -m spine_vision_b200 dataset classification ...`` end to end: same tree, same CSV as the reference's driver."""
    import spine_vision_b200.__main__ as cli

    g = np.load(GOLDEN / "host_dataset.npz")
    synthetic.make_spider_tree(tmp_path, seed=0)
    synthetic.make_phenikaa_tree(tmp_path, seed=0)
    rc = cli.main(["dataset", "classification", "--base-path", str(tmp_path), "--output-name", "cls", "--crop-size", "128", "128",
                   "--crop-delta-mm", *[str(float(v)) for v in g["delta_mm"]], "--last-disc-angle-boost", "1.5", "--device", dev()])
    assert rc == 0 and "39 records" in capsys.readouterr().out
    assert (tmp_path / "processed" / "cls" / "annotations.csv").read_text() == g["horizontal_csv"].item()
    _check_tree(tmp_path / "processed" / "cls", [str(n) for n in g["horizontal_names"]], g["horizontal_images"])
