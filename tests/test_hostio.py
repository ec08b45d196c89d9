"""CPU: the host input / output stage -- native MetaImage decoder vs oracle/metaimage.py, native PNG
encoder vs Pillow's decoder, and the dataset driver's host logic (work list, records, CSV, resume) vs the CSVs the
reference's OWN driver wrote (tests/golden/host_dataset.npz, oracle/make_golden_host.py).  No GPU work here."""
import io
from pathlib import Path
import zlib

import numpy as np
import pytest
from PIL import Image

from conftest import GOLDEN
from oracle import dicom, metaimage
from spine_vision_b200 import _lib, dataset, hostio, synthetic


@pytest.mark.parametrize("dtype", ["int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "float32", "float64"])
@pytest.mark.parametrize("layout", ["local-z", "local-raw", "mhd-raw", "mhd-zraw", "big-endian"])
def test_metaimage_reader_matches_oracle(tmp_path, dtype, layout):
    rng = np.random.default_rng(zlib.crc32(f"{dtype}-{layout}".encode()))
    info = np.iinfo(dtype) if np.issubdtype(np.dtype(dtype), np.integer) else None
    lo, hi = (max(info.min, -(2**23)), min(info.max, 2**23)) if info else (-1000, 1000)
    arr = (rng.integers(lo, hi, size=(5, 13, 11)) if info else rng.normal(0, 300, size=(5, 13, 11))).astype(dtype)
    direction = (0.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, -1.0, 0.0)
    path = tmp_path / ("v.mhd" if layout.startswith("mhd") else "v.mha")
    synthetic.write_metaimage(path, arr, (0.7, 0.65, 4.4), direction, origin=(1.5, -2.0, 3.25), compressed=layout in ("local-z", "mhd-zraw", "big-endian"),
                              big_endian=layout == "big-endian", separate_raw=layout.startswith("mhd"))
    want = metaimage.read(path)
    got = hostio.read_medical_image(path)
    assert got.array.dtype == np.float32 and got.array.shape == arr.shape
    assert np.array_equal(got.array, want.array.astype(np.float32)) and np.array_equal(want.array, arr)
    assert got.GetSpacing() == want.GetSpacing() == (0.7, 0.65, 4.4) and got.GetOrigin() == want.GetOrigin()
    assert np.array_equal(np.array(got.GetDirection()), np.array(want.GetDirection())) and got.GetDirection() == direction
    assert got.GetSize() == want.GetSize() == (11, 13, 5)
    assert got.integer_pixels == (info is not None)


def test_metaimage_2d_and_errors(tmp_path):
    a2 = np.arange(12, dtype=np.uint16).reshape(3, 4)
    synthetic.write_metaimage(tmp_path / "p.mha", a2, (0.5, 0.25), compressed=False)
    v = hostio.read_medical_image(tmp_path / "p.mha")
    assert v.array.shape == (1, 3, 4) and np.array_equal(v.array[0], a2) and v.spacing[:2] == (0.5, 0.25)
    with pytest.raises(FileNotFoundError):
        hostio.read_medical_image(tmp_path / "missing.mha")  # io/readers.py:143-144
    (tmp_path / "x.bin").write_bytes(b"123")
    with pytest.raises(ValueError, match="Unsupported format"):
        hostio.read_medical_image(tmp_path / "x.bin")  # io/readers.py:160-161
    (tmp_path / "v.nii.gz").write_bytes(b"\x1f\x8b")
    with pytest.raises(Exception):
        hostio.read_medical_image(tmp_path / "v.nii.gz")  # a truncated gzip stream: an exception the drivers catch (spider.py:131-133)
    (tmp_path / "bad.mha").write_text("hello\nworld\n")
    with pytest.raises(_lib.SvbError):
        hostio.read_medical_image(tmp_path / "bad.mha")
    # truncated compressed payload
    synthetic.write_metaimage(tmp_path / "t.mha", np.zeros((4, 8, 8), np.int16), (1, 1, 1))
    blob = (tmp_path / "t.mha").read_bytes()
    (tmp_path / "t.mha").write_bytes(blob[:-7])
    with pytest.raises(_lib.SvbError):
        hostio.read_medical_image(tmp_path / "t.mha")


def test_read_volumes_batch_skips_bad_files(tmp_path):
    pids = synthetic.make_spider_tree(tmp_path, n_patients=3, seed=3)
    img = tmp_path / "raw" / "SPIDER" / "images"
    paths = [img / f"{pids[0]}_t2.mha", tmp_path / "nope.mha", img / f"{pids[1]}_t1.mha", tmp_path / "raw" / "SPIDER" / "radiological_gradings.csv"]
    vols, errs = hostio.read_volumes(paths, n_threads=3, pin=False)
    assert [v is not None for v in vols] == [True, False, True, False] and errs[0] is None and "does not exist" in errs[1]
    for p, v in zip(paths, vols):
        if v is not None:
            want = metaimage.read(p)
            assert np.array_equal(v.array, want.array.astype(np.float32)) and v.spacing == want.GetSpacing() and v.integer_pixels


@pytest.mark.parametrize("shape", [(128, 128), (256, 256), (1, 1), (3, 1), (1, 5), (37, 129)])
def test_png_encoder_decodes_identically(shape, tmp_path):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    h, w = shape
    ramp = (np.add.outer(np.arange(h), 2 * np.arange(w)) % 256).astype(np.uint8)
    for img in (rng.integers(0, 256, size=shape, dtype=np.uint8), ramp, np.zeros(shape, np.uint8), np.full(shape, 255, np.uint8)):
        blob = hostio.encode_png(img)
        pil = Image.open(io.BytesIO(blob))
        assert pil.mode == "L" and pil.size == (w, h)  # what ClassificationDataset re-opens (classification.py:288-294)
        assert np.array_equal(np.asarray(pil), img)
    imgs = np.stack([np.roll(ramp, k, axis=0) for k in range(9)])
    paths = [tmp_path / f"c{k}.png" for k in range(9)]
    hostio.write_png_batch(imgs, paths, n_threads=4)
    for k, p in enumerate(paths):
        assert np.array_equal(np.asarray(Image.open(p).convert("L")), imgs[k])
    with pytest.raises(_lib.SvbError):
        hostio.write_png_batch(imgs[:1], [tmp_path / "no_such_dir" / "x.png"])


def test_filename_parsing_and_levels():
    assert dataset.parse_image_filename("spider_12_sag_t2_L3.png") == dataset.ParsedImageInfo("spider", "12", "sag_t2", 3, "spider_12_sag_t2_L3.png")
    p = dataset.parse_image_filename("phenikaa_A_b_01_sag_t1_L5.png")
    assert p is not None and p.patient_id == "A_b_01" and p.series_type == "sag_t1"
    for bad in ("spider_1_ax_t2_L3.png", "other_1_sag_t2_L3.png", "spider_1_sag_t2_L3.jpg", "spider_1_sag_t2_L33.png"):
        assert dataset.parse_image_filename(bad) is None
    assert [dataset.convert_spider_to_phenikaa_level(k) for k in (1, 5)] == [5, 1]  # spider.py:31-42
    row = {"Modic_0": "0", "Modic_1": "0", "Modic_2": "1", "Modic_3": "1", "Pfirrman grade": "4"}
    rec = dataset.make_record("phenikaa", "phenikaa_p_sag_t2_L1.png", "p", 1, "sag_t2", row)
    assert rec.modic == 2 and rec.pfirrmann_grade == 4 and rec.disc_bulging == 0  # phenikaa.py:88-108


def _lines(text):
    return text.strip().split("\n")


def _records(jobs):
    return [dataset.make_record(j.source, dataset.output_filename(j.source, j.patient_id, j.series_type, lvl), j.patient_id, lvl, j.series_type, row)
            for j in jobs for lvl, row in j.levels.items()]


def test_work_list_records_and_csv_match_reference_driver(tmp_path):
    """The host logic alone (no pixels): jobs in the reference's iteration order, records, CSV text == what the
    reference's create_classification_dataset wrote for the same trees; then its resume behaviour."""
    g = np.load(GOLDEN / "host_dataset.npz")
    synthetic.make_spider_tree(tmp_path, seed=0)
    synthetic.make_phenikaa_tree(tmp_path, seed=0)
    cfg = dataset.ClassificationDatasetConfig(base_path=tmp_path, output_name="cls", crop_size=(128, 128))
    assert cfg.spider_path == tmp_path / "raw" / "SPIDER" and cfg.output_path == tmp_path / "processed" / "cls"
    assert cfg.phenikaa_path == tmp_path / "interim" / "Phenikaa"
    sj = dataset.collect_spider_jobs(cfg, set())
    assert [(j.patient_id, j.series_type) for j in sj] == [("1", "sag_t1"), ("1", "sag_t2"), ("4", "sag_t1"), ("4", "sag_t2"), ("7", "sag_t2")]
    assert sorted(sj[-1].levels) == [1, 2, 3, 4]  # SPIDER level 7 -> -1 is dropped, dataset level 5 has no row
    pj = dataset.collect_phenikaa_jobs(cfg, set())
    assert [(j.patient_id, j.series_type, j.path.name) for j in pj] == [("PK00000", "sag_t1", "SAG  T1"), ("PK00000", "sag_t2", "Sag T2"),
                                                                        ("PK00001", "sag_t2", "Sag T2")]  # phenikaa.py:48-65 folder match
    dataset.write_annotations(tmp_path / "a.csv", _records(pj + sj))  # __init__.py:199-214: Phenikaa first, then SPIDER
    assert (tmp_path / "a.csv").read_text() == g["horizontal_csv"].item()
    empty = dataset.ClassificationDatasetConfig(base_path=tmp_path / "nowhere", output_name="x")
    assert dataset.collect_phenikaa_jobs(empty, set()) == [] and dataset.collect_spider_jobs(empty, set()) == []  # labels missing: warn, no jobs

    # resume: every image exists except three -> only those levels are jobs; recovered records come from the label files
    names = [str(n) for n in g["horizontal_names"]]
    deleted = [str(n) for n in g["delete_for_resume"]]
    images = cfg.output_path / "images"
    images.mkdir(parents=True)
    for n in names:
        if n not in deleted:
            (images / n).write_bytes(b"")
    existing = dataset.scan_existing_images(images)
    assert sorted(e.filename for e in existing) == sorted(set(names) - set(deleted))
    have = {f"images/{e.filename}" for e in existing}
    jobs2 = dataset.collect_phenikaa_jobs(cfg, have) + dataset.collect_spider_jobs(cfg, have)
    assert [(j.patient_id, j.series_type, sorted(j.levels)) for j in jobs2] == [("PK00000", "sag_t2", [3]), ("4", "sag_t1", [2]), ("7", "sag_t2", [4])]
    ph, sp = dataset.recover_annotations(existing, cfg.spider_path / "radiological_gradings.csv", cfg.phenikaa_path / "radiological_labels.csv")
    dataset.write_annotations(tmp_path / "b.csv", ph + sp + _records(jobs2))
    got, want = _lines((tmp_path / "b.csv").read_text()), _lines(g["resume_csv"].item())
    # recovered rows follow the directory listing order (spider.py:237: glob), which is filesystem-defined: compare as a set;
    # the new rows come last, in job order
    assert got[0] == want[0] and got[-3:] == want[-3:] and sorted(got) == sorted(want) and len(ph) == 14


@pytest.mark.parametrize("explicit", [True, False])
@pytest.mark.parametrize("dtype,rescale", [("uint16", None), ("int16", None), ("uint8", None), ("uint16", (2.0, -1024.0)), ("uint16", (0.5, 0.25))])
def test_dicom_series_reader_matches_oracle(tmp_path, explicit, dtype, rescale):
    """Native DICOM slice decoder + the ITK series conventions (hostio.read_dicom_series) vs oracle/dicom.py: scrambled
    file names, a second series with a larger UID and a stray text file in the folder, sequences of undefined length,
    oblique orientation, rescale."""
    rng = np.random.default_rng(zlib.crc32(f"{dtype}-{explicit}-{rescale}".encode()))
    info = np.iinfo(dtype)
    arr = rng.integers(max(info.min, -2000), min(info.max, 4000), size=(6, 17, 23)).astype(dtype)
    th = 0.2
    direction = np.array([[0.0, 0.0, 1.0], [np.cos(th), np.sin(th), 0.0], [np.sin(th), -np.cos(th), 0.0]])  # columns = axes, oblique
    folder = tmp_path / "Sag T2"
    synthetic.write_dicom_series(folder, arr, (0.61, 0.73, 3.3), direction.ravel(), origin=(12.5, -40.0, 7.0), series_uid="1.2.3.50",
                                 explicit=explicit, shuffle_seed=3, rescale=rescale)
    synthetic.write_dicom_series(tmp_path / "other", arr[:2], (0.61, 0.73, 3.3), direction.ravel(), series_uid="1.2.3.77", shuffle_seed=None)
    for f in (tmp_path / "other").iterdir():
        f.rename(folder / f"A_{f.name}")  # sorts in front by name, behind by series id
    (folder / "readme.txt").write_text("x")
    want = dicom.read_series(folder)
    got = hostio.read_medical_image(folder)
    assert got.array.shape == (6, 17, 23) and got.meta["series_uid"] == "1.2.3.50"
    assert np.array_equal(got.array, want.array.astype(np.float32))
    assert np.allclose(got.spacing, want.GetSpacing(), rtol=0, atol=1e-12) and np.allclose(got.spacing, (0.61, 0.73, 3.3), atol=1e-9)
    assert np.allclose(got.direction, want.GetDirection(), atol=1e-12) and np.allclose(got.origin, want.GetOrigin(), atol=1e-12)
    assert got.integer_pixels == (rescale is None or rescale == (2.0, -1024.0))
    # slices come back ordered along the normal (row x col), whatever the file names say
    normal = np.cross(direction[:, 0], direction[:, 1])
    stack = np.array(got.direction).reshape(3, 3)[:, 2]
    assert abs(float(np.dot(stack, normal))) > 0.999
    vals = arr.astype(np.float64) * (rescale[0] if rescale else 1.0) + (rescale[1] if rescale else 0.0)
    order = slice(None) if float(np.dot(direction[:, 2], normal)) > 0 else slice(None, None, -1)
    assert np.array_equal(got.array, vals[order].astype(np.float32))


def test_dicom_errors(tmp_path):
    (tmp_path / "empty").mkdir()
    with pytest.raises(ValueError, match="No DICOM series found"):  # io/readers.py:66-67
        hostio.read_medical_image(tmp_path / "empty")
    (tmp_path / "junk").mkdir()
    (tmp_path / "junk" / "a.dcm").write_bytes(b"\x00" * 200)
    with pytest.raises(ValueError, match="No DICOM series found"):
        hostio.read_medical_image(tmp_path / "junk")


def test_metaimage_writer_is_plain_zlib(tmp_path):
    arr = np.arange(2 * 3 * 4, dtype=np.int16).reshape(2, 3, 4)
    synthetic.write_metaimage(tmp_path / "w.mha", arr, (1, 2, 3))
    blob = (tmp_path / "w.mha").read_bytes()
    head, data = blob.split(b"ElementDataFile = LOCAL\n", 1)
    assert b"DimSize = 4 3 2" in head and b"ElementType = MET_SHORT" in head
    assert np.array_equal(np.frombuffer(zlib.decompress(data), dtype="<i2").reshape(2, 3, 4), arr)


_RECORDS_WORKER = r"""
import sys, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from spine_vision_b200 import dataset, synthetic
from pathlib import Path
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
cfg = dataset.ClassificationDatasetConfig(base_path=Path(sys.argv[4]), output_name="cls")
jobs = dataset.collect_spider_jobs(cfg, set())
mine = jobs[rank::2]                                   # the split create_classification_dataset(rank, world_size) uses
recs = [dataset.make_record(j.source, dataset.output_filename(j.source, j.patient_id, j.series_type, l), j.patient_id, l, j.series_type, row)
        for j in mine for l, row in j.levels.items()]
allr, failures = dataset.gather_records(recs, failure="boom" if rank == 1 else None)
assert failures == ["boom"]
want = [dataset.make_record(j.source, dataset.output_filename(j.source, j.patient_id, j.series_type, l), j.patient_id, l, j.series_type, row)
        for part in (jobs[0::2], jobs[1::2]) for j in part for l, row in j.levels.items()]
assert allr == want and len(allr) == 24, (len(allr), len(want))
# the plan (resume scan + job list) is made on rank 0 and broadcast: rank 1 arrives late and, by then, sees a PNG another rank
# "has already written" -- it must still get rank 0's job list, not a shorter one of its own (ADVICE r01, dataset.py:407)
import time
images = cfg.output_path / "images"
images.mkdir(parents=True, exist_ok=True)
if rank == 1:
    time.sleep(1.0)
    j0 = jobs[0]
    (images / dataset.output_filename(j0.source, j0.patient_id, j0.series_type, 1)).write_bytes(b"x")
recovered, planned = dataset.plan_dataset(cfg, images, rank, 2)
assert recovered == [] and planned == jobs, (rank, len(planned), len(jobs))
dist.barrier()
if rank == 1:
    assert len(dataset._plan_local(cfg, images)[1][0].levels) == len(jobs[0].levels) - 1   # what a per-rank scan would have seen
dist.destroy_process_group()
print("ok", rank)
"""


def test_gather_records_gloo_world2(tmp_path):
    """Multi-GPU dataset creation shards the job list by rank and gathers the records once (no data-path collective)."""
    import socket
    import subprocess
    import sys

    from conftest import ROOT

    synthetic.make_spider_tree(tmp_path, seed=0)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "w.py"
    script.write_text(_RECORDS_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), str(ROOT), str(port), str(r), str(tmp_path)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_ragged_png_writer_and_single_dicom_reader(tmp_path):
    """Host pieces of the localization dataset builder: images of different sizes from one pool -> PNG (threaded), and
    single DICOM slices decoded in a batch with per-file errors (datasets/localization.py:262-271)."""
    rng = np.random.default_rng(5)
    shapes = [(40, 33), (7, 120), (64, 64)]
    offs, o = [], 0
    for h, w in shapes:
        offs.append(o)
        o += (h * w + 3) // 4 * 4
    pool = rng.integers(0, 256, size=o, dtype=np.uint8)
    paths = [tmp_path / f"r{k}.png" for k in range(3)]
    hostio.write_png_ragged(pool, offs, shapes, paths, n_threads=2)
    for (h, w), off, p in zip(shapes, offs, paths):
        assert np.array_equal(np.asarray(Image.open(p)), pool[off : off + h * w].reshape(h, w))
    a = rng.integers(0, 4000, size=(30, 22)).astype(np.uint16)
    b = rng.integers(-500, 500, size=(16, 48)).astype(np.int16)
    synthetic.write_dicom_slice(tmp_path / "a.dcm", a, (0, 0, 0), (1, 0, 0), (0, 1, 0), (0.5, 0.5), "1.2.3", 1, rescale=(2.0, -10.0))
    synthetic.write_dicom_slice(tmp_path / "b.dcm", b, (0, 0, 0), (1, 0, 0), (0, 1, 0), (0.5, 0.5), "1.2.3", 2, explicit=False)
    (tmp_path / "c.dcm").write_bytes(b"nope")
    arrays, errors = hostio.read_dicom_files([tmp_path / "a.dcm", tmp_path / "c.dcm", tmp_path / "b.dcm", tmp_path / "missing.dcm"])
    assert np.array_equal(arrays[0], a.astype(np.float32) * 2 - 10) and np.array_equal(arrays[2], b.astype(np.float32))
    assert arrays[1] is None and arrays[3] is None and errors[0] is None and errors[1] and errors[3]
    assert np.array_equal(arrays[0], dicom.read_slice(tmp_path / "a.dcm")["px"].astype(np.float64) * 2 - 10)


def test_localization_series_lookup(tmp_path):
    from spine_vision_b200 import localization_dataset as loc

    synthetic.make_localization_tree(tmp_path, seed=0)
    m = loc.load_series_mapping(tmp_path / "raw" / "RSNA" / "train_series_descriptions.csv")
    assert m[101]["Sagittal T1"] == 1001 and loc.get_series_type(1002, 101, m) == "Sagittal T2/STIR"  # datasets/rsna.py:7-61
    assert loc.get_series_type(1, 101, m) is None and loc.get_series_type(1001, 999, m) is None
    cfg = loc.LocalizationDatasetConfig(base_path=tmp_path)
    assert cfg.lumbar_coords_path == tmp_path / "raw" / "Lumbar Coords" and cfg.rsna_path == tmp_path / "raw" / "RSNA"
    with pytest.raises(ValueError, match="empty records"):
        loc.write_records_csv([], tmp_path / "x.csv")  # io/tabular.py:28-29


def test_cli_flags_build_the_reference_config():
    """``python -m spine_vision_b200 dataset classification ...``: the flags tyro derives from the reference's config
    (cli/__init__.py:30-56) -- kebab-case, tuple fields as several values, --flag / --no-flag booleans."""
    import argparse

    import spine_vision_b200.__main__ as cli

    p = argparse.ArgumentParser()
    cli._add_fields(p, dataset.ClassificationDatasetConfig)
    ns = vars(p.parse_args(["--base-path", "/data", "--localization-model-path", "/w/best_model.pt", "--crop-size", "128", "128",
                            "--crop-delta-mm", "50", "20", "30", "30", "--crop-mode", "rotated", "--no-include-phenikaa",
                            "--model-variant", "v2_base", "--last-disc-angle-boost", "1.5"]))
    cfg = dataset.ClassificationDatasetConfig(**{k: tuple(v) if isinstance(v, list) else v for k, v in ns.items()})
    assert cfg.crop_size == (128, 128) and cfg.crop_delta_mm == (50.0, 20.0, 30.0, 30.0) and cfg.crop_mode == "rotated"
    assert cfg.include_phenikaa is False and cfg.include_spider is True and cfg.model_variant == "v2_base"
    assert str(cfg.localization_model_path) == "/w/best_model.pt" and cfg.last_disc_angle_boost == 1.5
    assert dataset.ClassificationDatasetConfig(**vars(p.parse_args([]))) == dataset.ClassificationDatasetConfig()  # defaults = the reference's
    with pytest.raises(SystemExit):
        p.parse_args(["--crop-mode", "diagonal"])


def test_dicom_parser_survives_truncated_and_corrupt_files(tmp_path):
    """Every prefix of a valid slice file, zero-length US elements and random byte flips: the native parser answers with an
    error (or a decoded slice), never with a crash or an out-of-bounds read -- the reference's ``except`` around
    ``sitk.ReadImage`` (datasets/localization.py:262-270) turns such files into skipped images."""
    import struct

    px = (np.arange(12 * 10, dtype=np.uint16).reshape(12, 10) * 7) % 4000
    good = tmp_path / "good.dcm"
    synthetic.write_dicom_slice(good, px, (0, 0, 0), (1, 0, 0), (0, 1, 0), (0.5, 0.5), "1.2.3.9", 1)
    blob = good.read_bytes()
    arrays, errors = hostio.read_dicom_files([good])
    assert errors == [None] and np.array_equal(arrays[0], px.astype(np.float32))
    paths = []
    for cut in list(range(0, 200)) + list(range(200, len(blob), 7)):
        p = tmp_path / f"cut_{cut}.dcm"
        p.write_bytes(blob[:cut])
        paths.append(p)
    rows_elem = struct.pack("<HH", 0x0028, 0x0010) + b"US" + struct.pack("<H", 2)
    at = blob.index(rows_elem)
    zero_len = blob[:at] + struct.pack("<HH", 0x0028, 0x0010) + b"US" + struct.pack("<H", 0) + blob[at + len(rows_elem) + 2 :]
    (tmp_path / "zero_len.dcm").write_bytes(zero_len)
    paths.append(tmp_path / "zero_len.dcm")
    rng = np.random.default_rng(0)
    for k in range(200):
        b = bytearray(blob)
        for pos in rng.integers(132, len(blob) - px.nbytes, size=3):
            b[pos] = int(rng.integers(0, 256))
        p = tmp_path / f"flip_{k}.dcm"
        p.write_bytes(bytes(b))
        paths.append(p)
    arrays, errors = hostio.read_dicom_files(paths)
    n_cut = len(paths) - 201
    assert all(a is None and e for a, e in zip(arrays[:n_cut], errors[:n_cut]))  # every strict prefix is an error
    assert arrays[n_cut] is None and errors[n_cut]                               # no Rows value -> error, not 0 x cols
    for a, e in zip(arrays[n_cut + 1 :], errors[n_cut + 1 :]):
        assert (a is None) == bool(e)
        assert a is None or (a.ndim == 2 and a.dtype == np.float32)


def test_native_decoders_under_address_sanitizer(tmp_path):
    """tests/native/fuzz_hostio.cpp: svb_hostio.cpp itself rebuilt with -fsanitize=address,undefined and driven over mutants
    (truncations, byte flips, extreme 32-bit fields, rewritten digits, dropped / doubled chunks) of valid MetaImage and DICOM
    files, plus the PNG encoder at and below its buffer bound.  A sanitizer report aborts the run."""
    import shutil
    import subprocess

    root = Path(__file__).resolve().parent.parent
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = tmp_path / "fuzz_hostio"
    cc = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-o", str(exe),
                         str(root / "tests/native/fuzz_hostio.cpp"), str(root / "spine_vision_b200/csrc/svb_hostio.cpp"), "-lz", "-lpthread"],
                        capture_output=True, text=True)
    if cc.returncode != 0 and "asan" in cc.stderr.lower():
        pytest.skip("libasan not installed")
    assert cc.returncode == 0, cc.stderr[-2000:]
    seeds = tmp_path / "seeds"
    work = tmp_path / "work"
    seeds.mkdir()
    work.mkdir()
    rng = np.random.default_rng(0)
    vol = rng.integers(-100, 3000, size=(5, 14, 11)).astype(np.int16)
    synthetic.write_metaimage(seeds / "a.mha", vol, (0.6, 0.7, 3.3), compressed=True)
    synthetic.write_metaimage(seeds / "b.mha", vol.astype(np.float32), (0.6, 0.7, 3.3), compressed=False)
    synthetic.write_metaimage(seeds / "c.mhd", vol.astype(np.uint16), (0.6, 0.7, 3.3), compressed=True, separate_raw=True)
    synthetic.write_metaimage(seeds / "d.mha", vol.astype(np.float64), (0.6, 0.7, 3.3), compressed=False, big_endian=True)
    px = (np.arange(12 * 10, dtype=np.uint16).reshape(12, 10) * 7) % 4000
    synthetic.write_dicom_slice(seeds / "e.dcm", px, (0, 0, 0), (1, 0, 0), (0, 1, 0), (0.5, 0.5), "1.2.3.9", 1)
    synthetic.write_dicom_slice(seeds / "i.dcm", px.astype(np.int16), (0, 0, 0), (1, 0, 0), (0, 1, 0), (0.5, 0.5), "1.2.3.9", 1,
                                explicit=False, rescale=(2.0, -1024.0))
    # the encapsulated decoders (RLE Lossless, JPEG Lossless with a restart interval, two fragments) get the same treatment
    synthetic.write_dicom_slice(seeds / "r.dcm", px, (0, 0, 0), (1, 0, 0), (0, 1, 0), (0.5, 0.5), "1.2.3.9", 1, compress="rle")
    synthetic.write_dicom_slice(seeds / "j.dcm", px.astype(np.int16), (0, 0, 0), (1, 0, 0), (0, 1, 0), (0.5, 0.5), "1.2.3.9", 1,
                                compress="jpeg", codec_kw={"predictor": 4, "restart_lines": 3}, fragments=2)
    files = [str(seeds / n) for n in ("a.mha", "b.mha", "c.mhd", "d.mha", "e.dcm", "i.dcm", "r.dcm", "j.dcm")]
    run = subprocess.run([str(exe), str(work), "1500", *files], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, (run.stdout + run.stderr)[-3000:]
    assert "no sanitizer report" in run.stdout


@pytest.mark.parametrize("dtype", ["int16", "uint8", "float32", "float64"])
@pytest.mark.parametrize("layout", ["local_zlib", "local_raw", "mhd_zraw", "big_endian_raw"])
def test_midplane_only_decode_equals_whole_decode_where_it_matters(tmp_path, dtype, layout):
    """``read_volumes(midplane_only=True)`` (the dataset driver's reader): for a sagittal acquisition only the two source slices
    around the middle plane are decoded -- ``volumes.plan_midplane`` must cut the same slab out of it as out of the whole
    volume; a volume in another orientation is decoded whole.  Even and odd slice counts, 1 and 2 slices."""
    from spine_vision_b200 import volumes

    rng = np.random.default_rng(zlib.crc32(f"{dtype}-{layout}".encode()))
    kw = dict(compressed=layout in ("local_zlib", "mhd_zraw"), separate_raw=layout == "mhd_zraw", big_endian=layout == "big_endian_raw")
    sag = (0.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, -1.0, 0.0)  # image z = Left: the slowest array axis holds the sagittal slices
    paths, dirs = [], []
    for k, nz in enumerate((15, 14, 2, 1, 9)):
        arr = (rng.random((nz, 23, 31)) * 200).astype(dtype)
        d = sag if k != 4 else None  # the last one is axial (identity direction): Left = the fastest array axis
        p = tmp_path / (f"v{k}.mhd" if layout == "mhd_zraw" else f"v{k}.mha")
        synthetic.write_metaimage(p, arr, (0.7, 0.7, 4.0), direction=d, **kw)
        paths.append(p); dirs.append(d)
    whole, e0 = hostio.read_volumes(paths)
    part, e1 = hostio.read_volumes(paths, midplane_only=True)
    assert e0 == e1 == [None] * 5
    for k, (a, b) in enumerate(zip(whole, part)):
        z0, z1 = b.meta["decoded_z"]
        nz = a.array.shape[0]
        if k == 4:
            assert (z0, z1) == (0, nz) and np.array_equal(a.array, b.array)
            continue
        assert z1 - z0 == min(2, nz) and a.meta["decoded_z"] == (0, nz)
        assert np.array_equal(a.array[z0:z1], b.array[z0:z1])
        pa = volumes.plan_midplane(a.array, a.spacing, a.direction, integer_pixels=a.integer_pixels)
        pb = volumes.plan_midplane(b.array, b.spacing, b.direction, integer_pixels=b.integer_pixels)
        assert pa.desc == pb.desc and pa.out_hw == pb.out_hw and np.array_equal(pa.slab, pb.slab)
    # a truncated compressed stream that still covers the slab decodes; one that ends in front of it is an error
    if layout == "local_zlib":
        blob = paths[0].read_bytes()
        cut = tmp_path / "cut.mha"
        cut.write_bytes(blob[: len(blob) // 4])
        vols, errs = hostio.read_volumes([cut], midplane_only=True)
        assert vols == [None] and errs[0]


@pytest.mark.parametrize("n_slices", [15, 14, 2, 1])
def test_dicom_series_midplane_only_feeds_k0_the_same_slab(tmp_path, n_slices):
    """``read_medical_image(folder, midplane_only=True)`` (what the dataset driver calls for the Phenikaa series): headers of
    every file are read, pixel data only of the two slices around the middle sagittal plane -- ``volumes.plan_midplane`` cuts the
    same slab and descriptor out of it as out of the fully decoded series; an axial stack is decoded whole."""
    from spine_vision_b200 import volumes

    rng = np.random.default_rng(n_slices)
    arr = rng.integers(0, 3000, size=(n_slices, 19, 27)).astype(np.uint16)
    sag = np.array([[0.0, 0.0, 1.0], [1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])  # columns: x -> P, y -> I, z (stack) -> L
    synthetic.write_dicom_series(tmp_path / "sag", arr, (0.7, 0.7, 4.0), sag.ravel(), series_uid="1.2.3.60", shuffle_seed=1)
    whole = hostio.read_medical_image(tmp_path / "sag")
    part = hostio.read_medical_image(tmp_path / "sag", midplane_only=True)
    z0, z1 = part.meta["decoded_z"]
    assert whole.meta["decoded_z"] == (0, n_slices) and z1 - z0 == min(2, n_slices)
    assert (part.spacing, part.direction, part.origin, part.integer_pixels) == (whole.spacing, whole.direction, whole.origin, whole.integer_pixels)
    assert np.array_equal(part.array[z0:z1], whole.array[z0:z1])
    pa = volumes.plan_midplane(whole.array, whole.spacing, whole.direction, integer_pixels=whole.integer_pixels)
    pb = volumes.plan_midplane(part.array, part.spacing, part.direction, integer_pixels=part.integer_pixels)
    assert pa.desc == pb.desc and pa.out_hw == pb.out_hw and np.array_equal(pa.slab, pb.slab)
    synthetic.write_dicom_series(tmp_path / "ax", arr, (0.7, 0.7, 4.0), np.eye(3).ravel(), series_uid="1.2.3.61", shuffle_seed=2)
    ax = hostio.read_medical_image(tmp_path / "ax", midplane_only=True)
    assert ax.meta["decoded_z"] == (0, n_slices) and np.array_equal(ax.array, hostio.read_medical_image(tmp_path / "ax").array)


def test_corrupt_header_skips_the_series_not_the_chunk(tmp_path):
    """ADVICE r01 (hostio.py:262): one corrupt header must cost one series (spider.py:131-133 except -> continue), never the chunk
    or the process.  DimSize beyond what the file can hold (60000^3), a non-numeric / negative / NaN DimSize, and a compressed
    stream 1000x too short for its header are format errors for THAT file; the good volume beside them still decodes."""
    good = np.arange(3 * 8 * 8, dtype=np.int16).reshape(3, 8, 8)
    synthetic.write_metaimage(tmp_path / "good.mha", good, (1.0, 1.0, 4.0))
    blob = (tmp_path / "good.mha").read_bytes()
    head, data = blob.split(b"ElementDataFile = LOCAL\n", 1)
    cases = {"huge": b"DimSize = 60000 60000 60000", "neg": b"DimSize = -4 8 3", "nan": b"DimSize = nan 8 3",
             "big": b"DimSize = 1e300 8 3", "more": b"DimSize = 4096 4096 64"}
    for name, dim in cases.items():
        lines = [dim if ln.startswith(b"DimSize") else ln for ln in head.split(b"\n")]
        (tmp_path / f"{name}.mha").write_bytes(b"\n".join(lines) + b"ElementDataFile = LOCAL\n" + data)
    paths = [tmp_path / "huge.mha", tmp_path / "good.mha"] + [tmp_path / f"{n}.mha" for n in ("neg", "nan", "big", "more")]
    vols, errs = hostio.read_volumes(paths, n_threads=2)
    assert vols[1] is not None and errs[1] is None and np.array_equal(vols[1].array, good.astype(np.float32))
    for i in (0, 2, 3, 4, 5):
        assert vols[i] is None and errs[i], (i, errs[i])
    for p in paths[:1] + paths[2:]:
        with pytest.raises(Exception):
            hostio.read_medical_image(p)


@pytest.mark.parametrize("name", ["u16", "i16", "u8", "noise"])
def test_compressed_dicom_rle_and_jpeg_lossless(tmp_path, name):
    """VERDICT r01 missing #1: the lossless ENCAPSULATED transfer syntaxes GDCM decodes for the reference -- RLE Lossless
    (1.2.840.10008.1.2.5, PS3.5 Annex G) and JPEG Lossless process 14 (1.2.840.10008.1.2.4.57 / .70, T.81 Annex H: all seven
    predictors, point transform, restart intervals, frames cut into several fragments).  Lossless means exact: the native decoder
    must return the very pixels that were encoded, and agree with the plain-Python restatement in oracle/dicom.py
    (parity with GDCM itself stays unpinned: it is not in the image)."""
    from oracle import dicom as od

    rng = np.random.default_rng(1)
    base = synthetic.make_iso_slice(5, 48, 40)
    img = {"u16": np.rint(base * 20).astype(np.uint16), "i16": (np.rint(base) - 300).astype(np.int16),
           "u8": np.rint(base * 255 / base.max()).astype(np.uint8), "noise": rng.integers(0, 65536, (48, 40)).astype(np.uint16)}[name]
    cases = [("rle", {}, 1), ("rle", {}, 3)] + [("jpeg", {"predictor": p}, 1) for p in range(1, 8)]
    cases += [("jpeg", {"predictor": 4, "restart_lines": 5}, 2), ("jpeg57", {"predictor": 6, "restart_lines": 1}, 1)]
    if img.dtype != np.uint8:
        cases.append(("jpeg", {"predictor": 1, "point_transform": 2}, 1))
    files, want = [], []
    for k, (comp, kw, frags) in enumerate(cases):
        px = (img >> 2 << 2).astype(img.dtype) if kw.get("point_transform") else img
        f = tmp_path / f"{k:02d}.dcm"
        synthetic.write_dicom_slice(f, px, (0, 0, float(k)), (0, 1, 0), (0, 0, -1), (0.6, 0.6), "1.2.3", k + 1, compress=comp,
                                    codec_kw=kw, fragments=frags, rescale=(2.0, -100.0) if k == 3 else None)
        files.append(f)
        want.append(px.astype(np.float64) * 2.0 - 100.0 if k == 3 else px.astype(np.float64))
        assert np.array_equal(od.read_slice(f)["px"], px)
    arrs, errs = hostio.read_dicom_files(files, n_threads=3)
    for k, f in enumerate(files):
        assert errs[k] is None, (cases[k], errs[k])
        assert np.array_equal(arrs[k], want[k].astype(np.float32)), cases[k]


def test_compressed_dicom_series_and_refusals(tmp_path):
    """A whole RLE / JPEG-lossless series folder reads like the native one (same volume, geometry, midplane slab); what is NOT
    decoded is refused with a message, never returned wrong: JPEG 2000, MONOCHROME1; BitsStored < BitsAllocated is masked /
    sign-extended; every truncation of a compressed file is an error, never a crash."""
    vol = (np.rint(synthetic.make_volume(4, 6, 40, 36)[0]) - 200).astype(np.int16)
    sp, d = (0.7, 0.7, 4.0), (0.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, -1.0, 0.0)
    synthetic.write_dicom_series(tmp_path / "native", vol, sp, d)
    ref_v = hostio.read_medical_image(tmp_path / "native")
    for comp in ("rle", "jpeg"):
        synthetic.write_dicom_series(tmp_path / comp, vol, sp, d, compress=comp)
        v = hostio.read_medical_image(tmp_path / comp)
        assert np.array_equal(v.array, ref_v.array) and v.spacing == ref_v.spacing and v.direction == ref_v.direction
        assert v.integer_pixels and v.pixel_kind == ref_v.pixel_kind
        slab = hostio.read_medical_image(tmp_path / comp, midplane_only=True)
        lo, hi = slab.meta["decoded_z"]
        assert np.array_equal(slab.array[lo:hi], ref_v.array[lo:hi])
    px = vol[0]
    args = ((0, 0, 0), (0, 1, 0), (0, 0, -1), (0.6, 0.6), "1.2.9", 1)
    synthetic.write_dicom_slice(tmp_path / "j2k.dcm", px, *args, compress="jpeg2000")
    synthetic.write_dicom_slice(tmp_path / "mono1.dcm", px, *args, photometric=b"MONOCHROME1 ")
    arrs, errs = hostio.read_dicom_files([tmp_path / "j2k.dcm", tmp_path / "mono1.dcm"])
    assert arrs == [None, None] and "1.2.840.10008.1.2.4.90" in errs[0] and "MONOCHROME1" in errs[1]
    # 12 bits stored in 16: the four high bits are junk (unsigned: masked; signed: sign-extended from bit 11)
    junk = (np.arange(40 * 36, dtype=np.int64).reshape(40, 36) * 37) % 4096
    u = (junk | 0xA000).astype(np.uint16)
    synthetic.write_dicom_slice(tmp_path / "u12.dcm", u, *args, bits_stored=12)
    synthetic.write_dicom_slice(tmp_path / "s12.dcm", u.view(np.int16), *args, bits_stored=12, compress="jpeg")
    arrs, errs = hostio.read_dicom_files([tmp_path / "u12.dcm", tmp_path / "s12.dcm"])
    assert errs == [None, None]
    assert np.array_equal(arrs[0], junk.astype(np.float32))
    assert np.array_equal(arrs[1], np.where(junk >= 2048, junk - 4096, junk).astype(np.float32))
    # truncations and byte flips of compressed files: an error or a decode, never a crash
    rng = np.random.default_rng(3)
    for comp in ("rle", "jpeg"):
        blob = (tmp_path / comp / sorted(p.name for p in (tmp_path / comp).iterdir())[0]).read_bytes()
        start = blob.index(b"\xe0\x7f\x10\x00")
        for cut in list(range(start, len(blob), 97)) + [len(blob) - 1]:
            (tmp_path / "t.dcm").write_bytes(blob[:cut])
            a, e = hostio.read_dicom_files([tmp_path / "t.dcm"])
            assert a[0] is None and e[0]
        for _ in range(60):
            m = bytearray(blob)
            for pos in rng.integers(start, len(blob), 3):
                m[pos] = int(rng.integers(0, 256))
            (tmp_path / "m.dcm").write_bytes(bytes(m))
            hostio.read_dicom_files([tmp_path / "m.dcm"])


def test_nifti_nrrd_and_single_dicom_readers(tmp_path):
    """VERDICT r01 missing #1: ``read_medical_image`` dispatches NIfTI, NRRD and single ``.dcm`` files too (io/readers.py:23-28,
    76-126, 128-161).  Arrays round-trip exactly for every pixel type; the geometry follows ITK's documented conventions (RAS
    files come back as LPS); parity with SimpleITK itself is UNPINNED (absent from the image)."""
    rng = np.random.default_rng(2)
    d_sag = (0.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, -1.0, 0.0)
    th = np.deg2rad(12.0)
    d_obl = (np.cos(th), -np.sin(th), 0.0, np.sin(th), np.cos(th), 0.0, 0.0, 0.0, 1.0)
    sp, org = (0.6, 0.75, 3.5), (-12.5, 40.25, 7.0)
    for dt in ("int16", "uint16", "uint8", "float32", "int32"):
        vol = (rng.random((5, 9, 7)) * 200).astype(dt)
        for k, (d, kw) in enumerate([(None, {}), (d_sag, {}), (d_obl, {}), (d_obl, {"use_sform": True}), (d_sag, {"big_endian": True})]):
            f = tmp_path / f"n_{dt}_{k}.nii{'.gz' if k % 2 else ''}"
            synthetic.write_nifti(f, vol, sp, d, org, **kw)
            v = hostio.read_medical_image(f)
            assert np.array_equal(v.array, vol.astype(np.float32)) and v.integer_pixels == (dt != "float32")
            assert np.allclose(v.spacing, sp, atol=1e-6) and np.allclose(v.origin, org, atol=1e-5)
            assert np.allclose(np.array(v.direction).reshape(3, 3), np.eye(3) if d is None else np.array(d).reshape(3, 3), atol=1e-6)
        for k, (d, kw) in enumerate([(None, {}), (d_sag, {"gz": True}), (d_obl, {"space": "right-anterior-superior"}),
                                     (d_sag, {"detached": True, "gz": True})]):
            f = tmp_path / f"r_{dt}_{k}.nrrd"
            synthetic.write_nrrd(f, vol, sp, d, org, **kw)
            v = hostio.read_medical_image(f)
            assert np.array_equal(v.array, vol.astype(np.float32))
            assert np.allclose(v.spacing, sp) and np.allclose(v.origin, org)
            assert np.allclose(np.array(v.direction).reshape(3, 3), np.eye(3) if d is None else np.array(d).reshape(3, 3))
    # the conventions on a case written out by hand: identity qform, offsets (10, 20, 30) mm in RAS -> LPS origin (-10, -20, 30),
    # direction diag(-1, -1, 1); scl_slope / scl_inter make the image float
    import struct

    h = bytearray(352)
    struct.pack_into("<i", h, 0, 348)
    struct.pack_into("<8h", h, 40, 3, 4, 3, 2, 1, 1, 1, 1)
    struct.pack_into("<2h", h, 70, 4, 16)
    struct.pack_into("<8f", h, 76, 1.0, 0.5, 0.5, 2.0, 0, 0, 0, 0)
    struct.pack_into("<f", h, 108, 352.0)
    struct.pack_into("<2f", h, 112, 2.0, -3.0)
    struct.pack_into("<2h", h, 252, 1, 0)
    struct.pack_into("<6f", h, 256, 0.0, 0.0, 0.0, 10.0, 20.0, 30.0)
    h[344:348] = b"n+1\x00"
    raw = np.arange(24, dtype="<i2")
    (tmp_path / "hand.nii").write_bytes(bytes(h) + raw.tobytes())
    v = hostio.read_medical_image(tmp_path / "hand.nii")
    assert v.array.shape == (2, 3, 4) and np.array_equal(v.array.ravel(), raw * 2.0 - 3.0) and not v.integer_pixels
    assert v.spacing == (0.5, 0.5, 2.0) and v.origin == (-10.0, -20.0, 30.0)
    assert np.array_equal(np.array(v.direction).reshape(3, 3), np.diag([-1.0, -1.0, 1.0]))
    # a single .dcm file: a one-slice volume
    px = (rng.random((12, 10)) * 3000).astype(np.uint16)
    synthetic.write_dicom_slice(tmp_path / "one.dcm", px, (1.5, -2.0, 33.0), (0, 1, 0), (0, 0, -1), (0.6, 0.7), "1.2.7", 1, compress="jpeg")
    v = hostio.read_medical_image(tmp_path / "one.dcm")
    assert v.array.shape == (1, 12, 10) and np.array_equal(v.array[0], px.astype(np.float32))
    assert v.spacing[:2] == (0.7, 0.6) and v.origin == (1.5, -2.0, 33.0) and v.pixel_kind == 2
    assert np.allclose(np.array(v.direction).reshape(3, 3)[:, 2], np.cross((0, 1, 0), (0, 0, -1)))
    # errors keep the reference's shape: unknown suffix -> ValueError, junk content -> an exception the drivers catch
    (tmp_path / "x.nii").write_bytes(b"junk" * 100)
    (tmp_path / "x.nrrd").write_bytes(b"junk" * 100)
    for f in ("x.nii", "x.nrrd"):
        with pytest.raises(Exception):
            hostio.read_medical_image(tmp_path / f)
