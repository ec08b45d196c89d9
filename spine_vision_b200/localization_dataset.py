"""``create_localization_dataset`` on the GPU path (second half): the reference's builder of the
localizer's TRAINING set (``spine_vision/datasets/localization.py:326-382``) with the same configuration fields, output tree
and ``annotations.csv``.  Its pixel work is ``normalize_to_uint8`` + PNG save per image (localization.py:147-151 for the
Lumbar-Coords ``.npy`` arrays, :262-267 for every annotated RSNA DICOM slice); here the label files are walked first (no
pixels), the distinct images are decoded on a thread pool, normalised in ragged batches by ``svb_normalize_u8`` and written
by the threaded PNG encoder.  Records, their order, the de-duplication of images and the skip rules are the reference's.
"""

from __future__ import annotations

import csv
import logging
import shutil
from pathlib import Path

import numpy as np
from pydantic import BaseModel, ConfigDict, computed_field

from . import hostio, ops
from .dataset import ProcessingResult

logger = logging.getLogger("spine_vision_b200.localization_dataset")


class LocalizationDatasetConfig(BaseModel):
    """Field-for-field twin of the reference's ``LocalizationDatasetConfig`` (localization.py:30-66 + ``BaseConfig``);
    ``device``, ``chunk_images``, ``io_threads`` and ``png_level`` are the only additions."""

    model_config = ConfigDict(arbitrary_types_allowed=True)

    verbose: bool = False
    enable_file_log: bool = False
    log_path: Path = Path.cwd() / "logs"

    base_path: Path = Path.cwd() / "data"
    output_name: str = "localization"
    include_neural_foraminal: bool = True
    include_spinal_canal: bool = True
    skip_invalid_instances: bool = True

    device: str = "cuda:0"
    chunk_images: int = 256
    io_threads: int = 0
    png_level: int = 6

    @computed_field
    @property
    def lumbar_coords_path(self) -> Path:
        return self.base_path / "raw" / "Lumbar Coords"

    @computed_field
    @property
    def rsna_path(self) -> Path:
        return self.base_path / "raw" / "RSNA"

    @computed_field
    @property
    def output_path(self) -> Path:
        path = self.base_path / "processed" / self.output_name
        path.mkdir(parents=True, exist_ok=True)
        return path


class AnnotationRecord(BaseModel):
    """One row of the localization ``annotations.csv`` (localization.py:69-77)."""

    image_path: str
    level: str
    relative_x: float
    relative_y: float
    series_type: str
    source: str


# ------------------------------------------------------------------------------------------ RSNA series lookup (datasets/rsna.py:7-61)
def load_series_mapping(series_desc_path: Path) -> dict[int, dict[str, int]]:
    mapping: dict[int, dict[str, int]] = {}
    with open(series_desc_path, newline="") as f:
        for row in csv.DictReader(f):
            mapping.setdefault(int(row["study_id"]), {})[row["series_description"]] = int(row["series_id"])
    return mapping


def get_series_type(series_id: int, study_id: int, series_mapping: dict[int, dict[str, int]]) -> str | None:
    for desc, sid in series_mapping.get(study_id, {}).items():
        if sid == series_id:
            return desc
    return None


# ------------------------------------------------------------------------------------------ the batched pixel stage
def normalize_and_save(arrays: list[np.ndarray], out_paths: list[Path], config: LocalizationDatasetConfig) -> None:
    """``Image.fromarray(normalize_to_uint8(arr)).save(path)`` for a list of 2-D arrays of any sizes: ragged batches through
    ``svb_normalize_u8`` (one min/max pass + one normalise pass on the device), threaded native PNG encoding on the host.
    A target that is not ``.png`` (the reference keeps the ``.jpg`` name of a Lumbar-Coords array, so Pillow writes a JPEG
    there, localization.py:124-151) goes through Pillow's encoder like in the reference -- JPEG is lossy, only the same
    libjpeg gives the same file."""
    step = max(1, config.chunk_images)
    for c0 in range(0, len(arrays), step):
        chunk = [np.ascontiguousarray(a, dtype=np.float32) for a in arrays[c0 : c0 + step]]
        paths = out_paths[c0 : c0 + step]
        pool = ops.SlicePool.from_numpy(chunk, config.device)
        out = ops.normalize_u8(pool).cpu().numpy()
        offs = pool.offs.cpu().numpy()
        png = [k for k, p in enumerate(paths) if str(p).lower().endswith(".png")]
        if png:
            hostio.write_png_ragged(out, offs[png], [chunk[k].shape for k in png], [paths[k] for k in png], config.png_level, config.io_threads)
        for k, p in enumerate(paths):
            if k not in png:
                from PIL import Image  # same encoder as the reference for non-PNG targets

                h, w = chunk[k].shape
                Image.fromarray(out[int(offs[k]) : int(offs[k]) + h * w].reshape(h, w)).save(p)


def process_lumbar_coords_pretrain(coords_csv_path: Path, data_path: Path, output_images_path: Path,
                                   config: LocalizationDatasetConfig | None = None) -> list[AnnotationRecord]:
    """localization.py:80-178: ready-made JPGs are copied, ``.npy`` arrays are normalised and saved as PNG (batched)."""
    config = config or LocalizationDatasetConfig()
    folders = {"spider": "processed_spider_jpgs", "lsd": "processed_lsd_jpgs", "osf": "processed_osf_jpgs", "tseg": "processed_tseg_jpgs"}
    npy_folders = {"spider": None, "lsd": "processed_lsd", "osf": "processed_osf", "tseg": "processed_tseg"}
    series_types = {"spider": "sag_t2", "lsd": "sag_t2", "osf": "sag_t1", "tseg": "ct"}
    records: list[AnnotationRecord] = []
    done: set[str] = set()
    to_norm: list[tuple[Path, Path]] = []
    with open(coords_csv_path, newline="") as f:
        for row in csv.DictReader(f):
            filename, source = row["filename"], row["source"]
            folder = folders.get(source)
            if folder is None:
                logger.warning("Unknown source: %s", source)
                continue
            out_name = f"pretrain_{source}_{filename}"
            if not out_name.endswith((".jpg", ".png")):
                out_name = out_name.replace(".npy", ".png")
            if out_name not in done:
                src = data_path / folder / filename
                if src.exists():
                    shutil.copy(src, output_images_path / out_name)
                    done.add(out_name)
                else:
                    npy_folder = npy_folders.get(source)
                    npy = data_path / npy_folder / filename.replace(".jpg", ".npy") if npy_folder else None
                    if npy is not None and npy.exists():
                        to_norm.append((npy, output_images_path / out_name))
                        done.add(out_name)
                    else:
                        logger.warning("File not found: %s%s", src, f" or {npy}" if npy is not None else "")
                        continue
            records.append(AnnotationRecord(image_path=f"images/{out_name}", level=row["level"], relative_x=float(row["relative_x"]),
                                            relative_y=float(row["relative_y"]), series_type=series_types[source],
                                            source=f"pretrain_{source}"))
    if to_norm:
        normalize_and_save([np.load(p) for p, _ in to_norm], [o for _, o in to_norm], config)
    return records


def process_rsna_improved(coords_csv_path: Path, series_desc_path: Path, rsna_images_path: Path, output_images_path: Path,
                          config: LocalizationDatasetConfig) -> list[AnnotationRecord]:
    """localization.py:181-287.  Two passes: the rows are filtered exactly as the reference filters them and the distinct
    DICOM files collected; those are decoded + normalised + saved in batches; a row whose image failed is dropped (the
    reference's ``except ... continue``)."""
    mapping = load_series_mapping(series_desc_path)
    with open(coords_csv_path, newline="") as f:
        rows = list(csv.DictReader(f))
    kept: list[tuple[dict, str, str]] = []  # (row, series_type, output filename)
    wanted: dict[str, Path] = {}
    for row in rows:
        series_id, study_id, instance = int(row["series_id"]), int(row["study_id"]), int(row["instance_number"])
        condition = row["condition"]
        if "Subarticular" in condition:
            continue
        if "Spinal Canal" in condition and not config.include_spinal_canal:
            continue
        if "Neural Foraminal" in condition and not config.include_neural_foraminal:
            continue
        if config.skip_invalid_instances and instance < 0:
            continue
        desc = get_series_type(series_id, study_id, mapping)
        if desc is None:
            continue
        if "Sagittal T1" in desc:
            series_type = "sag_t1"
        elif "Sagittal T2" in desc:
            series_type = "sag_t2"
        else:
            continue
        dcm = rsna_images_path / str(study_id) / str(series_id) / f"{instance}.dcm"
        if not dcm.exists():
            continue
        name = f"rsna_{study_id}_{series_id}_{instance}.png"
        wanted.setdefault(name, dcm)
        kept.append((row, series_type, name))
    names = list(wanted)
    failed: set[str] = set()
    step = max(1, config.chunk_images)
    for c0 in range(0, len(names), step):
        chunk = names[c0 : c0 + step]
        arrays, errors = hostio.read_dicom_files([wanted[n] for n in chunk], config.io_threads)
        good = [k for k, a in enumerate(arrays) if a is not None]
        for k, a in enumerate(arrays):
            if a is None:
                logger.error("Error processing %s: %s", wanted[chunk[k]], errors[k])
                failed.add(chunk[k])
        if good:
            normalize_and_save([arrays[k] for k in good], [output_images_path / chunk[k] for k in good], config)
    return [AnnotationRecord(image_path=f"images/{name}", level=row["level"], relative_x=float(row["relative_x"]),
                             relative_y=float(row["relative_y"]), series_type=st, source="rsna")
            for row, st, name in kept if name not in failed]


def write_records_csv(records: list[BaseModel], csv_path: Path) -> None:
    """io/tabular.py:18-36."""
    if not records:
        raise ValueError("Cannot write empty records list")
    fieldnames = list(type(records[0]).model_fields.keys())
    with open(csv_path, "w", newline="") as f:
        writer = csv.DictWriter(f, fieldnames=fieldnames)
        writer.writeheader()
        for rec in records:
            writer.writerow(rec.model_dump())


def create_localization_dataset(config: LocalizationDatasetConfig) -> ProcessingResult:
    """Drop-in for ``create_localization_dataset`` (localization.py:326-382)."""
    if config.verbose:
        logger.setLevel(logging.DEBUG)
    images = config.output_path / "images"
    images.mkdir(parents=True, exist_ok=True)
    records = process_lumbar_coords_pretrain(config.lumbar_coords_path / "coords_pretrain.csv", config.lumbar_coords_path / "data",
                                             images, config)
    logger.info("Processed %d pretrain annotation records", len(records))
    rsna = process_rsna_improved(config.lumbar_coords_path / "coords_rsna_improved.csv", config.rsna_path / "train_series_descriptions.csv",
                                 config.rsna_path / "train_images", images, config)
    logger.info("Processed %d RSNA annotation records", len(rsna))
    records = records + rsna
    write_records_csv(records, config.output_path / "annotations.csv")
    return ProcessingResult(num_samples=len(records), output_path=config.output_path,
                            summary=f"Created {len(records)} IVD coordinate annotations")
