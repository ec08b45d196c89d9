"""Host input / output stage either side of the GPU path, over the C ABI's host entries.

* ``read_medical_image`` mirrors ``spine_vision.io.readers.read_medical_image`` (io/readers.py:128-161) for the formats
  that can be decoded without SimpleITK: MetaImage ``.mha`` / ``.mhd`` (what SPIDER ships).  It returns a
  ``MedicalVolume`` carrying what the reference reads off the ``sitk.Image`` further down the path:
  ``GetArrayFromImage`` (float32 ``[z, y, x]``), ``GetSpacing``, ``GetDirection``, ``GetOrigin``.
* ``read_volumes`` decodes a batch on a thread pool into ONE float32 buffer (K0 then stages only the two source planes per
  volume it needs into pinned memory).
* ``write_png_batch`` mirrors ``Image.fromarray(crop).save(path)`` (spider.py:158, phenikaa.py:213) for a batch of
  equally sized uint8 crops on a thread pool.

* ``read_dicom_series`` mirrors ``read_dicom_series`` (io/readers.py:48-73: ``sitk.ImageSeriesReader`` over GDCM) for
  native-pixel-data series, one slice per file (what the Phenikaa folders hold): the C side parses and decodes the files
  on a thread pool, this module restates the ITK conventions around them (first series id, slices ordered along the
  normal, origin / spacing / direction of the stack).  Parity unpinned (GDCM absent), see ``oracle/dicom.py``.

Formats that need a decoder this build does not have (NIfTI, NRRD, compressed DICOM) raise ``UnsupportedFormatError`` /
``SvbError``; the dataset drivers skip such series exactly like the reference skips a series whose reader raises.
"""

from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np
import torch

from . import _lib

_INTEGER_TYPES = {0, 1, 2, 3, 4, 5, 6, 7}  # svb_mha_type ids of the integer element types
# svb_mha_type -> SVB_PIXEL_* (what cv2.warpAffine does to a slice of that type in the rotated crop mode); the types OpenCV's
# remap does not take (int8, 32 / 64-bit integers: the reference raises and skips the series) stay "float"
_PIXEL_KIND = {1: _lib.PIXEL_UINT8, 2: _lib.PIXEL_INT16, 3: _lib.PIXEL_UINT16}


class UnsupportedFormatError(ValueError):
    """``read_medical_image``: the path exists but no decoder for its format is available here."""


@dataclass
class MedicalVolume:
    """What the hot path needs of the reference's ``sitk.Image``."""

    array: np.ndarray  # float32 [z, y, x] = sitk.GetArrayFromImage(image)
    spacing: tuple[float, float, float]  # image.GetSpacing()  (x, y, z)
    direction: tuple[float, ...] = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)  # image.GetDirection()
    origin: tuple[float, float, float] = (0.0, 0.0, 0.0)
    integer_pixels: bool = False  # the file's pixel type is integral (ITK casts resampled values back to it)
    pixel_kind: int = 0  # SVB_PIXEL_*: how cv2.warpAffine treats a slice of the file's pixel type (rotated crop mode)
    meta: dict = field(default_factory=dict)

    def GetSize(self):
        return (self.array.shape[2], self.array.shape[1], self.array.shape[0])

    def GetSpacing(self):
        return self.spacing

    def GetDirection(self):
        return self.direction

    def GetOrigin(self):
        return self.origin


def detect_format(path: Path) -> str:
    """``detect_format`` (io/readers.py:24-62) reduced to a tag."""
    path = Path(path)
    if path.is_dir():
        return "DICOM"
    name = path.name.lower()
    if name.endswith(".mha"):
        return "MHA"
    if name.endswith(".mhd"):
        return "MHD"
    if name.endswith(".nii") or name.endswith(".nii.gz"):
        return "NIFTI"
    if name.endswith(".nrrd"):
        return "NRRD"
    if name.endswith(".dcm") or name.endswith(".ima"):
        return "DICOM_FILE"
    return "UNKNOWN"


def _header(path: Path) -> _lib.MhaInfo:
    info = _lib.MhaInfo()
    _lib.check(_lib.load().svb_mha_read_header(os.fsencode(str(path)), C.byref(info)))
    return info


def _volume_from(info: _lib.MhaInfo, arr: np.ndarray) -> MedicalVolume:
    nx, ny, nz = info.dim[0], info.dim[1], info.dim[2]
    return MedicalVolume(array=arr.reshape(nz, ny, nx), spacing=tuple(info.spacing), direction=tuple(info.direction),
                         origin=tuple(info.origin), integer_pixels=info.element_type in _INTEGER_TYPES,
                         pixel_kind=_PIXEL_KIND.get(int(info.element_type), _lib.PIXEL_FLOAT),
                         meta={"element_type": int(info.element_type), "compressed": bool(info.compressed), "ndim": int(info.ndim)})


def read_dicom_series(folder_path: Path, n_threads: int = 0, pin: bool = False, midplane_only: bool = False) -> MedicalVolume:
    """``read_dicom_series`` (io/readers.py:48-73).  ITK conventions restated here (host arithmetic on a few numbers per slice):

    * the directory's regular files are scanned, non-DICOM files are ignored; ``ValueError`` when none is left (:66-67);
    * series ids = the distinct SeriesInstanceUIDs in lexicographic order, the FIRST one is read (:65, :69);
    * slices are ordered by the projection of ImagePositionPatient on the slice normal ``row x col`` (gdcm::IPPSorter),
      ties / missing positions by file name;
    * ``GetOrigin`` = position of the first slice; ``GetSpacing`` = (column spacing, row spacing, |last - first| / (n - 1));
      ``GetDirection`` columns = row cosines, column cosines, (last - first) normalised (the normal when n == 1);
    * stored values with RescaleSlope / Intercept applied; the pixel type counts as integral when both are whole numbers.

    ``midplane_only`` (the dataset driver): every header is still read (series selection and slice order need them), but when
    the stack direction is the left-right axis only the two slices ``volumes.plan_midplane`` reads are decoded; the rest of the
    array is UNINITIALISED and ``meta["decoded_z"]`` says which slices are real.
    """
    lib = _lib.load()
    folder_path = Path(folder_path)
    files = sorted(p for p in folder_path.iterdir() if p.is_file())
    n = len(files)
    infos = (_lib.DicomInfo * max(n, 1))()
    rcs = (C.c_int32 * max(n, 1))()
    if n:
        c_paths = (C.c_char_p * n)(*[os.fsencode(str(p)) for p in files])
        lib.svb_dicom_read_headers(c_paths, n, infos, int(n_threads), C.addressof(rcs))
    ok = [i for i in range(n) if rcs[i] == 0]
    if not ok:
        raise ValueError(f"No DICOM series found in {folder_path}")
    uid = min(infos[i].series_uid for i in ok)
    sel = [i for i in ok if infos[i].series_uid == uid]
    first = infos[sel[0]]
    row = np.array(first.orientation[0:3], dtype=np.float64)
    col = np.array(first.orientation[3:6], dtype=np.float64)
    normal = np.cross(row, col)
    sel.sort(key=lambda i: (float(np.dot(np.array(infos[i].position[:]), normal)) if infos[i].has_position else 0.0, files[i].name))
    rows, cols = first.rows, first.cols
    for i in sel:
        if (infos[i].rows, infos[i].cols) != (rows, cols):
            raise ValueError(f"DICOM series in {folder_path} has slices of different sizes")
    m = len(sel)
    p0 = np.array(infos[sel[0]].position[:], dtype=np.float64)
    p1 = np.array(infos[sel[-1]].position[:], dtype=np.float64)
    if m > 1 and float(np.linalg.norm(p1 - p0)) > 0.0:
        dz = float(np.linalg.norm(p1 - p0)) / (m - 1)
        third = (p1 - p0) / np.linalg.norm(p1 - p0)
    else:
        dz = float(first.spacing_between_slices or first.slice_thickness or 1.0)
        third = normal
    direction = np.stack([row, col, third], axis=1)  # columns = axes
    spacing = (float(first.pixel_spacing[1]), float(first.pixel_spacing[0]), dz)
    k_lo, k_hi = 0, m
    if midplane_only:
        from . import volumes as _vol

        try:
            axis, lo, hi = _vol.midplane_source_planes((cols, rows, m), spacing, tuple(float(v) for v in direction.ravel()))[:3]
            if axis == 0:
                k_lo, k_hi = int(lo), int(hi) + 1
        except ValueError:  # direction cosines that do not resolve: the driver reports it, decode everything
            pass
    host = torch.empty(m * rows * cols, dtype=torch.float32)
    if pin and torch.cuda.is_available():
        host = host.pin_memory()
    part = sel[k_lo:k_hi]
    mp = len(part)
    c_sel = (C.c_char_p * mp)(*[os.fsencode(str(files[i])) for i in part])
    c_infos = (_lib.DicomInfo * mp)(*[infos[i] for i in part])
    c_dsts = (C.c_void_p * mp)(*[host.data_ptr() + 4 * k * rows * cols for k in range(k_lo, k_hi)])
    c_sizes = (C.c_size_t * mp)(*([rows * cols] * mp))
    _lib.check(lib.svb_dicom_read_slices_f32(c_sel, mp, c_infos, c_dsts, c_sizes, int(n_threads), None))
    integral = all(float(infos[i].rescale_slope).is_integer() and float(infos[i].rescale_intercept).is_integer() for i in sel)
    kind = _lib.PIXEL_FLOAT
    if integral:  # the integer type ITK / GDCM give the series: 16-bit stored values (signed when stored signed or shifted below zero)
        signed = any(infos[i].pixel_representation == 1 or infos[i].rescale_intercept < 0 for i in sel)
        if first.bits_allocated == 16:
            kind = _lib.PIXEL_INT16 if signed else _lib.PIXEL_UINT16
        elif first.bits_allocated == 8 and not signed:
            kind = _lib.PIXEL_UINT8
    return MedicalVolume(pixel_kind=kind, array=host.numpy().reshape(m, rows, cols), spacing=spacing,
                         direction=tuple(float(v) for v in direction.ravel()), origin=tuple(float(v) for v in p0), integer_pixels=integral,
                         meta={"series_uid": uid.decode("ascii", "replace"), "files": [files[i].name for i in sel], "decoded_z": (k_lo, k_hi)})


def read_dicom_files(paths, n_threads: int = 0):
    """Single DICOM slices (``sitk.ReadImage(dcm)`` + ``GetArrayFromImage(...)[0]``, datasets/localization.py:262-266) decoded
    on a thread pool.  Returns ``(arrays, errors)``: ``arrays[i]`` float32 ``[rows, cols]`` (stored value * slope + intercept),
    or ``None`` with the reason in ``errors[i]`` -- the caller logs and skips, like the reference's ``except``."""
    lib = _lib.load()
    paths = [Path(p) for p in paths]
    n = len(paths)
    infos = (_lib.DicomInfo * max(n, 1))()
    rcs = (C.c_int32 * max(n, 1))()
    errors: list[str | None] = [None] * n
    arrays: list[np.ndarray | None] = [None] * n
    if n == 0:
        return arrays, errors
    c_paths = (C.c_char_p * n)(*[os.fsencode(str(p)) for p in paths])
    lib.svb_dicom_read_headers(c_paths, n, infos, int(n_threads), C.addressof(rcs))
    ok = [i for i in range(n) if rcs[i] == 0]
    for i in range(n):
        if rcs[i] != 0:
            errors[i] = f"libspine_b200 error {rcs[i]}: not a supported DICOM slice: {paths[i]}"
    if not ok:
        return arrays, errors
    m = len(ok)
    sizes = [infos[i].rows * infos[i].cols for i in ok]
    offs = np.concatenate([[0], np.cumsum([(s + 3) // 4 * 4 for s in sizes])]).astype(np.int64)
    host = np.empty(int(offs[-1]), dtype=np.float32)
    c_sel = (C.c_char_p * m)(*[os.fsencode(str(paths[i])) for i in ok])
    c_infos = (_lib.DicomInfo * m)(*[infos[i] for i in ok])
    c_dsts = (C.c_void_p * m)(*[host.ctypes.data + 4 * int(offs[k]) for k in range(m)])
    c_sizes = (C.c_size_t * m)(*sizes)
    rcs2 = (C.c_int32 * m)()
    lib.svb_dicom_read_slices_f32(c_sel, m, c_infos, c_dsts, c_sizes, int(n_threads), C.addressof(rcs2))
    for k, i in enumerate(ok):
        if rcs2[k] != 0:
            errors[i] = f"libspine_b200 error {rcs2[k]} while decoding {paths[i]}"
        else:
            arrays[i] = host[int(offs[k]) : int(offs[k]) + sizes[k]].reshape(infos[i].rows, infos[i].cols)
    return arrays, errors


def read_medical_image(path: Path, midplane_only: bool = False) -> MedicalVolume:
    """``read_medical_image`` (io/readers.py:128-161): same error behaviour (``FileNotFoundError`` for a missing path,
    ``ValueError`` for an unknown format).  ``midplane_only`` is the dataset driver's switch for DICOM series (see
    ``read_dicom_series``); the default decodes everything, like the reference."""
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError(f"Path does not exist: {path}")
    fmt = detect_format(path)
    if fmt in ("MHA", "MHD"):
        info = _header(path)
        n = info.dim[0] * info.dim[1] * info.dim[2]
        arr = np.empty(n, dtype=np.float32)
        _lib.check(_lib.load().svb_mha_read_f32(os.fsencode(str(path)), C.byref(info), arr.ctypes.data, n))
        return _volume_from(info, arr)
    if fmt == "DICOM":
        return read_dicom_series(path, midplane_only=midplane_only)
    if fmt == "UNKNOWN":
        raise ValueError(f"Unsupported format for path: {path}")
    raise UnsupportedFormatError(f"{fmt} decoding needs SimpleITK, which this build does not link; path: {path}")


def read_volumes(paths, n_threads: int = 0, pin: bool = False, midplane_only: bool = False):
    """Decode many MetaImage volumes on a thread pool into one float32 buffer (``pin`` is accepted and ignored: whole volumes
    never travel to the device).
    ``midplane_only`` (the dataset driver): when the left-right axis of a volume is its slowest array axis -- a sagittal
    acquisition, what SPIDER ships -- only the two source slices ``volumes.plan_midplane`` reads are decoded (the zlib stream is
    inflated up to them and no further); the REST OF THAT ARRAY IS UNINITIALISED, ``meta["decoded_z"]`` says which slices are
    real.  Volumes in any other orientation are decoded whole.
    Returns ``(volumes, errors)``: ``volumes[i]`` is a ``MedicalVolume`` whose array is a view into the shared buffer, or
    ``None`` when file i could not be read (``errors[i]`` holds the reason) -- the drivers skip those series
    (spider.py:139-141)."""
    lib = _lib.load()
    paths = [Path(p) for p in paths]
    n = len(paths)
    infos = (_lib.MhaInfo * max(n, 1))()
    errors: list[str | None] = [None] * n
    ok = []
    for i, p in enumerate(paths):
        try:
            if not p.exists():
                raise FileNotFoundError(f"Path does not exist: {p}")
            if detect_format(p) not in ("MHA", "MHD"):
                raise UnsupportedFormatError(f"{detect_format(p)} decoding is not available: {p}")
            _lib.check(lib.svb_mha_read_header(os.fsencode(str(p)), C.byref(infos[i])))
            ok.append(i)
        except Exception as e:  # noqa: BLE001 -- mirrored: any reader error skips the series
            errors[i] = str(e)
    # One plain host buffer for the chunk (virtual until touched: with ``midplane_only`` only the two decoded slices of a volume
    # ever become resident).  The C header parser has already bounded every volume by the bytes its file really holds; should
    # the chunk's buffer still not be obtainable, the largest volumes are dropped one at a time -- each with ITS error, so the
    # driver skips that series (spider.py:131-133) instead of losing the whole chunk.
    host = None
    while host is None:
        sizes = [infos[i].dim[0] * infos[i].dim[1] * infos[i].dim[2] for i in ok]
        offs, total = [], 0
        for s in sizes:
            offs.append(total)
            total += (s + 3) // 4 * 4
        try:
            host = torch.empty(max(total, 4), dtype=torch.float32)
        except (RuntimeError, MemoryError) as e:
            if not ok:
                raise
            worst = max(range(len(ok)), key=lambda k: sizes[k])
            errors[ok[worst]] = f"cannot allocate {sizes[worst] * 4} bytes for {paths[ok[worst]]}: {e}"
            del ok[worst]
    hv = host.numpy()
    m = len(ok)
    volumes: list[MedicalVolume | None] = [None] * n
    if m:
        c_paths = (C.c_char_p * m)(*[os.fsencode(str(paths[i])) for i in ok])
        c_infos = (_lib.MhaInfo * m)(*[infos[i] for i in ok])
        c_dsts = (C.c_void_p * m)(*[host.data_ptr() + 4 * o for o in offs])
        c_sizes = (C.c_size_t * m)(*sizes)
        rcs = (C.c_int32 * m)()
        z_lo = [0] * m
        z_hi = [int(infos[i].dim[2]) for i in ok]
        if midplane_only:
            from . import volumes as _vol

            for j, i in enumerate(ok):
                try:
                    axis, lo, hi = _vol.midplane_source_planes(tuple(infos[i].dim), tuple(infos[i].spacing), tuple(infos[i].direction))[:3]
                except ValueError:  # direction cosines that do not resolve: the driver reports it, decode whole
                    continue
                if axis == 0 and infos[i].ndim == 3:
                    z_lo[j], z_hi[j] = int(lo), int(hi) + 1
        if any(z_lo[j] != 0 or z_hi[j] != infos[i].dim[2] for j, i in enumerate(ok)):
            lib.svb_mha_read_batch_slab_f32(c_paths, m, c_infos, c_dsts, c_sizes, (C.c_int32 * m)(*z_lo), (C.c_int32 * m)(*z_hi),
                                            int(n_threads), C.addressof(rcs))
        else:
            lib.svb_mha_read_batch_f32(c_paths, m, c_infos, c_dsts, c_sizes, int(n_threads), C.addressof(rcs))
        for j, i in enumerate(ok):
            if rcs[j] != 0:
                errors[i] = f"libspine_b200 error {rcs[j]} while decoding {paths[i]}"
            else:
                volumes[i] = _volume_from(infos[i], hv[offs[j] : offs[j] + sizes[j]])
                volumes[i].meta["decoded_z"] = (z_lo[j], z_hi[j])
    return volumes, errors


def encode_png(image: np.ndarray, level: int = 6) -> bytes:
    """One 8-bit greyscale PNG in memory."""
    lib = _lib.load()
    img = np.ascontiguousarray(image, dtype=np.uint8)
    if img.ndim != 2:
        raise ValueError("encode_png takes a 2-D uint8 array (PIL mode 'L')")
    h, w = img.shape
    cap = lib.svb_png_bound(h, w)
    buf = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t(0)
    _lib.check(lib.svb_png_encode_gray8(img.ctypes.data, h, w, int(level), buf.ctypes.data, cap, C.byref(n)))
    return buf[: n.value].tobytes()


def write_png_ragged(pool_u8: np.ndarray, offs, shapes, paths, level: int = 6, n_threads: int = 0) -> None:
    """PNGs of different sizes from one flat uint8 pool: image i = ``pool_u8[offs[i] : offs[i] + h*w].reshape(h, w)``."""
    lib = _lib.load()
    pool = np.ascontiguousarray(pool_u8, dtype=np.uint8)
    n = len(paths)
    if n == 0:
        return
    c_offs = np.ascontiguousarray(offs, dtype=np.int64)
    c_hw = np.ascontiguousarray(shapes, dtype=np.int32).reshape(n, 2)
    assert c_offs.shape == (n,) and int((c_offs + c_hw[:, 0].astype(np.int64) * c_hw[:, 1]).max()) <= pool.size
    c_paths = (C.c_char_p * n)(*[os.fsencode(str(p)) for p in paths])
    _lib.check(lib.svb_png_write_gray8_ragged(pool.ctypes.data, c_offs.ctypes.data, c_hw.ctypes.data, n, c_paths, int(level),
                                              int(n_threads), None))


def write_png_batch(images: np.ndarray, paths, level: int = 6, n_threads: int = 0) -> None:
    """``Image.fromarray(crop).save(path)`` for ``images`` uint8 ``[n, h, w]`` -> ``paths[n]``, encoded and written on a
    thread pool.  Raises ``SvbError`` if any file could not be written."""
    lib = _lib.load()
    imgs = np.ascontiguousarray(images, dtype=np.uint8)
    if imgs.ndim != 3 or imgs.shape[0] != len(paths):
        raise ValueError("write_png_batch takes uint8 [n, h, w] and n paths")
    n, h, w = imgs.shape
    if n == 0:
        return
    c_paths = (C.c_char_p * n)(*[os.fsencode(str(p)) for p in paths])
    _lib.check(lib.svb_png_write_gray8_batch(imgs.ctypes.data, n, h, w, c_paths, int(level), int(n_threads), None))
