"""Host input / output stage either side of the GPU path, over the C ABI's host entries.

* ``read_medical_image`` mirrors ``spine_vision.io.readers.read_medical_image`` (io/readers.py:128-161) for the formats
  that can be decoded without SimpleITK: MetaImage ``.mha`` / ``.mhd`` (what SPIDER ships).  It returns a
  ``MedicalVolume`` carrying what the reference reads off the ``sitk.Image`` further down the path:
  ``GetArrayFromImage`` (float32 ``[z, y, x]``), ``GetSpacing``, ``GetDirection``, ``GetOrigin``.
* ``read_volumes`` decodes a batch on a thread pool straight into ONE pinned float32 buffer (the H2D staging area).
* ``write_png_batch`` mirrors ``Image.fromarray(crop).save(path)`` (spider.py:158, phenikaa.py:213) for a batch of
  equally sized uint8 crops on a thread pool.

Formats that need a decoder this image does not have (DICOM series, NIfTI, NRRD) raise ``UnsupportedFormatError``;
the dataset drivers skip such series exactly like the reference skips a series whose reader raises.
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
                         meta={"element_type": int(info.element_type), "compressed": bool(info.compressed), "ndim": int(info.ndim)})


def read_medical_image(path: Path) -> MedicalVolume:
    """``read_medical_image`` (io/readers.py:128-161): same error behaviour (``FileNotFoundError`` for a missing path,
    ``ValueError`` for an unknown format)."""
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
    if fmt == "UNKNOWN":
        raise ValueError(f"Unsupported format for path: {path}")
    raise UnsupportedFormatError(f"{fmt} decoding needs SimpleITK, which this build does not link; path: {path}")


def read_volumes(paths, n_threads: int = 0, pin: bool = True):
    """Decode many MetaImage volumes on a thread pool into one (pinned) float32 buffer.
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
    sizes = [infos[i].dim[0] * infos[i].dim[1] * infos[i].dim[2] for i in ok]
    offs, total = [], 0
    for s in sizes:
        offs.append(total)
        total += (s + 3) // 4 * 4
    host = torch.empty(max(total, 4), dtype=torch.float32)
    if pin and torch.cuda.is_available():
        host = host.pin_memory()
    hv = host.numpy()
    m = len(ok)
    volumes: list[MedicalVolume | None] = [None] * n
    if m:
        c_paths = (C.c_char_p * m)(*[os.fsencode(str(paths[i])) for i in ok])
        c_infos = (_lib.MhaInfo * m)(*[infos[i] for i in ok])
        c_dsts = (C.c_void_p * m)(*[host.data_ptr() + 4 * o for o in offs])
        c_sizes = (C.c_size_t * m)(*sizes)
        rcs = (C.c_int32 * m)()
        lib.svb_mha_read_batch_f32(c_paths, m, c_infos, c_dsts, c_sizes, int(n_threads), C.addressof(rcs))
        for j, i in enumerate(ok):
            if rcs[j] != 0:
                errors[i] = f"libspine_b200 error {rcs[j]} while decoding {paths[i]}"
            else:
                volumes[i] = _volume_from(infos[i], hv[offs[j] : offs[j] + sizes[j]])
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
