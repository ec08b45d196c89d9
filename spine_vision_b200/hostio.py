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
            # the batch call keeps only the codes (the messages are thread-local in its workers): ask again for the reason, so that
            # a refused transfer syntax / MONOCHROME1 / truncated file is NAMED in the log instead of silently costing a series
            one = (_lib.DicomInfo * 1)()
            lib.svb_dicom_read_headers((C.c_char_p * 1)(os.fsencode(str(paths[i]))), 1, one, 1, None)
            why = lib.svb_last_error().decode("utf-8", "replace")
            errors[i] = f"libspine_b200 error {rcs[i]}: not a supported DICOM slice ({why})"
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


# ------------------------------------------------------------------------------------------ NIfTI / NRRD / single DICOM file
# Neither dataset builder reads these (SPIDER ships .mha, Phenikaa DICOM folders); read_medical_image dispatches them
# (io/readers.py:23-28, 76-126), so they are here for callers of that function.  Plain Python + NumPy, like the reference's own
# reader layer; ITK's conventions restated from its documentation -- PARITY UNPINNED (SimpleITK is absent from the image).
_NIFTI_TYPES = {2: "u1", 4: "i2", 8: "i4", 16: "f4", 64: "f8", 256: "i1", 512: "u2", 768: "u4", 1024: "i8", 1280: "u8"}


def read_nifti(file_path: Path) -> MedicalVolume:
    """``read_nifti`` (io/readers.py:76-86) = ``sitk.ReadImage`` = itk::NiftiImageIO for a 3-D scalar NIfTI-1 file (``.nii`` /
    ``.nii.gz``, single file).  Restated conventions: sizes ``dim[1..3]``; voxel order x fastest -> ``[z, y, x]``;
    ``scl_slope`` / ``scl_inter`` applied when they are not the identity (the image then is float32); orientation from the
    qform when ``qform_code > 0`` (quaternion b, c, d and ``qfac = pixdim[0]``, spacing ``pixdim[1..3]``), else from the
    sform rows (spacing = column norms), else identity; NIfTI is RAS+, ITK is LPS+: the x and y rows of the rotation and of
    the offset change sign."""
    import gzip
    import struct

    blob = Path(file_path).read_bytes()
    if blob[:2] == b"\x1f\x8b":
        blob = gzip.decompress(blob)
    if len(blob) < 352:
        raise ValueError(f"{file_path}: too short for a NIfTI-1 header")
    end = "<" if struct.unpack_from("<i", blob, 0)[0] == 348 else ">"
    if struct.unpack_from(end + "i", blob, 0)[0] != 348 or blob[344:347] not in (b"n+1", b"ni1"):
        raise ValueError(f"{file_path}: not a NIfTI-1 file")
    if blob[344:347] == b"ni1":
        raise UnsupportedFormatError(f"{file_path}: two-file NIfTI (.hdr / .img) is not supported")
    dim = struct.unpack_from(end + "8h", blob, 40)
    if dim[0] < 3 or any(d > 1 for d in dim[4 : dim[0] + 1]):
        raise UnsupportedFormatError(f"{file_path}: {dim[0]}-D NIfTI (3-D scalar volumes only)")
    nx, ny, nz = (max(1, int(d)) for d in dim[1:4])
    datatype = struct.unpack_from(end + "h", blob, 70)[0]
    if datatype not in _NIFTI_TYPES:
        raise UnsupportedFormatError(f"{file_path}: NIfTI datatype {datatype}")
    pixdim = struct.unpack_from(end + "8f", blob, 76)
    vox_offset = int(struct.unpack_from(end + "f", blob, 108)[0])
    slope, inter = struct.unpack_from(end + "2f", blob, 112)
    qform_code, sform_code = struct.unpack_from(end + "2h", blob, 252)
    qb, qc, qd, qx, qy, qz = struct.unpack_from(end + "6f", blob, 256)
    srow = np.array(struct.unpack_from(end + "12f", blob, 280), dtype=np.float64).reshape(3, 4)
    dt = np.dtype(_NIFTI_TYPES[datatype]).newbyteorder(end)
    n = nx * ny * nz
    if len(blob) < vox_offset + n * dt.itemsize:
        raise ValueError(f"{file_path}: voxel data truncated")
    raw = np.frombuffer(blob, dtype=dt, count=n, offset=vox_offset).reshape(nz, ny, nx)
    scaled = slope != 0.0 and not (slope == 1.0 and inter == 0.0)
    arr = (raw.astype(np.float64) * float(slope) + float(inter)).astype(np.float32) if scaled else raw.astype(np.float32)
    spacing = [float(abs(pixdim[k])) or 1.0 for k in (1, 2, 3)]
    if qform_code > 0:
        a2 = 1.0 - (qb * qb + qc * qc + qd * qd)
        a = float(np.sqrt(a2)) if a2 > 1e-7 else 0.0
        b, c, d = float(qb), float(qc), float(qd)
        if a == 0.0:
            nrm = 1.0 / float(np.sqrt(b * b + c * c + d * d))
            b, c, d = b * nrm, c * nrm, d * nrm
        R = np.array([[a * a + b * b - c * c - d * d, 2 * b * c - 2 * a * d, 2 * b * d + 2 * a * c],
                      [2 * b * c + 2 * a * d, a * a + c * c - b * b - d * d, 2 * c * d - 2 * a * b],
                      [2 * b * d - 2 * a * c, 2 * c * d + 2 * a * b, a * a + d * d - c * c - b * b]])
        if pixdim[0] < 0:
            R[:, 2] = -R[:, 2]
        off = np.array([qx, qy, qz], dtype=np.float64)
    elif sform_code > 0:
        M = srow[:, :3]
        norms = np.linalg.norm(M, axis=0)
        spacing = [float(v) or 1.0 for v in norms]
        R = M / np.where(norms == 0, 1.0, norms)
        off = srow[:, 3].copy()
    else:
        R, off = np.eye(3), np.zeros(3)
    flip = np.array([-1.0, -1.0, 1.0])
    direction = R * flip[:, None]
    origin = off * flip
    integral = (not scaled) and dt.kind in "iu"
    return MedicalVolume(array=arr, spacing=tuple(spacing), direction=tuple(float(v) for v in direction.ravel()),
                         origin=tuple(float(v) for v in origin), integer_pixels=integral,
                         pixel_kind=ops_kind(dt) if integral else _lib.PIXEL_FLOAT, meta={"format": "NIFTI", "datatype": int(datatype)})


def ops_kind(dt) -> int:
    dt = np.dtype(dt)
    return {("i", 2): _lib.PIXEL_INT16, ("u", 2): _lib.PIXEL_UINT16, ("u", 1): _lib.PIXEL_UINT8}.get((dt.kind, dt.itemsize), _lib.PIXEL_FLOAT)


_NRRD_TYPES = {"signed char": "i1", "int8": "i1", "int8_t": "i1", "uchar": "u1", "unsigned char": "u1", "uint8": "u1", "uint8_t": "u1",
               "short": "i2", "short int": "i2", "signed short": "i2", "int16": "i2", "int16_t": "i2", "ushort": "u2",
               "unsigned short": "u2", "uint16": "u2", "uint16_t": "u2", "int": "i4", "signed int": "i4", "int32": "i4", "int32_t": "i4",
               "uint": "u4", "unsigned int": "u4", "uint32": "u4", "uint32_t": "u4", "longlong": "i8", "long long": "i8", "int64": "i8",
               "int64_t": "i8", "ulonglong": "u8", "unsigned long long": "u8", "uint64": "u8", "uint64_t": "u8", "float": "f4", "double": "f8"}


def read_nrrd(file_path: Path) -> MedicalVolume:
    """``read_nrrd`` (io/readers.py:100-110) = itk::NrrdImageIO for a 3-D scalar NRRD (attached ``.nrrd`` or detached header
    with ``data file``; ``raw`` / ``gzip`` encoding).  Restated conventions: ``sizes`` x fastest; ``space directions`` give spacing
    (vector norms) and direction (normalised vectors, columns = axes), ``space origin`` the origin; a ``right-anterior-superior``
    space is turned into ITK's LPS by negating x and y (``left-anterior-superior``: y only); without ``space directions`` the
    ``spacings`` field and an identity direction."""
    import gzip

    path = Path(file_path)
    blob = path.read_bytes()
    if not blob.startswith(b"NRRD"):
        raise ValueError(f"{path}: not a NRRD file")
    sep = blob.find(b"\n\n")
    head = (blob if sep < 0 else blob[:sep]).decode("latin-1").splitlines()
    fields = {}
    for ln in head[1:]:
        if ln.startswith("#") or ":" not in ln:
            continue
        k, v = ln.split(":", 1)
        fields[k.strip().lower()] = v.lstrip("=").strip()
    if int(fields.get("dimension", "0")) != 3:
        raise UnsupportedFormatError(f"{path}: {fields.get('dimension')}-D NRRD (3-D scalar volumes only)")
    t = fields.get("type", "").lower()
    if t not in _NRRD_TYPES:
        raise UnsupportedFormatError(f"{path}: NRRD type {t!r}")
    nx, ny, nz = (int(v) for v in fields["sizes"].split())
    end = ">" if fields.get("endian", "little").lower() == "big" else "<"
    dt = np.dtype(_NRRD_TYPES[t]).newbyteorder(end)
    enc = fields.get("encoding", "raw").lower()
    if "data file" in fields or "datafile" in fields:
        data = (path.parent / fields.get("data file", fields.get("datafile"))).read_bytes()
    else:
        if sep < 0:
            raise ValueError(f"{path}: no data after the header")
        data = blob[sep + 2 :]
    if enc in ("gzip", "gz"):
        data = gzip.decompress(data)
    elif enc != "raw":
        raise UnsupportedFormatError(f"{path}: NRRD encoding {enc!r}")
    n = nx * ny * nz
    if len(data) < n * dt.itemsize:
        raise ValueError(f"{path}: voxel data truncated")
    arr = np.frombuffer(data, dtype=dt, count=n).reshape(nz, ny, nx)
    spacing, direction, origin = [1.0, 1.0, 1.0], np.eye(3), np.zeros(3)
    vec = lambda s: [float(v) for v in s.strip("() ").split(",")]  # noqa: E731
    if "space directions" in fields:
        import re

        cols = [vec(g) for g in re.findall(r"\(([^)]*)\)", fields["space directions"])]
        if len(cols) != 3:
            raise UnsupportedFormatError(f"{path}: space directions {fields['space directions']!r}")
        D = np.array(cols, dtype=np.float64).T  # columns = axes
        norms = np.linalg.norm(D, axis=0)
        spacing = [float(v) or 1.0 for v in norms]
        direction = D / np.where(norms == 0, 1.0, norms)
    elif "spacings" in fields:
        spacing = [float(v) for v in fields["spacings"].split()]
    if "space origin" in fields:
        origin = np.array(vec(fields["space origin"]), dtype=np.float64)
    space = fields.get("space", "left-posterior-superior").lower()
    flip = {"right-anterior-superior": (-1.0, -1.0, 1.0), "ras": (-1.0, -1.0, 1.0), "left-anterior-superior": (1.0, -1.0, 1.0),
            "las": (1.0, -1.0, 1.0)}.get(space, (1.0, 1.0, 1.0))
    direction = direction * np.array(flip)[:, None]
    origin = origin * np.array(flip)
    integral = dt.kind in "iu"
    return MedicalVolume(array=arr.astype(np.float32), spacing=tuple(spacing), direction=tuple(float(v) for v in direction.ravel()),
                         origin=tuple(float(v) for v in origin), integer_pixels=integral,
                         pixel_kind=ops_kind(dt) if integral else _lib.PIXEL_FLOAT, meta={"format": "NRRD", "type": t})


def read_dicom_file(file_path: Path) -> MedicalVolume:
    """``read_dicom_file`` (io/readers.py:113-123) = ``sitk.ReadImage`` of ONE DICOM file: a volume of one slice.  Geometry as
    itk::GDCMImageIO reports it for a single file: spacing (column spacing, row spacing, SpacingBetweenSlices or 1.0),
    direction columns = row cosines, column cosines, their cross product, origin = ImagePositionPatient."""
    arrs, errs = read_dicom_files([Path(file_path)])
    if arrs[0] is None:
        raise ValueError(errs[0])
    info = _lib.DicomInfo()
    _lib.load().svb_dicom_read_headers((C.c_char_p * 1)(os.fsencode(str(file_path))), 1, C.byref(info), 1, None)
    row, col = np.array(info.orientation[0:3]), np.array(info.orientation[3:6])
    direction = np.stack([row, col, np.cross(row, col)], axis=1)
    integral = float(info.rescale_slope).is_integer() and float(info.rescale_intercept).is_integer()
    kind = _lib.PIXEL_FLOAT
    if integral and info.bits_allocated == 16:
        kind = _lib.PIXEL_INT16 if (info.pixel_representation == 1 or info.rescale_intercept < 0) else _lib.PIXEL_UINT16
    elif integral and info.bits_allocated == 8 and info.pixel_representation == 0 and info.rescale_intercept >= 0:
        kind = _lib.PIXEL_UINT8
    return MedicalVolume(array=arrs[0][None], spacing=(float(info.pixel_spacing[1]), float(info.pixel_spacing[0]),
                                                       float(info.spacing_between_slices) or 1.0),
                         direction=tuple(float(v) for v in direction.ravel()), origin=tuple(float(v) for v in info.position[:]),
                         integer_pixels=integral, pixel_kind=kind, meta={"format": "DICOM_FILE"})


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
    if fmt == "NIFTI":
        return read_nifti(path)
    if fmt == "NRRD":
        return read_nrrd(path)
    if fmt == "DICOM_FILE":
        return read_dicom_file(path)
    raise ValueError(f"Unsupported format for path: {path}")


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
