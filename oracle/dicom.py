"""Pure-Python restatement of what ``sitk.ImageSeriesReader`` over GDCM does for a folder of single-frame DICOM slices
(``spine_vision/io/readers.py:48-73``), for the subset the Phenikaa series need (native pixel data, monochrome).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  **PARITY UNPINNED**: SimpleITK 2.5.3 / GDCM are absent from the
image.  Restated conventions: series ids = distinct SeriesInstanceUIDs in lexicographic order, the first is read; slices
ordered by the projection of ImagePositionPatient on the normal ``row x col`` (gdcm::IPPSorter); origin = first position;
spacing = (column spacing, row spacing, |last - first| / (n - 1)); direction columns = row cosines, column cosines,
(last - first) normalised; RescaleSlope / Intercept applied.  ``svb_dicom_*`` + ``hostio.read_dicom_series`` are tested
against this module.
"""

from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

from oracle.metaimage import Image

_LONG = {b"OB", b"OW", b"OF", b"OD", b"OL", b"SQ", b"UC", b"UR", b"UT", b"UN"}


def _elements(blob: bytes):
    """Yield (tag, value bytes) of the top-level data elements; undefined-length sequences are skipped as a whole."""
    if blob[128:132] != b"DICM":
        raise ValueError("not a DICOM Part-10 file")
    pos, explicit, ts = 132, True, "1.2.840.10008.1.2.1"

    def header(pos, explicit):
        g, e = struct.unpack_from("<HH", blob, pos)
        if g == 0xFFFE:
            return g, e, struct.unpack_from("<I", blob, pos + 4)[0], pos + 8
        if explicit:
            vr = blob[pos + 4 : pos + 6]
            if vr in _LONG:
                return g, e, struct.unpack_from("<I", blob, pos + 8)[0], pos + 12
            return g, e, struct.unpack_from("<H", blob, pos + 6)[0], pos + 8
        return g, e, struct.unpack_from("<I", blob, pos + 4)[0], pos + 8

    def skip(pos, explicit):
        while True:
            g, e, ln, pos = header(pos, explicit)
            if g == 0xFFFE and e in (0xE0DD, 0xE00D):
                return pos
            pos = skip(pos, explicit) if ln == 0xFFFFFFFF else pos + ln

    while pos + 8 <= len(blob):
        if struct.unpack_from("<H", blob, pos)[0] != 0x0002 and explicit and ts == "1.2.840.10008.1.2":
            explicit = False
        g, e, ln, vpos = header(pos, explicit if struct.unpack_from("<H", blob, pos)[0] != 0x0002 else True)
        if ln == 0xFFFFFFFF:
            pos = skip(vpos, explicit)
            continue
        val = blob[vpos : vpos + ln]
        if (g, e) == (0x0002, 0x0010):
            ts = val.rstrip(b"\x00 ").decode()
            if ts not in ("1.2.840.10008.1.2", "1.2.840.10008.1.2.1"):
                raise ValueError(f"unsupported transfer syntax {ts}")
        yield (g, e), val
        pos = vpos + ln


def read_slice(path: Path) -> dict:
    el = dict(_elements(Path(path).read_bytes()))
    txt = lambda t: el[t].rstrip(b"\x00 ").decode().strip()  # noqa: E731
    nums = lambda t: [float(v) for v in txt(t).split("\\")]  # noqa: E731
    u16 = lambda t: struct.unpack("<H", el[t][:2])[0]  # noqa: E731
    rows, cols, bits, signed = u16((0x0028, 0x0010)), u16((0x0028, 0x0011)), u16((0x0028, 0x0100)), u16((0x0028, 0x0103))
    dt = np.dtype({8: "i1" if signed else "u1", 16: "<i2" if signed else "<u2", 32: "<i4" if signed else "<u4"}[bits])
    px = np.frombuffer(el[(0x7FE0, 0x0010)], dtype=dt, count=rows * cols).reshape(rows, cols)
    slope = float(txt((0x0028, 0x1053))) if (0x0028, 0x1053) in el else 1.0
    inter = float(txt((0x0028, 0x1052))) if (0x0028, 0x1052) in el else 0.0
    return dict(uid=txt((0x0020, 0x000E)), pos=np.array(nums((0x0020, 0x0032))), iop=np.array(nums((0x0020, 0x0037))),
                spacing=nums((0x0028, 0x0030)), px=px, slope=slope, inter=inter, name=Path(path).name)


def read_series(folder: Path) -> Image:
    slices = []
    for f in sorted(Path(folder).iterdir()):
        if f.is_file():
            try:
                slices.append(read_slice(f))
            except Exception:  # noqa: BLE001 -- GDCM's directory scan ignores what it cannot parse
                continue
    if not slices:
        raise ValueError(f"No DICOM series found in {folder}")
    uid = min(s["uid"] for s in slices)
    sel = [s for s in slices if s["uid"] == uid]
    row, col = sel[0]["iop"][:3], sel[0]["iop"][3:]
    normal = np.cross(row, col)
    sel.sort(key=lambda s: (float(np.dot(s["pos"], normal)), s["name"]))
    n = len(sel)
    identity = all(s["slope"] == 1.0 and s["inter"] == 0.0 for s in sel)
    integral = all(float(s["slope"]).is_integer() and float(s["inter"]).is_integer() for s in sel)
    if identity:
        arr = np.stack([s["px"] for s in sel])
    else:
        arr = np.stack([s["px"].astype(np.float64) * s["slope"] + s["inter"] for s in sel])
        if integral:
            arr = arr.astype(np.int32)
    p0, p1 = sel[0]["pos"], sel[-1]["pos"]
    if n > 1 and np.linalg.norm(p1 - p0) > 0:
        dz, third = float(np.linalg.norm(p1 - p0)) / (n - 1), (p1 - p0) / np.linalg.norm(p1 - p0)
    else:
        dz, third = 1.0, normal
    direction = np.stack([row, col, third], axis=1)
    return Image(arr, (sel[0]["spacing"][1], sel[0]["spacing"][0], dz), direction.ravel(), p0)
