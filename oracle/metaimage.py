"""NumPy restatement of the MetaImage (.mha / .mhd) reading that ``sitk.ReadImage`` performs for the SPIDER volumes
(``spine_vision/io/readers.py:65-73``; ITK ``MetaImageIO``): header ``key = value`` lines up to ``ElementDataFile``,
then raw or zlib-compressed voxels, x fastest.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  **PARITY UNPINNED**: SimpleITK 2.5.3 is absent from the image;
this restates the published MetaIO layout and ITK's conventions (row i of ``TransformMatrix`` = direction cosine of
image axis i = column i of ``GetDirection()``).  ``svb_mha_read_f32`` is tested against it.
"""

from __future__ import annotations

import zlib
from pathlib import Path

import numpy as np

_TYPES = {"MET_CHAR": "i1", "MET_UCHAR": "u1", "MET_SHORT": "i2", "MET_USHORT": "u2", "MET_INT": "i4", "MET_UINT": "u4",
          "MET_LONG": "i4", "MET_ULONG": "u4", "MET_LONG_LONG": "i8", "MET_ULONG_LONG": "u8", "MET_FLOAT": "f4", "MET_DOUBLE": "f8"}


class Image:
    """The handful of ``sitk.Image`` accessors the path uses."""

    def __init__(self, array_zyx, spacing, direction, origin):
        self.array, self.spacing, self.direction, self.origin = array_zyx, tuple(spacing), tuple(direction), tuple(origin)

    def GetSize(self):
        return tuple(int(s) for s in self.array.shape[::-1])

    def GetSpacing(self):
        return self.spacing

    def GetDirection(self):
        return self.direction

    def GetOrigin(self):
        return self.origin


def read(path) -> Image:
    path = Path(path)
    blob = path.read_bytes()
    hdr: dict[str, str] = {}
    pos = 0
    while True:
        end = blob.index(b"\n", pos)
        line = blob[pos:end].decode("ascii")
        pos = end + 1
        if "=" not in line:
            continue
        k, v = (t.strip() for t in line.split("=", 1))
        hdr[k] = v
        if k == "ElementDataFile":
            break
    nd = int(hdr["NDims"])
    dims = [int(t) for t in hdr["DimSize"].split()]
    dt = np.dtype(_TYPES[hdr["ElementType"]]).newbyteorder(">" if hdr.get("BinaryDataByteOrderMSB", "False")[0] in "Tt1" else "<")
    data = blob[pos:] if hdr["ElementDataFile"] == "LOCAL" else (path.parent / hdr["ElementDataFile"]).read_bytes()
    if hdr.get("CompressedData", "False")[0] in "Tt1":
        n = int(hdr.get("CompressedDataSize", len(data)))
        data = zlib.decompress(data[:n])
    arr = np.frombuffer(data, dtype=dt, count=int(np.prod(dims))).reshape(dims[::-1])
    arr = arr.astype(dt.newbyteorder("="))
    spacing = [float(t) for t in hdr.get("ElementSpacing", " ".join(["1"] * nd)).split()]
    origin = [float(t) for t in hdr.get("Offset", hdr.get("Position", " ".join(["0"] * nd))).split()]
    tm = np.array([float(t) for t in hdr.get("TransformMatrix", " ".join(str(float(i == j)) for i in range(nd) for j in range(nd))).split()])
    direction = tm.reshape(nd, nd).T.ravel()
    return Image(arr, spacing, direction, origin)
