"""Pure-Python restatement of what ``sitk.ImageSeriesReader`` over GDCM does for a folder of single-frame DICOM slices
(``spine_vision/io/readers.py:48-73``), for the subset the Phenikaa series need (monochrome; native pixel data, and the two
lossless encapsulated transfer syntaxes GDCM decodes for MR exports: RLE Lossless, PS3.5 Annex G, and JPEG Lossless process 14,
ITU-T T.81 Annex H -- both restated here from the published algorithms, in plain Python).

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
        if ln == 0xFFFFFFFF and (g, e) == (0x7FE0, 0x0010):  # encapsulated pixel data: offset table item, then the fragments
            frags, first = [], True
            while True:
                ig, ie, iln = struct.unpack_from("<HHI", blob, vpos)
                vpos += 8
                if (ig, ie) == (0xFFFE, 0xE0DD):
                    break
                if not first:
                    frags.append(blob[vpos : vpos + iln])
                first = False
                vpos += iln
            yield (g, e), b"".join(frags)
            return
        if ln == 0xFFFFFFFF:
            pos = skip(vpos, explicit)
            continue
        val = blob[vpos : vpos + ln]
        if (g, e) == (0x0002, 0x0010):
            ts = val.rstrip(b"\x00 ").decode()
            if ts not in ("1.2.840.10008.1.2", "1.2.840.10008.1.2.1") + tuple(ENCAPSULATED):
                raise ValueError(f"unsupported transfer syntax {ts}")
            yield (0xFFFF, 0x0001), ts.encode()
        yield (g, e), val
        pos = vpos + ln


ENCAPSULATED = {"1.2.840.10008.1.2.5": "rle", "1.2.840.10008.1.2.4.57": "jpeg", "1.2.840.10008.1.2.4.70": "jpeg"}


def decode_rle(frame: bytes, n_px: int, nbytes: int) -> np.ndarray:
    """PS3.5 Annex G: header of 16 little-endian uint32 (segment count, offsets); segment k is the PackBits coding of byte
    plane ``nbytes - 1 - k`` (most significant first).  Returns little-endian bytes ``[n_px, nbytes]``."""
    head = struct.unpack_from("<16I", frame, 0)
    nseg = head[0]
    if nseg != nbytes:
        raise ValueError(f"RLE frame with {nseg} segments for {nbytes}-byte samples")
    out = np.zeros((n_px, nbytes), dtype=np.uint8)
    for k in range(nseg):
        lo, hi = head[1 + k], (head[2 + k] if k + 1 < nseg else len(frame))
        plane = bytearray()
        q = lo
        while len(plane) < n_px and q < hi:
            c = frame[q] if frame[q] < 128 else frame[q] - 256
            q += 1
            if c >= 0:
                plane += frame[q : q + c + 1]
                q += c + 1
            elif c != -128:
                plane += bytes([frame[q]]) * (1 - c)
                q += 1
        out[:, nbytes - 1 - k] = np.frombuffer(bytes(plane[:n_px]), dtype=np.uint8)
    return out


def decode_jpeg_lossless(frame: bytes, rows: int, cols: int) -> np.ndarray:
    """ITU-T T.81 Annex H, process 14 (SOF3, Huffman): ``Px`` from predictor ``Ss`` (first line: left neighbour, first sample
    2^(P-Pt-1); first sample of other lines: the one above), difference = Huffman category SSSS + SSSS raw bits (EXTEND),
    SSSS = 16 means 32768, arithmetic modulo 2^16, restart intervals re-initialise the prediction.  Returns uint16 ``[rows, cols]``
    already shifted left by the point transform."""
    assert frame[:2] == b"\xff\xd8"
    pos, tables, restart = 2, {}, 0
    while True:
        assert frame[pos] == 0xFF
        m = frame[pos + 1]
        ln = struct.unpack_from(">H", frame, pos + 2)[0]
        d = frame[pos + 4 : pos + 2 + ln]
        if m == 0xC3:
            P, h, w, nc = struct.unpack_from(">BHHB", d, 0)
            assert (h, w, nc) == (rows, cols, 1)
        elif m == 0xC4:
            q = 0
            while q < len(d):
                th = d[q] & 15
                counts = list(d[q + 1 : q + 17])
                vals = d[q + 17 : q + 17 + sum(counts)]
                code, k, tab = 0, 0, {}
                for length, cnt in enumerate(counts, start=1):
                    for _ in range(cnt):
                        tab[(length, code)] = vals[k]
                        code += 1
                        k += 1
                    code <<= 1
                tables[th] = tab
                q += 17 + sum(counts)
        elif m == 0xDD:
            restart = struct.unpack_from(">H", d, 0)[0]
        elif m == 0xDA:
            tab, predictor, pt = tables[d[2] >> 4], d[3], d[5] & 15
            pos += 2 + ln
            break
        pos += 2 + ln
    # un-stuff the entropy-coded segment(s) into a bit string per restart interval
    intervals, cur = [], bytearray()
    while pos < len(frame):
        b = frame[pos]
        if b == 0xFF:
            nxt = frame[pos + 1]
            if nxt == 0x00:
                cur.append(0xFF)
                pos += 2
                continue
            if 0xD0 <= nxt <= 0xD7:
                intervals.append(bytes(cur))
                cur = bytearray()
                pos += 2
                continue
            break  # EOI
        cur.append(b)
        pos += 1
    intervals.append(bytes(cur))
    img = np.zeros((rows, cols), dtype=np.int64)
    lines_per = restart // cols if restart else rows
    y = 0
    for chunk in intervals:
        bits = "".join(f"{b:08b}" for b in chunk)
        bp = 0
        for yy in range(y, min(y + lines_per, rows)):
            for x in range(cols):
                length, code = 0, 0
                while True:
                    code = (code << 1) | int(bits[bp]); bp += 1; length += 1
                    if (length, code) in tab:
                        s = tab[(length, code)]
                        break
                if s == 0:
                    diff = 0
                elif s == 16:
                    diff = 32768
                else:
                    diff = int(bits[bp : bp + s], 2); bp += s
                    if diff < (1 << (s - 1)):
                        diff -= (1 << s) - 1
                if yy == y:
                    px = (1 << (P - pt - 1)) if x == 0 else img[yy, x - 1]
                elif x == 0:
                    px = img[yy - 1, 0]
                else:
                    ra, rb, rc = int(img[yy, x - 1]), int(img[yy - 1, x]), int(img[yy - 1, x - 1])
                    px = {1: ra, 2: rb, 3: rc, 4: ra + rb - rc, 5: ra + ((rb - rc) >> 1), 6: rb + ((ra - rc) >> 1), 7: (ra + rb) >> 1}[predictor]
                img[yy, x] = (int(px) + diff) & 0xFFFF
        y += lines_per
    return ((img << pt) & 0xFFFF).astype(np.uint16)


def read_slice(path: Path) -> dict:
    el = dict(_elements(Path(path).read_bytes()))
    txt = lambda t: el[t].rstrip(b"\x00 ").decode().strip()  # noqa: E731
    nums = lambda t: [float(v) for v in txt(t).split("\\")]  # noqa: E731
    u16 = lambda t: struct.unpack("<H", el[t][:2])[0]  # noqa: E731
    rows, cols, bits, signed = u16((0x0028, 0x0010)), u16((0x0028, 0x0011)), u16((0x0028, 0x0100)), u16((0x0028, 0x0103))
    dt = np.dtype({8: "i1" if signed else "u1", 16: "<i2" if signed else "<u2", 32: "<i4" if signed else "<u4"}[bits])
    codec = ENCAPSULATED.get(el.get((0xFFFF, 0x0001), b"").decode())
    if codec == "rle":
        raw = decode_rle(el[(0x7FE0, 0x0010)], rows * cols, bits // 8).tobytes()
    elif codec == "jpeg":
        u = decode_jpeg_lossless(el[(0x7FE0, 0x0010)], rows, cols)
        raw = (u.astype("<u2") if bits == 16 else u.astype("u1")).tobytes()
    else:
        raw = el[(0x7FE0, 0x0010)]
    px = np.frombuffer(raw, dtype=dt, count=rows * cols).reshape(rows, cols)
    stored = u16((0x0028, 0x0101)) if (0x0028, 0x0101) in el else bits
    if stored < bits:  # GDCM: keep BitsStored bits, sign-extend signed data from the high stored bit
        u = px.astype(np.int64) & ((1 << stored) - 1)
        if signed:
            u = np.where(u >> (stored - 1) & 1, u - (1 << stored), u)
        px = u.astype(dt)
    if txt((0x0028, 0x0004)) == "MONOCHROME1":
        raise ValueError("MONOCHROME1 is not restated")
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
