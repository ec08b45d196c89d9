"""Deterministic synthetic RSNA/SPIDER-shaped inputs (there is no dataset access).

Shapes follow what the reference produces after ``resample_to_isotropic`` +
``extract_middle_slice`` (``datasets/classification/cropping.py:37-79``): a
float32 sagittal plane at 0.3 mm isotropic spacing, e.g. 512 px @ 0.7 mm ->
``int(round(512*0.7/0.3)) = 1195`` px.  Intensities are
MRI-like: non-negative, smooth anatomy + a bright vertebral column with five
darker discs + noise, background exactly 0.

Only NumPy's PCG64 ``Generator`` streams are used, which are stable across
NumPy versions, so a seed names the same slice in the build container and on
the GPU box.
"""

from __future__ import annotations

import numpy as np

ISO_SPACING = 0.3
LEVEL_Y = (0.28, 0.38, 0.48, 0.58, 0.66)
COLUMN_X = 0.48


def iso_size(size_px: int, spacing_mm: float, iso: float = ISO_SPACING) -> int:
    """cropping.py:45-48: ``int(round(osz * osp / nsp))``."""
    return int(round(size_px * spacing_mm / iso))


def _upsample_linear(low: np.ndarray, h: int, w: int) -> np.ndarray:
    ly, lx = low.shape
    ys = np.linspace(0, ly - 1, h, dtype=np.float32)
    xs = np.linspace(0, lx - 1, w, dtype=np.float32)
    y0 = np.minimum(ys.astype(np.int32), ly - 2)
    x0 = np.minimum(xs.astype(np.int32), lx - 2)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    a = low[y0][:, x0]
    b = low[y0][:, x0 + 1]
    c = low[y0 + 1][:, x0]
    d = low[y0 + 1][:, x0 + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def make_iso_slice(seed: int, h: int = 1195, w: int = 1195, dtype=np.float32) -> np.ndarray:
    """One synthetic isotropic middle sagittal slice, ``[h, w]`` float32, range ~0..1500."""
    rng = np.random.default_rng(seed)
    low = rng.random((20, 20), dtype=np.float32)
    img = 250.0 + 450.0 * _upsample_linear(low, h, w)
    yy = np.linspace(0, 1, h, dtype=np.float32)[:, None]
    xx = np.linspace(0, 1, w, dtype=np.float32)[None, :]
    cx = COLUMN_X + 0.02 * np.float32(rng.standard_normal())
    # vertebral column: bright band, gently curved
    curve = cx + 0.04 * np.sin((yy - 0.2) * 3.0)
    img += 500.0 * np.exp(-(((xx - curve) / 0.06) ** 2))
    # five discs: darker flattened blobs
    for ly in LEVEL_Y:
        y0 = ly + 0.01 * np.float32(rng.standard_normal())
        cxl = cx + 0.04 * np.sin((y0 - 0.2) * 3.0)
        img -= 420.0 * np.exp(-(((xx - cxl) / 0.05) ** 2) - (((yy - y0) / 0.012) ** 2))
    # body outline: zero background outside an ellipse
    body = (((xx - 0.5) / 0.47) ** 2 + ((yy - 0.5) / 0.49) ** 2) < 1.0
    noise = rng.gamma(4.0, 12.0, size=(h, w)).astype(np.float32)
    img = np.maximum(img + noise - 48.0, 0.0) * body
    return np.ascontiguousarray(img.astype(dtype))


def make_volume(seed: int, n_slices: int = 15, h: int = 512, w: int = 512, spacing=(0.7, 0.7, 4.0)):
    """A synthetic sagittal SERIES as the reference reads it: array ``[z, y, x] = [L, I, P]``
    float32 (slices stacked Left-Right, rows superior->inferior, columns anterior->posterior), ``GetSpacing()`` =
    (P, I, L) mm, and the direction matrix of such an acquisition (image x -> Posterior, y -> Inferior, z -> Left).
    Slices drift smoothly with z so that the interpolation between the two middle slices matters."""
    rng = np.random.default_rng(10_000 + seed)
    base = make_iso_slice(seed, h, w)
    drift = _upsample_linear(rng.random((6, 6), dtype=np.float32), h, w)
    vol = np.stack([base * (1.0 + 0.04 * (z - n_slices / 2) * (drift - 0.5)) for z in range(n_slices)]).astype(np.float32)
    direction = (0.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, -1.0, 0.0)  # columns: x -> +y_LPS (P), y -> -z_LPS (I), z -> +x_LPS (L)
    return vol, tuple(float(s) for s in spacing), direction


def make_batch(seeds, h: int = 1195, w: int = 1195) -> list[np.ndarray]:
    return [make_iso_slice(int(s), h, w) for s in seeds]


def ragged_shapes(n: int, seed: int = 0) -> list[tuple[int, int]]:
    """Config 3: in-plane H,W ~ U{320..1024}, spacing ~ U(0.30, 0.95) mm -> iso sizes."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        hh, ww = int(rng.integers(320, 1025)), int(rng.integers(320, 1025))
        sp = float(rng.uniform(0.30, 0.95))
        out.append((iso_size(hh, sp), iso_size(ww, sp)))
    return out


def make_coords(n_series: int, seed: int = 0, border_frac: float = 0.01, hw=(1195, 1195)) -> np.ndarray:
    """Config 4: ``[n,5,2]`` float32 (x,y); x ~ N(0.48,0.03), y ~ level mean + N(0,0.02);
    ``border_frac`` of the points are forced within 40 px of a border to exercise clipping."""
    rng = np.random.default_rng(seed)
    xy = np.empty((n_series, 5, 2), dtype=np.float32)
    xy[:, :, 0] = rng.normal(COLUMN_X, 0.03, size=(n_series, 5))
    xy[:, :, 1] = np.asarray(LEVEL_Y, dtype=np.float32)[None, :] + rng.normal(0, 0.02, size=(n_series, 5))
    force = rng.random((n_series, 5)) < border_frac
    side = rng.integers(0, 4, size=(n_series, 5))
    off = rng.random((n_series, 5)) * 40.0
    h, w = hw
    xy[:, :, 0] = np.where(force & (side == 0), off / w, xy[:, :, 0])
    xy[:, :, 0] = np.where(force & (side == 1), 1.0 - (off + 1) / w, xy[:, :, 0])
    xy[:, :, 1] = np.where(force & (side == 2), off / h, xy[:, :, 1])
    xy[:, :, 1] = np.where(force & (side == 3), 1.0 - (off + 1) / h, xy[:, :, 1])
    return np.clip(xy, 0.0, np.float32(0.99999)).astype(np.float32)


CONVNEXT_VARIANTS = {
    "base": ((3, 3, 27, 3), (128, 256, 512, 1024)),
    "tiny": ((3, 3, 9, 3), (96, 192, 384, 768)),
    "small": ((3, 3, 27, 3), (96, 192, 384, 768)),
    "large": ((3, 3, 27, 3), (192, 384, 768, 1536)),
    "xlarge": ((3, 3, 27, 3), (256, 512, 1024, 2048)),
}


def random_state_dict(variant: str = "base", seed: int = 0, num_levels: int = 5, trained_like: bool = False):
    """Random-init ``CoordinateRegressor`` state dict under the reference's key names
    (timm ConvNeXt scheme + ``head.{0,2,5}``): conv/linear ~ trunc_normal(std=0.02),
    zero biases, LayerNorm (1, 0), layer-scale 1e-6 -- timm's init.  ``trained_like`` widens the
    layer scale to U(0.1, 1) so that every block contributes."""
    import torch

    depths, dims = CONVNEXT_VARIANTS[variant]
    g = torch.Generator().manual_seed(seed)

    def tn(*shape):
        return torch.nn.init.trunc_normal_(torch.empty(*shape), std=0.02, generator=g)

    sd = {}
    sd["backbone.stem.0.weight"], sd["backbone.stem.0.bias"] = tn(dims[0], 3, 4, 4), torch.zeros(dims[0])
    sd["backbone.stem.1.weight"], sd["backbone.stem.1.bias"] = torch.ones(dims[0]), torch.zeros(dims[0])
    for s, (d, c) in enumerate(zip(depths, dims)):
        p = f"backbone.stages.{s}."
        if s > 0:
            sd[p + "downsample.0.weight"], sd[p + "downsample.0.bias"] = torch.ones(dims[s - 1]), torch.zeros(dims[s - 1])
            sd[p + "downsample.1.weight"], sd[p + "downsample.1.bias"] = tn(c, dims[s - 1], 2, 2), torch.zeros(c)
        for j in range(d):
            q = p + f"blocks.{j}."
            sd[q + "gamma"] = (torch.rand(c, generator=g) * 0.9 + 0.1) if trained_like else torch.full((c,), 1e-6)
            sd[q + "conv_dw.weight"], sd[q + "conv_dw.bias"] = tn(c, 1, 7, 7), torch.zeros(c)
            sd[q + "norm.weight"], sd[q + "norm.bias"] = torch.ones(c), torch.zeros(c)
            sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"] = tn(4 * c, c), torch.zeros(4 * c)
            sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"] = tn(c, 4 * c), torch.zeros(c)
    sd["backbone.head.norm.weight"], sd["backbone.head.norm.bias"] = torch.ones(dims[3]), torch.zeros(dims[3])
    sd["head.0.weight"], sd["head.0.bias"] = torch.ones(dims[3]), torch.zeros(dims[3])
    sd["head.2.weight"], sd["head.2.bias"] = tn(256, dims[3]), torch.zeros(256)
    sd["head.5.weight"], sd["head.5.bias"] = tn(num_levels * 2, 256), torch.zeros(num_levels * 2)
    return sd


# ------------------------------------------------------------------------------------------ synthetic dataset trees
_MET_TYPES = {"int8": "MET_CHAR", "uint8": "MET_UCHAR", "int16": "MET_SHORT", "uint16": "MET_USHORT", "int32": "MET_INT",
              "uint32": "MET_UINT", "int64": "MET_LONG_LONG", "uint64": "MET_ULONG_LONG", "float32": "MET_FLOAT",
              "float64": "MET_DOUBLE"}


def write_metaimage(path, array_zyx: np.ndarray, spacing_xyz, direction=None, origin=(0.0, 0.0, 0.0), compressed: bool = True,
                    big_endian: bool = False, separate_raw: bool = False) -> None:
    """Write a MetaImage the way ITK's MetaImageIO lays it out (for synthetic test / bench trees only).  ``direction`` is
    ``image.GetDirection()`` (row-major, column a = cosine of image axis a); the file's TransformMatrix is its transpose."""
    import zlib
    from pathlib import Path

    path = Path(path)
    a = np.ascontiguousarray(array_zyx)
    nd = a.ndim
    d = np.eye(nd) if direction is None else np.asarray(direction, dtype=np.float64).reshape(nd, nd)
    raw = a.astype(a.dtype.newbyteorder(">" if big_endian else "<"), copy=False).tobytes()
    payload = zlib.compress(raw, 6) if compressed else raw
    fmt = lambda vals: " ".join(repr(float(v)) if not float(v).is_integer() else str(int(v)) for v in vals)  # noqa: E731
    lines = ["ObjectType = Image", f"NDims = {nd}", "BinaryData = True", f"BinaryDataByteOrderMSB = {big_endian}",
             f"CompressedData = {compressed}"]
    if compressed:
        lines.append(f"CompressedDataSize = {len(payload)}")
    lines += [f"TransformMatrix = {fmt(d.T.ravel())}", f"Offset = {fmt(origin[:nd])}", f"CenterOfRotation = {fmt([0] * nd)}",
              "AnatomicalOrientation = RAI", f"ElementSpacing = {fmt(spacing_xyz[:nd])}",
              f"DimSize = {' '.join(str(s) for s in a.shape[::-1])}", f"ElementType = {_MET_TYPES[a.dtype.name]}"]
    if separate_raw:
        raw_name = path.with_suffix(".zraw" if compressed else ".raw").name
        lines.append(f"ElementDataFile = {raw_name}")
        path.write_text("\n".join(lines) + "\n")
        (path.parent / raw_name).write_bytes(payload)
    else:
        lines.append("ElementDataFile = LOCAL")
        path.write_bytes(("\n".join(lines) + "\n").encode("ascii") + payload)


SPIDER_LABEL_COLUMNS = ["Patient", "IVD label", "Modic", "UP endplate", "LOW endplate", "Spondylolisthesis", "Disc herniation",
                        "Disc narrowing", "Disc bulging", "Pfirrman grade"]


def make_spider_tree(base_path, n_patients: int = 3, seed: int = 0, in_plane=(96, 88), n_slices: int = 9, dtype="int16",
                     spacing=(0.8, 0.75, 3.3), missing_t1=(2,)):
    """A SPIDER-shaped raw dataset under ``base_path/raw/SPIDER`` (what ``process_spider`` walks, spider.py:62-108):
    ``images/{patient}_{t1,t2}.mha`` (sagittal stacks: image x -> Posterior, y -> Inferior, z -> Left, integer pixels,
    zlib-compressed) and ``radiological_gradings.csv`` with SPIDER's level numbering (1 = L5/S1).  Patients listed in
    ``missing_t1`` have no T1 file; the last patient carries a level 7 row (ignored by the drivers) and lacks level 1.
    Deterministic in (seed, sizes); returns the list of patient ids."""
    import csv
    from pathlib import Path

    root = Path(base_path) / "raw" / "SPIDER"
    (root / "images").mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(50_000 + seed)
    h, w = in_plane
    pids = [int(p) for p in (1 + np.arange(n_patients) * 3)]
    rows = []
    for k, pid in enumerate(pids):
        for si, suffix in enumerate(("t1", "t2")):
            if suffix == "t1" and k in missing_t1:
                continue
            vol, _, direction = make_volume(1000 * seed + 10 * pid + si, n_slices, h, w, spacing)
            vol = vol * (1.0 if suffix == "t2" else 0.6)
            arr = np.clip(np.rint(vol), 0, 32000).astype(dtype) if np.issubdtype(np.dtype(dtype), np.integer) else vol.astype(dtype)
            write_metaimage(root / "images" / f"{pid}_{suffix}.mha", arr, spacing, direction, origin=(-12.5, 30.0, 7.25))
        levels = list(range(1, 6))
        if k == n_patients - 1:
            levels = [2, 3, 4, 5, 7]
        for lvl in levels:
            rows.append({"Patient": pid, "IVD label": lvl, "Modic": int(rng.integers(0, 4)), "UP endplate": int(rng.integers(0, 2)),
                         "LOW endplate": int(rng.integers(0, 2)), "Spondylolisthesis": int(rng.integers(0, 2)),
                         "Disc herniation": int(rng.integers(0, 2)), "Disc narrowing": int(rng.integers(0, 2)),
                         "Disc bulging": int(rng.integers(0, 2)), "Pfirrman grade": int(rng.integers(1, 6))})
    with open(root / "radiological_gradings.csv", "w", newline="") as f:
        wr = csv.DictWriter(f, fieldnames=SPIDER_LABEL_COLUMNS)
        wr.writeheader()
        wr.writerows(rows)
    return pids


# ------------------------------------------------------------------------------------------ synthetic DICOM series
def _dcm_elem(group: int, elem: int, vr: str, value: bytes, explicit: bool = True) -> bytes:
    import struct

    if len(value) % 2:
        value += b"\x00" if vr in ("UI", "OB", "OW") else b" "
    tag = struct.pack("<HH", group, elem)
    if not explicit:
        return tag + struct.pack("<I", len(value)) + value
    if vr in ("OB", "OW", "SQ", "UN", "UT"):
        return tag + vr.encode() + b"\x00\x00" + struct.pack("<I", len(value)) + value
    return tag + vr.encode() + struct.pack("<H", len(value)) + value


def rle_encode_frame(px: np.ndarray) -> bytes:
    """DICOM RLE Lossless (PS3.5 Annex G) of one frame: a 64-byte header (segment count, 15 offsets) and one PackBits-coded
    byte plane per byte of the sample, most significant first.  Synthetic test data only."""
    import struct

    a = np.ascontiguousarray(px)
    nbytes = a.dtype.itemsize
    planes = a.astype(a.dtype.newbyteorder("<"), copy=False).view(np.uint8).reshape(-1, nbytes)
    segs = []
    for k in range(nbytes - 1, -1, -1):  # most significant byte first
        data = planes[:, k].tobytes()
        out = bytearray()
        i, n = 0, len(data)
        while i < n:
            run = 1
            while i + run < n and run < 128 and data[i + run] == data[i]:
                run += 1
            if run >= 3:
                out += bytes([257 - run, data[i]])
                i += run
                continue
            j = i
            while j < n and j - i < 128:
                if j + 2 < n and data[j] == data[j + 1] == data[j + 2]:
                    break
                j += 1
            out += bytes([j - i - 1]) + data[i:j]
            i = j
        if len(out) % 2:
            out += b"\x00"
        segs.append(bytes(out))
    offs, o = [], 64
    for sg in segs:
        offs.append(o)
        o += len(sg)
    return struct.pack("<16I", len(segs), *(offs + [0] * (15 - len(offs)))) + b"".join(segs)


def jpeg_lossless_encode_frame(px: np.ndarray, predictor: int = 1, point_transform: int = 0, restart_lines: int = 0,
                               precision: int | None = None) -> bytes:
    """JPEG Lossless, process 14 (ITU-T T.81 Annex H, SOF3 + Huffman) of one monochrome frame -- what DICOM transfer syntaxes
    1.2.840.10008.1.2.4.57 / .70 carry.  ``predictor`` 1-7, ``point_transform`` Pt, ``restart_lines`` > 0 adds a DRI segment and
    RSTn markers every that many lines.  Samples are taken modulo 2^16 (two's-complement int16 as stored).  Synthetic test
    data only (pure Python: small frames)."""
    import struct

    a = np.ascontiguousarray(px)
    rows, cols = a.shape
    P = precision or a.dtype.itemsize * 8
    v = (a.astype(np.int64) & ((1 << (a.dtype.itemsize * 8)) - 1)) >> point_transform
    init = 1 << (P - point_transform - 1)
    pred = np.zeros_like(v)
    ra = np.zeros_like(v); ra[:, 1:] = v[:, :-1]
    rb = np.zeros_like(v); rb[1:] = v[:-1]
    rc = np.zeros_like(v); rc[1:, 1:] = v[:-1, :-1]
    table = {1: ra, 2: rb, 3: rc, 4: ra + rb - rc, 5: ra + ((rb - rc) >> 1), 6: rb + ((ra - rc) >> 1), 7: (ra + rb) >> 1}
    pred[:] = table[predictor]
    pred[:, 0] = rb[:, 0]            # first sample of a line: the one above
    first = [0] if not restart_lines else list(range(0, rows, restart_lines))
    for y in first:                  # first line of the scan / of a restart interval: left neighbour, then 2^(P-Pt-1)
        pred[y] = ra[y]
        pred[y, 0] = init
    diff = (v - pred) & 0xFFFF
    diff = np.where(diff > 32768, diff - 65536, diff)  # -32767 .. 32768
    mag = np.abs(diff)
    cat = np.where(mag == 0, 0, np.floor(np.log2(np.maximum(mag, 1))).astype(np.int64) + 1)
    # canonical Huffman table over the 17 categories: lengths 2, 3 x5, 4 .. 14 (Kraft sum < 1, no all-ones code), most frequent first
    order = [int(c) for c in np.argsort(-np.bincount(cat.ravel(), minlength=17), kind="stable")]
    bits = [0, 1, 5, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0]
    codes, code, k = {}, 0, 0
    for ln, cnt in enumerate(bits, start=1):
        for _ in range(cnt):
            codes[order[k]] = (code, ln)
            code += 1
            k += 1
        code <<= 1
    out = bytearray(b"\xff\xd8")
    out += b"\xff\xc3" + struct.pack(">HBHHB", 11, P, rows, cols, 1) + bytes([1, 0x11, 0])
    out += b"\xff\xc4" + struct.pack(">H", 2 + 1 + 16 + 17) + bytes([0x00]) + bytes(bits) + bytes(order)
    if restart_lines:
        out += b"\xff\xdd" + struct.pack(">HH", 4, restart_lines * cols)
    out += b"\xff\xda" + struct.pack(">HB", 8, 1) + bytes([1, 0x00, predictor, 0, point_transform])
    acc = nbits = 0

    def flush_bytes():
        nonlocal acc, nbits
        while nbits >= 8:
            b = (acc >> (nbits - 8)) & 0xFF
            out.append(b)
            if b == 0xFF:
                out.append(0)
            nbits -= 8
        acc &= (1 << nbits) - 1

    rst = 0
    for y in range(rows):
        if restart_lines and y and y % restart_lines == 0:
            if nbits:
                acc = (acc << (8 - nbits)) | ((1 << (8 - nbits)) - 1)
                nbits = 8
                flush_bytes()
            out += bytes([0xFF, 0xD0 + rst])
            rst = (rst + 1) & 7
        for x in range(cols):
            c = int(cat[y, x])
            cd, ln = codes[c]
            acc = (acc << ln) | cd
            nbits += ln
            if 0 < c < 16:
                d = int(diff[y, x])
                if d < 0:
                    d += (1 << c) - 1
                acc = (acc << c) | d
                nbits += c
            flush_bytes()
    if nbits:
        acc = (acc << (8 - nbits)) | ((1 << (8 - nbits)) - 1)
        nbits = 8
        flush_bytes()
    return bytes(out) + b"\xff\xd9"


def write_dicom_slice(path, pixels: np.ndarray, position, row_cos, col_cos, pixel_spacing, series_uid: str, instance: int,
                      slice_thickness: float = 4.0, rescale=None, explicit: bool = True, with_sequence: bool = True,
                      compress: str | None = None, codec_kw: dict | None = None, fragments: int = 1, bits_stored: int | None = None,
                      photometric: bytes = b"MONOCHROME2") -> None:
    """One single-frame MR slice as a DICOM Part-10 file (synthetic test / bench trees only).
    ``explicit`` selects Explicit VR Little Endian, else Implicit VR Little Endian.  ``with_sequence`` adds a sequence and
    an item of undefined length in front of the pixel data (real files carry them; the parser has to skip them).
    ``compress`` = "rle" (1.2.840.10008.1.2.5) or "jpeg" (JPEG Lossless SV1, 1.2.840.10008.1.2.4.70; ``codec_kw`` goes to the
    encoder) writes encapsulated pixel data: an empty Basic Offset Table and the frame cut into ``fragments`` items."""
    import struct
    from pathlib import Path

    px = np.ascontiguousarray(pixels)
    assert px.ndim == 2 and px.dtype in (np.uint16, np.int16, np.uint8)
    ds = lambda vals: "\\".join(repr(float(v)) for v in vals).encode()  # noqa: E731
    us = lambda v: struct.pack("<H", v)  # noqa: E731
    ts = "1.2.840.10008.1.2.1" if explicit else "1.2.840.10008.1.2"
    if compress is not None:
        assert explicit, "encapsulated transfer syntaxes are explicit VR little endian"
        ts = {"rle": "1.2.840.10008.1.2.5", "jpeg": "1.2.840.10008.1.2.4.70", "jpeg57": "1.2.840.10008.1.2.4.57",
              "jpeg2000": "1.2.840.10008.1.2.4.90"}[compress]
    meta = (_dcm_elem(0x0002, 0x0001, "OB", b"\x00\x01") + _dcm_elem(0x0002, 0x0002, "UI", b"1.2.840.10008.5.1.4.1.1.4") +
            _dcm_elem(0x0002, 0x0003, "UI", f"{series_uid}.{instance}".encode()) + _dcm_elem(0x0002, 0x0010, "UI", ts.encode()) +
            _dcm_elem(0x0002, 0x0012, "UI", b"1.2.826.0.1.3680043.9.7433.1"))
    meta = _dcm_elem(0x0002, 0x0000, "UL", struct.pack("<I", len(meta))) + meta
    e = lambda g, el, vr, v: _dcm_elem(g, el, vr, v, explicit)  # noqa: E731
    body = e(0x0008, 0x0060, "CS", b"MR")
    if with_sequence:
        inner = e(0x0008, 0x1150, "UI", b"1.2.840.10008.5.1.4.1.1.4")
        item = struct.pack("<HHI", 0xFFFE, 0xE000, 0xFFFFFFFF) + inner + struct.pack("<HHI", 0xFFFE, 0xE00D, 0)
        seq = item + struct.pack("<HHI", 0xFFFE, 0xE0DD, 0)
        body += struct.pack("<HH", 0x0008, 0x1140) + ((b"SQ\x00\x00") if explicit else b"") + struct.pack("<I", 0xFFFFFFFF) + seq
    body += e(0x0018, 0x0050, "DS", repr(float(slice_thickness)).encode())
    body += e(0x0020, 0x000D, "UI", b"1.2.3.4.5") + e(0x0020, 0x000E, "UI", series_uid.encode())
    body += e(0x0020, 0x0013, "IS", str(instance).encode())
    body += e(0x0020, 0x0032, "DS", ds(position)) + e(0x0020, 0x0037, "DS", ds(list(row_cos) + list(col_cos)))
    body += e(0x0028, 0x0002, "US", us(1)) + e(0x0028, 0x0004, "CS", photometric)
    body += e(0x0028, 0x0010, "US", us(px.shape[0])) + e(0x0028, 0x0011, "US", us(px.shape[1]))
    body += e(0x0028, 0x0030, "DS", ds(pixel_spacing))
    bits = px.dtype.itemsize * 8
    stored = bits_stored or bits
    body += e(0x0028, 0x0100, "US", us(bits)) + e(0x0028, 0x0101, "US", us(stored)) + e(0x0028, 0x0102, "US", us(stored - 1))
    body += e(0x0028, 0x0103, "US", us(1 if px.dtype == np.int16 else 0))
    if rescale is not None:
        body += e(0x0028, 0x1052, "DS", repr(float(rescale[1])).encode()) + e(0x0028, 0x1053, "DS", repr(float(rescale[0])).encode())
    if compress is None:
        body += e(0x7FE0, 0x0010, "OW" if bits > 8 else "OB", px.astype(px.dtype.newbyteorder("<"), copy=False).tobytes())
    else:
        if compress == "rle":
            frame = rle_encode_frame(px)
        elif compress == "jpeg2000":
            frame = b"\xff\x4f\xff\x51" + b"\x00" * 60  # not a decodable stream: only the transfer syntax matters (it is refused)
        else:
            frame = jpeg_lossless_encode_frame(px, **(codec_kw or {}))
        if len(frame) % 2:
            frame += b"\x00"
        item = lambda b: struct.pack("<HHI", 0xFFFE, 0xE000, len(b)) + b  # noqa: E731
        cut = [len(frame) * k // fragments // 2 * 2 for k in range(fragments)] + [len(frame)]
        body += struct.pack("<HH", 0x7FE0, 0x0010) + b"OB\x00\x00" + struct.pack("<I", 0xFFFFFFFF) + item(b"")
        body += b"".join(item(frame[cut[k] : cut[k + 1]]) for k in range(fragments)) + struct.pack("<HHI", 0xFFFE, 0xE0DD, 0)
    Path(path).write_bytes(b"\x00" * 128 + b"DICM" + meta + body)


def write_dicom_series(folder, array_zyx: np.ndarray, spacing_xyz, direction, origin=(0.0, 0.0, 0.0), series_uid="1.2.826.1.100",
                       explicit: bool = True, shuffle_seed: int | None = 0, rescale=None, compress: str | None = None) -> None:
    """A volume as one DICOM file per z index.  File names deliberately do NOT follow the slice order (a reader has to sort
    by position); ``direction`` as ``image.GetDirection()`` (columns = axes)."""
    from pathlib import Path

    folder = Path(folder)
    folder.mkdir(parents=True, exist_ok=True)
    d = np.asarray(direction, dtype=np.float64).reshape(3, 3)
    n = array_zyx.shape[0]
    names = np.arange(n)
    if shuffle_seed is not None:
        names = np.random.default_rng(shuffle_seed).permutation(n)
    for k in range(n):
        pos = np.asarray(origin, dtype=np.float64) + d[:, 2] * spacing_xyz[2] * k
        write_dicom_slice(folder / f"IM{int(names[k]):04d}.dcm", array_zyx[k], pos, d[:, 0], d[:, 1], (spacing_xyz[1], spacing_xyz[0]),
                          series_uid, instance=n - k, slice_thickness=spacing_xyz[2], rescale=rescale, explicit=explicit, compress=compress)


def write_nifti(path, array_zyx: np.ndarray, spacing_xyz, direction=None, origin=(0.0, 0.0, 0.0), use_sform: bool = False,
                scl=None, big_endian: bool = False) -> None:
    """A single-file NIfTI-1 volume (``.nii``, gzip when the name ends in ``.gz``) holding the LPS geometry given the way ITK
    writes it: RAS rotation (x and y rows negated) as a quaternion qform, or as sform rows with ``use_sform``.  Test data only."""
    import gzip
    import struct
    from pathlib import Path

    a = np.ascontiguousarray(array_zyx)
    codes = {"uint8": 2, "int16": 4, "int32": 8, "float32": 16, "float64": 64, "int8": 256, "uint16": 512, "uint32": 768}
    e = ">" if big_endian else "<"
    nz, ny, nx = a.shape
    D = np.eye(3) if direction is None else np.asarray(direction, dtype=np.float64).reshape(3, 3)
    flip = np.array([-1.0, -1.0, 1.0])
    R = D * flip[:, None]
    off = np.asarray(origin, dtype=np.float64) * flip
    qfac = 1.0
    if np.linalg.det(R) < 0:
        R = R.copy()
        R[:, 2] = -R[:, 2]
        qfac = -1.0
    # rotation matrix -> quaternion (a >= 0), nifti_mat44_to_quatern
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        qa = 0.5 * np.sqrt(1 + tr); qb = 0.25 * (R[2, 1] - R[1, 2]) / qa; qc = 0.25 * (R[0, 2] - R[2, 0]) / qa; qd = 0.25 * (R[1, 0] - R[0, 1]) / qa
    else:
        xd, yd, zd = 1 + R[0, 0] - (R[1, 1] + R[2, 2]), 1 + R[1, 1] - (R[0, 0] + R[2, 2]), 1 + R[2, 2] - (R[0, 0] + R[1, 1])
        if xd > 1:
            qb = 0.5 * np.sqrt(xd); qc = 0.25 * (R[0, 1] + R[1, 0]) / qb; qd = 0.25 * (R[0, 2] + R[2, 0]) / qb; qa = 0.25 * (R[2, 1] - R[1, 2]) / qb
        elif yd > 1:
            qc = 0.5 * np.sqrt(yd); qb = 0.25 * (R[0, 1] + R[1, 0]) / qc; qd = 0.25 * (R[1, 2] + R[2, 1]) / qc; qa = 0.25 * (R[0, 2] - R[2, 0]) / qc
        else:
            qd = 0.5 * np.sqrt(zd); qb = 0.25 * (R[0, 2] + R[2, 0]) / qd; qc = 0.25 * (R[1, 2] + R[2, 1]) / qd; qa = 0.25 * (R[1, 0] - R[0, 1]) / qd
        if qa < 0:
            qb, qc, qd = -qb, -qc, -qd
    h = bytearray(352)
    struct.pack_into(e + "i", h, 0, 348)
    struct.pack_into(e + "8h", h, 40, 3, nx, ny, nz, 1, 1, 1, 1)
    struct.pack_into(e + "h", h, 70, codes[a.dtype.name])
    struct.pack_into(e + "h", h, 72, a.dtype.itemsize * 8)
    struct.pack_into(e + "8f", h, 76, qfac, float(spacing_xyz[0]), float(spacing_xyz[1]), float(spacing_xyz[2]), 0, 0, 0, 0)
    struct.pack_into(e + "f", h, 108, 352.0)
    struct.pack_into(e + "2f", h, 112, *(scl if scl is not None else (1.0, 0.0)))
    if use_sform:
        struct.pack_into(e + "2h", h, 252, 0, 1)
        M = (D * flip[:, None]) * np.asarray(spacing_xyz, dtype=np.float64)[None, :]
        struct.pack_into(e + "12f", h, 280, *np.concatenate([M, off[:, None]], axis=1).ravel())
    else:
        struct.pack_into(e + "2h", h, 252, 1, 0)
        struct.pack_into(e + "6f", h, 256, float(qb), float(qc), float(qd), *off)
    h[344:348] = b"n+1\x00"
    blob = bytes(h) + a.astype(a.dtype.newbyteorder(e), copy=False).tobytes()
    Path(path).write_bytes(gzip.compress(blob) if str(path).endswith(".gz") else blob)


def write_nrrd(path, array_zyx: np.ndarray, spacing_xyz, direction=None, origin=(0.0, 0.0, 0.0), gz: bool = False,
               space: str = "left-posterior-superior", detached: bool = False) -> None:
    """A 3-D NRRD (attached data, or a ``.nhdr``-style detached ``data file``) with ``space directions`` / ``space origin``;
    ``space`` = "right-anterior-superior" stores the RAS form of the LPS geometry given.  Test data only."""
    import gzip
    from pathlib import Path

    a = np.ascontiguousarray(array_zyx)
    names = {"uint8": "uchar", "int8": "signed char", "int16": "short", "uint16": "ushort", "int32": "int", "float32": "float", "float64": "double"}
    D = np.eye(3) if direction is None else np.asarray(direction, dtype=np.float64).reshape(3, 3)
    flip = np.array([-1.0, -1.0, 1.0]) if space.startswith("right") else np.ones(3)
    V = (D * flip[:, None]) * np.asarray(spacing_xyz, dtype=np.float64)[None, :]
    o = np.asarray(origin, dtype=np.float64) * flip
    fmt = lambda v: "(" + ",".join(repr(float(x)) for x in v) + ")"  # noqa: E731
    nz, ny, nx = a.shape
    data = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
    if gz:
        data = gzip.compress(data)
    head = ["NRRD0004", "# synthetic", f"type: {names[a.dtype.name]}", "dimension: 3", f"space: {space}", f"sizes: {nx} {ny} {nz}",
            "space directions: " + " ".join(fmt(V[:, k]) for k in range(3)), "kinds: domain domain domain", "endian: little",
            f"encoding: {'gzip' if gz else 'raw'}", "space origin: " + fmt(o)]
    path = Path(path)
    if detached:
        raw = path.with_suffix(".raw.gz" if gz else ".raw")
        raw.write_bytes(data)
        path.write_bytes(("\n".join(head + [f"data file: {raw.name}"]) + "\n").encode())
    else:
        path.write_bytes(("\n".join(head) + "\n\n").encode() + data)


PHENIKAA_LABEL_COLUMNS = ["Patient ID", "IVD label", "Modic_0", "Modic_1", "Modic_2", "Modic_3", "UP endplate", "LOW endplate",
                          "Spondylolisthesis", "Disc herniation", "Disc narrowing", "Disc bulging", "Pfirrman grade"]


def make_phenikaa_tree(base_path, n_patients: int = 2, seed: int = 0, in_plane=(90, 84), n_slices: int = 7, spacing=(0.78, 0.82, 4.4)):
    """A Phenikaa-shaped interim dataset under ``base_path/interim/Phenikaa`` (what ``process_phenikaa`` walks,
    phenikaa.py:131-176): ``images/<patient>/<series folder>/*.dcm`` with the series folders named like the scanner
    exports them ("Sag T2", "SAG  T1" -- matched case- and space-insensitively), uint16 pixels, and
    ``radiological_labels.csv`` (levels 1 = L1/L2 ... 5 = L5/S1, one-hot Modic columns).  The first patient's T2 folder
    also holds a second, shorter series with a lexicographically LARGER SeriesInstanceUID and a non-DICOM file; the last
    patient has no T1 folder.  Deterministic; returns the patient ids."""
    import csv
    from pathlib import Path

    root = Path(base_path) / "interim" / "Phenikaa"
    (root / "images").mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(70_000 + seed)
    h, w = in_plane
    pids = [f"PK{seed:02d}{k:03d}" for k in range(n_patients)]
    rows = []
    for k, pid in enumerate(pids):
        for si, folder in enumerate(("SAG  T1", "Sag T2")):
            if si == 0 and k == n_patients - 1:
                continue
            vol, _, direction = make_volume(2000 * seed + 20 * k + si, n_slices, h, w, spacing)
            arr = np.clip(np.rint(vol * (0.6 if si == 0 else 1.0)), 0, 4095).astype(np.uint16)
            sdir = root / "images" / pid / folder
            write_dicom_series(sdir, arr, spacing, direction, origin=(-40.0 + k, -95.5, 210.25), series_uid=f"1.2.826.1.{100 + 10 * k + si}",
                               explicit=(k + si) % 2 == 0, shuffle_seed=seed + k + si)
            if k == 0 and si == 1:
                write_dicom_series(sdir / "_tmp", arr[:2] // 2, spacing, direction, series_uid="1.2.826.1.999", shuffle_seed=None)
                for f in (sdir / "_tmp").iterdir():
                    f.rename(sdir / f"ZZ_{f.name}")
                (sdir / "_tmp").rmdir()
                (sdir / "notes.txt").write_text("not a DICOM file\n")
        for lvl in range(1, 6):
            modic = int(rng.integers(0, 4))
            row = {"Patient ID": pid, "IVD label": lvl, "UP endplate": int(rng.integers(0, 2)), "LOW endplate": int(rng.integers(0, 2)),
                   "Spondylolisthesis": int(rng.integers(0, 2)), "Disc herniation": int(rng.integers(0, 2)),
                   "Disc narrowing": int(rng.integers(0, 2)), "Disc bulging": int(rng.integers(0, 2)), "Pfirrman grade": int(rng.integers(1, 6))}
            row.update({f"Modic_{i}": int(i == modic) for i in range(4)})
            rows.append(row)
    with open(root / "radiological_labels.csv", "w", newline="") as f:
        wr = csv.DictWriter(f, fieldnames=PHENIKAA_LABEL_COLUMNS)
        wr.writeheader()
        wr.writerows(rows)
    return pids


def make_localization_tree(base_path, seed: int = 0):
    """Raw inputs of ``create_localization_dataset`` (datasets/localization.py:326-382) in miniature, deterministic:

    * ``raw/Lumbar Coords/coords_pretrain.csv`` + ``data/processed_*_jpgs`` (ready-made JPGs: copied byte for byte, so any
      bytes do) + ``data/processed_{lsd,osf}/*.npy`` (arrays that have to be normalised), incl. a repeated file, an unknown
      source and a missing file;
    * ``raw/Lumbar Coords/coords_rsna_improved.csv`` + ``raw/RSNA/train_series_descriptions.csv`` +
      ``raw/RSNA/train_images/<study>/<series>/<instance>.dcm`` (single-slice DICOMs of different sizes, one with a
      rescale, one constant image), incl. rows that every skip rule of ``process_rsna_improved`` removes.
    """
    import csv
    from pathlib import Path

    rng = np.random.default_rng(90_000 + seed)
    lc = Path(base_path) / "raw" / "Lumbar Coords"
    rs = Path(base_path) / "raw" / "RSNA"
    for d in ("processed_spider_jpgs", "processed_lsd_jpgs", "processed_osf_jpgs", "processed_tseg_jpgs", "processed_lsd", "processed_osf",
              "processed_tseg"):
        (lc / "data" / d).mkdir(parents=True, exist_ok=True)
    (lc / "data" / "processed_spider_jpgs" / "sp_001.jpg").write_bytes(b"\xff\xd8\xff\xe0" + bytes(rng.integers(0, 256, 300, dtype=np.uint8)))
    (lc / "data" / "processed_tseg_jpgs" / "ct_7.jpg").write_bytes(b"\xff\xd8\xff\xe0" + bytes(rng.integers(0, 256, 200, dtype=np.uint8)))
    np.save(lc / "data" / "processed_lsd" / "lsd_10.npy", make_iso_slice(300 + seed, 96, 80))
    np.save(lc / "data" / "processed_lsd" / "lsd_11.npy", (make_iso_slice(301 + seed, 70, 131) * 3).astype(np.int16))
    np.save(lc / "data" / "processed_osf" / "osf_3.npy", make_iso_slice(302 + seed, 64, 64).astype(np.float64))
    pre = [("sp_001.jpg", "spider"), ("sp_001.jpg", "spider"), ("lsd_10.jpg", "lsd"), ("lsd_11.jpg", "lsd"), ("lsd_10.jpg", "lsd"),
           ("osf_3.jpg", "osf"), ("ct_7.jpg", "tseg"), ("nope.jpg", "osf"), ("x.jpg", "mystery"), ("sp_404.jpg", "spider")]
    levels = ["L1/L2", "L2/L3", "L3/L4", "L4/L5", "L5/S1"]
    with open(lc / "coords_pretrain.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["filename", "source", "level", "relative_x", "relative_y"])
        for k, (fn, src) in enumerate(pre):
            w.writerow([fn, src, levels[k % 5], round(float(rng.uniform(0.3, 0.7)), 6), round(float(rng.uniform(0.2, 0.8)), 6)])
    series = {(101, 1001): "Sagittal T1", (101, 1002): "Sagittal T2/STIR", (101, 1003): "Axial T2", (202, 2001): "Sagittal T2/STIR"}
    rs.mkdir(parents=True, exist_ok=True)
    with open(rs / "train_series_descriptions.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["study_id", "series_id", "series_description"])
        for (study, sid), desc in series.items():
            w.writerow([study, sid, desc])
    shapes = {(101, 1001, 5): (88, 72), (101, 1001, 6): (88, 72), (101, 1002, 8): (120, 100), (202, 2001, 3): (64, 96), (202, 2001, 4): (64, 96)}
    for k, ((study, sid, inst), (h, wd)) in enumerate(shapes.items()):
        d = rs / "train_images" / str(study) / str(sid)
        d.mkdir(parents=True, exist_ok=True)
        img = np.clip(np.rint(make_iso_slice(400 + seed + k, h, wd)), 0, 4095).astype(np.uint16)
        if k == 4:
            img[:] = 777  # constant image: normalize_to_uint8 casts the raw values (wraps mod 256)
        write_dicom_slice(d / f"{inst}.dcm", img, (0.0, 0.0, float(inst)), (0, 1, 0), (0, 0, -1), (0.6, 0.6), f"1.2.840.{sid}", inst,
                          rescale=(2.0, -100.0) if k == 2 else None, explicit=k % 2 == 0)
    (rs / "train_images" / "202" / "2001" / "9.dcm").write_bytes(b"broken")
    rows = [(1001, 101, 5, "Left Neural Foraminal Narrowing"), (1001, 101, 5, "Right Neural Foraminal Narrowing"),
            (1001, 101, 6, "Left Neural Foraminal Narrowing"), (1002, 101, 8, "Spinal Canal Stenosis"),
            (1003, 101, 2, "Left Subarticular Stenosis"), (1002, 101, -1, "Spinal Canal Stenosis"), (2001, 202, 3, "Spinal Canal Stenosis"),
            (2001, 202, 4, "Spinal Canal Stenosis"), (2001, 202, 9, "Spinal Canal Stenosis"), (2001, 202, 77, "Spinal Canal Stenosis"),
            (5555, 202, 1, "Spinal Canal Stenosis"), (1003, 101, 2, "Spinal Canal Stenosis"), (2001, 202, 3, "Spinal Canal Stenosis")]
    with open(lc / "coords_rsna_improved.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["series_id", "study_id", "instance_number", "relative_x", "relative_y", "level", "condition"])
        for k, (sid, study, inst, cond) in enumerate(rows):
            w.writerow([sid, study, inst, round(float(rng.uniform(0.3, 0.7)), 6), round(float(rng.uniform(0.2, 0.8)), 6), levels[k % 5], cond])
