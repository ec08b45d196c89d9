"""K0 host side: from source volumes to the middle isotropic sagittal slices the rest of the path consumes.

Mirrors ``resample_to_isotropic`` + ``extract_middle_slice`` + ``get_slice_spacing``
(``spine_vision/datasets/classification/cropping.py:37-101``) for the one plane that survives: the host resolves
orientation and sizes (integer / double arithmetic on a handful of numbers per series) and uploads only the two
source planes around the middle Left-Right index; the device interpolates (``svb_k0_midplane_resample``).
Parity with SimpleITK is unpinned (see include/spine_b200.h); volumes are taken as float32.
"""

from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib, ops

ISOTROPIC_SPACING = (0.3, 0.3, 0.3)  # cropping.py:22


class K0Series(C.Structure):
    """``svb_k0_series`` (include/spine_b200.h)."""

    _fields_ = [("vol_off", C.c_int64), ("out_off", C.c_int64), ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
                ("ax_row", C.c_int32), ("ax_col", C.c_int32), ("ax_fix", C.c_int32), ("flip_row", C.c_int32),
                ("flip_col", C.c_int32), ("out_h", C.c_int32), ("out_w", C.c_int32), ("fix_lo", C.c_int32),
                ("fix_hi", C.c_int32), ("fix_inside", C.c_int32), ("integer_pixels", C.c_int32), ("fix_frac", C.c_double),
                ("sp_row", C.c_double), ("sp_col", C.c_double), ("new_sp_row", C.c_double), ("new_sp_col", C.c_double)]


def lpi_axes(direction=None):
    """``sitk.DICOMOrient(image, "LPI")`` as an axis assignment: for oriented axis o = 0 (Left), 1 (Posterior),
    2 (Inferior) the source image axis and whether it is reversed.  ``direction`` = ``image.GetDirection()``
    (9 numbers, column a = cosine of image axis a in LPS space); None = identity."""
    d = np.eye(3) if direction is None else np.asarray(direction, dtype=np.float64).reshape(3, 3)
    dom = [int(np.argmax(np.abs(d[:, a]))) for a in range(3)]
    if sorted(dom) != [0, 1, 2]:
        raise ValueError("direction cosines do not resolve to three distinct anatomical axes")
    want_sign = (1.0, 1.0, -1.0)
    axis_of, flip = [0, 0, 0], [False, False, False]
    for a in range(3):
        o = dom[a]
        axis_of[o] = a
        flip[o] = bool(np.sign(d[o, a]) != want_sign[o])
    return axis_of, flip


@dataclass
class MidplanePlan:
    """Everything K0 needs for one series, and what the reference would report for it."""

    out_hw: tuple[int, int]
    spacing: tuple[float, float]  # get_slice_spacing(resampled): (row, col) = (0.3, 0.3)
    slab: np.ndarray  # float32 [z, y, x] -- the source array cut to the two planes around the fixed index
    desc: dict


def midplane_source_planes(size_xyz, spacing_xyz, direction=None, new_spacing=ISOTROPIC_SPACING):
    """Which source planes the middle sagittal plane of the resampled, LPI-oriented volume is interpolated from
    (cropping.py:45-48, 63-79): ``(array axis of [z, y, x], lo, hi, inside, continuous index, resampled sizes, axis_of, flip)``.
    Depends on the header only (size, spacing, direction), so a reader can decode just those planes."""
    size = tuple(int(v) for v in size_xyz)
    ns = [int(round(osz * osp / nsp)) for osz, osp, nsp in zip(size, spacing_xyz, new_spacing)]  # cropping.py:45-48
    axis_of, flip = lpi_axes(direction)
    a_fix = axis_of[0]
    # middle index along Left-Right in the oriented, resampled volume (cropping.py:78-79), back in image-axis order
    mid = ns[a_fix] // 2
    idx = ns[a_fix] - 1 - mid if flip[0] else mid
    u = (idx * float(new_spacing[a_fix])) / float(spacing_xyz[a_fix])
    inside = -0.5 <= u < size[a_fix] - 0.5
    base = math.floor(u)
    lo = min(max(base, 0), size[a_fix] - 1)
    hi = min(max(base + 1, 0), size[a_fix] - 1)
    return 2 - a_fix, lo, hi, inside, u, ns, axis_of, flip


def plan_midplane(volume_zyx: np.ndarray, spacing_xyz, direction=None, new_spacing=ISOTROPIC_SPACING,
                  integer_pixels: bool | None = None) -> MidplanePlan:
    """``integer_pixels``: the image's pixel type is integral, so ITK casts every resampled value back to it (truncation);
    None = decide from the array dtype."""
    v = np.asarray(volume_zyx)
    if integer_pixels is None:
        integer_pixels = bool(np.issubdtype(v.dtype, np.integer))
    if v.ndim != 3:
        raise ValueError("plan_midplane takes a 3-D volume in sitk.GetArrayFromImage order [z, y, x]")
    size = (v.shape[2], v.shape[1], v.shape[0])
    arr_axis, lo, hi, inside, u, ns, axis_of, flip = midplane_source_planes(size, spacing_xyz, direction, new_spacing)
    a_fix, a_col, a_row = axis_of[0], axis_of[1], axis_of[2]
    base = math.floor(u)
    slab = np.ascontiguousarray(np.take(v, [lo, hi] if hi != lo else [lo], axis=arr_axis).astype(np.float32, copy=False))
    dims = [size[0], size[1], size[2]]
    dims[a_fix] = slab.shape[arr_axis]
    desc = dict(nx=dims[0], ny=dims[1], nz=dims[2], ax_row=a_row, ax_col=a_col, ax_fix=a_fix, flip_row=int(flip[2]),
                flip_col=int(flip[1]), out_h=ns[a_row], out_w=ns[a_col], fix_lo=0, fix_hi=slab.shape[arr_axis] - 1,
                fix_inside=int(inside), integer_pixels=int(bool(integer_pixels)), fix_frac=float(u - base), sp_row=float(spacing_xyz[a_row]), sp_col=float(spacing_xyz[a_col]),
                new_sp_row=float(new_spacing[a_row]), new_sp_col=float(new_spacing[a_col]))
    return MidplanePlan((ns[a_row], ns[a_col]), (float(new_spacing[a_row]), float(new_spacing[a_col])), slab, desc)


class PinnedVolumes:
    """A batch of series as the host hands it to the streamed driver when K0 runs on the device: per series only the (at most)
    two source planes around the middle Left-Right index (``plan_midplane``), float32, in ONE pinned buffer, plus the
    ``svb_k0_series`` rows and the layout of the isotropic planes K0 will produce.  512 x 512 sources at 0.7 mm: 2.1 MB per
    series cross PCIe instead of the 5.7 MB of the resampled 1195 x 1195 plane."""

    def __init__(self, volumes, spacings, directions=None, integer_pixels=None, pixel_kinds=None, plans=None, cache_tag: str | None = None):
        if plans is None:
            B = len(volumes)
            directions = directions if directions is not None else [None] * B
            integer_pixels = integer_pixels if integer_pixels is not None else [None] * B
            plans = [plan_midplane(v, s, d, integer_pixels=ip) for v, s, d, ip in zip(volumes, spacings, directions, integer_pixels)]
            if pixel_kinds is None:
                pixel_kinds = [ops.SlicePool.kind_of(np.asarray(v).dtype) for v in volumes]
        B = len(plans)
        self.shapes = [p.out_hw for p in plans]                       # isotropic planes (H', W')
        self.spacings = [p.spacing for p in plans]                    # get_slice_spacing of each
        self.pixel_kinds = list(pixel_kinds) if pixel_kinds is not None else [0] * B
        self.out_offs, self.out_total = ops.SlicePool.layout(self.shapes)
        self.vol_offs, total = [], 0
        for p in plans:
            self.vol_offs.append(total)
            total += (p.slab.size + 3) // 4 * 4
        self.vol_ends = [o + (p.slab.size + 3) // 4 * 4 for o, p in zip(self.vol_offs, plans)]
        if cache_tag is not None:  # a driver that makes one of these per chunk re-uses its pinned staging (cudaHostAlloc is slow)
            self.host = ops.PinnedCache.get(cache_tag, max(total, 4))
        else:
            self.host = torch.empty(max(total, 4), dtype=torch.float32)
            if torch.cuda.is_available():
                self.host = self.host.pin_memory()
        hv = self.host.numpy()
        self.descs = (K0Series * max(B, 1))()
        for i, p in enumerate(plans):
            hv[self.vol_offs[i] : self.vol_offs[i] + p.slab.size] = p.slab.ravel()
            for k, val in p.desc.items():
                setattr(self.descs[i], k, val)
            self.descs[i].vol_off, self.descs[i].out_off = self.vol_offs[i], self.out_offs[i]

    @classmethod
    def from_plans(cls, plans, pixel_kinds=None) -> "PinnedVolumes":
        """From ``plan_midplane`` results (a decode stage that never holds more than one whole volume at a time)."""
        return cls(None, None, plans=list(plans), pixel_kinds=pixel_kinds)

    @property
    def n(self) -> int:
        return len(self.shapes)

    @property
    def nbytes(self) -> int:
        return (self.vol_ends[-1] if self.vol_ends else 0) * 4 + C.sizeof(K0Series) * self.n

    def resident_pool(self, device="cuda:0") -> "ops.SlicePool":
        """K0 over the whole batch into a fresh pool of isotropic planes resident in HBM (one call, no streaming)."""
        dev = ops._require_cuda(device)
        data = torch.empty(max(self.out_total, 4), dtype=torch.float32, device=dev)
        pool = ops.SlicePool(data, torch.tensor(self.out_offs, dtype=torch.int64).to(dev),
                             torch.tensor(self.shapes, dtype=torch.int32).reshape(-1, 2).to(dev), list(self.shapes), h2d_bytes=self.nbytes)
        pool.set_pixel_kinds(self.pixel_kinds)
        if self.n:
            ops.midplane_resample_into(self.host.to(dev, non_blocking=True), self.chunk_descs(0, self.n).to(dev, non_blocking=True), pool)
        return pool

    def chunk_descs(self, i0: int, i1: int) -> torch.Tensor:
        """``svb_k0_series`` rows of series [i0, i1) with offsets relative to the chunk's own buffers, as pinned bytes."""
        rows = (K0Series * (i1 - i0))()
        for k, i in enumerate(range(i0, i1)):
            C.memmove(C.byref(rows[k]), C.byref(self.descs[i]), C.sizeof(K0Series))
            rows[k].vol_off = self.vol_offs[i] - self.vol_offs[i0]
            rows[k].out_off = self.out_offs[i] - self.out_offs[i0]
        t = torch.frombuffer(bytearray(bytes(rows)), dtype=torch.uint8)
        return t.pin_memory() if torch.cuda.is_available() else t


_STAGE_FLIP = [0]


def midplane_resample(volumes, spacings, directions=None, device="cuda:0", integer_pixels=None, pixel_kinds=None):
    """K0 over a batch: list of ``[z, y, x]`` arrays + ``image.GetSpacing()`` (+ ``GetDirection()``) per series ->
    ``(ops.SlicePool of the middle isotropic sagittal slices, [(row_spacing, col_spacing)])``.
    ``pixel_kinds`` (SVB_PIXEL_* per series; default: from the arrays' dtypes) travels with the pool: the resampled slice
    keeps the file's pixel type in the reference, which is what its rotated crop mode warps."""
    dev = ops._require_cuda(device)
    lib = _lib.load()
    B = len(volumes)
    directions = directions if directions is not None else [None] * B
    integer_pixels = integer_pixels if integer_pixels is not None else [None] * B
    plans = [plan_midplane(v, s, d, integer_pixels=ip) for v, s, d, ip in zip(volumes, spacings, directions, integer_pixels)]
    shapes = [p.out_hw for p in plans]
    out_offs, out_total = ops.SlicePool.layout(shapes)
    vol_offs, vol_total = [], 0
    for p in plans:
        vol_offs.append(vol_total)
        vol_total += (p.slab.size + 3) // 4 * 4
    # two alternating pinned staging buffers: the copy of call k may still be in flight when call k+1 fills its slabs
    _STAGE_FLIP[0] ^= 1
    host = ops.PinnedCache.get(f"k0_slabs_{_STAGE_FLIP[0]}", max(vol_total, 4))
    hv = host.numpy()
    descs = (K0Series * max(B, 1))()
    for i, p in enumerate(plans):
        hv[vol_offs[i] : vol_offs[i] + p.slab.size] = p.slab.ravel()
        for k, val in p.desc.items():
            setattr(descs[i], k, val)
        descs[i].vol_off, descs[i].out_off = vol_offs[i], out_offs[i]
    vols_d = host.to(dev, non_blocking=True)
    desc_host = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).pin_memory()
    desc_d = desc_host.to(dev, non_blocking=True)
    data = torch.empty(max(out_total, 4), dtype=torch.float32, device=dev)
    pool = ops.SlicePool(data, torch.tensor(out_offs, dtype=torch.int64).to(dev), torch.tensor(shapes, dtype=torch.int32).reshape(-1, 2).to(dev),
                         list(shapes), h2d_bytes=host.numel() * 4)
    pool.set_pixel_kinds(pixel_kinds if pixel_kinds is not None else [ops.SlicePool.kind_of(np.asarray(v).dtype) for v in volumes])
    if B:
        mh, mw = pool.max_hw
        need = lib.svb_k0_workspace_bytes(B, mh, mw)
        ws = ops._Workspace.get("k0", need, dev)
        _lib.check(lib.svb_k0_midplane_resample(vols_d.data_ptr(), desc_d.data_ptr(), B, mh, mw, data.data_ptr(), ws.data_ptr(), ws.numel(),
                                                _lib.current_stream()))
    return pool, [p.spacing for p in plans]
