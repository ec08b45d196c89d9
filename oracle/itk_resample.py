"""CPU restatement of the SimpleITK steps in front of the hot path:
``resample_to_isotropic`` + ``extract_middle_slice`` + ``get_slice_spacing``
(``spine_vision/datasets/classification/cropping.py:37-101``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

**PARITY UNPINNED.**  SimpleITK 2.5.3 (``uv.lock:3756-3757``) is not installed here and is not under
``/root/reference``; the reference has no test or recorded output for these functions.  What follows restates
the documented behaviour of the ITK classes the calls resolve to, and is the definition K0 is tested against:

* ``ResampleImageFilter`` with an identity transform, same origin/direction, ``sitkLinear``, default pixel 0:
  output index ``i`` on axis ``a`` is the physical offset ``i * new_spacing[a]`` along that axis, i.e. the
  continuous input index ``u = (i * new_spacing[a]) / spacing[a]``; a point is inside the buffer when
  ``-0.5 <= u < size - 0.5`` on every axis, otherwise the default pixel is written
  (``ImageBase::IsInsideBuffer`` / ``TransformPhysicalPointToContinuousIndex``).
* ``LinearInterpolateImageFunction`` (3-D): ``base = floor(u)``, ``d = u - base``; neighbours outside the buffer
  are clamped to the nearest valid index; the blend is the nested lerp ``v0 + (v1 - v0) * d`` along x, then y,
  then z, in double; the result is cast to the pixel type: float32 as is, integer pixel types by ``static_cast``
  (``ResampleImageFilter::CastPixelWithBoundsChecking``), i.e. truncation toward zero.
* ``DICOMOrient(image, "LPI")``: axes are permuted / flipped (no resampling) so that index 0 increases towards
  the patient's Left, index 1 towards Posterior, index 2 towards Inferior; each image axis is assigned the
  anatomical axis its direction cosine is largest along (ITK physical space is LPS).
* ``GetArrayFromImage`` returns ``[z, y, x]``; the middle sagittal slice is ``arr[:, :, nx // 2]``: rows run
  superior -> inferior, columns anterior -> posterior, at the middle Left-Right index.

Only the one plane that survives is computed (``resample_middle_sagittal``); ``resample_volume`` +
``orient_lpi`` restate the whole-volume path for small cases so that the shortcut can be checked against it.
"""

from __future__ import annotations

import numpy as np

ISO = (0.3, 0.3, 0.3)  # cropping.py:22


def new_size(size, spacing, new_spacing=ISO):
    """cropping.py:45-48 (Python ``round`` = banker's rounding)."""
    return [int(round(osz * osp / nsp)) for osz, osp, nsp in zip(size, spacing, new_spacing)]


def lpi_axes(direction=None):
    """For the LPI-oriented image return ``(axis_of[o], flip[o])`` for o = 0 (L), 1 (P), 2 (I): which source image
    axis becomes oriented axis o and whether it is reversed.  ``direction``: 3x3, column a = direction cosine of
    image axis a in LPS space (``image.GetDirection()`` reshaped 3x3); None = identity (an LPS image)."""
    d = np.eye(3) if direction is None else np.asarray(direction, dtype=np.float64).reshape(3, 3)
    dom = [int(np.argmax(np.abs(d[:, a]))) for a in range(3)]
    if sorted(dom) != [0, 1, 2]:
        raise ValueError("direction cosines do not resolve to three distinct anatomical axes")
    want_sign = (1.0, 1.0, -1.0)  # L = +x_LPS, P = +y_LPS, I = -z_LPS
    axis_of, flip = [0, 0, 0], [False, False, False]
    for a in range(3):
        o = dom[a]
        axis_of[o] = a
        flip[o] = bool(np.sign(d[o, a]) != want_sign[o])
    return axis_of, flip


def _axis_samples(n_out: int, size: int, spacing: float, new_sp: float):
    """Per output index: (inside, i0, i1, frac) of the continuous source index on one axis."""
    idx = np.arange(n_out, dtype=np.float64)
    u = (idx * new_sp) / spacing
    inside = (u >= -0.5) & (u < size - 0.5)
    base = np.floor(u)
    frac = u - base
    i0 = np.clip(base, 0, size - 1).astype(np.int64)
    i1 = np.clip(base + 1, 0, size - 1).astype(np.int64)
    return inside, i0, i1, frac


def resample_volume(vol_zyx: np.ndarray, spacing_xyz, new_spacing=ISO) -> np.ndarray:
    """Whole-volume ``resample_to_isotropic`` (cropping.py:37-60) for SMALL volumes; array order [z, y, x]."""
    v = np.asarray(vol_zyx)
    size = (v.shape[2], v.shape[1], v.shape[0])
    ns = new_size(size, spacing_xyz, new_spacing)
    ax = [_axis_samples(ns[a], size[a], float(spacing_xyz[a]), float(new_spacing[a])) for a in range(3)]
    (inx, x0, x1, fx), (iny, y0, y1, fy), (inz, z0, z1, fz) = ax
    vd = v.astype(np.float64)

    def g(zi, yi, xi):
        return vd[zi[:, None, None], yi[None, :, None], xi[None, None, :]]

    fx_, fy_, fz_ = fx[None, None, :], fy[None, :, None], fz[:, None, None]
    def plane(zi):
        a = g(zi, y0, x0) + (g(zi, y0, x1) - g(zi, y0, x0)) * fx_
        b = g(zi, y1, x0) + (g(zi, y1, x1) - g(zi, y1, x0)) * fx_
        return a + (b - a) * fy_
    p0, p1 = plane(z0), plane(z1)
    out = p0 + (p1 - p0) * fz_
    inside = inz[:, None, None] & iny[None, :, None] & inx[None, None, :]
    out = np.where(inside, out, 0.0)
    if np.issubdtype(v.dtype, np.integer):
        return np.trunc(out).astype(v.dtype)
    return out.astype(v.dtype)


def orient_lpi(arr_zyx: np.ndarray, direction=None) -> np.ndarray:
    """``sitk.GetArrayFromImage(sitk.DICOMOrient(image, "LPI"))`` on an array in [z, y, x] order."""
    axis_of, flip = lpi_axes(direction)
    a = np.asarray(arr_zyx)
    # array axis of image axis k is (2 - k); oriented array is [I, P, L] = image axes (axis_of[2], axis_of[1], axis_of[0])
    out = np.transpose(a, (2 - axis_of[2], 2 - axis_of[1], 2 - axis_of[0]))
    for arr_ax, o in ((0, 2), (1, 1), (2, 0)):
        if flip[o]:
            out = np.flip(out, axis=arr_ax)
    return np.ascontiguousarray(out)


def extract_middle_slice_full(vol_zyx, spacing_xyz, direction=None) -> np.ndarray:
    """The reference's order of operations on a whole (small) volume: resample, orient, ``arr[:, :, n // 2]``."""
    o = orient_lpi(resample_volume(vol_zyx, spacing_xyz), direction)
    return o[:, :, o.shape[2] // 2]


def resample_middle_sagittal(vol_zyx: np.ndarray, spacing_xyz, direction=None, new_spacing=ISO):
    """Only the plane ``extract_middle_slice(resample_to_isotropic(image))`` keeps, plus ``get_slice_spacing``.
    Returns ``(slice [nI, nP] of float32 -- or of the volume's integer dtype --, (row_spacing, col_spacing))``."""
    v = np.asarray(vol_zyx)
    size = (v.shape[2], v.shape[1], v.shape[0])
    ns = new_size(size, spacing_xyz, new_spacing)
    axis_of, flip = lpi_axes(direction)
    a_fix, a_col, a_row = axis_of[0], axis_of[1], axis_of[2]  # L (fixed), P (columns), I (rows)
    samples = [_axis_samples(ns[a], size[a], float(spacing_xyz[a]), float(new_spacing[a])) for a in range(3)]

    def pick(a, o, n_sel=None):
        inside, i0, i1, fr = samples[a]
        sel = np.arange(ns[a]) if n_sel is None else np.array([n_sel])
        if flip[o]:
            sel = ns[a] - 1 - sel
        return inside[sel], i0[sel], i1[sel], fr[sel]

    rows, cols, fixed = pick(a_row, 2), pick(a_col, 1), pick(a_fix, 0, ns[a_fix] // 2)
    per_axis = {a_row: rows, a_col: cols, a_fix: fixed}
    shape = {a_row: (-1, 1), a_col: (1, -1), a_fix: (1, 1)}
    ins, lo, hi, fr = {}, {}, {}, {}
    for a in range(3):
        i, l, h, f = per_axis[a]
        ins[a], lo[a], hi[a], fr[a] = (t.reshape(shape[a]) for t in (i, l, h, f))
    vd = v.astype(np.float64)

    def g(zi, yi, xi):
        return vd[zi, yi, xi]

    def plane(zi):
        a = g(zi, lo[1], lo[0]) + (g(zi, lo[1], hi[0]) - g(zi, lo[1], lo[0])) * fr[0]
        b = g(zi, hi[1], lo[0]) + (g(zi, hi[1], hi[0]) - g(zi, hi[1], lo[0])) * fr[0]
        return a + (b - a) * fr[1]
    p0, p1 = plane(lo[2]), plane(hi[2])
    out = p0 + (p1 - p0) * fr[2]
    inside = ins[0] & ins[1] & ins[2]
    out = np.where(inside, out, 0.0)
    out = (np.trunc(out).astype(v.dtype) if np.issubdtype(v.dtype, np.integer) else out.astype(np.float32))
    out = np.broadcast_to(out, (ns[a_row], ns[a_col]))
    return np.ascontiguousarray(out), (float(new_spacing[a_row]), float(new_spacing[a_col]))
