"""Integer/fp32 restatements of the third-party arithmetic on the hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Every function here is a NumPy restatement of an algorithm the reference
*calls* but does not contain; the CUDA kernels implement the same arithmetic
step for step, so a mismatch against these functions is a kernel bug.

* ``normalize_to_uint8``  follows ``spine_vision/io/__init__.py:15-30``.
* ``pillow_resize_u8``    restates Pillow ``Image.resize(size, BILINEAR)`` on 8-bit
  data (``src/libImaging/Resample.c``: ``precompute_coeffs``,
  ``normalize_coeffs_8bpc``, ``ImagingResampleHorizontal_8bpc`` /
  ``Vertical_8bpc``), the code ``torchvision.transforms.Resize`` runs at
  ``cropping.py:463-472`` and ``training/datasets/classification.py:250``.
* ``cv_resize_u8``        restates OpenCV ``cv2.resize(uint8, INTER_LINEAR)``
  (``modules/imgproc/src/resize.cpp``: ``resizeGeneric_`` with
  ``HResizeLinear`` / ``VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>``),
  the call at ``cropping.py:132``.
* ``letterbox_geometry`` / ``resize_with_padding`` follow ``cropping.py:104-146``.
"""

from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Pillow: 22
CV_COEF_BITS = 11  # OpenCV INTER_RESIZE_COEF_BITS
CV_COEF_SCALE = 1 << CV_COEF_BITS


# --------------------------------------------------------------------------- normalise
def cast_f32_to_u8(a: np.ndarray) -> np.ndarray:
    """``float32 -> uint8`` exactly as NumPy on x86-64 does it: truncate to
    int32 (out-of-range and NaN give 0x80000000), keep the low byte."""
    a = np.asarray(a, dtype=np.float32)
    ok = np.isfinite(a) & (a >= -2147483648.0) & (a < 2147483648.0)
    t = np.where(ok, a, 0.0).astype(np.int64)  # trunc toward zero
    return (t & 0xFF).astype(np.uint8)


def normalize_to_uint8(arr: np.ndarray) -> np.ndarray:
    """io/__init__.py:15-30 -- global min-max to [0,255], truncating cast.

    fp32 throughout, operation order sub -> true divide -> mul; when
    max == min the unscaled values are cast directly (wrap mod 256)."""
    a = np.asarray(arr).astype(np.float32)
    if a.size == 0:
        return a.astype(np.uint8)
    mn = np.float32(a.min())
    mx = np.float32(a.max())
    rng = np.float32(mx - mn)
    if rng > 0:
        a = ((a - mn) / rng) * np.float32(255)
    return cast_f32_to_u8(a)


# --------------------------------------------------------------------------- Pillow
def pillow_coeffs(in_size: int, out_size: int):
    """Pillow ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for BILINEAR.

    Returns ``(bounds[out,2] int32 (xmin, n), kk[out,ksize] int32, ksize)``;
    all intermediate math is IEEE double with C truncating casts."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size  # box is float32
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        n = xmax - xmin
        w = np.zeros(n, dtype=np.float64)
        ww = 0.0
        for x in range(n):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            v = 1.0 - a if a < 1.0 else 0.0
            w[x] = v
            ww += v
        if ww != 0.0:
            w = w / ww
        for x in range(n):
            v = w[x]
            if v < 0:
                kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS))
            else:
                kk[xx, x] = int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx, 0] = xmin
        bounds[xx, 1] = n
    return bounds, kk, ksize


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _pillow_pass(src: np.ndarray, out_size: int) -> np.ndarray:
    """One separable pass along axis 1 (uint8 in, uint8 out)."""
    in_size = src.shape[1]
    bounds, kk, ksize = pillow_coeffs(in_size, out_size)
    s = src.astype(np.int64)
    out = np.empty((src.shape[0], out_size), dtype=np.uint8)
    for xx in range(out_size):
        xmin, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = (s[:, xmin : xmin + n] * kk[xx, :n].astype(np.int64)).sum(axis=1)
        out[:, xx] = _clip8(acc + (1 << (PRECISION_BITS - 1)))
    return out


def pillow_resize_u8(img: np.ndarray, out_hw: tuple[int, int]) -> np.ndarray:
    """``Image.fromarray(img).resize((out_w, out_h), Image.BILINEAR)`` on an
    8-bit single-channel image: horizontal pass, uint8 round, vertical pass.
    Pillow skips a pass whose size is unchanged (box == full image)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    oh, ow = out_hw
    h, w = img.shape
    tmp = img if ow == w else _pillow_pass(img, ow)
    if oh == h:
        return tmp.copy()
    return np.ascontiguousarray(_pillow_pass(np.ascontiguousarray(tmp.T), oh).T)


# --------------------------------------------------------------------------- OpenCV
def cv_axis_table(src: int, dst: int, clamp_fraction: bool):
    """Per-axis source index and 11-bit weights of OpenCV's 8U INTER_LINEAR.

    ``scale = 1/(dst/src)`` in double; ``f = float32((i+0.5)*scale - 0.5)``;
    ``si = floor(f)``; ``f -= si``.  Horizontally the fraction is zeroed when
    the index is clamped (``clamp_fraction``); vertically only the row indices
    are clamped.  Weights are ``saturate_cast<short>(w * 2048)`` = rint."""
    inv = dst / src
    scale = 1.0 / inv
    idx = np.zeros(dst, dtype=np.int32)
    a0 = np.zeros(dst, dtype=np.int32)
    a1 = np.zeros(dst, dtype=np.int32)
    for i in range(dst):
        f = np.float32((i + 0.5) * scale - 0.5)
        si = int(math.floor(float(f)))
        f = np.float32(f - np.float32(si))
        if clamp_fraction:
            if si < 0:
                f, si = np.float32(0.0), 0
            if si >= src - 1:
                f, si = np.float32(0.0), src - 1
        w0 = np.float32(np.float32(1.0) - f) * np.float32(CV_COEF_SCALE)
        w1 = np.float32(f) * np.float32(CV_COEF_SCALE)
        a0[i] = int(np.rint(w0))
        a1[i] = int(np.rint(w1))
        idx[i] = si
    return idx, a0, a1


def cv_resize_u8(img: np.ndarray, out_hw: tuple[int, int]) -> np.ndarray:
    """``cv2.resize(img, (out_w, out_h), interpolation=cv2.INTER_LINEAR)`` for
    a single-channel uint8 image."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    sh, sw = img.shape
    dh, dw = out_hw
    if (dh, dw) == (sh, sw):
        return img.copy()
    sx, ax0, ax1 = cv_axis_table(sw, dw, clamp_fraction=True)
    sy, ay0, ay1 = cv_axis_table(sh, dh, clamp_fraction=False)
    s = img.astype(np.int32)
    sx1 = np.minimum(sx + 1, sw - 1)
    hbuf = s[:, sx] * ax0[None, :] + s[:, sx1] * ax1[None, :]  # [sh, dw], scaled by 2048
    r0 = np.clip(sy, 0, sh - 1)
    r1 = np.clip(sy + 1, 0, sh - 1)
    v = ((ay0[:, None] * (hbuf[r0] >> 4)) >> 16) + ((ay1[:, None] * (hbuf[r1] >> 4)) >> 16)
    return np.clip((v + 2) >> 2, 0, 255).astype(np.uint8)


# --------------------------------------------------------------------------- letterbox
def letterbox_geometry(h: int, w: int, target_hw: tuple[int, int]):
    """cropping.py:118-141 -- ``(new_h, new_w, y_off, x_off)`` with Python
    float math and banker's ``round``."""
    th, tw = target_hw
    scale = min(th / h, tw / w)
    new_h = int(round(h * scale))
    new_w = int(round(w * scale))
    return new_h, new_w, (th - new_h) // 2, (tw - new_w) // 2


def resize_with_padding(image_u8: np.ndarray, target_hw: tuple[int, int]) -> np.ndarray:
    """cropping.py:104-146 for uint8 input, with ``cv_resize_u8`` as the resize."""
    h, w = image_u8.shape[:2]
    th, tw = target_hw
    new_h, new_w, yo, xo = letterbox_geometry(h, w, target_hw)
    canvas = np.zeros((th, tw), dtype=np.uint8)
    canvas[yo : yo + new_h, xo : xo + new_w] = cv_resize_u8(image_u8, (new_h, new_w))
    return canvas


def mm_to_pixels(delta_mm, spacing):
    """cropping.py:149-169."""
    row_spacing, col_spacing = spacing
    left_mm, right_mm, top_mm, bottom_mm = delta_mm
    return (
        int(round(left_mm / col_spacing)),
        int(round(right_mm / col_spacing)),
        int(round(top_mm / row_spacing)),
        int(round(bottom_mm / row_spacing)),
    )


def crop_box(h: int, w: int, x: float, y: float, delta_px):
    """cropping.py:338-348 -- truncating centre, clipped box ``(x1, x2, y1, y2)``."""
    cx = int(x * w)
    cy = int(y * h)
    left, right, top, bottom = delta_px
    return max(0, cx - left), min(w, cx + right), max(0, cy - top), min(h, cy + bottom)


def crop_region_horizontal(image: np.ndarray, x: float, y: float, crop_size, delta_px) -> np.ndarray:
    """cropping.py:316-354 with the restated normalise / resize."""
    h, w = image.shape[:2]
    x1, x2, y1, y2 = crop_box(h, w, x, y, delta_px)
    crop = image[y1:y2, x1:x2]
    return resize_with_padding(normalize_to_uint8(crop), crop_size)
