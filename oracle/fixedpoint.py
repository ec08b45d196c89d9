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


# --------------------------------------------------------------------------- OpenCV warpAffine (rotated crop mode)
CV_AB_BITS = 10  # imgwarp.cpp: AB_BITS = MAX(10, INTER_BITS)
CV_INTER_BITS = 5  # 1/32-pixel source coordinates
CV_INTER_TAB = 1 << CV_INTER_BITS


def rotation_matrix_2d(center, angle_deg: float, scale: float = 1.0) -> np.ndarray:
    """``cv2.getRotationMatrix2D`` (imgwarp.cpp): double arithmetic, angle in degrees."""
    a = angle_deg * (math.pi / 180.0)  # OpenCV: angle *= CV_PI/180 (one multiply by the constant)
    alpha = math.cos(a) * scale
    beta = math.sin(a) * scale
    cx, cy = float(center[0]), float(center[1])
    return np.array([[alpha, beta, (1 - alpha) * cx - beta * cy], [-beta, alpha, beta * cx + (1 - alpha) * cy]], dtype=np.float64)


def invert_affine(m: np.ndarray) -> np.ndarray:
    """The in-place inversion ``cv::warpAffine`` applies to a forward map (no WARP_INVERSE_MAP), same op order."""
    M = [float(v) for v in np.asarray(m, dtype=np.float64).ravel()]
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[4] * D, M[0] * D
    M[0] = A11
    M[1] *= -D
    M[3] *= -D
    M[4] = A22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2], M[5] = b1, b2
    return np.array(M, dtype=np.float64).reshape(2, 3)


def _cv_round(v: np.ndarray) -> np.ndarray:
    """saturate_cast<int>(double) = cvRound: round half to even, saturating."""
    return np.clip(np.rint(v), -2147483648, 2147483647).astype(np.int64)


def warp_affine_coords(inv: np.ndarray, xs: np.ndarray, ys: np.ndarray):
    """Source pixel (sx, sy) and 5-bit fractions (fx, fy) OpenCV uses for destination pixels (xs, ys):
    fixed-point AB_SCALE = 1024 coordinates rounded to 1/32 px (WarpAffineInvoker)."""
    AB = float(1 << CV_AB_BITS)
    rd = (1 << CV_AB_BITS) // CV_INTER_TAB // 2  # round_delta for INTER_LINEAR = 16
    adelta = _cv_round(inv[0, 0] * xs.astype(np.float64) * AB)
    bdelta = _cv_round(inv[1, 0] * xs.astype(np.float64) * AB)
    X0 = _cv_round((inv[0, 1] * ys.astype(np.float64) + inv[0, 2]) * AB) + rd
    Y0 = _cv_round((inv[1, 1] * ys.astype(np.float64) + inv[1, 2]) * AB) + rd
    X = (X0 + adelta) >> (CV_AB_BITS - CV_INTER_BITS)
    Y = (Y0 + bdelta) >> (CV_AB_BITS - CV_INTER_BITS)
    # saturate_cast<short> of the integer part
    sx = np.clip(X >> CV_INTER_BITS, -32768, 32767)
    sy = np.clip(Y >> CV_INTER_BITS, -32768, 32767)
    return sx, sy, X & (CV_INTER_TAB - 1), Y & (CV_INTER_TAB - 1)


def warp_affine_f32(img: np.ndarray, m: np.ndarray, region=None) -> np.ndarray:
    """``cv2.warpAffine(img, m, (w, h), flags=INTER_LINEAR, borderMode=BORDER_REPLICATE)`` for a float32 image
    (the call at ``cropping.py:292-301``).  ``region=(x1, x2, y1, y2)`` evaluates only that window of the output.
    remapBilinear<Cast<float,float>, RemapNoVec, float>: 32x32 table of exact bilinear weights, four taps
    summed left to right in float32, replicated border."""
    a = np.asarray(img, dtype=np.float32)
    h, w = a.shape
    inv = invert_affine(m)
    x1, x2, y1, y2 = region if region is not None else (0, w, 0, h)
    xs = np.arange(x1, x2, dtype=np.int64)[None, :]
    ys = np.arange(y1, y2, dtype=np.int64)[:, None]
    sx, sy, fx, fy = warp_affine_coords(inv, xs, ys)
    sx, sy, fx, fy = np.broadcast_arrays(sx, sy, fx, fy)
    x0, x1c = np.clip(sx, 0, w - 1), np.clip(sx + 1, 0, w - 1)
    y0, y1c = np.clip(sy, 0, h - 1), np.clip(sy + 1, 0, h - 1)
    one = np.float32(1.0)
    s = np.float32(1.0 / CV_INTER_TAB)
    vx1 = fx.astype(np.float32) * s
    vy1 = fy.astype(np.float32) * s
    vx0, vy0 = one - vx1, one - vy1
    w0, w1, w2, w3 = vy0 * vx0, vy0 * vx1, vy1 * vx0, vy1 * vx1
    return ((a[y0, x0] * w0 + a[y0, x1c] * w1) + a[y1c, x0] * w2) + a[y1c, x1c] * w3


def warp_affine(img: np.ndarray, m: np.ndarray, region=None) -> np.ndarray:
    """``cv2.warpAffine(img, m, (w, h), INTER_LINEAR, BORDER_REPLICATE)`` in the image's OWN pixel type -- the reference hands
    OpenCV the middle slice as ``sitk.GetArrayFromImage`` gave it (cropping.py:63-79, 292-301), int16 for SPIDER / MR DICOM:
      float32         remapBilinear<Cast<float,float>>: four taps blended in float32 (``warp_affine_f32``)
      int16 / uint16  the same float32 blend, then Cast<float,short/ushort> = cvRound (half to even), saturating
      uint8           15-bit fixed-point weights (32-fx)(32-fy)*32 ... (exact for bilinear, they sum to 2^15), (acc + 2^14) >> 15
    Checked bit-identical to cv2.warpAffine for all four types in tests/test_oracle.py."""
    a = np.asarray(img)
    if a.dtype == np.uint8:
        h, w = a.shape
        inv = invert_affine(m)
        x1, x2, y1, y2 = region if region is not None else (0, w, 0, h)
        xs = np.arange(x1, x2, dtype=np.int64)[None, :]
        ys = np.arange(y1, y2, dtype=np.int64)[:, None]
        sx, sy, fx, fy = np.broadcast_arrays(*warp_affine_coords(inv, xs, ys))
        x0, x1c = np.clip(sx, 0, w - 1), np.clip(sx + 1, 0, w - 1)
        y0, y1c = np.clip(sy, 0, h - 1), np.clip(sy + 1, 0, h - 1)
        ai = a.astype(np.int64)
        acc = (ai[y0, x0] * ((32 - fy) * (32 - fx) * 32) + ai[y0, x1c] * ((32 - fy) * fx * 32)
               + ai[y1c, x0] * (fy * (32 - fx) * 32) + ai[y1c, x1c] * (fy * fx * 32))
        return np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)
    f = warp_affine_f32(a.astype(np.float32), m, region)
    if a.dtype in (np.dtype(np.int16), np.dtype(np.uint16)):
        info = np.iinfo(a.dtype)
        return np.clip(np.rint(f), info.min, info.max).astype(a.dtype)
    return f


def rotation_angles(ivd_locations: dict, image_shape, last_disc_angle_boost: float = 1.0) -> dict:
    """``get_rotation_angles`` (cropping.py:172-255): tangent of the disc chain by finite differences, the last
    point from a 3-point quadratic fit; same NumPy calls, same Python float arithmetic."""
    if len(ivd_locations) < 2:
        return {level: 0.0 for level in ivd_locations}
    h, w = image_shape
    pts = sorted(((lvl, nx * w, ny * h) for lvl, (nx, ny) in ivd_locations.items()), key=lambda p: p[2])
    n = len(pts)
    out = {}
    for i, (lvl, px, py) in enumerate(pts):
        if i == 0:
            dx, dy = pts[1][1] - px, pts[1][2] - py
            dxdy = dx / dy if dy != 0 else 0.0
        elif i == n - 1:
            if n >= 3:
                last = pts[-3:]
                a, b, _ = np.polyfit(np.array([p[2] for p in last]), np.array([p[1] for p in last]), deg=2)
                dxdy = 2 * a * py + b
            else:
                dx, dy = px - pts[i - 1][1], py - pts[i - 1][2]
                dxdy = dx / dy if dy != 0 else 0.0
        else:
            dx, dy = pts[i + 1][1] - pts[i - 1][1], pts[i + 1][2] - pts[i - 1][2]
            dxdy = dx / dy if dy != 0 else 0.0
        ang = float(np.degrees(np.arctan(dxdy)))
        if i == n - 1:
            ang *= last_disc_angle_boost
        out[lvl] = -ang
    return out


def crop_region_rotated(image: np.ndarray, x: float, y: float, crop_size, delta_px, angle_deg: float) -> np.ndarray:
    """cropping.py:258-313: rotate about the disc centre IN THE SLICE'S PIXEL TYPE (``warp_affine``), cut the box, per-crop
    min-max, letterbox.  Types OpenCV's remap does not take are warped as float32."""
    a = np.asarray(image)
    if a.dtype not in (np.dtype(np.int16), np.dtype(np.uint16), np.dtype(np.uint8)):
        a = a.astype(np.float32)
    h, w = a.shape
    cx, cy = int(x * w), int(y * h)
    x1, x2, y1, y2 = crop_box(h, w, x, y, delta_px)
    rot = warp_affine(a, rotation_matrix_2d((cx, cy), angle_deg), (x1, x2, y1, y2))
    return resize_with_padding(normalize_to_uint8(rot), crop_size)
