"""Port of the reference's own Python for the hot path, calling the same
third-party libraries (NumPy, Pillow, OpenCV, torch) the reference calls.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  This is the CPU
baseline ``bench.py`` times (``cpu_baseline.kind == "port"``) and the checker
the GPU parity tests compare against.  Each function cites what it follows.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from oracle.fixedpoint import mm_to_pixels  # noqa: F401  (cropping.py:149-169, pure Python)

IMAGENET_MEAN = [0.485, 0.456, 0.406]  # cropping.py:23
IMAGENET_STD = [0.229, 0.224, 0.225]  # cropping.py:24


def normalize_to_uint8(arr: np.ndarray) -> np.ndarray:
    """io/__init__.py:15-30, verbatim semantics (NumPy does the cast)."""
    arr = arr.astype(np.float32)
    arr_min, arr_max = arr.min(), arr.max()
    if arr_max - arr_min > 0:
        arr = (arr - arr_min) / (arr_max - arr_min) * 255
    with np.errstate(invalid="ignore"):
        return arr.astype(np.uint8)


def resize_with_padding(image: np.ndarray, target_size: tuple[int, int]) -> np.ndarray:
    """cropping.py:104-146 (OpenCV does the resize)."""
    import cv2

    h, w = image.shape[:2]
    target_h, target_w = target_size
    scale = min(target_h / h, target_w / w)
    new_h = int(round(h * scale))
    new_w = int(round(w * scale))
    resized = cv2.resize(image, (new_w, new_h), interpolation=cv2.INTER_LINEAR)
    if resized.dtype != np.uint8:
        resized = normalize_to_uint8(resized)
    canvas = np.zeros((target_h, target_w), dtype=np.uint8)
    y_offset = (target_h - new_h) // 2
    x_offset = (target_w - new_w) // 2
    canvas[y_offset : y_offset + new_h, x_offset : x_offset + new_w] = resized
    return canvas


def crop_region_horizontal(image, center_x, center_y, crop_size, crop_delta) -> np.ndarray:
    """cropping.py:316-354."""
    h, w = image.shape[:2]
    cx = int(center_x * w)
    cy = int(center_y * h)
    left, right, top, bottom = crop_delta
    x1 = max(0, cx - left)
    x2 = min(w, cx + right)
    y1 = max(0, cy - top)
    y2 = min(h, cy + bottom)
    crop = image[y1:y2, x1:x2]
    return resize_with_padding(normalize_to_uint8(crop), crop_size)


def get_rotation_angles(ivd_locations, image_shape, last_disc_angle_boost: float = 1.0):
    """cropping.py:172-255 (same NumPy calls)."""
    from oracle.fixedpoint import rotation_angles

    return rotation_angles(ivd_locations, image_shape, last_disc_angle_boost)


def crop_region_rotated(image, center_x, center_y, crop_size, crop_delta, rotation_angle) -> np.ndarray:
    """cropping.py:258-313 (OpenCV does the rotation)."""
    import cv2

    h, w = image.shape[:2]
    cx = int(center_x * w)
    cy = int(center_y * h)
    left, right, top, bottom = crop_delta
    rotation_matrix = cv2.getRotationMatrix2D((cx, cy), rotation_angle, 1.0)
    rotated = cv2.warpAffine(image, rotation_matrix, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
    x1 = max(0, cx - left)
    x2 = min(w, cx + right)
    y1 = max(0, cy - top)
    y2 = min(h, cy + bottom)
    crop = rotated[y1:y2, x1:x2]
    return resize_with_padding(normalize_to_uint8(crop), crop_size)


@dataclass
class CropContext:
    """cropping.py:357-404: horizontal (the default, config.py:44) and rotated modes."""

    image: np.ndarray
    ivd_locations: dict
    crop_size: tuple
    crop_delta_px: tuple
    mode: str = "horizontal"
    last_disc_angle_boost: float = 1.0
    rotation_angles: dict | None = None

    def __post_init__(self) -> None:
        if self.mode == "rotated" and self.rotation_angles is None:  # cropping.py:369-375
            self.rotation_angles = get_rotation_angles(self.ivd_locations, self.image.shape[:2], self.last_disc_angle_boost)

    def crop(self, level_idx: int):
        if level_idx not in self.ivd_locations:
            return None
        cx, cy = self.ivd_locations[level_idx]
        if self.mode == "rotated" and self.rotation_angles:
            return crop_region_rotated(self.image, cx, cy, self.crop_size, self.crop_delta_px, self.rotation_angles.get(level_idx, 0.0))
        return crop_region_horizontal(self.image, cx, cy, self.crop_size, self.crop_delta_px)


def preprocess_slice(image: np.ndarray, image_size: tuple[int, int]) -> tuple[np.ndarray, torch.Tensor]:
    """cropping.py:463-472 -- normalise, PIL RGB, torchvision Resize (Pillow
    antialiased BILINEAR), ToTensor, Normalize.  Returns the resized uint8
    plane (the K1 parity target) and the ``[3,H,W]`` fp32 model input."""
    from PIL import Image

    u8 = normalize_to_uint8(image)
    pil = Image.fromarray(u8).convert("RGB")
    # transforms.Resize((h, w)) on a PIL image == pil.resize((w, h), BILINEAR)
    pil = pil.resize((image_size[1], image_size[0]), Image.BILINEAR)
    rgb = np.asarray(pil)
    plane = np.ascontiguousarray(rgb[:, :, 0])
    t = torch.from_numpy(rgb.copy()).permute(2, 0, 1).to(torch.float32).div(255)  # ToTensor
    mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32).view(3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=torch.float32).view(3, 1, 1)
    return plane, (t - mean) / std


def predict_ivd_locations(model: torch.nn.Module, image: np.ndarray, device: str, image_size: tuple[int, int]):
    """cropping.py:444-483 -- batch 1, fp32, one ``.cpu()`` per series."""
    _, tensor = preprocess_slice(image, image_size)
    tensor = tensor.unsqueeze(0).to(device)
    with torch.no_grad():
        output_np = model(tensor).cpu().numpy()[0]
    return {i: (float(output_np[i, 0]), float(output_np[i, 1])) for i in range(output_np.shape[0])}


def classifier_input(t2_crop: np.ndarray | None, t1_crop: np.ndarray | None, output_size=(256, 256)):
    """training/datasets/classification.py:40-68 + 247-278 (no augmentation):
    [T2,T1,T2] stack -> Pillow bilinear Resize -> ToTensor -> Normalize.
    Returns ``(uint8 [H,W,3], float32 [3,H,W])``."""
    from PIL import Image

    if t2_crop is not None and t1_crop is not None:
        rgb = np.stack([t2_crop, t1_crop, t2_crop], axis=-1)
    elif t2_crop is not None:
        rgb = np.stack([t2_crop] * 3, axis=-1)
    elif t1_crop is not None:
        rgb = np.stack([t1_crop] * 3, axis=-1)
    else:
        raise ValueError("At least one of t2_crop or t1_crop must be provided")
    pil = Image.fromarray(rgb).resize((output_size[1], output_size[0]), Image.BILINEAR)
    u8 = np.asarray(pil).copy()
    t = torch.from_numpy(u8).permute(2, 0, 1).to(torch.float32).div(255)
    mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32).view(3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=torch.float32).view(3, 1, 1)
    return u8, (t - mean) / std


def localize_and_crop_series(model, slices, crop_size, crop_delta_mm, image_size=(512, 512), spacing=(0.3, 0.3), device="cpu"):
    """The per-series loop of ``process_spider`` (spider.py:90-152) from the
    middle slice onwards: predict -> mm_to_pixels -> CropContext -> 5 crops."""
    coords, crops = [], []
    delta_px = mm_to_pixels(crop_delta_mm, spacing)
    for sl in slices:
        locs = predict_ivd_locations(model, sl, device, image_size)
        ctx = CropContext(sl, locs, crop_size, delta_px)
        crops.append(np.stack([ctx.crop(i) for i in range(5)]))
        coords.append(np.array([locs[i] for i in range(5)], dtype=np.float64))
    return np.stack(coords), np.stack(crops)
