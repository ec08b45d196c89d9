"""GPU parity: K4 (classifier-input producer) through the C ABI vs the reference's ClassificationDataset transform as
restated in oracle/reference_path.classifier_input (pinned by tests/golden/classifier_input.npz in test_oracle.py):
float32 output bit-exact; 16-bit outputs equal the rounded float32."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN
from gpu_util import dev, requires_gpu
from oracle import reference_path as ref
from spine_vision_b200 import ops, pipeline, synthetic


@requires_gpu
def test_k4_all_byte_values_bit_exact():
    """Every uint8 value in every channel position: [T2,T1,T2], T2 only, T1 only; normalised and not."""
    d = dev()
    t2 = np.arange(256, dtype=np.uint8).reshape(16, 16)
    t1 = t2[::-1, ::-1].copy()
    planes = torch.from_numpy(np.stack([t2, t1])).to(d)
    i2 = torch.tensor([0, 0, -1], dtype=torch.int32, device=d)
    i1 = torch.tensor([1, -1, 1], dtype=torch.int32, device=d)
    got = ops.classifier_input(planes, i2, i1).cpu().numpy()
    for p, (a, b) in enumerate(((t2, t1), (t2, None), (None, t1))):
        _, want = ref.classifier_input(a, b, output_size=(16, 16))
        assert np.array_equal(got[p], want.numpy()), f"sample {p}: max diff {np.abs(got[p] - want.numpy()).max()}"
    raw = ops.classifier_input(planes, i2, i1, normalize=False).cpu().numpy()
    assert np.array_equal(raw[0, 0], (torch.from_numpy(t2).float() / 255).numpy()) and np.array_equal(raw[0, 1], (torch.from_numpy(t1).float() / 255).numpy())
    for dt in (torch.bfloat16, torch.float16):
        g16 = ops.classifier_input(planes, i2, i1, dtype=dt).float().cpu().numpy()
        assert np.array_equal(g16, torch.from_numpy(got).to(dt).float().numpy())


@requires_gpu
def test_k4_from_k3_crops_matches_reference_golden():
    """End of the path: 128x128 crops -> (K3 second output) 256x256 -> K4 == the reference's transform on the PNG pair."""
    d = dev()
    g = np.load(GOLDEN / "classifier_input.npz")
    # golden pair straight from the reference (construct_3channel + transforms.Resize)
    up = torch.from_numpy(np.ascontiguousarray(g["up"].transpose(2, 0, 1))).to(d)  # [3,256,256] u8: T2, T1, T2 resized by Pillow
    planes = torch.stack([up[0], up[1]]).contiguous()
    got = ops.classifier_input(planes, torch.tensor([0], dtype=torch.int32, device=d), torch.tensor([1], dtype=torch.int32, device=d)).cpu()
    _, want = ref.classifier_input(g["t2"], g["t1"])
    assert np.array_equal(got[0].numpy(), want.numpy())
    # and through K3: crops2 of a T2 / T1 slice pair at the same coordinates
    sl = [synthetic.make_iso_slice(40, 700, 640), synthetic.make_iso_slice(41, 700, 640) * 0.5]
    pool = ops.SlicePool.from_numpy(sl, d)
    xy = torch.from_numpy(synthetic.make_coords(1, seed=4, hw=(700, 640))).to(d).repeat(2, 1, 1).contiguous()
    crops, crops2, _ = pipeline.crop_levels(pool, xy, (50, 20, 30, 30), crop_size=(128, 128), second_size=(256, 256))
    flat = crops2.reshape(10, 256, 256)
    i2 = torch.arange(0, 5, dtype=torch.int32, device=d)
    i1 = torch.arange(5, 10, dtype=torch.int32, device=d)
    out = ops.classifier_input(flat, i2, i1).cpu().numpy()
    c = crops.cpu().numpy()
    for lvl in range(5):
        _, want = ref.classifier_input(c[0, lvl], c[1, lvl])
        assert np.array_equal(out[lvl], want.numpy()), f"level {lvl}"


@requires_gpu
def test_k4_argument_checks():
    d = dev()
    planes = torch.zeros((1, 5, 3), dtype=torch.uint8, device=d)  # H*W not a multiple of 4
    idx = torch.zeros(1, dtype=torch.int32, device=d)
    from spine_vision_b200 import _lib

    with pytest.raises(_lib.SvbError):
        ops.classifier_input(planes, idx, idx)
    empty = ops.classifier_input(torch.zeros((1, 4, 4), dtype=torch.uint8, device=d), idx[:0], idx[:0])
    assert tuple(empty.shape) == (0, 3, 4, 4)
