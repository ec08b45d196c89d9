"""BASELINE.json configs[4], the localizer-only sweep: uint8 [B, S, S] -> coordinates, convnext_base random init, bf16 operands.

    python scripts/bench_forward.py [batch=512] [size=768] [micro_batch=64]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from spine_vision_b200 import synthetic  # noqa: E402
from spine_vision_b200.cropping import LocalizationModel  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
S = int(sys.argv[2]) if len(sys.argv) > 2 else 768
mb = int(sys.argv[3]) if len(sys.argv) > 3 else 64
dev = "cuda:0"
model = LocalizationModel(synthetic.random_state_dict("base", seed=0), dev, dtype="bf16", micro_batch=mb)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randint(0, 256, (B, S, S), generator=g, device=dev, dtype=torch.uint8)
out = torch.empty((B, 5, 2), dtype=torch.float32, device=dev)
flops, launches = model.engine.cost(B, S, S)
for _ in range(2):
    model.predict_u8(x, out=out)
torch.cuda.synchronize()
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); model.predict_u8(x, out=out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[1]
print(f"config 5: {B} x {S}x{S} forward in {ms:.1f} ms = {B / ms * 1e3:.0f} img/s; GEMM flops {flops / 1e12:.1f} TFLOP -> {flops / ms / 1e9:.0f} TFLOP/s "
      f"over the WHOLE forward (depthwise conv, LayerNorms, head included in the time); {launches} launches; finite: {bool(torch.isfinite(out).all())}")
