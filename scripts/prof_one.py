"""One launch of one layer kernel for ncu: python scripts/prof_one.py <kind> <C> <hw> <nb> [fp16|bf16]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from spine_vision_b200 import ops  # noqa: E402

kind, C, hw, NB = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
dt = torch.float16 if (len(sys.argv) > 5 and sys.argv[5] == "fp16") else torch.bfloat16
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
x = torch.randn(NB, hw, hw, C, generator=g).to(dt).to(dev)
taps = (torch.randn(49, C, generator=g) * 0.1).to(dev)
bias = torch.zeros(C, device=dev)
M = NB * hw * hw
if kind == "dwtc":
    wtc = ops.dwconv_tc_pack(taps, dt)
    out = torch.empty_like(x)
    for _ in range(2):
        ops.dwconv_raw_tc(x, wtc, bias, out=out)
elif kind == "dwraw":
    for _ in range(2):
        ops.dwconv_raw(x, taps, bias)
elif kind == "fused":
    a = torch.randn(M, C, generator=g).to(dt).to(dev)
    w1 = (torch.randn(4 * C, C, generator=g) / C ** 0.5).to(dt).to(dev)
    w2 = (torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5).to(dt).to(dev)
    stat = torch.ones((M, 2), device=dev)
    xo = a.clone()
    for _ in range(2):
        ops.mlp_fused_ln(a, w1, torch.zeros(4 * C, device=dev), w1.float().sum(1), stat, w2, torch.zeros(C, device=dev), torch.ones(C, device=dev), xo)
torch.cuda.synchronize()
