"""Launch representative layer kernels once each (micro-batch 32 shapes) -- the ncu --set full target.
Also prints CUDA-event timings (meaningful only when NOT under ncu)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from spine_vision_b200 import ops  # noqa: E402

dev = "cuda:0"
NB = int(sys.argv[1]) if len(sys.argv) > 1 else 32
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dt = torch.float16 if (len(sys.argv) > 3 and sys.argv[3] == "fp16") else torch.bfloat16
g = torch.Generator().manual_seed(0)


def timeit(name, fn, flops=None, nbytes=None):
    if REPS == 0:  # profiling mode: exactly one launch per layer
        fn()
        torch.cuda.synchronize()
        return
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / REPS
    extra = ""
    if flops:
        extra += f" {flops / ms / 1e9:8.1f} TFLOP/s"
    if nbytes:
        extra += f" {nbytes / ms / 1e6:8.1f} GB/s"
    print(f"{name:34s} {ms * 1e3:9.1f} us{extra}", flush=True)


for s, (C, hw) in enumerate([(128, 128), (256, 64), (512, 32), (1024, 16)]):
    M = NB * hw * hw
    x = torch.randn(NB, hw, hw, C, generator=g).to(dt).to(dev)
    taps = (torch.randn(49, C, generator=g) * 0.1).to(dev)
    bias = torch.zeros(C, device=dev)
    lnw, lnb = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    raw = torch.empty_like(x)
    # the default block: raw depthwise conv + token statistics, fc1 with the LayerNorm folded in
    timeit(f"dwconv_raw C={C} {hw}x{hw}", lambda: ops.dwconv_raw(x, taps, bias, out=raw), flops=2.0 * M * C * 49, nbytes=M * C * 4)
    _, stat = ops.dwconv_raw(x, taps, bias, out=raw)
    if C % 64 == 0:  # the tensor-core depthwise kernel (what the model runs for fp16) and fc1 fed by its partial statistics
        wtc = ops.dwconv_tc_pack(taps, dt)
        raw_tc = torch.empty_like(x)
        part = torch.empty((M, C // 64, 2), dtype=torch.float32, device=dev)
        timeit(f"dwconv_raw_tc C={C} {hw}x{hw}", lambda: ops.dwconv_raw_tc(x, wtc, bias, out=raw_tc, part=part), flops=2.0 * M * C * 49, nbytes=M * C * 4)
    if REPS != 0:  # the round-1 block (SVB_LN_FOLD=0), for comparison; not launched in profiling mode
        timeit(f"dwconv_ln C={C} {hw}x{hw}", lambda: ops.dwconv_ln(x, taps, bias, lnw, lnb), flops=2.0 * M * C * 49, nbytes=M * C * 4)
    a = raw.view(M, C)
    w1 = (torch.randn(4 * C, C, generator=g) / C ** 0.5).to(dt).to(dev)
    b1 = torch.zeros(4 * C, device=dev)
    s1 = w1.float().sum(1)
    hd = torch.empty((M, 4 * C), dtype=dt, device=dev)
    timeit(f"fc1+LN+gelu M={M} N={4*C} K={C}", lambda: ops.gemm(a, w1, b1, 3, resid=stat, gamma=s1, out=hd), flops=2.0 * M * 4 * C * C, nbytes=M * C * 2 * 5)
    if REPS != 0:
        timeit(f"fc1+gelu  M={M} N={4*C} K={C}", lambda: ops.gemm(a, w1, b1, 0, out=hd), flops=2.0 * M * 4 * C * C, nbytes=M * C * 2 * 5)
    w2 = (torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5).to(dt).to(dev)
    b2 = torch.zeros(C, device=dev)
    gam = torch.ones(C, device=dev)
    xo = a.clone()
    timeit(f"fc2+resid M={M} N={C} K={4*C}", lambda: ops.gemm(hd, w2, b2, 1, resid=xo, gamma=gam, out=xo), flops=2.0 * M * 4 * C * C,
           nbytes=M * C * 2 * 6)
    if C in (128, 256) and REPS != 0:
        xf = a.clone()
        timeit(f"fused MLP M={M} C={C}", lambda: ops.mlp_fused(a, w1, b1, w2, b2, gam, xf), flops=4.0 * M * 4 * C * C, nbytes=M * C * 2 * 3)
    del x, hd, xo, raw
u8 = torch.randint(0, 256, (NB, 512, 512), generator=g, dtype=torch.uint8).to(dev)
wf, bf = (torch.randn(128, 16, generator=g) * 0.01).to(dev), torch.zeros(128, device=dev)
timeit("stem_ln", lambda: ops.stem_ln(u8, wf, bf, torch.ones(128, device=dev), torch.zeros(128, device=dev)), nbytes=NB * (512 * 512 + 16384 * 256))
