"""Layer kernels of the localizer at micro-batch NB shapes, timed alone with CUDA events -- or, with REPS = 0, launched ONCE each in
the order the default forward runs them (the `ncu --set full` target: scripts/prof_r02.sh, scripts/summarise_ncu_layers.py).

    python scripts/prof_layers.py [NB=32] [REPS=5] [fp16|bf16]

Default forward per ConvNeXt block (fp16): dwconv_rawtc_kernel + ln_stat_finalize_kernel, then mlp_fused_kernel<LNF> at C = 128 / 256
or gemm_kernel<LNGELU> + gemm_kernel<RESID> at C = 512 / 1024.  Timing mode also times the alternatives (FP32-pipe depthwise
kernels, un-fused pair at C = 128 / 256, fc1 without the LayerNorm fold)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from spine_vision_b200 import ops  # noqa: E402

dev = "cuda:0"
NB = int(sys.argv[1]) if len(sys.argv) > 1 else 32
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dt = torch.float16 if (len(sys.argv) <= 3 or sys.argv[3] == "fp16") else torch.bfloat16
PROFILE = REPS == 0
g = torch.Generator().manual_seed(0)


def timeit(name, fn, flops=None, nbytes=None, always=False):
    if PROFILE:  # profiling mode: exactly one launch per layer of the default forward
        if always:
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
    print(f"{name:38s} {ms * 1e3:9.1f} us{extra}", flush=True)


for s, (C, hw) in enumerate([(128, 128), (256, 64), (512, 32), (1024, 16)]):
    M = NB * hw * hw
    fused = C in (128, 256)
    x = torch.randn(NB, hw, hw, C, generator=g).to(dt).to(dev)
    taps = (torch.randn(49, C, generator=g) * 0.1).to(dev)
    bias = torch.zeros(C, device=dev)
    lnw, lnb = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    raw = torch.empty_like(x)
    wtc = ops.dwconv_tc_pack(taps, dt)
    part = torch.empty((M, C // 64, 2), dtype=torch.float32, device=dev)
    dw_flops, dw_bytes = 2.0 * M * C * 49, M * C * 4
    # ---- depthwise 7x7 + token statistics
    timeit(f"dwconv_raw_tc+finalize C={C} {hw}x{hw}", lambda: ops.dwconv_raw_tc(x, wtc, bias, out=raw, part=part), dw_flops, dw_bytes, always=dt == torch.float16)
    timeit(f"dwconv_raw (FP32 pipe) C={C} {hw}x{hw}", lambda: ops.dwconv_raw(x, taps, bias, out=raw), dw_flops, dw_bytes, always=dt != torch.float16)
    timeit(f"dwconv_ln (round 1) C={C} {hw}x{hw}", lambda: ops.dwconv_ln(x, taps, bias, lnw, lnb), dw_flops, dw_bytes)
    if PROFILE:  # no extra launches under ncu: any finite statistics will do for the kernels that follow
        stat = torch.ones((M, 2), device=dev)
    elif dt == torch.float16:
        _, stat = ops.dwconv_raw_tc(x, wtc, bias, out=raw, part=part)
    else:
        _, stat = ops.dwconv_raw(x, taps, bias, out=raw)
    # ---- MLP
    a = raw.view(M, C)
    w1 = (torch.randn(4 * C, C, generator=g) / C ** 0.5).to(dt).to(dev)
    b1 = torch.zeros(4 * C, device=dev)
    s1 = w1.float().sum(1)
    w2 = (torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5).to(dt).to(dev)
    b2 = torch.zeros(C, device=dev)
    gam = torch.ones(C, device=dev)
    hd = torch.empty((M, 4 * C), dtype=dt, device=dev)
    xo = a.clone()
    mlp_flops = 2.0 * M * 4 * C * C
    if fused:
        timeit(f"fused MLP + LN fold M={M} C={C}", lambda: ops.mlp_fused_ln(a, w1, b1, s1, stat, w2, b2, gam, xo), 2 * mlp_flops, M * C * 2 * 3, always=True)
        timeit(f"fused MLP (no fold) M={M} C={C}", lambda: ops.mlp_fused(a, w1, b1, w2, b2, gam, xo), 2 * mlp_flops, M * C * 2 * 3)
    timeit(f"fc1+LN+gelu M={M} N={4*C} K={C}", lambda: ops.gemm(a, w1, b1, 3, resid=stat, gamma=s1, out=hd), mlp_flops, M * C * 2 * 5, always=not fused)
    timeit(f"fc1+gelu  M={M} N={4*C} K={C}", lambda: ops.gemm(a, w1, b1, 0, out=hd), mlp_flops, M * C * 2 * 5)
    timeit(f"fc2+resid M={M} N={C} K={4*C}", lambda: ops.gemm(hd, w2, b2, 1, resid=xo, gamma=gam, out=xo), mlp_flops, M * C * 2 * 6, always=not fused)
    del x, hd, xo, raw
u8 = torch.randint(0, 256, (NB, 512, 512), generator=g, dtype=torch.uint8).to(dev)
wf, bf = (torch.randn(128, 16, generator=g) * 0.01).to(dev), torch.zeros(128, device=dev)
timeit("stem_ln", lambda: ops.stem_ln(u8, wf, bf, torch.ones(128, device=dev), torch.zeros(128, device=dev), dt), nbytes=NB * (512 * 512 + 16384 * 256), always=True)
