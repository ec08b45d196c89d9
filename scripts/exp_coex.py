"""Experiment: do a half-SM tcgen05 GEMM CTA and a depthwise-conv CTA of another stream overlap on the same SMs?

    python scripts/exp_coex.py [NB] [REPS]

Per stage shape (micro-batch NB): time on CUDA events
  gemm_full   [fc1, fc2] x REPS, full-SM GEMM footprint, one stream
  gemm_half   the same with the co-resident footprint (SVB_GEMM_HALF=1: BN = 128, 113.5 KB, 256 TMEM columns)
  dw          dwconv_ln x REPS, one stream
  both        gemm_half on stream A and dw on stream B at the same time (REPS each)
If the SM really runs both pipes at once, both ~= max(gemm_half, dw) instead of gemm_half + dw.
"""
import os
import sys
import threading
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from spine_vision_b200 import ops  # noqa: E402

dev = "cuda:0"
NB = int(sys.argv[1]) if len(sys.argv) > 1 else 64
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dt = torch.bfloat16
g = torch.Generator().manual_seed(0)


class Clocks(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.s, self.p, self.stop = [], [], False

    def run(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(0)
        while not self.stop:
            self.s.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            self.p.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
            time.sleep(0.02)


def run(name, fa, fb, reps_a, reps_b):
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(2):
        if fa:
            with torch.cuda.stream(sa):
                fa()
        if fb:
            with torch.cuda.stream(sb):
                fb()
    torch.cuda.synchronize()
    ck = Clocks()
    ck.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea, eb = torch.cuda.Event(), torch.cuda.Event()
    cur = torch.cuda.current_stream()
    OUTER = 12  # long enough for NVML to see the steady state
    e0.record()
    sa.wait_event(e0)
    sb.wait_event(e0)
    for _ in range(OUTER):
        if fa:
            with torch.cuda.stream(sa):
                for _ in range(reps_a):
                    fa()
        if fb:
            with torch.cuda.stream(sb):
                for _ in range(reps_b):
                    fb()
    ea.record(sa)
    eb.record(sb)
    cur.wait_event(ea)
    cur.wait_event(eb)
    e1.record()
    torch.cuda.synchronize()
    ck.stop = True
    ck.join()
    ms = e0.elapsed_time(e1) / OUTER
    s = sorted(ck.s[len(ck.s) // 4:]) or [0]
    p = sorted(ck.p[len(ck.p) // 4:]) or [0]
    print(f"  {name:34s} {ms * 1e3:9.1f} us   sm {s[len(s) // 2]} MHz  {p[len(p) // 2]:.0f} W", flush=True)
    return ms


STAGES = [(512, 32), (256, 64), (128, 128)]
if len(sys.argv) > 3:
    STAGES = [st for st in STAGES if str(st[0]) in sys.argv[3].split(",")]
for C, hw in STAGES:
    M = NB * hw * hw
    x = torch.randn(NB, hw, hw, C, generator=g).to(dt).to(dev)
    taps = (torch.randn(49, C, generator=g) * 0.1).to(dev)
    bias = torch.zeros(C, device=dev)
    lnw, lnb = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    a = torch.randn(M, C, generator=g).to(dt).to(dev)
    w1 = (torch.randn(4 * C, C, generator=g) / C ** 0.5).to(dt).to(dev)
    b1 = torch.zeros(4 * C, device=dev)
    hd = torch.empty((M, 4 * C), dtype=dt, device=dev)
    w2 = (torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5).to(dt).to(dev)
    b2 = torch.zeros(C, device=dev)
    gam = torch.full((C,), 1e-3, device=dev)
    xo = a.clone()
    dwo = torch.empty_like(x)

    def mlp():
        ops.gemm(a, w1, b1, 0, out=hd)
        ops.gemm(hd, w2, b2, 1, resid=xo, gamma=gam, out=xo)

    def dw():
        ops.dwconv_ln(x, taps, bias, lnw, lnb, out=dwo)

    print(f"C={C} {hw}x{hw} NB={NB} (per iteration: {REPS} x [fc1+fc2] and/or {REPS} x dwconv_ln)")
    os.environ["SVB_GEMM_HALF"] = "0"
    tf = run("gemm_full", mlp, None, REPS, 0)
    os.environ["SVB_GEMM_HALF"] = "1"
    th = run("gemm_half", mlp, None, REPS, 0)
    td = run("dw", None, dw, 0, REPS)
    tb = run("both (gemm_half || dw)", mlp, dw, REPS, REPS)
    os.environ["SVB_GEMM_HALF"] = "0"
    tb2 = run("both (gemm_full || dw)", mlp, dw, REPS, REPS)
    print(f"  -> sum full+dw {1e3 * (tf + td):.1f}  sum half+dw {1e3 * (th + td):.1f}  overlapped(half) {1e3 * tb:.1f}  overlapped(full) {1e3 * tb2:.1f}  "
          f"gain vs serial full {(tf + td) / tb:.3f}x", flush=True)
    del x, a, hd, xo, dwo
