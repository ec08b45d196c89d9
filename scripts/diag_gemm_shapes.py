"""Diagnostic (not a bench): cuBLAS bf16 timing of the ConvNeXt-base pointwise GEMM shapes at micro-batch 32,
next to svb_gemm, to know how much head-room each shape has.  CUDA events, L2 flushed between reps."""
import sys, torch
sys.path.insert(0, ".")
from spine_vision_b200 import ops

dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
shapes = [(524288, 512, 128), (524288, 128, 512), (131072, 1024, 256), (131072, 256, 1024),
          (32768, 2048, 512), (32768, 512, 2048), (8192, 4096, 1024), (8192, 1024, 4096)]

def timeit(fn, reps=15):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]

for M, N, K in shapes:
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.05
    bias = torch.zeros(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    t_cublas = timeit(lambda: torch.matmul(a, w.t(), out=out))
    t_svb = timeit(lambda: ops.gemm(a, w, bias, 2, out=out))
    t_gelu = timeit(lambda: ops.gemm(a, w, bias, 0, out=out))
    fl = 2.0 * M * N * K
    print(f"M={M:7d} N={N:5d} K={K:5d}  cublas {t_cublas*1e3:7.1f} us {fl/t_cublas/1e9:7.1f} TF/s | svb bias {t_svb*1e3:7.1f} us {fl/t_svb/1e9:7.1f} TF/s | svb gelu {t_gelu*1e3:7.1f} us {fl/t_gelu/1e9:7.1f} TF/s", flush=True)
