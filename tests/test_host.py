"""CPU: host logic, the C-ABI surface, and the multi-process (gloo) shard/gather path."""
import os
import re
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import ROOT
from spine_vision_b200 import _lib, cropping, ops, pipeline, synthetic


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()  # built by __graft_entry__.build(); no compute calls without a GPU
    header = (ROOT / "include" / "spine_b200.h").read_text()
    declared = set(re.findall(r"\b(svb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed from the header"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), f"libspine_b200.so does not export {sym}"
    assert lib.svb_version() == 100
    assert lib.svb_k1_workspace_bytes(2, 1195, 1195, 512, 512) > 0
    assert lib.svb_k3_workspace_bytes(128, 128, 256, 256) > 0


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.SlicePool.from_numpy([np.zeros((8, 8), np.float32)], "cuda:0")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.LocalizationEngine({}, "cpu")


def test_product_never_imports_oracle():
    for f in (ROOT / "spine_vision_b200").rglob("*.py"):
        src = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_mm_to_pixels_mirror():
    assert cropping.mm_to_pixels((35, 5, 20, 20), (0.3, 0.3)) == (117, 17, 67, 67)
    assert cropping.mm_to_pixels((50, 20, 30, 30), (0.3, 0.3)) == (167, 67, 100, 100)
    assert cropping.mm_to_pixels((55, 15, 17.5, 20), (0.3, 0.3)) == (183, 50, 58, 67)
    assert cropping.get_center_fallback_locations()[4] == (0.5, 0.65)


def test_slice_pool_layout_and_synthetic_determinism():
    offs, total = ops.SlicePool.layout([(3, 5), (4, 4), (7, 1)])
    assert offs == [0, 16, 32] and total == 40
    a, b = synthetic.make_iso_slice(5, 200, 180), synthetic.make_iso_slice(5, 200, 180)
    assert a.dtype == np.float32 and np.array_equal(a, b) and a.min() == 0.0 and a.max() > 500
    assert synthetic.iso_size(512, 0.7) == 1195
    xy = synthetic.make_coords(50, seed=1)
    assert xy.shape == (50, 5, 2) and xy.min() >= 0 and xy.max() < 1


def test_shard_series_balanced_and_complete():
    shapes = synthetic.ragged_shapes(200, seed=0)
    sizes = [h * w for h, w in shapes]
    for world in (1, 2, 4, 8):
        shards = pipeline.shard_series(sizes, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(200))
        loads = [sum(sizes[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(sizes)


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from spine_vision_b200 import pipeline
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
n = 11
sizes = [(i * 37) % 13 + 1 for i in range(n)]
mine = pipeline.shard_series(sizes, 2)[rank]
coords = torch.stack([torch.full((5, 2), float(i)) for i in mine]) if mine else torch.zeros((0, 5, 2))
crops = torch.stack([torch.full((5, 4, 4), i, dtype=torch.uint8) for i in mine]) if mine else torch.zeros((0, 5, 4, 4), dtype=torch.uint8)
c, k = pipeline.gather_results(mine, coords, crops, n)
assert c.shape == (n, 5, 2) and k.shape == (n, 5, 4, 4)
for i in range(n):
    assert float(c[i, 0, 0]) == float(i) and int(k[i, 0, 0, 0]) == i, (rank, i)
dist.destroy_process_group()
print("ok", rank)
"""


def test_gather_results_gloo_world2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), str(ROOT), str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_gather_results_single_process():
    c, k = pipeline.gather_results([2, 0], torch.ones((2, 5, 2)), torch.ones((2, 5, 3, 3), dtype=torch.uint8), 3)
    assert c[1].abs().sum() == 0 and c[0].sum() == 10 and k[2].sum() == 45
