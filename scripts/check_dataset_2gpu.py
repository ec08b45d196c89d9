"""Multi-GPU check of the dataset driver (run under torchrun, one process per GPU): every rank takes every world-size-th series,
writes its own PNGs, the records meet on rank 0 (one all_gather_object over NCCL), and the tree equals the single-GPU golden.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_dataset_2gpu.py
"""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
from PIL import Image

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from spine_vision_b200 import dataset, synthetic  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
box = [tempfile.mkdtemp() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
base = Path(box[0])
if rank == 0:
    synthetic.make_spider_tree(base, seed=0)
    synthetic.make_phenikaa_tree(base, seed=0)
dist.barrier()
g = np.load(Path(__file__).resolve().parent.parent / "tests" / "golden" / "host_dataset.npz")
cfg = dataset.ClassificationDatasetConfig(base_path=base, output_name="cls", crop_size=(128, 128), crop_delta_mm=tuple(float(v) for v in g["delta_mm"]),
                                          last_disc_angle_boost=1.5, device=f"cuda:{local}", chunk_series=3)
res = dataset.create_classification_dataset(cfg, rank=rank, world_size=world)
dist.barrier()
if rank == 0:
    names = [str(n) for n in g["horizontal_names"]]
    found = sorted(p.name for p in (cfg.output_path / "images").glob("*.png"))
    assert found == names, (found, names)
    for n, want in zip(names, g["horizontal_images"]):
        assert np.array_equal(np.asarray(Image.open(cfg.output_path / "images" / n)), want), n
    got = sorted((cfg.output_path / "annotations.csv").read_text().strip().split("\n"))
    assert got == sorted(g["horizontal_csv"].item().strip().split("\n")) and res.num_samples == len(names)
    print(f"ok: {world} ranks, {res.num_samples} records, PNG tree and CSV rows equal the single-GPU golden")
dist.destroy_process_group()
