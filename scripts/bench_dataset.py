"""Files-to-PNG throughput of the drop-in entry point (NOT the bench.py headline, which starts from decoded slices):
`create_classification_dataset` over a synthetic SPIDER tree of config-1-sized volumes (15 x 512 x 512 int16, zlib .mha),
timed as a whole -- label walk, threaded MetaImage decode, K0, K1, localizer, K3, D2H, threaded PNG encode + write, CSV.

    python scripts/bench_dataset.py [n_patients] [chunk_series]
"""
import sys
import tempfile
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from spine_vision_b200 import dataset, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 64
base = Path(tempfile.mkdtemp())
t0 = time.perf_counter()
synthetic.make_spider_tree(base, n_patients=n, seed=1, in_plane=(512, 512), n_slices=15, spacing=(0.7, 0.7, 4.0), missing_t1=())
print(f"tree: {2 * n} volumes written in {time.perf_counter() - t0:.1f} s", flush=True)
sd = synthetic.random_state_dict("base", seed=0)
ckpt = base / "model.pt"
torch.save({"model_state_dict": sd}, ckpt)
for tag, kw in (("warm-up", dict(output_name="w")), ("timed", dict(output_name="t"))):
    cfg = dataset.ClassificationDatasetConfig(base_path=base, localization_model_path=ckpt, crop_size=(128, 128), crop_delta_mm=(50, 20, 30, 30),
                                              chunk_series=chunk, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = dataset.create_classification_dataset(cfg)
    dt = time.perf_counter() - t0
    series = res.num_samples / 5.0
    print(f"{tag}: {res.num_samples} crops from {2 * n} volumes in {dt:.2f} s = {2 * n / dt:.1f} series/s files -> PNG + CSV "
          f"(chunk {chunk}, {torch.get_num_threads()} torch threads; includes model load {'' if tag == 'timed' else '+ first-call setup'})", flush=True)
