import sys, tempfile, time, cProfile, pstats
from pathlib import Path
import torch
sys.path.insert(0, ".")
from spine_vision_b200 import dataset, synthetic
n=48
base = Path(tempfile.mkdtemp())
synthetic.make_spider_tree(base, n_patients=n, seed=1, in_plane=(512, 512), n_slices=15, spacing=(0.7, 0.7, 4.0), missing_t1=())
ckpt = base / "model.pt"; torch.save({"model_state_dict": synthetic.random_state_dict("base", seed=0)}, ckpt)
cfgw = dataset.ClassificationDatasetConfig(base_path=base, localization_model_path=ckpt, crop_size=(128,128), crop_delta_mm=(50,20,30,30), output_name="w")
dataset.create_classification_dataset(cfgw)
cfg = dataset.ClassificationDatasetConfig(base_path=base, localization_model_path=ckpt, crop_size=(128,128), crop_delta_mm=(50,20,30,30), output_name="t")
pr = cProfile.Profile(); t0=time.perf_counter(); pr.enable(); dataset.create_classification_dataset(cfg); pr.disable(); print("total", time.perf_counter()-t0)
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
