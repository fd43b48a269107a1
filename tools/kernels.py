"""Dev helper: run the stand-alone kernel timing loop of bench.py (for ncu captures)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import bench
dev = torch.device('cuda', 0)
print(json.dumps(bench.kernel_rooflines(dev, bench.load_peaks())))
