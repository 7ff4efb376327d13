"""GPU box: the camera-motion-estimation sub-benchmark alone (bench_sub.gmc_estimation), one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench_sub  # noqa: E402

hbm = 6577.4
try:
    hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", hbm))
except Exception:
    pass
print(json.dumps(bench_sub.gmc_estimation(torch, torch.device("cuda:0"), hbm, steps=int(sys.argv[1]) if len(sys.argv) > 1 else 40)))
