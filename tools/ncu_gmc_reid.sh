#!/bin/bash
# GPU box: ncu --set full of the camera-motion estimator's and the ReID glue kernels (one call each), raw page exported to CSV
set -u
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
cat > /tmp/gmc_one.py <<'PY'
import os, sys
import numpy as np, torch
ROOT = os.environ.get("GRAFT_REPO_ROOT", ".")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "yolov7-tracker_b200"))
from b200track.gmc import GmcEstimator
from b200track.reid import ReidExtractor
from b200track.synth import make_stream, pack_frames, textured_frame
from oracle import reid as R
S, size = 8, 1280
base = [textured_frame(7000 + s, size, size, n_rect=1200) for s in range(S)]
dets_np, cnt_np = pack_frames(make_stream(7100, S, 300, img=size)[0], 320)
dets, cnt = torch.from_numpy(dets_np).cuda(), torch.from_numpy(cnt_np).cuda()
est = GmcEstimator(S, size, size, 2, max_kp=32768)
for k in range(3):
    fr = torch.from_numpy(np.stack([np.roll(b, (3 * k, -2 * k), (0, 1)) for b in base])).cuda()
    if k == 2:
        torch.cuda.profiler.start()
    est.estimate(fr, dets, cnt, det_thresh=0.2)
    torch.cuda.synchronize()
ext = ReidExtractor(R.seeded_state_dict(3), bn_mode="batch")
rng = np.random.default_rng(9)
x1 = rng.uniform(0, 1100, 256); y1 = rng.uniform(0, 1000, 256)
tlbr = np.stack([x1, y1, x1 + rng.uniform(20, 80, 256), y1 + rng.uniform(40, 160, 256)], 1)
ext.features_from_frame(fr[0], tlbr)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
PY
timeout 400 ncu --profile-from-start off --set full --clock-control none -k regex:"gray_kernel|fast_score|blur_kernel|nms_flag|scan_rows|compact_kernel|describe|match_kernel|filter_kernel|ransac|fit_kernel|reid_crop|maxpool|add_relu|bn_stats|bn_apply|avgpool" -c 24 -f -o /tmp/prof_gmc python /tmp/gmc_one.py > gpurun_out/ncu_gmc.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/prof_gmc.ncu-rep --page raw --csv > gpurun_out/prof_gmc_raw.csv 2> gpurun_out/prof_gmc_raw.err; echo "export rc=$?"
ls -la gpurun_out/prof_gmc_raw.csv
