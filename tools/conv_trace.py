"""GPU box, diagnostic build (python yolov7-tracker_b200/build.py --trace; B2T_LIB_PATH=.../libb200track_trace.so):
where does the MMA warp of the conv kernel spend its cycles?  python tools/conv_trace.py L23 halo=1 mt=1 bn=256 ..."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "yolov7-tracker_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
from conv_layer_bench import LAYERS  # noqa: E402
from b200track.conv import ConvPlan, pack_conv_weight  # noqa: E402

name = sys.argv[1]
kw = dict(halo=0, mt=1, bn=0, stages=0, producers=0, splits=1, reps=5, batch=8, halo_bufs=0)
for a in sys.argv[2:]:
    k, v = a.split("=")
    kw[k] = int(v)
cin, cout, k, s, hw = LAYERS[name]
n = kw["batch"]
dt = torch.float16
x = torch.randn((n, hw, hw, cin), device="cuda").to(dt)
w = torch.randn((cout, cin, k, k), device="cuda") * (1.5 / (cin * k * k) ** 0.5)
b = torch.randn(cout, device="cuda") * 0.5
y = torch.zeros((n, hw // s, hw // s, cout), device="cuda", dtype=dt)
plan = ConvPlan(x, pack_conv_weight(w, dtype=dt), b, y, n, hw, hw, cin, 0, cout, k, s, 0, block_n=kw["bn"], stages=kw["stages"], halo=bool(kw["halo"]),
                mt=kw["mt"], splits=kw["splits"], producers=kw["producers"], halo_bufs=kw["halo_bufs"])
plan.run(); plan.run()
buf = (C.c_longlong * (16 * 1024))()
plan.lib.b2t_conv_plan_trace(plan.handle, buf, 1024)          # clear
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(kw["reps"]):
    plan.run()
e1.record(); torch.cuda.synchronize()
ncta = plan.lib.b2t_conv_plan_trace(plan.handle, buf, 1024)
t = np.array(buf[:ncta * 16], dtype=np.float64).reshape(ncta, 16)
us = e0.elapsed_time(e1) * 1e3 / kw["reps"]
print(name, {k_: v for k_, v in kw.items() if v}, plan.info, "%.1f us" % us)
if ncta:
    names = ["mma warp total", "wait tile ring", "wait TMEM free", "wait halo tile", "wait operand stage", "issue MMAs"]
    # the MMA-warp counters are overwritten per launch (last launch), the epilogue counters accumulate over reps
    for i, nm in enumerate(names):
        print("  %-20s mean %9.0f clk  (%5.1f %% of total)   min %9.0f max %9.0f" % (nm, t[:, i].mean(), 100 * t[:, i].mean() / t[:, 0].mean(), t[:, i].min(), t[:, i].max()))
    for g in range(2):
        if t[:, 9 + 2 * g].sum() > 0:
            print("  epilogue group %d: waits for the accumulator %.0f clk per tile (%.0f tiles per CTA per launch)" %
                  (g, t[:, 8 + 2 * g].sum() / t[:, 9 + 2 * g].sum(), t[:, 9 + 2 * g].mean() / (kw["reps"] + 0)))
