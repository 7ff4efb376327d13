"""GPU: sweep ring depth / BLOCK_N / tile width of the conv kernel on the YOLOv7-w6 shapes that dominate the step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "yolov7-tracker_b200"))
import torch
from b200track.conv import ConvPlan, pack_conv_weight

def bench(n, h, cin, cout, k, s, reps=10, **kw):
    x = torch.randn((n, h, h, cin), device="cuda").to(torch.bfloat16)
    w = torch.randn((cout, cin, k, k), device="cuda") * 0.05
    b = torch.zeros(cout, device="cuda")
    ho = h // s
    y = torch.zeros((n, ho, ho, (cout + 7) // 8 * 8), dtype=torch.bfloat16, device="cuda")
    try:
        p = ConvPlan(x, pack_conv_weight(w), b, y, n, h, h, cin, 0, cout, k, s, 0, **kw)
    except Exception as e:
        return None, str(e)[:60]
    for _ in range(2): p.run()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): p.run()
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / reps
    return ms, p.flops / ms / 1e9

shapes = [(640, 16, 64, 3, 1), (640, 64, 128, 3, 2), (320, 128, 64, 1, 1), (320, 64, 64, 3, 1), (320, 256, 128, 1, 1), (320, 128, 256, 3, 2),
          (160, 128, 128, 3, 1), (160, 256, 128, 1, 1), (160, 512, 256, 1, 1), (80, 256, 256, 3, 1), (160, 128, 256, 3, 1), (80, 1024, 512, 1, 1)]
if len(sys.argv) > 1 and sys.argv[1] == "nostore":
    for h, ci, co, k, s in shapes:
        a = bench(8, h, ci, co, k, s, act=1)
        b = bench(8, h, ci, co, k, s, act=3)          # bit 1: debug -- skip the global stores
        print("%4d^2 %4d->%-4d k%d s%d  with stores %.3f ms (%.0f TF)   without %.3f ms (%.0f TF)" % (h, ci, co, k, s, a[0], a[1], b[0], b[1]))
    sys.exit(0)
for h, ci, co, k, s in shapes:
    res = []
    for st in (2, 3, 4, 6):
        for bn in ((0,) if co <= 64 else (0, 64, 128) if co <= 128 else (0, 128, 256)):
            if bn > co: continue
            ms, tf = bench(8, h, ci, co, k, s, stages=st, block_n=bn)
            if ms: res.append((ms, st, bn, tf))
    res.sort()
    print("%4d^2 %4d->%-4d k%d s%d  best: %s" % (h, ci, co, k, s, "  ".join("st%d bn%d %.3fms %.0fTF" % (st, bn, ms, tf) for ms, st, bn, tf in res[:4])),
          "| worst %.3fms" % res[-1][0])
