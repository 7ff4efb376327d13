"""profiles/<tag>_conv_layer_table.txt: every conv launch of one bench step (from the ncu launch list in gpurun_out/launches.csv) next
to the layer's two floors -- flops / measured bf16 peak and algorithmic bytes / measured HBM bandwidth.  CPU only: the layer shapes
come from the planner dry run (tests/plan_dryrun.py), the times from the committed launch list.

    python tools/conv_layer_table.py [tag]
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "yolov7-tracker_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main(tag="r01", batch=8, size=1280):
    rows = []
    with open(os.path.join(ROOT, "gpurun_out", "launches.csv")) as f:
        for d in csv.DictReader([l for l in f if l.startswith('"')]):
            rows.append((d["Kernel Name"], d["Grid Size"], float(d["Metric Value"].replace(",", "")) / 1e3))
    starts = [i for i, r in enumerate(rows) if "image_reorg" in r[0]]
    step = rows[starts[1]:starts[2]]                                  # the second profiled step
    convs = [(g, t) for n, g, t in step if "conv_bias_act" in n]
    from plan_dryrun import dry_run_plan
    det, plan = dry_run_plan(1, size)
    names = [n for _, f, n in det.ops if f > 0]
    assert len(names) == len(convs) == len(plan), (len(names), len(convs), len(plan))
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak, hbm = float(peaks.get("bf16_tflops_sustained", 1361.1)) * 1e12, float(peaks.get("hbm_gbs", 6577.4)) * 1e9
    out, tot, floors = [], 0.0, 0.0
    for (g, t), name, c in zip(convs, names, plan):
        ho, wo = c["h"] // c["stride"], c["w"] // c["stride"]
        kin = 16 if c["rowpack"] else c["cin"]
        fl = 2.0 * batch * ho * wo * c["cout"] * c["kh"] * c["kw"] * kin
        by = batch * (c["h"] * c["w"] * c["cin"] * 2 + ho * wo * c["cout"] * (4 if c["out_f32"] else 2)) + c["kh"] * c["kw"] * kin * c["cout"] * 2
        tf, th = fl / peak * 1e6, by / hbm * 1e6
        out.append((name.replace(".conv", "").replace("model.", "L"), c["cin"], c["cout"], c["kh"], c["stride"], c["h"], c["w"], g, t, fl / t / 1e6, tf, th))
        tot += t
        floors += max(tf, th)
    path = os.path.join(ROOT, "profiles", tag + "_conv_layer_table.txt")
    with open(path, "w") as f:
        f.write("# per conv launch of one bench step (batch %d, %dx%d): ncu gpu__time_duration (serialised, cold L2: read it as a profile of\n" % (batch, size, size))
        f.write("# where the time goes, the in-graph total is lower) vs the two floors of the layer: flops / %.0f TFLOP/s and algorithmic\n" % (peak / 1e12))
        f.write("# (input + output + weight) bytes / %.0f GB/s.  bound = the larger floor; x = measured / that floor.\n" % (hbm / 1e9))
        f.write("%-28s %5s %5s %2s %2s %9s %-12s %9s %8s %9s %9s %-6s %5s\n" % ("layer", "cin", "cout", "k", "s", "in HxW", "grid", "us", "TFLOP/s", "flop us", "hbm us", "bound", "x"))
        for name, cin, cout, k, s, h, w, g, t, tfs, tf, th in out:
            f.write("%-28s %5d %5d %2d %2d %4dx%-4d %-12s %9.1f %8.0f %9.1f %9.1f %-6s %5.1f\n"
                    % (name, cin, cout, k, s, h, w, g, t, tfs, tf, th, "tensor" if tf >= th else "hbm", t / max(tf, th)))
        f.write("# total %.1f us over %d launches; sum of floors %.1f us\n" % (tot, len(out), floors))
    print("wrote", path)


if __name__ == "__main__":
    main(*(sys.argv[1:2] or ["r01"]))
