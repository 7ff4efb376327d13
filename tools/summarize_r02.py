"""Turns the raw ncu exports of tools/r02_profile.sh (gpurun_out/r02_*.csv) into the summaries committed under profiles/."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from summarize_profiles import full  # noqa: E402

OUT, GO = os.path.join(ROOT, "profiles"), os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"


def launch_table(path):
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next(i for i, r in enumerate(rows) if r[0] == "ID")
    H = rows[hdr]; ix = {h: i for i, h in enumerate(H)}
    by_id = {}
    for r in rows[hdr + 1:]:
        if len(r) < len(H):
            continue
        d = by_id.setdefault(int(r[ix["ID"]]), {"kernel": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]})
        v = float(r[ix["Metric Value"]].replace(",", "")); u = r[ix["Metric Unit"]]
        m = r[ix["Metric Name"]]
        if m == "gpu__time_duration.sum":
            d["us"] = v / 1000 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000)
        elif m.startswith("sm__pipe_tensor"):
            d["tensor_pct"] = v
        elif m.startswith("dram__bytes"):
            d[m] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    return [by_id[k] for k in sorted(by_id)]


def main():
    L = launch_table(os.path.join(GO, "r02_launches.csv"))
    # one step = from one letterbox_reorg launch to the next
    starts = [i for i, d in enumerate(L) if "letterbox_reorg" in d["kernel"]]
    step = L[starts[1]:starts[2]]
    convs = [d for d in step if "conv_bias_act" in d["kernel"]]
    agg = {}
    for d in step:
        k = d["kernel"].split("(")[0][-70:]
        a = agg.setdefault(k, [0, 0.0, 0.0]); a[0] += 1; a[1] += d.get("us", 0.0); a[2] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(OUT, tag + "_pipeline_launch_list.txt"), "w") as f:
        f.write("# ncu --profile-from-start off --clock-control none, metrics gpu__time_duration.sum / sm__pipe_tensor_cycles_active / dram bytes: ONE step of\n"
                "# `python bench.py` (batch 8, 1280x1280, uint8 ingest).  Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n")
        f.write("%-72s %6s %12s %10s %7s %12s\n" % ("kernel", "count", "total_us", "avg_us", "share", "dram_MB"))
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%-72s %6d %12.1f %10.2f %6.1f%% %12.1f\n" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot, a[2] / 1e6))
        f.write("# step total %.1f us over %d launches\n" % (tot, len(step)))
    rd = sum(d.get("dram__bytes_read.sum", 0) for d in convs); wr = sum(d.get("dram__bytes_write.sum", 0) for d in convs)
    with open(os.path.join(OUT, tag + "_conv_launches.txt"), "w") as f:
        f.write("# every conv launch of one bench step, in launch order: grid, block, ncu time (cold, serialised), tensor-pipe active %, DRAM MB read / written\n")
        for i, d in enumerate(convs):
            f.write("%3d %-14s %-12s %9.1f us  tensor %5.1f %%  dram rd %8.1f MB  wr %8.1f MB\n" % (i, d["grid"], d["block"], d.get("us", 0), d.get("tensor_pct", 0),
                    d.get("dram__bytes_read.sum", 0) / 1e6, d.get("dram__bytes_write.sum", 0) / 1e6))
        tw = sum(d.get("us", 0) * d.get("tensor_pct", 0) for d in convs) / max(sum(d.get("us", 0) for d in convs), 1e-9)
        f.write("# %d launches, %.1f us, time-weighted tensor-pipe activity %.1f %%, DRAM %.2f GB read + %.2f GB written\n" % (len(convs), sum(d.get("us", 0) for d in convs), tw, rd / 1e9, wr / 1e9))
    for src, name, filt in (("r02_conv_raw.csv", "_conv_ncu_full.txt", "conv_bias_act"), ("r02_misc_raw.csv", "_track_step_ncu_full.txt", "track_step"),
                            ("r02_misc_raw.csv", "_nms_ingest_ncu_full.txt", ""), ("r02_assoc_raw.csv", "_assoc_ops_ncu_full.txt", "")):
        p = os.path.join(GO, src)
        if os.path.exists(p) and os.path.getsize(p) > 1000:
            full(p, os.path.join(OUT, tag + name.replace("_ncu_full.txt", "X_ncu_full.txt") if False else os.path.join(OUT, tag + name)), filt)
    json.dump({"kernel": "conv_bias_act", "launches": len(convs), "dram_bytes_read": rd, "dram_bytes_write": wr,
               "source": "ncu --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum summed over the %d conv launches of one bench step "
                         "(profiles/%s_conv_launches.txt)" % (len(convs), tag)}, open(os.path.join(OUT, tag + "_conv_traffic.json"), "w"), indent=1)
    for name in ("r02_micro_f64.log", "r02_phase.log"):
        p = os.path.join(GO, name)
        if os.path.exists(p):
            open(os.path.join(OUT, name), "w").write(open(p).read())
    print("conv launches", len(convs), "dram GB", (rd + wr) / 1e9)


if __name__ == "__main__":
    main()
