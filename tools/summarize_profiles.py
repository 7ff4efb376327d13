"""Turns the raw ncu outputs in gpurun_out/ into the small text summaries committed under profiles/."""
import csv, io, os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles"); GO = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next(i for i, r in enumerate(rows) if r[0] == "ID")
    H = rows[hdr]; ix = {h: i for i, h in enumerate(H)}
    agg = {}
    for r in rows[hdr + 1:]:
        if len(r) < len(H) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", "")); u = r[ix["Metric Unit"]]
        us = v / 1000 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000)
        k = r[ix["Kernel Name"]][:90]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("# ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none  (timed region of python bench.py, short run)\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write("%-92s %6s %12s %10s %7s\n" % ("kernel", "count", "total_us", "avg_us", "share"))
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%-92s %6d %12.1f %10.2f %6.1f%%\n" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))

def full(rep, out, kernel_filter):
    # rep: an .ncu-rep, or the `ncu -i ... --page raw --csv` export of one made on the GPU box (reports with source pages
    # outgrow gpurun's 64 MiB copy-back limit)
    txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    H = rows[0]; units = rows[1]
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread", "lts__t_bytes.sum",
            "sm__inst_executed_pipe_tensor.sum", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max",
            "smsp__average_warp_latency_issue_stalled_barrier.pct", "l1tex__data_bank_conflicts_pipe_lsu.sum",
            "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_no_instructions",
            "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_sample_buffer_full"]
    ix = {h: i for i, h in enumerate(H)}

    def to_bytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    # DRAM traffic of all captured launches of the kernel (one step of the bench = every conv layer once)
    sel = [r for r in rows[2:] if kernel_filter in r[ix["Kernel Name"]]]
    if sel and "dram__bytes_read.sum" in ix:
        rd = sum(to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) for r in sel)
        wr = sum(to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]]) for r in sel)
        json.dump({"kernel": kernel_filter, "launches": len(sel), "dram_bytes_read": rd, "dram_bytes_write": wr,
                   "source": "ncu --set full --clock-control none, %s (dram__bytes_read.sum + dram__bytes_write.sum summed over the launches)" % os.path.basename(out)},
                  open(out.replace("_ncu_full.txt", "_traffic.json"), "w"), indent=1)
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on  (steady-state launches of the timed region)\n")
        for r in rows[2:]:
            if kernel_filter not in r[ix["Kernel Name"]]:
                continue
            for w in want:
                if w in ix:
                    f.write("%-72s %s %s\n" % (w, r[ix[w]], units[ix[w]]))
            f.write("\n")

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if os.path.exists(os.path.join(GO, "launches.csv")):
        launches(os.path.join(GO, "launches.csv"), os.path.join(OUT, tag + "_pipeline_launch_list.txt"))
    if os.path.exists(os.path.join(GO, "launches_tracker.csv")):
        launches(os.path.join(GO, "launches_tracker.csv"), os.path.join(OUT, tag + "_tracker_launch_list.txt"))
    rep = os.path.join(GO, "prof_track_step.ncu-rep")
    if os.path.exists(rep):
        full(rep, os.path.join(OUT, tag + "_track_step_ncu_full.txt"), "track_step")
    for rep in (os.path.join(GO, "prof_conv_raw.csv"), os.path.join(GO, "prof_conv.ncu-rep")):
        if os.path.exists(rep) and os.path.getsize(rep) > 1000:
            full(rep, os.path.join(OUT, tag + "_conv_ncu_full.txt"), "conv_bias_act")
            break
    for name in ("bench.json", "bench_ref.json", "bench_tracker.json", "phase.log", "conv_sweep.log", "nostore.log"):
        p = os.path.join(GO, name)
        if os.path.exists(p):
            open(os.path.join(OUT, tag + "_" + name.replace(".json", "_line.json")), "w").write(open(p).read())
