"""Turns the ncu captures of tools/profile_lightcone.py into the committed summaries under profiles/:
   gpurun_out/r02_launches_<wl>.csv + r02_profile_meta_<wl>.json -> profiles/r02_traffic_<wl>.json, profiles/r02_launches_<wl>.csv
   gpurun_out/r02_full_<wl>.ncu-rep                              -> profiles/r02_ncu_full_summary_<wl>.md
usage: python tools/summarize_r02.py [workload, default c3] [suffix]
   with a suffix (e.g. _sparse) only gpurun_out/r02_full_<wl><suffix>.ncu-rep -> profiles/r02_ncu_full_summary_<wl><suffix>.md is made
   (a full-set capture of other groups of the same light cone: PGROUPS=0 ... -o gpurun_out/r02_full_c3_sparse)"""
import collections, csv, json, os, subprocess, sys

wlname = sys.argv[1] if len(sys.argv) > 1 else "c3"
suffix = sys.argv[2] if len(sys.argv) > 2 else ""
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
meta = json.load(open(f"{go}/r02_profile_meta_{wlname}.json"))
src = f"{go}/r02_launches_{wlname}.csv"
if os.path.exists(src) and not suffix:
    rows = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
    launch = collections.OrderedDict()
    for r in rows:
        d = launch.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0].replace("void ", "")})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    passes, cur = [], []
    for l in launch.values():
        if "sum_types_kernel" in l["name"]:
            passes.append(cur)
            cur = []
        elif "finalize_map" not in l["name"]:
            cur.append(l)
    assert len(passes) == len(meta["groups"]), (len(passes), meta["groups"])
    summ, share = [], collections.defaultdict(float)
    for g, acc, p in zip(meta["groups"], meta["accepted_pairs"], passes):
        by = collections.defaultdict(float)
        for l in p:
            by[l["name"].split("::")[-1].split("<")[0]] += l["gpu__time_duration.sum"] / 1e6
        dr = sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in p)
        summ.append(dict(group=g, accepted_pairs=acc, launches=len(p), ms=round(sum(by.values()), 3), dram_GB=round(dr / 1e9, 2),
                         dram_over_algorithmic=round(dr / meta["bytes_per_pass"], 3), by_kernel={k: round(v, 3) for k, v in by.items()}))
        for k, v in by.items():
            share[k] += v
    n = len(summ)
    tot_ms, tot_dram = sum(s["ms"] for s in summ), sum(s["dram_GB"] for s in summ)
    json.dump({"dram_bytes_per_launch": tot_dram / n * 1e9, "csrc_sha": meta["csrc_sha"], "workload": wlname,
               "unit": "bytes per pass: dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of one pass, averaged over the groups of the light cone",
               "algorithmic_bytes_per_pass": meta["bytes_per_pass"], "dram_over_algorithmic": tot_dram * 1e9 / n / meta["bytes_per_pass"],
               "avg_ms_per_pass_under_ncu": tot_ms / n, "per_group": summ, "kernel_time_share": {k: round(v / tot_ms, 3) for k, v in share.items()},
               "source": f"profiles/r02_launches_{wlname}.csv: ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                         "dram__bytes_write.sum --clock-control none python tools/profile_lightcone.py (cold-cache, serialised: compare shares, not absolutes)"},
              open(f"{pr}/r02_traffic_{wlname}.json", "w"), indent=1)
    subprocess.run(["cp", src, f"{pr}/r02_launches_{wlname}.csv"])
    print("groups", n, "avg ms/pass", round(tot_ms / n, 3), "avg dram GB/pass", round(tot_dram / n, 2), "x algorithmic", round(tot_dram * 1e9 / n / meta["bytes_per_pass"], 3),
          {k: round(v / tot_ms, 3) for k, v in share.items()})

rep = f"{go}/r02_full_{wlname}{suffix}.ncu-rep"
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_global_red.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    out = [f"# Round 02 — `ncu --set full` of the pass kernels ({wlname}, CUDA sources {meta['csrc_sha']})", "",
           "Command (after the same command exited 0 without ncu): `PGROUPS=<dense group> ncu --set full --clock-control none --import-source on "
           "--profile-from-start off -k regex:\"deposit_pipelined|bin_histogram|bin_scatter|tile_deposit\" -c 4 python tools/profile_lightcone.py`", "",
           ("Captured launches: the kernels of the densest group's pass (one slice)." if not suffix else f"Captured launches: the pass kernels of the groups named by the capture ({suffix.strip('_')}).") + "  Times under ncu are cold-cache and serialised.", "",
           "| metric | " + " | ".join(f"launch {i}" for i in range(len(rows) - 2)) + " |", "|---|" + "---|" * (len(rows) - 2)]

    def fmt(v):
        try:
            return "%.4g" % float(v)
        except ValueError:
            return v.split("(")[0].replace("void ", "")[:44]

    for w in want + stalls:
        if w in hdr:
            i = hdr.index(w)
            label = w.replace("smsp__average_warps_issue_stalled_", "stall: ").replace("_per_issue_active.ratio", " (warps per issue)")
            out.append("| " + label + (" [" + units[i] + "]" if units[i] else "") + " | " + " | ".join(fmt(r[i]) for r in rows[2:]) + " |")
    open(f"{pr}/r02_ncu_full_summary_{wlname}{suffix}.md", "w").write("\n".join(out) + "\n")
    print("\n".join(out[6:14]))
