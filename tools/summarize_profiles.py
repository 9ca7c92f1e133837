"""Turns the ncu captures of `bench.py` (gpurun_out/r01_launches.csv, gpurun_out/r01_full.ncu-rep) into the committed summaries
under profiles/.  Measurement aid.  usage: python tools/summarize_profiles.py [round tag, default r01]"""
import collections, csv, json, os, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
rows = list(csv.DictReader(l for l in open(f"{go}/{tag}_launches.csv") if l.startswith('"')))
launch = collections.OrderedDict()
for r in rows:
    d = launch.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
L = list(launch.values())
names = [l["name"] for l in L]
# bench.py launches, after the synthetic-data kernel: 3 warm-up passes (groups 0,1,2), 9 timed passes (groups 0..8), then one extra
# pass per group.  A pass is one direct K1 launch, or 1 slice (sparse) / 4 slices (dense) of (K1 + 5 bin kernels).  Slices of one
# pass and consecutive one-slice passes look alike in the launch list, so the per-group slice counts are taken from the AUTO
# rule of slicer_capi.cu (est_accept: direct < 1.5 % <= one slice < 12 % <= 2^28-particle slices) for this workload.
SLICES = {0: 0, 1: 1, 2: 1, 3: 1, 4: 4, 5: 4, 6: 4, 7: 4, 8: 4}
order = [0, 1, 2] + list(range(9))
passes, i = [], 0
while i < len(L) and "deposit_pipelined" not in names[i]:
    i += 1
for g in order:
    n = 1 if SLICES[g] == 0 else 6 * SLICES[g]
    grp = L[i:i + n]
    assert "deposit_pipelined" in grp[0]["name"] and (n == 1 or "bin_histogram" in grp[1]["name"]), (g, [x["name"] for x in grp[:3]])
    assert i + n >= len(L) or "deposit_pipelined" in names[i + n], (g, names[i + n])
    passes.append(grp)
    i += n
timed = passes[3:12]  # bench: 3 warm-up passes, then the 9 timed ones
summ, share = [], collections.defaultdict(float)
for g, p in enumerate(timed):
    by = collections.defaultdict(float)
    for l in p:
        by[l["name"].split("::")[-1].split("<")[0]] += l["gpu__time_duration.sum"] / 1e6
    dr = sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in p)
    summ.append(dict(group=g, launches=len(p), ms=round(sum(by.values()), 3), dram_GB=round(dr / 1e9, 2), by_kernel={k: round(v, 3) for k, v in by.items()}))
    for k, v in by.items():
        share[k] += v
tot_ms, tot_dram = sum(s["ms"] for s in summ), sum(s["dram_GB"] for s in summ)
json.dump({"dram_bytes_per_launch": tot_dram / 9 * 1e9,
           "unit": "bytes per pass: dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of one pass, averaged over the 9 groups of the bench light cone",
           "algorithmic_bytes_per_pass": 12 * 1024 ** 3, "avg_ms_per_pass_under_ncu": tot_ms / 9, "per_group": summ,
           "kernel_time_share": {k: round(v / tot_ms, 3) for k, v in share.items()},
           "source": f"profiles/{tag}_launches.csv: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 "
                     "python bench.py --steps 9 --warmup 3 --no-cpu --no-e2e (cold-cache, serialised: compare shares, not absolutes)"},
          open(f"{pr}/{tag}_traffic.json", "w"), indent=1)
subprocess.run(["cp", f"{go}/{tag}_launches.csv", f"{pr}/{tag}_launches.csv"])
print("passes", len(passes), "avg ms/pass", round(tot_ms / 9, 3), "avg dram GB/pass", round(tot_dram / 9, 2), {k: round(v / tot_ms, 3) for k, v in share.items()})

raw = subprocess.run(["ncu", "-i", f"{go}/{tag}_full.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
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
out = [f"# Round {tag[1:]} — `ncu --set full` summary of the bench pass kernels", "",
       "Command (after the same command exited 0 without ncu): `ncu --set full --clock-control none --import-source on -k "
       "regex:\"deposit_pipelined|bin_scatter|tile_deposit\" -s 26 -c 8 python bench.py --steps 9 --warmup 3 --no-cpu --no-e2e`", "",
       "Captured launches: consecutive kernels of the timed light cone of the C3 workload (a binned slice = 2^28 particles, a direct pass = 2^30). "
       "Times under ncu are cold-cache and serialised.", "",
       "| metric | " + " | ".join(f"launch {i}" for i in range(len(rows) - 2)) + " |", "|---|" + "---|" * (len(rows) - 2)]


def fmt(v):
    try:
        return "%.4g" % float(v)
    except ValueError:
        return v.split("(")[0].replace("void ", "")[:44]


for w in want:
    if w in hdr:
        i = hdr.index(w)
        out.append("| " + w + (" [" + units[i] + "]" if units[i] else "") + " | " + " | ".join(fmt(r[i]) for r in rows[2:]) + " |")
open(f"{pr}/{tag}_ncu_full_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[6:16]))
