"""Measurement aid: one light cone of a bench workload, group by group, for ncu.  Run under
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --csv --log-file gpurun_out/r02_launches_<workload>.csv python tools/profile_lightcone.py
(after the same command exited 0 without ncu).  A warm-up light cone runs unprofiled; then every group's pass is followed by a
slicer_fetch_fixed, whose `sum_types_kernel` marks the end of the group in the launch list.  Writes gpurun_out/r02_profile_meta.json
(hash of the CUDA sources, accepted pairs per group) for tools/summarize_r02.py.
Env: WORKLOAD (c3), PGROUPS (comma list, default all)."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

name = os.environ.get("WORKLOAD", "c3")
wl = bench.Workload(name, "strong", 0, 1, 0)
W = wl.W
groups = [int(v) for v in os.environ.get("PGROUPS", ",".join(str(i) for i in range(W["ngroups"]))).split(",")]
for g in groups:  # warm-up (also allocates the record buffers)
    wl.s.deposit_slots(wl.groups[g], 0)
wl.s.synchronize()
cudart = ctypes.CDLL("libcudart.so")
cudart.cudaProfilerStart()
acc = []
for g in groups:
    wl.s.deposit_slots(wl.groups[g], 0)
    wl.s.synchronize()
    wl.s.fetch_fixed(0, -1, W["npix"])  # marker: sum_types_kernel
    acc.append(int(sum(wl.s.fetch(k, -1, W["npix"], want_map=False)[1].sum() for k in range(bench.LENS_PER_SNAP))))
cudart.cudaProfilerStop()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"workload": name, "groups": groups, "accepted_pairs": acc, "particles": wl.n, "bytes_per_pass": wl.bytes_per_pass, "csrc_sha": bench.csrc_sha()},
          open(os.path.join(ROOT, "gpurun_out", f"r02_profile_meta_{name}{'_partial' if 'PGROUPS' in os.environ else ''}.json"), "w"))
print("profiled groups", groups, "accepted", acc)
