#!/usr/bin/env python
"""bench.py — throughput of the light-cone mass-map hot path on B200 (and of the reference's CPU path beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[2] ("C3") — synthetic 1024^3 DM particles per GPU, box 256 Mpc/h,
2048^2 map, 5 deg field, zs = 1.0, TSC.  One STEP = one pass of the hot path over one snapshot: every particle goes
through box transform -> shell selection -> projection -> FoV cut -> TSC deposit into the 4 lens planes of one
randomisation group (numberOfLensPerSnap = 4, densitymaps.h:23).  Step i uses group i mod 9 of the C3 light cone
(planes 4g .. 4g+3, pile g), so 9 steps are one whole light cone and near/far planes are both weighted in.
The reference needs 4 passes over the snapshot for the same 4 planes (slicer-v2.cpp:138-207); both arms report
`particles / second` = snapshot particles turned into their 4 planes per second.

Multi-GPU (weak scaling): every rank holds its own 1024^3 shard of the snapshot, deposits into private planes and
the planes are summed onto rank 0 with ncclReduce(int64) inside the step (replaces slicer-v2.cpp:214-217).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tsc_deposited_particles_per_sec"
UNIT = "particles/s"

# ---- C3 workload ---------------------------------------------------------------------------------------------
NG = 1024                    # particles per GPU = NG^3
BOX = 256000.0               # kpc/h (POS_U 1.0, gadget2io.h:14)
NPIX = 2048
FOV_DEG = 5.0
ZS = 1.0
MASS = 5.2                   # 1e10 Msun/h, massarr[1]
SEEDS = (-229, -230, -231)   # examples/InputParams.ini
NGROUPS = 9                  # full randomisation groups of the zs=1 light cone (37 planes, Ds = 2329.5 Mpc/h)
LENS_PER_SNAP = 4


def c3_planes(BOX=BOX, NPIX=NPIX, FOV_DEG=FOV_DEG, NGROUPS=NGROUPS):
    """Plane descriptors of the C3 light cone, group by group (plan arithmetic of buildPlanes/randomizeBox).
    The keyword arguments exist for tools/probe_groups.py (other geometries, e.g. C5's 8192^2 maps); bench.py uses the defaults."""
    from slicer_b200 import capi, plan

    nplanes = NGROUPS * LENS_PER_SNAP
    randomize = [1 if i % LENS_PER_SNAP == 0 else 0 for i in range(nplanes)]
    rnd = plan.randomize_box(*SEEDS, randomize)
    fov = float(np.float32(FOV_DEG))  # data.cpp:29 parses fov with stof
    fovrad = fov / 180.0 * math.pi
    thick = BOX / 1e3 / LENS_PER_SNAP
    groups, raw = [], []
    ld2 = 0.0
    for i in range(nplanes):
        ld = ld2
        ld2 = ld + thick  # ldbut += box/1e3/numOfLensPerSnap (densitymaps.cpp:93-100): a running sum
        g = i // LENS_PER_SNAP
        # rcase = floor(ld/box*1e3) (slicer-v2.cpp:137,184-185)
        d = dict(sgn=[int(rnd["sgnX"][i]), int(rnd["sgnY"][i]), int(rnd["sgnZ"][i])], face=int(rnd["face"][i]),
                 centre=[float(rnd["x0"][i]), float(rnd["y0"][i]), float(rnd["z0"][i])], rcase=float(g), ld=ld, ld2=ld2,
                 nrepperp=0, fovradiants=fovrad, boxsize=BOX)
        raw.append(d)
        if i % LENS_PER_SNAP == 0:
            groups.append([])
        groups[-1].append(capi.plane_desc(d["sgn"], d["face"], d["centre"], d["rcase"], d["ld"], d["ld2"], fovrad, NPIX))
    return groups, raw


# ---- clocks --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.15:
                continue
            p = [v.strip() for v in line.split(",")]
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---- the reference on the host cores -------------------------------------------------------------------------
def _write_sample_file(path_base, n, seed, start=0):
    from slicer_b200 import synth

    pos = synth.hash_positions(n, BOX, seed, start=start)
    synth.write_snapshot(path_base, {1: pos}, [0, MASS, 0, 0, 0, 0], 0.0, BOX, numfiles=1, with_vel_id=False)
    return path_base


def _ref_worker(args):
    """One 'MPI rank' of the reference: createDensityMaps (densitymaps.cpp:419) on its own sub-file, for each of the
    4 planes of each group in `groups` (a full snapshot pass per plane, exactly as slicer-v2.cpp:138-207 drives it)."""
    path_base, raw_planes, groups = args
    from oracle.ref_bindings import RefLib

    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)  # the reference prints per-type min/max for rank 0; keep the bench's stdout clean
    ref = RefLib(ngp=False)
    t0 = time.perf_counter()
    checksum = 0.0
    for g in groups:
        for p in raw_planes[g * LENS_PER_SNAP:(g + 1) * LENS_PER_SNAP]:
            m = ref.create_density_maps(path_base, 0, 1, NPIX, p["fovradiants"], p["sgn"], p["face"], p["centre"], p["rcase"],
                                        p["ld"], p["ld2"], p["nrepperp"])
            checksum += float(m.sum(dtype=np.float64))
    return time.perf_counter() - t0, checksum


def run_reference_sample(ncores, n_per_core, groups, raw_planes, tmpdir, seed=1234):
    """-> (particles/s for the 4-planes-per-snapshot job, wall seconds).  Files are written before timing."""
    import multiprocessing as mp

    bases = []
    for r in range(ncores):
        bases.append(_write_sample_file(os.path.join(tmpdir, f"sample_{r}"), n_per_core, seed, start=r * n_per_core))
    jobs = [(b, raw_planes, groups) for b in bases]
    t0 = time.perf_counter()
    if ncores == 1:
        saved = os.dup(1)
        try:
            res = [_ref_worker(jobs[0])]
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    else:
        with mp.get_context("fork").Pool(ncores) as pool:
            res = pool.map(_ref_worker, jobs)
    wall = max(r[0] for r in res)  # slowest rank, as an MPI job would finish; excludes process start-up
    # every group = one snapshot turned into its 4 planes; all ranks work on disjoint sub-files of that snapshot
    particles = ncores * n_per_core * len(groups)
    return particles / wall, wall, res


def reference_arm(args, rank, world):
    if rank != 0:
        return
    from oracle import ref_bindings

    if not ref_bindings.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libslicer_ref.so not built"}))
        return
    _, raw = c3_planes()
    ncores = os.cpu_count() or 1
    # a step = a bounded sample of the workload: each core streams `n_per_core` particles of the snapshot for the 4
    # planes of one group.  ~0.15 us per particle-pass -> 4 planes x 2^22 particles ~ 2.5 s per step per core.
    n_per_core = 1 << 22
    with tempfile.TemporaryDirectory() as td:
        for i in range(args.warmup):
            run_reference_sample(ncores, n_per_core, [i % NGROUPS], raw, td)
        wall = 0.0  # the sample files are written outside the timed part of run_reference_sample
        particles = 0
        for i in range(args.steps):
            _, w, _ = run_reference_sample(ncores, n_per_core, [i % NGROUPS], raw, td)
            wall += w
            particles += ncores * n_per_core
    value = particles / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "config": workload_config(1) | {"sample": f"{ncores} sub-files x {n_per_core} particles per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ncores, "kind": "reference",
                         "sample": f"{ncores} processes (one per host core, as MPI ranks over sub-files), each "
                                   f"createDensityMaps on a {n_per_core}-particle sub-file for the 4 planes of one group per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    return {
        "workload": "C3: synthetic 1024^3 DM particles per GPU (uniform, counter-hash), box 256 Mpc/h, 2048^2 map, 5 deg, zs=1.0, "
                    "TSC; step = 1 snapshot pass -> 4 lens planes of randomisation group (step mod 9)",
        "particles_per_gpu": NG ** 3, "npix": NPIX, "fov_deg": FOV_DEG, "zs": ZS, "planes_per_pass": LENS_PER_SNAP,
        "mas": "TSC", "l2": "inputs (12.9 GB per pass) larger than L2; no flush needed", "parallelism": f"particle shards x{n_gpus}, "
        "ncclReduce(int64) of 4 planes per step" if n_gpus > 1 else "1 GPU",
    }


def bind_to_gpu_numa_node(index):
    """Best effort: run this rank (and first-touch its pinned staging memory) on the CPUs NVML reports as local to its GPU,
    so that 8 concurrent host->device streams do not cross the socket interconnect.  Returns the CPU list or None."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64 + 8)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None


# ---- our arm -------------------------------------------------------------------------------------------------
def ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from slicer_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU path (use --impl reference for the CPU reference)")
    torch.cuda.set_device(local_rank)
    affinity = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    groups, raw = c3_planes()
    npart = NG ** 3
    massarr = [0, MASS, 0, 0, 0, 0]

    # ------------------------------------------------------------------ resident phase: `value` and the roofline
    s = capi.Slicer(npix_max=NPIX, max_planes=LENS_PER_SNAP, mas=capi.MAS_TSC, particle_capacity=npart + 64, device=local_rank,
                    deposit_mode=args.deposit_mode, record_capacity=npart)
    if world > 1:
        uid = [capi.Slicer.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        s.comm_init_rank(uid[0], world, rank)
    s.begin_snapshot(BOX, massarr, False)
    s.stage_synthetic(1, npart, 1000 + rank)
    s.synchronize()

    def step(i):
        s.deposit(groups[i % NGROUPS])
        if world > 1:
            s.reduce(LENS_PER_SNAP, 0)

    for i in range(args.warmup):
        step(i)
    s.synchronize()
    s.reset_stats()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    s.synchronize()
    t0 = time.time()
    s.timer_begin()
    for i in range(args.steps):
        step(i)
    ms = s.timer_end()
    barrier()
    t1 = time.time()
    ms = max_over_ranks(ms)
    st = s.stats()
    clocks = None
    if sampler:
        time.sleep(0.25)
        sampler.stop()
        clocks = sampler.summary(t0, t1)
    value = world * npart * args.steps / (ms * 1e-3)
    kernel_ms = st.deposit_ms_sum / max(1, st.deposit_passes)
    launches = int(st.launches)
    # per-group pass time (one extra pass each, outside the timed region): shows the near/far spread behind the average
    per_group = []
    for g in range(NGROUPS):
        s.deposit(groups[g])
        per_group.append(round(s.stats().last_deposit_ms, 3))
    s.deposit(groups[(args.steps - 1) % NGROUPS])
    accepted = []
    for k in range(LENS_PER_SNAP):
        _, c, _ = s.fetch(k, -1, NPIX, want_map=False)
        accepted.append(int(c[1]))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = 12.0 * npart / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "one pass = pipe::deposit_pipelined_kernel<TSC,AOS,SINGLE,*> (stream + screen + exact projection) "
                "[+ binned::bin_* sort + binned::tile_deposit_kernel when > 3 % of the snapshot is inside the field]", "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_launch": 12 * npart, "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else
                "fallback 6650 GB/s"}
    prof = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass

    # ------------------------------------------------------------------ e2e phase: host buffers through the C ABI
    e2e = None
    if not args.no_e2e:
        nb = 8
        per = npart // nb
        pin = capi.PinnedBuffer(npart * 12)
        capi._check(capi.lib().slicer_download_segment(s.h, 0, ctypes.c_void_p(pin.ptr), None))  # same particles, now on the host
        s.close()
        e = capi.Slicer(npix_max=NPIX, max_planes=LENS_PER_SNAP, mas=capi.MAS_TSC, particle_capacity=per + 64, staging_buffers=2,
                        device=local_rank, deposit_mode=args.deposit_mode)
        if world > 1:
            uid = [capi.Slicer.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            e.comm_init_rank(uid[0], world, rank)
        maps = [np.empty((NPIX, NPIX), np.float32) for _ in range(LENS_PER_SNAP)]
        cnt = np.zeros(6, np.int64)

        def e2e_step(i):
            e.begin_snapshot(BOX, massarr, False)
            for b in range(nb):
                if b:
                    e.next_batch()
                e.stage_ptr(1, pin.ptr + b * per * 12, per)
                e.deposit(groups[i % NGROUPS], accumulate=b > 0)
            if world > 1:
                e.reduce(LENS_PER_SNAP, 0)
            if rank == 0:
                for k in range(LENS_PER_SNAP):
                    capi._check(capi.lib().slicer_fetch(e.h, k, -1, maps[k].ctypes.data, cnt.ctypes.data, None))

        ew = min(args.warmup, 3)
        for i in range(ew):
            e2e_step(i)
        e.synchronize()
        barrier()
        e.synchronize()
        w0 = time.perf_counter()
        e.timer_begin()
        for i in range(args.steps):
            e2e_step(i)
        ems = e.timer_end()
        wall_ms = (time.perf_counter() - w0) * 1e3
        ems = max_over_ranks(max(ems, wall_ms))  # host wall clock includes the synchronous fetches
        e2e = {"value": world * npart * args.steps / (ems * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 12 * npart * world,
               "d2h_bytes_per_step": LENS_PER_SNAP * NPIX * NPIX * 4 + 48, "ms_per_step": ems / args.steps,
               "api": "slicer_begin_snapshot/next_batch/stage_particles(pinned)/deposit[_accumulate]/reduce/fetch, "
                      f"{nb} sub-file batches per snapshot, 2 staging pools"}
        launches_e2e = int(e.stats().launches)
        e.close()
        pin.free()
    else:
        s.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import ref_bindings

        if ref_bindings.available():
            with tempfile.TemporaryDirectory() as td:
                n = 1 << 21
                v, wall, _ = run_reference_sample(1, n, list(range(NGROUPS)), raw, td)
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": f"reference createDensityMaps (oracle/_ref) on a {n}-particle sub-file of the same synthetic snapshot, "
                             f"all 36 planes of the 9 groups ({wall:.1f} s)"}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64->int64",
            "data": "synthetic", "config": workload_config(world), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks,
            "extra": {"per_group_kernel_ms": per_group, "stream_rate_particles_per_s": value, "ref_equiv_particle_passes_per_s": value * LENS_PER_SNAP,
                      "accepted_pairs_last_step": accepted, "kernel_ms_avg": kernel_ms,
                      "gpu_launches_e2e": launches_e2e if e2e else None, "cpu_affinity_rank0": affinity},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=9)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--deposit-mode", type=int, default=0, help="0 auto, 1 direct map atomics, 2 binned shared-memory tiles")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if world != args.gpus and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", os.environ["MASTER_PORT"], os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        reference_arm(args, rank, world)
    else:
        ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
