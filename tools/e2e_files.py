"""Measurement aid: the whole product, files to FITS, on configuration-shaped synthetic snapshots — wall clock of the C++ driver
(`SLICER_b200 --gpus ...`: ini -> plan -> GADGET-2 sub-files through the threaded reader -> H2D -> passes -> reduce -> FITS).
usage: python tools/e2e_files.py [--ng 512] [--numfiles 16] [--gpus 0] [--shape c3|c1] [--work /tmp/e2e] [--ref]
Prints one JSON line; with --ref the reference executable (oracle/_ref/SLICER_ref, 1 rank) runs on the same files first and
every plane is compared (only sensible for small --ng).  The snapshot files are written before timing (page cache warm)."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from slicer_b200 import host, synth  # noqa: E402
from test_gpu_driver import INI, read_shim_fits  # noqa: E402

SHAPES = {"c1": dict(box=128000.0, npix=256, fov=2.0, zs=0.5, nsnap=7), "c3": dict(box=256000.0, npix=2048, fov=5.0, zs=1.0, nsnap=12)}
ap = argparse.ArgumentParser()
ap.add_argument("--ng", type=int, default=512)
ap.add_argument("--numfiles", type=int, default=16)
ap.add_argument("--gpus", default="0")
ap.add_argument("--shape", default="c3", choices=sorted(SHAPES))
ap.add_argument("--work", default="/tmp/e2e")
ap.add_argument("--ref", action="store_true")
a = ap.parse_args()
S = SHAPES[a.shape]
os.makedirs(a.work + "/snaps", exist_ok=True)
names = []
t0 = time.time()
for i in range(S["nsnap"]):
    base = f"{a.work}/snaps/snap_{i:03d}"
    if not os.path.exists(base + ".0"):
        synth.write_snapshot(base, {1: synth.hash_positions(a.ng ** 3, S["box"], 1000 + i)}, [0, 1.0375, 0, 0, 0, 0], 0.1 * i, S["box"],
                             numfiles=a.numfiles, with_vel_id=False)
    names.append(f"snap_{i:03d}")
open(a.work + "/snapshot_list.txt", "w").write("\n".join(names))
t_write = time.time() - t0
res = {}
arms = ([("ref", [os.path.join(ROOT, "oracle", "_ref", "SLICER_ref")])] if a.ref else []) + [("gpu", [host.EXE_PATH, "--quiet", "--gpus", a.gpus])]
timing = ""
for tag, exe in arms:
    out = f"{a.work}/out_{tag}"
    subprocess.run(["rm", "-rf", out])
    os.makedirs(out)
    ini = f"{a.work}/{tag}.ini"
    open(ini, "w").write(INI.format(npix=S["npix"], zs=S["zs"], fov=S["fov"], list=a.work + "/snapshot_list.txt", snapdir=a.work + "/snaps/",
                                    outdir=out + "/test_", pip=0, snopt=0))
    t0 = time.time()
    r = subprocess.run(exe + [ini], cwd=a.work, capture_output=True, text=True, env=dict(os.environ, SLICER_B200_TIMING="1"))
    res[tag] = time.time() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    if tag == "gpu":
        timing = " ".join(l for l in r.stderr.splitlines() if l.startswith("[timing] total"))
nplanes = len([f for f in os.listdir(a.work + "/out_gpu") if f.endswith(".fits")])
line = dict(shape=a.shape, ng=a.ng, snapshots=S["nsnap"], numfiles=a.numfiles, gpus=a.gpus, planes=nplanes, wall_s=round(res["gpu"], 3),
            pos_payload_GB=round(a.ng ** 3 * 12 * S["nsnap"] / 1e9, 2), files_write_s=round(t_write, 1), driver_timing=timing)
if a.ref:
    worst = 0.0
    for f in sorted(f for f in os.listdir(a.work + "/out_ref") if f.endswith(".fits")):
        _, rimg = read_shim_fits(f"{a.work}/out_ref/{f}")
        _, gimg = host.read_fits(f"{a.work}/out_gpu/{f}")
        np.testing.assert_allclose(gimg, rimg, rtol=1e-6, atol=1e-9)
        nz = rimg > 1e-6
        if nz.any():
            worst = max(worst, float(np.max(np.abs(gimg[nz] - rimg[nz]) / rimg[nz])))
    line.update(ref_wall_s=round(res["ref"], 2), speedup=round(res["ref"] / res["gpu"], 1), worst_pixel_rel_diff=worst)
print(json.dumps(line))
