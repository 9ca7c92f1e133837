"""TEST INFRASTRUCTURE (not collected by pytest; run by hand on the GPU box) — Config C1/C2 end to end (BASELINE.json configs[0..1]): synthetic L128N256-shaped snapshots (ng^3 DM particles, box 128 Mpc/h,
snapshots z = 0 .. 0.6), examples/InputParams.ini values (256^2 map, 2 deg, zs = 0.5, seeds -229/-230/-231), TSC.
Runs the reference executable (oracle/_ref/SLICER_ref, 1 rank) and SLICER_b200 on the same files, times both (wall clock,
including file I/O and FITS output), and compares every plane.  The reference executable is the checker here, as in tests/test_gpu_driver.py.
usage: python tests/e2e_c1.py [ng=256] [numfiles=4] [workdir=/tmp/c1]   (E2E_SKIP_REF=1: time SLICER_b200 only, for sizes the reference needs hours for)"""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from slicer_b200 import host, synth
from test_gpu_driver import INI, read_shim_fits

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 256
numfiles = int(sys.argv[2]) if len(sys.argv) > 2 else 4
work = sys.argv[3] if len(sys.argv) > 3 else "/tmp/c1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
box = 128000.0
os.makedirs(work + "/snaps", exist_ok=True)
names = []
t0 = time.time()
for i in range(7):
    synth.write_snapshot(f"{work}/snaps/snap_{i:03d}", {1: synth.hash_positions(ng ** 3, box, 1000 + i)}, [0, 1.0375, 0, 0, 0, 0], 0.1 * i, box,
                         numfiles=numfiles)
    names.append(f"snap_{i:03d}")
open(work + "/snapshot_list.txt", "w").write("\n".join(names))
print(f"wrote 7 snapshots of {ng}^3 particles in {numfiles} sub-files ({time.time() - t0:.1f} s)", flush=True)
res = {}
arms = [("gpu", [host.EXE_PATH, "--quiet"]), ("ref", [os.path.join(ROOT, "oracle", "_ref", "SLICER_ref")])]
if os.environ.get("E2E_SKIP_REF"):
    arms = arms[:1]
for tag, exe in arms:
    out = f"{work}/out_{tag}"
    subprocess.run(["rm", "-rf", out]); os.makedirs(out)
    ini = f"{work}/{tag}.ini"
    open(ini, "w").write(INI.format(npix=256, zs=0.5, fov=2.0, list=work + "/snapshot_list.txt", snapdir=work + "/snaps/", outdir=out + "/test_", pip=0, snopt=0))
    t0 = time.time()
    r = subprocess.run(exe + [ini], cwd=work, capture_output=True, text=True, env=dict(os.environ, SLICER_B200_TIMING="1"))
    res[tag] = time.time() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    print(f"{tag}: {res[tag]:.2f} s wall", [l for l in r.stderr.splitlines() if l.startswith('[timing]')], flush=True)
if os.environ.get("E2E_SKIP_REF"):
    print(f"{ng ** 3 * 7 * 12 / 1e9:.1f} GB of POS payload in {res['gpu']:.2f} s = {ng ** 3 * 7 * 12 / 1e9 / res['gpu']:.2f} GB/s end to end, {len(os.listdir(work + '/out_gpu'))} files written")
    sys.exit(0)
files = sorted(f for f in os.listdir(work + "/out_ref") if f.endswith(".fits"))
assert files == sorted(f for f in os.listdir(work + "/out_gpu") if f.endswith(".fits"))
worst = 0.0
tot_ref = tot_gpu = 0.0
for f in files:
    _, rimg = read_shim_fits(f"{work}/out_ref/{f}")
    _, gimg = host.read_fits(f"{work}/out_gpu/{f}")
    np.testing.assert_allclose(gimg, rimg, rtol=1e-6, atol=1e-9)
    nz = rimg > 1e-6
    if nz.any():
        worst = max(worst, float(np.max(np.abs(gimg[nz] - rimg[nz]) / rimg[nz])))
    tot_ref += float(rimg.sum(dtype=np.float64)); tot_gpu += float(gimg.sum(dtype=np.float64))
npass = len(files) * ng ** 3
print(f"{len(files)} planes identical to rtol 1e-6 (worst pixel rel diff {worst:.2e}); total mass ref {tot_ref:.6f} gpu {tot_gpu:.6f}")
print(f"particle-passes {npass:.3e}: reference {npass / res['ref'] / 1e6:.2f} M/s (1 core), SLICER_b200 {npass / res['gpu'] / 1e6:.1f} M/s, speed-up {res['ref'] / res['gpu']:.1f}x")
