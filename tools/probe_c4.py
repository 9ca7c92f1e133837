"""Measurement aid: C4-like hydro pass (gas + DM + stars, per-particle masses with the MAX_M cut, per-type maps)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from slicer_b200 import capi, synth

n = int(os.environ.get("N", str(1 << 26)))  # particles per type
NPIX, BOX = 1024, bench.BOX
groups, raw = bench.c3_planes(BOX, NPIX, bench.FOV_DEG, bench.NGROUPS)
s = capi.Slicer(npix_max=NPIX, max_planes=4, mas=capi.MAS_TSC, particle_capacity=3 * n + 64, mass_capacity=3 * n + 64, per_type_maps=True,
                record_capacity=3 * n, deposit_mode=int(os.environ.get("DMODE", "0")))
s.begin_snapshot(BOX, [0, bench.MASS, 0, 0, 0, 0], True)
rng = np.random.default_rng(1)
for t, seed in ((0, 11), (1, 12), (4, 13)):
    pos = synth.uniform_positions(n, BOX, seed)
    m = None
    if t != 1:
        m = rng.random(n, dtype=np.float32) * 2 + 0.1
        m[:: 97] = 5e3  # ~1 % above MAX_M = 1e3: counted, deposited with mass 0 (densitymaps.cpp:368)
    s.stage(t, pos, m)
s.synchronize()
for g in [int(v) for v in os.environ.get("PGROUPS", "0,2,4,6,8").split(",")]:
    for rep in range(2):
        s.deposit(groups[g])
    st = s.stats()
    acc = sum(int(s.fetch(k, -1, NPIX, want_map=False)[1].sum()) for k in range(4))
    by = (12 * 3 + 4 * 2) * n
    print(f"group {g}: {st.last_deposit_ms:8.3f} ms  accepted {acc:11d} ({acc / (3 * n) * 100:5.2f} %)  {3 * n / st.last_deposit_ms / 1e6:7.1f} Gpart/s  "
          f"{by / st.last_deposit_ms / 1e6 / 6551.7 * 100:5.1f} % roofline (12 B + 4 B mass for two of three types)", flush=True)
