"""Measurement aid: per-randomisation-group kernel time of the bench workload (not part of the product)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from slicer_b200 import capi

ng = int(os.environ.get("NG", "1024"))
mas = capi.MAS_NGP if os.environ.get("MAS") == "NGP" else capi.MAS_TSC
# other geometries: BOX [kpc/h], NPIX, FOV [deg], NGRP (e.g. C5: BOX=1e6 NPIX=8192 FOV=5 NGRP=6)
BOX, NPIX = float(os.environ.get("BOX", bench.BOX)), int(os.environ.get("NPIX", bench.NPIX))
NGRP = int(os.environ.get("NGRP", bench.NGROUPS))
groups, raw = bench.c3_planes(BOX, NPIX, float(os.environ.get("FOV", bench.FOV_DEG)), NGRP)
n = ng ** 3
s = capi.Slicer(npix_max=NPIX, max_planes=4, mas=mas, particle_capacity=n + 64, record_capacity=int(os.environ.get('RECCAP', str(n))),
                deposit_mode=int(os.environ.get('DMODE', '0')))
s.begin_snapshot(BOX, [0, bench.MASS, 0, 0, 0, 0], False)
s.stage_synthetic(1, n, 1000)
s.synchronize()
for g in [int(v) for v in os.environ.get('PGROUPS', ','.join(str(i) for i in range(NGRP))).split(',')]:
    for rep in range(2):
        s.deposit(groups[g])
    st = s.stats()
    acc = sum(int(s.fetch(k, -1, NPIX, want_map=False)[1][1]) for k in range(4))
    print(f"group {g}: {st.last_deposit_ms:8.3f} ms  accepted {acc:11d} ({acc / n * 100:5.2f} %)  "
          f"{n / st.last_deposit_ms / 1e6:7.1f} Gpart/s  {12 * n / st.last_deposit_ms / 1e6 / 6551.7 * 100:5.1f} % roofline  "
          f"{acc / st.last_deposit_ms / 1e6:6.2f} Grec/s  flagged {st.flagged_pairs}+{st.flagged_void}", flush=True)
