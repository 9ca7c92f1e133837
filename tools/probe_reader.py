"""Measurement aid: GADGET-2 sub-file read rate (page cache -> page-locked buffer) against SLICER_B200_IO_THREADS."""
import ctypes as C, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    from slicer_b200 import host
    t, b = C.c_double(), C.c_longlong()
    lib = host.lib()
    lib.shost_time_read_subfile.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    rc = lib.shost_time_read_subfile(sys.argv[2].encode(), 0, int(sys.argv[3]), 4, C.byref(t), C.byref(b))
    print(f"threads {os.environ.get('SLICER_B200_IO_THREADS', 'default'):>7}  pinned {sys.argv[3]}  rc {rc}  {b.value / 1e9:.2f} GB in {t.value * 1e3:8.1f} ms  = {b.value / t.value / 1e9:6.2f} GB/s")
else:
    import numpy as np
    from slicer_b200 import synth
    n = int(os.environ.get("N", str(1 << 27)))
    base = os.path.join(os.environ.get("DIR", "/tmp"), "probe_reader_snap")
    synth.write_snapshot(base, {1: synth.uniform_positions(n, 256000.0, 5)}, [0, 1.0, 0, 0, 0, 0], 0.0, 256000.0, numfiles=1, with_vel_id=False)
    for pinned in (1,):
        for th in ("1", "2", "4", "8", "16"):
            subprocess.run([sys.executable, __file__, "child", base + ".0", str(pinned)], env=dict(os.environ, SLICER_B200_IO_THREADS=th))
    os.remove(base + ".0")
