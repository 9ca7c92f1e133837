"""Python-side helpers for plane plans (tests / bench).  The product's plan stage is the C++ host code in
slicer_b200/host/; this module only reproduces `randomizeBox` (densitymaps.cpp:166-248) with the C library's own
srand/rand, exactly like the reference binary does (glibc TYPE_3 generator)."""
from __future__ import annotations

import ctypes

import numpy as np

_libc = ctypes.CDLL("libc.so.6")
_libc.srand.argtypes = [ctypes.c_uint]
_libc.rand.restype = ctypes.c_int
RAND_MAX_F = np.float32(2147483647)  # float(RAND_MAX) == 2147483648.0f


def _r() -> np.float32:
    return np.float32(_libc.rand()) / RAND_MAX_F  # rand()/float(RAND_MAX), a float division


def randomize_box(seedcenter: int, seedface: int, seedsign: int, randomize, lens_per_snap: int = 4, fixed_vertex: bool = False):
    n = len(randomize)
    x0, y0, z0 = (np.zeros(n) for _ in range(3))
    face, sx, sy, sz = (np.zeros(n, np.int32) for _ in range(4))
    for i in range(n):
        if randomize[i]:
            _libc.srand(ctypes.c_uint((seedcenter + i // lens_per_snap * 13) & 0xFFFFFFFF))
            if fixed_vertex:  # -DFixedPLCVertex (densitymaps.cpp:191-195)
                x0[i], y0[i], z0[i] = 0.0, 0.0, 0.5
            else:
                x0[i], y0[i], z0[i] = float(_r()), float(_r()), float(_r())
            face[i] = 7
            _libc.srand(ctypes.c_uint((seedface + i // lens_per_snap * 5) & 0xFFFFFFFF))
            while face[i] > 6 or face[i] < 1:
                face[i] = int(1 + float(_r()) * 5.0 + 0.5)
            _libc.srand(ctypes.c_uint((seedsign + i // lens_per_snap * 8) & 0xFFFFFFFF))
            for s in (sx, sy, sz):
                s[i] = 2
                while s[i] > 1 or s[i] < 0:
                    s[i] = int(float(_r()) + 0.5)
                if s[i] == 0:
                    s[i] = -1
        else:
            for a in (x0, y0, z0, face, sx, sy, sz):
                a[i] = a[i - 1]
    return dict(x0=x0, y0=y0, z0=z0, face=face, sgnX=sx, sgnY=sy, sgnZ=sz)
