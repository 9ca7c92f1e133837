"""TEST INFRASTRUCTURE — ctypes bindings to oracle/_build/libslicer_oracle.so (oracle/slicer_oracle.c),
plus numpy-level compositions that mirror `createDensityMaps` (densitymaps.cpp:419-524) on in-memory arrays.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libslicer_oracle.so")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")

MAX_M = 1e3  # densitymaps.h:21
LENS_PER_SNAP = 4  # densitymaps.h:23


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "slicer_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return LIB_PATH


_libc = C.CDLL("libc.so.6")
_libc.srand.argtypes = [C.c_uint]
_libc.rand.restype = C.c_int


def libc_srand(seed: int) -> None:
    _libc.srand(C.c_uint(seed & 0xFFFFFFFF))


def libc_rand() -> int:
    return _libc.rand()


class Oracle:
    def __init__(self):
        build()
        self.lib = lib = C.CDLL(LIB_PATH)
        lib.orc_transform.argtypes = [_f32p, C.c_long, C.c_double, _i32p, C.c_int, _f64p, C.c_float, _f32p, _f32p, _f32p]
        lib.orc_getpolar.argtypes = [C.c_double] * 3 + [C.POINTER(C.c_double)] * 3
        lib.orc_weight.restype = C.c_float
        lib.orc_weight.argtypes = [C.c_float, C.c_float, C.c_double]
        lib.orc_select_project.restype = C.c_long
        lib.orc_select_project.argtypes = [
            _f32p, _f32p, _f32p, C.c_void_p, C.c_float, C.c_double, C.c_long, C.c_double, C.c_double, C.c_double,
            C.c_int, C.c_double, C.c_int, C.c_int, _f32p, _f32p, _f32p, C.c_long,
        ]
        lib.orc_gridist_w.argtypes = [_f32p, _f32p, _f32p, C.c_long, C.c_int, C.c_int, _f32p]
        lib.orc_gridist_w_f64.argtypes = [_f32p, _f32p, _f32p, C.c_long, C.c_int, C.c_int, _f64p]
        lib.orc_gridist_w_fixed.argtypes = [_f32p, _f32p, _f32p, C.c_long, C.c_int, C.c_int, C.c_int, _i64p]
        lib.orc_ngp_cell.restype = C.c_long
        lib.orc_ngp_cell.argtypes = [C.c_float, C.c_float, C.c_int]
        lib.orc_randomize_box.argtypes = [C.c_int] * 4 + [_i32p, C.c_int, C.c_int, _f64p, _f64p, _f64p, _i32p, _i32p, _i32p, _i32p]
        lib.orc_cosmo_table.argtypes = [C.c_double] * 4 + [C.c_int, _f64p, _f64p]

    # -- primitives -----------------------------------------------------------------------------
    def transform(self, raw, boxsize, sgn, face, centre, rcase):
        raw = np.ascontiguousarray(raw, np.float32).reshape(-1, 3)
        n = raw.shape[0]
        x, y, z = (np.empty(n, np.float32) for _ in range(3))
        self.lib.orc_transform(raw, n, float(boxsize), np.ascontiguousarray(sgn, np.int32), int(face),
                               np.ascontiguousarray(centre, np.float64), float(np.float32(rcase)), x, y, z)
        return x, y, z

    def getpolar(self, x, y, z):
        ra, dec, d = C.c_double(), C.c_double(), C.c_double()
        self.lib.orc_getpolar(x, y, z, C.byref(ra), C.byref(dec), C.byref(d))
        return ra.value, dec.value, d.value

    def weight(self, ixx, ixh, dx):
        return np.float32(self.lib.orc_weight(float(np.float32(ixx)), float(np.float32(ixh)), float(dx)))

    def select_project(self, x, y, z, ld, ld2, boxsize, nrepperp, fovradiants, npix, const_mass=1.0,
                       per_particle=None, max_m=MAX_M, snopt=0, cap=None):
        n = len(x)
        cap = cap or max(1024, n // 4)
        pp = None
        if per_particle is not None:
            per_particle = np.ascontiguousarray(per_particle, np.float32)
            pp = per_particle.ctypes.data
        while True:
            xs, ys, ms = (np.empty(cap, np.float32) for _ in range(3))
            # snopt > 0 consumes libc rand() (one per accepted pair): the caller seeds it (libc_srand) and, because a
            # too-small buffer forces a second call, must give a `cap` that is large enough (checked below)
            na = self.lib.orc_select_project(x, y, z, pp, float(np.float32(const_mass)), float(max_m), n, float(ld),
                                             float(ld2), float(boxsize), int(nrepperp), float(fovradiants), int(npix),
                                             int(snopt), xs, ys, ms, cap)
            if snopt and na > cap:
                raise RuntimeError("select_project(snopt>0): pass cap >= number of accepted pairs")
            if na <= cap:
                return xs[:na].copy(), ys[:na].copy(), ms[:na].copy()
            cap = int(na)

    def gridist_w(self, xs, ys, ms, nn, do_ngp=False):
        m = np.zeros(nn * nn, np.float32)
        self.lib.orc_gridist_w(np.ascontiguousarray(xs, np.float32), np.ascontiguousarray(ys, np.float32),
                               np.ascontiguousarray(ms, np.float32), len(xs), nn, int(do_ngp), m)
        return m

    def gridist_w_f64(self, xs, ys, ms, nn, do_ngp=False, out=None):
        m = np.zeros(nn * nn, np.float64) if out is None else out
        self.lib.orc_gridist_w_f64(np.ascontiguousarray(xs, np.float32), np.ascontiguousarray(ys, np.float32),
                                   np.ascontiguousarray(ms, np.float32), len(xs), nn, int(do_ngp), m)
        return m

    def gridist_w_fixed(self, xs, ys, ms, nn, frac_bits, do_ngp=False, out=None):
        m = np.zeros(nn * nn, np.int64) if out is None else out
        self.lib.orc_gridist_w_fixed(np.ascontiguousarray(xs, np.float32), np.ascontiguousarray(ys, np.float32),
                                     np.ascontiguousarray(ms, np.float32), len(xs), nn, int(do_ngp), int(frac_bits), m)
        return m

    def ngp_cells(self, xs, ys, nn):
        return np.array([self.lib.orc_ngp_cell(float(a), float(b), nn) for a, b in zip(xs, ys)], np.int64)

    def randomize_box(self, seedcenter, seedface, seedsign, randomize, lens_per_snap=LENS_PER_SNAP, fixed_vertex=False):
        randomize = np.ascontiguousarray(randomize, np.int32)
        n = len(randomize)
        x0, y0, z0 = (np.zeros(n) for _ in range(3))
        face, sx, sy, sz = (np.zeros(n, np.int32) for _ in range(4))
        self.lib.orc_randomize_box(seedcenter, seedface, seedsign, n, randomize, lens_per_snap, int(fixed_vertex),
                                   x0, y0, z0, face, sx, sy, sz)
        return dict(x0=x0, y0=y0, z0=z0, face=face, sgnX=sx, sgnY=sy, sgnZ=sz)

    def cosmo_table(self, om0, oml, w, zs, n=1000):
        zl, dl = np.zeros(n), np.zeros(n)
        self.lib.orc_cosmo_table(om0, oml, w, zs, n, zl, dl)
        return zl, dl

    # -- composition: one plane from in-memory particle arrays -------------------------------------
    def plane_from_particles(self, types, plane, npix, do_ngp=False, frac_bits=None):
        """Mirror of readPos + mapParticles + the accumulation in createDensityMaps for one sub-file.

        types: list of dicts {type, raw [n,3] f32, const_mass | masses (f32 [n]), cut (apply MAX_M)}
        plane: dict boxsize, sgn, face, centre, rcase, ld, ld2, nrepperp, fovradiants
        Returns dict with per-type float maps ('maps', in reference float order), 'counts' [6],
        'ingrid' [6] (NGP in-grid hits), 'f64' per-type exact sums, and (if frac_bits) 'fixed' int64 maps.
        """
        maps = np.zeros((6, npix * npix), np.float32)
        f64 = np.zeros((6, npix * npix), np.float64)
        fixed = np.zeros((6, npix * npix), np.int64) if frac_bits is not None else None
        counts = np.zeros(6, np.int64)
        ingrid = np.zeros(6, np.int64)
        accepted = {}
        for t in types:
            ty = t["type"]
            x, y, z = self.transform(t["raw"], plane["boxsize"], plane["sgn"], plane["face"], plane["centre"], plane["rcase"])
            xs, ys, ms = self.select_project(
                x, y, z, plane["ld"], plane["ld2"], plane["boxsize"], plane.get("nrepperp", 0), plane["fovradiants"],
                npix, const_mass=t.get("const_mass", 0.0), per_particle=t.get("masses"),
                max_m=MAX_M if t.get("cut", True) else -1.0,
            )
            counts[ty] += len(xs)
            accepted[ty] = (xs, ys, ms)
            if len(xs):
                # mapParticles assigns (not adds) gridist_w's result per type per sub-file (densitymaps.cpp:407)
                maps[ty] = self.gridist_w(xs, ys, ms, npix, do_ngp)
                self.gridist_w_f64(xs, ys, ms, npix, do_ngp, out=f64[ty])
                if fixed is not None:
                    self.gridist_w_fixed(xs, ys, ms, npix, frac_bits, do_ngp, out=fixed[ty])
                cells = self.ngp_cells_fast(xs, ys, npix)
                ingrid[ty] += int((cells >= 0).sum())
        return dict(maps=maps, counts=counts, ingrid=ingrid, f64=f64, fixed=fixed, accepted=accepted)

    @staticmethod
    def ngp_cells_fast(xs, ys, nn):
        """Vectorised utilities.cpp:69-76 (float / double division, floor, in-grid test)."""
        dl = 1.0 / float(nn)
        gx = np.floor(xs.astype(np.float64) / dl).astype(np.int64)
        gy = np.floor(ys.astype(np.float64) / dl).astype(np.int64)
        ok = (gx >= 0) & (gx < nn) & (gy >= 0) & (gy < nn)
        return np.where(ok, gx + nn * gy, -1)
