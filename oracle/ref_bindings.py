"""TEST INFRASTRUCTURE — ctypes bindings to oracle/_ref/libslicer_ref{,_ngp}.so (the reference's own
sources compiled by oracle/Makefile, entry points in oracle/ref_harness.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def available(ngp: bool = False) -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libslicer_ref_ngp.so" if ngp else "libslicer_ref.so"))


class RefLib:
    """The compiled reference. `ngp=True` loads the `#define DO_NGP true` build (densitymaps.h:22)."""

    def __init__(self, ngp: bool = False):
        path = os.path.join(REF_DIR, "libslicer_ref_ngp.so" if ngp else "libslicer_ref.so")
        self.lib = lib = C.CDLL(path)
        lib.ref_do_ngp.restype = C.c_int
        lib.ref_lens_per_snap.restype = C.c_int
        lib.ref_max_m.restype = C.c_double
        lib.ref_weight.restype = C.c_float
        lib.ref_weight.argtypes = [C.c_float, C.c_float, C.c_double]
        lib.ref_getpolar.argtypes = [C.c_double] * 3 + [C.POINTER(C.c_double)] * 3
        lib.ref_gridist_w.argtypes = [_f32p, _f32p, _f32p, C.c_long, C.c_int, C.c_int, _f32p]
        lib.ref_srand.argtypes = [C.c_uint]
        lib.ref_randomize_box.argtypes = [C.c_int] * 4 + [_i32p, _f64p, _f64p, _f64p, _i32p, _i32p, _i32p, _i32p]
        lib.ref_cosmo_table.argtypes = [C.c_double] * 4 + [C.c_int, _f64p, _f64p]
        lib.ref_plan.argtypes = (
            [C.c_double] * 4
            + [C.c_int, _f64p, _f64p, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
            + [_f64p] * 4
            + [_i32p] * 3
            + [C.POINTER(C.c_int)]
        )
        lib.ref_read_header.argtypes = [C.c_char_p, _i32p, _f64p, _f64p, C.POINTER(C.c_int)]
        lib.ref_read_pos.argtypes = [C.c_char_p, _i32p, C.c_int, _f64p, C.c_float, _f32p, _f32p, _f32p]
        lib.ref_map_subfile.argtypes = [
            C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_int, _i32p, C.c_int, _f64p, C.c_float,
            C.c_double, C.c_double, C.c_int, _f32p, _i32p,
        ]
        lib.ref_create_density_maps.argtypes = [
            C.c_char_p, C.c_uint, C.c_uint, C.c_int, C.c_double, C.c_int, C.c_int, _i32p, C.c_int, _f64p,
            C.c_double, C.c_double, C.c_double, C.c_int, _f32p, C.c_void_p,
        ]

    # -- scalar helpers -------------------------------------------------------------------------
    def do_ngp(self) -> bool:
        return bool(self.lib.ref_do_ngp())

    def weight(self, ixx, ixh, dx) -> np.float32:
        return np.float32(self.lib.ref_weight(float(np.float32(ixx)), float(np.float32(ixh)), float(dx)))

    def getpolar(self, x, y, z):
        ra, dec, d = C.c_double(), C.c_double(), C.c_double()
        self.lib.ref_getpolar(x, y, z, C.byref(ra), C.byref(dec), C.byref(d))
        return ra.value, dec.value, d.value

    def gridist_w(self, x, y, w, nn, do_ngp):
        x = np.ascontiguousarray(x, np.float32)
        y = np.ascontiguousarray(y, np.float32)
        w = np.ascontiguousarray(w, np.float32)
        out = np.zeros(nn * nn, np.float32)
        self.lib.ref_gridist_w(x, y, w, len(x), nn, int(do_ngp), out)
        return out

    def srand(self, seed):
        self.lib.ref_srand(C.c_uint(seed & 0xFFFFFFFF))

    def randomize_box(self, seedcenter, seedface, seedsign, randomize):
        randomize = np.ascontiguousarray(randomize, np.int32)
        n = len(randomize)
        x0, y0, z0 = (np.zeros(n) for _ in range(3))
        face, sx, sy, sz = (np.zeros(n, np.int32) for _ in range(4))
        self.lib.ref_randomize_box(seedcenter, seedface, seedsign, n, randomize, x0, y0, z0, face, sx, sy, sz)
        return dict(x0=x0, y0=y0, z0=z0, face=face, sgnX=sx, sgnY=sy, sgnZ=sz)

    def cosmo_table(self, om0, oml, w, zs, n=1000):
        zl, dl = np.zeros(n), np.zeros(n)
        self.lib.ref_cosmo_table(om0, oml, w, zs, n, zl, dl)
        return zl, dl

    def plan(self, om0, oml, w, zs, snapred, snapbox, directory, suffix="t", cap=4096):
        snapred = np.ascontiguousarray(snapred, np.float64)
        snapbox = np.ascontiguousarray(snapbox, np.float64)
        nplanes, nrepl, Ds = C.c_int(), C.c_int(), C.c_double()
        ld, ld2, zsim, zfs = (np.zeros(cap) for _ in range(4))
        fromsnapi, randomize, replication = (np.zeros(cap, np.int32) for _ in range(3))
        rc = self.lib.ref_plan(
            om0, oml, w, zs, len(snapred), snapred, snapbox, directory.encode(), suffix.encode(), cap,
            C.byref(nplanes), C.byref(Ds), ld, ld2, zsim, zfs, fromsnapi, randomize, replication, C.byref(nrepl),
        )
        if rc:
            raise RuntimeError(f"ref_plan rc={rc}")
        n = nplanes.value
        # buildPlanes can append more ld entries than nplanes (densitymaps.cpp:122-123,154); keep nplanes
        return dict(
            nplanes=n, Ds=Ds.value, ld=ld[:n].copy(), ld2=ld2[:n].copy(), zsimlens=zsim[:n].copy(),
            zfromsnap=zfs[:n].copy(), fromsnapi=fromsnapi[:n].copy(), randomize=randomize[:n].copy(),
            replication=replication[: nrepl.value].copy(),
        )

    def read_header(self, file):
        npart = np.zeros(6, np.int32)
        massarr = np.zeros(6)
        sc = np.zeros(6)
        nf = C.c_int()
        if self.lib.ref_read_header(file.encode(), npart, massarr, sc, C.byref(nf)):
            raise FileNotFoundError(file)
        return dict(npart=npart, massarr=massarr, time=sc[0], redshift=sc[1], boxsize=sc[2], om0=sc[3],
                    oml=sc[4], h=sc[5], numfiles=nf.value)

    def read_pos(self, file, sgn, face, centre, rcase):
        hdr = self.read_header(file)
        n = int(hdr["npart"].sum())
        x, y, z = (np.zeros(n, np.float32) for _ in range(3))
        rc = self.lib.ref_read_pos(file.encode(), np.ascontiguousarray(sgn, np.int32), int(face),
                                   np.ascontiguousarray(centre, np.float64), float(np.float32(rcase)), x, y, z)
        if rc:
            raise RuntimeError("ref_read_pos failed")
        return x, y, z, hdr

    def map_subfile(self, file, npix, fovradiants, sgn, face, centre, rcase, ld, ld2, nrepperp=0, snopt=0, hydro=0):
        """-> (maps float32 [6, npix, npix], counts int32 [6]); maps[t][gy, gx]."""
        maps = np.zeros((6, npix * npix), np.float32)
        counts = np.zeros(6, np.int32)
        rc = self.lib.ref_map_subfile(
            file.encode(), npix, float(fovradiants), snopt, hydro, np.ascontiguousarray(sgn, np.int32), int(face),
            np.ascontiguousarray(centre, np.float64), float(np.float32(rcase)), float(ld), float(ld2),
            int(nrepperp), maps, counts,
        )
        if rc:
            raise RuntimeError(f"ref_map_subfile rc={rc}")
        return maps.reshape(6, npix, npix), counts

    def create_density_maps(self, file_base, ffmin, ffmax, npix, fovradiants, sgn, face, centre, rcase, ld, ld2,
                            nrepperp=0, snopt=0, hydro=0, per_type=False):
        tot = np.zeros(npix * npix, np.float32)
        per = np.zeros((6, npix * npix), np.float32) if per_type else None
        rc = self.lib.ref_create_density_maps(
            file_base.encode(), ffmin, ffmax, npix, float(fovradiants), snopt, hydro,
            np.ascontiguousarray(sgn, np.int32), int(face), np.ascontiguousarray(centre, np.float64),
            float(np.float32(rcase)), float(ld), float(ld2), int(nrepperp), tot,
            per.ctypes.data if per is not None else None,
        )
        if rc:
            raise RuntimeError(f"ref_create_density_maps rc={rc}")
        if per_type:
            return tot.reshape(npix, npix), per.reshape(6, npix, npix)
        return tot.reshape(npix, npix)
