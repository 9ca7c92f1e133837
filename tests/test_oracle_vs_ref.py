"""Randomised cross-check of the C restatement against the reference's own compiled code
(oracle/_ref/libslicer_ref{,_ngp}.so, built by oracle/Makefile from /root/reference/SLICER).  CPU only; skipped
when the reference build is absent."""
import os

import numpy as np
import pytest

from slicer_b200 import synth


def test_ref_build_flags(reflib, reflib_ngp):
    assert not reflib.do_ngp() and reflib_ngp.do_ngp()
    assert reflib.lib.ref_lens_per_snap() == 4
    assert reflib.lib.ref_max_m() == 1e3


def test_weight_random(oracle, reflib):
    rng = np.random.default_rng(5)
    for nn in (7, 256, 1000, 8192):
        dl = 1.0 / nn
        for _ in range(300):
            x = np.float32(rng.random())
            g = int(np.floor(float(x) / dl)) + int(rng.integers(-2, 3))
            c = np.float32((g + 0.5) * dl)
            assert float(oracle.weight(x, c, dl)) == float(reflib.weight(x, c, dl))


def test_gridist_random(oracle, reflib):
    rng = np.random.default_rng(6)
    for nn in (16, 100):
        n = 3000
        x = (rng.random(n) * 1.1 - 0.05).astype(np.float32)
        y = (rng.random(n) * 1.1 - 0.05).astype(np.float32)
        w = (rng.random(n) * 3).astype(np.float32)
        for ngp in (False, True):
            a = oracle.gridist_w(x, y, w, nn, ngp)
            b = reflib.gridist_w(x, y, w, nn, ngp)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_randomize_box_random(oracle, reflib):
    rng = np.random.default_rng(7)
    for _ in range(5):
        seeds = [int(v) for v in rng.integers(-10000, 10000, 3)]
        randomize = [1] + [int(v) for v in rng.integers(0, 2, 19)]
        a = oracle.randomize_box(*seeds, randomize)
        b = reflib.randomize_box(*seeds, randomize)
        for k in a:
            assert np.array_equal(a[k], b[k]), k


def test_cosmo_table_random(oracle, reflib):
    for (om, ol, w, zs) in ((0.3, 0.7, -1.0, 0.5), (0.27, 0.7, -0.8, 3.0), (0.3, 0.7, -1.0, 4.0)):
        a = oracle.cosmo_table(om, ol, w, zs)
        b = reflib.cosmo_table(om, ol, w, zs)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("face", [1, 2, 3, 4, 5, 6])
def test_subfile_random(oracle, reflib, reflib_ngp, tmp_path, face):
    """readPos + mapParticles of the reference on a synthetic sub-file == oracle composition, bit for bit."""
    rng = np.random.default_rng(100 + face)
    box = [128000.0, 99999.9, 250000.0][face % 3]
    n = 8000
    pos = {1: synth.uniform_positions(n, box, 200 + face), 4: synth.uniform_positions(n // 4, box, 300 + face)}
    masses = {4: (rng.random(n // 4) * 5).astype(np.float32)}
    masses[4][::11] = 2e3
    massarr = [0, 0.75, 0, 0, 0, 0]
    base = str(tmp_path / "snap")
    synth.write_snapshot(base, pos, massarr, 0.0, box, masses=masses)
    sgn = [int(v) for v in rng.choice([-1, 1], 3)]
    centre = [float(np.float32(v)) for v in rng.random(3)]
    pile = int(rng.integers(0, 4))
    q = int(rng.integers(0, 4))
    ld = (pile + q / 4) * box / 1e3
    ld2 = (pile + (q + 1) / 4) * box / 1e3
    fovrad = float(np.float32(30.0 / (pile + 1))) / 180.0 * np.pi
    npix = 64
    plane = dict(boxsize=box, sgn=sgn, face=face, centre=centre, rcase=float(pile), ld=ld, ld2=ld2, nrepperp=face % 2,
                 fovradiants=fovrad)
    types = [dict(type=1, raw=pos[1], const_mass=massarr[1]), dict(type=4, raw=pos[4], masses=masses[4], cut=True)]
    for lib, ngp in ((reflib, False), (reflib_ngp, True)):
        maps, counts = lib.map_subfile(base + ".0", npix, fovrad, sgn, face, centre, float(pile), ld, ld2, face % 2, hydro=1)
        res = oracle.plane_from_particles(types, plane, npix, do_ngp=ngp)
        assert res["counts"].tolist() == counts.tolist()
        assert counts[1] > 20
        for t in (1, 4):
            assert np.array_equal(res["maps"][t].view(np.uint32), maps[t].reshape(-1).view(np.uint32))
