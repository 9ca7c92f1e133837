"""The C restatement (oracle/slicer_oracle.c) against the golden vectors in tests/golden/, which
oracle/make_golden.py generated from the reference's own compiled code.  CPU only."""
import numpy as np
import pytest

from conftest import unhex


def test_weight_bit_exact(oracle, kat):
    for x, c, dl, w in kat["weight"]:
        got = oracle.weight(np.float32(unhex(x)), np.float32(unhex(c)), unhex(dl))
        assert float(got) == unhex(w)


def test_getpolar_bit_exact(oracle, kat):
    for x, y, z, ra, dec, d in kat["getpolar"]:
        g = oracle.getpolar(unhex(x), unhex(y), unhex(z))
        assert g == (unhex(ra), unhex(dec), unhex(d))


def test_gridist_small(oracle, kat):
    g = kat["gridist_small"]
    x = np.array([unhex(v) for v in g["x"]], np.float32)
    y = np.array([unhex(v) for v in g["y"]], np.float32)
    w = np.array([unhex(v) for v in g["w"]], np.float32)
    tsc = oracle.gridist_w(x, y, w, g["nn"], False)
    ngp = oracle.gridist_w(x, y, w, g["nn"], True)
    assert tsc.tolist() == [unhex(v) for v in g["tsc"]]
    assert ngp.tolist() == [unhex(v) for v in g["ngp"]]
    # SURVEY.md App. A spot values
    assert ngp[36] == np.float32(1.03750002) and ngp[39] == np.float32(0.5) and ngp[56] == np.float32(2.0)


def test_randomize_box(oracle, kat):
    for r in kat["randomize_box"]:
        got = oracle.randomize_box(*r["seeds"], r["randomize"])
        for k in ("x0", "y0", "z0"):
            assert got[k].tolist() == [unhex(v) for v in r[k]]
        for k in ("face", "sgnX", "sgnY", "sgnZ"):
            assert got[k].tolist() == r[k]
    # example seeds, group 0
    got = oracle.randomize_box(-229, -230, -231, [1, 0, 0, 0, 1, 0, 0, 0, 1])
    assert got["face"].tolist() == [1, 1, 1, 1, 3, 3, 3, 3, 4]
    assert (got["sgnX"][0], got["sgnY"][0], got["sgnZ"][0]) == (1, -1, 1)
    assert abs(got["x0"][0] - 0.904897332) < 1e-9 and abs(got["z0"][8] - 0.54057765) < 1e-9


def test_cosmo_table(oracle, kat):
    for c in kat["cosmo_table"]:
        zl, dl = oracle.cosmo_table(c["om0"], c["oml"], c["w"], c["zs"])
        assert [zl[i] for i in c["idx"]] == [unhex(v) for v in c["zl"]]
        assert [dl[i] for i in c["idx"]] == [unhex(v) for v in c["dl"]]


@pytest.mark.parametrize("name", ["dm_face1", "dm_face3_pile2", "dm_face4_repl", "hydro_multi", "odd_box_face2", "face5_repl2"])
def test_particle_cases(oracle, golden, name):
    m = golden.meta[name]
    plane = golden.plane(name)
    types = golden.types(name)
    # readPos: transformed coordinates of all types, concatenated in type order
    xs = [oracle.transform(t["raw"], plane["boxsize"], plane["sgn"], plane["face"], plane["centre"], plane["rcase"]) for t in types]
    for k, key in enumerate("xyz"):
        got = np.concatenate([x[k] for x in xs])
        assert np.array_equal(got.view(np.uint32), golden.arr[f"{name}/{key}"].view(np.uint32)), key
    # mapParticles: counts and float maps (bit-exact: same float32 accumulation order)
    for do_ngp, tag in ((False, "tsc"), (True, "ngp")):
        res = oracle.plane_from_particles(types, plane, m["npix"], do_ngp=do_ngp, frac_bits=40)
        assert res["counts"].tolist() == m["counts"]
        for t in m["types_with_maps"]:
            want = golden.arr[f"{name}/{tag}{t}"].reshape(-1)
            assert np.array_equal(res["maps"][t].view(np.uint32), want.view(np.uint32)), (tag, t)
            # the exact accumulators agree with the float map to float accuracy, and with each other to 2^-40
            np.testing.assert_allclose(res["f64"][t], want, rtol=3e-6, atol=1e-6)
            np.testing.assert_allclose(res["fixed"][t] * 2.0**-40, res["f64"][t], rtol=0, atol=9 * m["counts"][t] * 2.0**-41 + 1e-300)
