"""Pins the CPU oracle (oracle/) to vectors produced by the unmodified reference
(oracle/make_golden.py) and to the reference's own known-answer tests
(tests/integration/test_exhaustive_fingerprinting.py:38-89)."""
import numpy as np
import pytest

from tests.conftest import SOLVER_CASE_NAMES
from oracle import oracle as orc


def test_boundary_cases_1d():
    # reference tests/integration/test_exhaustive_fingerprinting.py:38-59
    s2 = np.sqrt(2.0)
    A = np.array([[0.0], [1.0], [0.0]])
    Y = np.array([[1, 0, s2 / 2, 0, s2 / 2], [0, 0, -s2 / 2, 2, s2 / 2], [0, 1, 0, 0, 0]])
    w_exp = [0, 0, 0, 2, s2 / 2]
    obj_exp = [1, 1, 1, 0, 0.5]
    for i in range(5):
        w, sub, tot, obj, yrec = orc.solve(A, Y[:, i].copy(), np.array([1]))
        assert np.isclose(w[0], w_exp[i]) and np.isclose(obj, obj_exp[i])


def test_boundary_cases_2d():
    # reference tests/integration/test_exhaustive_fingerprinting.py:62-89
    s2, s3 = np.sqrt(2.0), np.sqrt(3.0)
    A = np.array([[0.5, s3 * 0.5], [s3 * 0.5, 0.5]])
    Y = np.array([[-s3 / 2, 0.5, -1, -s3 / 2, 0.5001, 0.5, s3 / 2, s2 / 2, -s2 / 2.0],
                  [0.5, -s3 / 2, 0, 0.5001, -s3 / 2, s3 / 2, 0.5, s2 / 2, -s2 / 2.0]])
    w_exp = np.array([[0, 0], [0, 0], [0, 0], [8.66025404e-05, 0], [0, 8.66025404e-05],
                      [1, 0], [0, 1], [0.51763809, 0.51763809], [0, 0]])
    obj_exp = np.array([1, 1, 1, 1.0001000025, 1.0001000025, 0, 0, 0, 1])
    for i in range(9):
        w, sub, tot, obj, yrec = orc.solve(A, Y[:, i].copy(), np.array([1, 1]))
        assert np.all(np.isclose(w, w_exp[i])), i
        assert np.isclose(obj, obj_exp[i]), i


@pytest.mark.parametrize("name", SOLVER_CASE_NAMES)
def test_solver_matches_reference(solver_cases, name):
    A = solver_cases[name + "_A"]
    Y = solver_cases[name + "_Y"]
    sizes = solver_cases[name + "_sizes"]
    for v in range(Y.shape[0]):
        w, sub, tot, obj, yrec = orc.solve(A, Y[v].copy(), sizes)
        assert np.array_equal(sub, solver_cases[name + "_sub"][v]), (name, v)
        if sizes.size <= 3:
            # same summation order, no FMA: bit-identical to the Numba code
            assert np.array_equal(w, solver_cases[name + "_w"][v]), (name, v)
            assert obj == solver_cases[name + "_obj"][v], (name, v)
        else:
            # >= 4 blocks: both sides call scipy.optimize.nnls
            assert np.allclose(w, solver_cases[name + "_w"][v], rtol=1e-12, atol=0)
            assert np.isclose(obj, solver_cases[name + "_obj"][v], rtol=1e-12)
        assert np.allclose(yrec, solver_cases[name + "_yrec"][v], rtol=1e-13, atol=1e-15)


def test_reference_synthetic(ref_synth):
    # reference test_synthetic_data (:94-153), shrunk; outputs from the reference
    A, Y, sizes = ref_synth["A"], ref_synth["Y"], ref_synth["sizes"]
    for v in range(Y.shape[0]):
        w, sub, tot, obj, _ = orc.solve(A, Y[v].copy(), sizes)
        assert np.array_equal(tot, ref_synth["tot"][v])
        assert np.array_equal(tot, ref_synth["ID"][:, v])
        assert np.array_equal(w, ref_synth["w"][v])
        assert obj == ref_synth["obj"][v]


def test_table_matches_reference_interpolators(ukbb):
    tab = orc.init_table(ukbb["dictionary"], ukbb["sch_mat"], ukbb["orientation"])
    assert np.array_equal(tab["off"], ukbb["ref_off"])
    assert np.array_equal(tab["nodes"], ukbb["ref_nodes"])
    assert np.array_equal(tab["table"], ukbb["ref_table"])
    assert np.array_equal(tab["Gms_un"], ukbb["ref_Gms_un"])


@pytest.mark.parametrize("mode", ["exact", "between"])
def test_rotation_matches_reference(ukbb, mode):
    tab = orc.init_table(ukbb["dictionary"], ukbb["sch_mat"], ukbb["orientation"])
    plan = orc.plan_scheme(tab, ukbb["sch_" + mode])
    if mode == "between":
        assert np.any(plan["shell_hi"] != plan["shell_lo"])
    for d, ref in zip(ukbb["dirs"], ukbb["rot_" + mode]):
        D = orc.rotate(tab, plan, d)
        # reference |g.u| goes through BLAS gemv -> a few ulp, not bitwise
        assert np.allclose(D, ref, rtol=1e-12, atol=1e-15)


def _oracle_fit_maps(ukbb, numfasc, csf, ear):
    tab = orc.init_table(ukbb["dictionary"], ukbb["sch_mat"], ukbb["orientation"])
    sch = ukbb["sch_exact"]
    plan = orc.plan_scheme(tab, sch)
    mask = ukbb["mask"] > 0
    sig_csf, sig_ear = orc.iso_signals(sch, float(ukbb["T2_csf"]), float(ukbb["DIFF_csf"]),
                                       float(ukbb["T2_ear"]), ukbb["DIFF_ear"])
    Kv = numfasc[mask].astype(int)
    maxfasc = int(Kv.max())
    csfv = np.zeros(Kv.size) if csf is None else csf[mask]
    earv = np.zeros(Kv.size) if ear is None else ear[mask]
    rows = orc.fit_rows(tab, plan, ukbb["data"][mask], Kv, csfv, earv,
                        ukbb["peaks"][mask][:, :3 * maxfasc], sig_csf, sig_ear)
    return rows, maxfasc, bool(np.any(csfv > 0)), bool(np.any(earv > 0)), mask


@pytest.mark.parametrize("tag", ["A", "B", "C", "D"])
def test_fit_rows_match_reference(ukbb, tag):
    numfasc, csf, ear = ukbb["numfasc"], ukbb["csf"], ukbb["ear"]
    if tag == "B":
        csf = ear = None
    elif tag == "A":
        ear = None
    elif tag == "C":
        numfasc = np.minimum(numfasc, 1)
    rows, maxfasc, csf_on, ear_on, mask = _oracle_fit_maps(ukbb, numfasc, csf, ear)

    def ref(name):
        return ukbb["fit%s_%s" % (tag, name)][mask]

    assert np.allclose(rows[:, 0], ref("M0"), rtol=1e-9, atol=0)
    for k in range(maxfasc):
        assert np.allclose(rows[:, 1 + k], ref("frac_f%d" % k), rtol=1e-9, atol=1e-300)
        # atom IDs through the derived property maps (rad is injective on the subset)
        ids = rows[:, 1 + maxfasc + k].astype(int)
        assert np.allclose(ukbb["rad"][ids] * (rows[:, 1 + k] > 0), ref("rad_f%d" % k),
                           rtol=0, atol=0)
    if csf_on:
        assert np.allclose(rows[:, 2 * maxfasc + 1], ref("frac_csf"), rtol=1e-9, atol=1e-300)
    if ear_on:
        i_ear = 2 * maxfasc + csf_on + 1
        assert np.allclose(rows[:, i_ear], ref("frac_ear"), rtol=1e-9, atol=1e-300)
        ids = rows[:, i_ear + 1].astype(int)
        assert np.array_equal(ukbb["DIFF_ear"][ids] * (rows[:, i_ear] > 0), ref("D_ear"))
    ysq = np.sum(ukbb["data"][mask] ** 2, axis=1) / ukbb["data"].shape[-1]
    assert np.all(np.abs(rows[:, -2] - ref("MSE")) <= 1e-12 * ysq + 1e-9 * ref("MSE"))
    assert np.allclose(rows[:, -1], ref("R2"), rtol=1e-9, atol=1e-12)


def test_mc_average_oracle_matches_reference():
    """monte_carlo_average (mfu:2758-2812): the sequential C restatement against the
    unmodified Numba kernel (tests/golden/mc_cases.npz, oracle/make_golden.py mc_cases)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mc_cases.npz"))
    for dim in (2, 3):
        for ds in (1.0, 0.73):
            got = orc.mc_average(g["phases"][:, :dim], g["pick"], g["gsc"][:, :dim], ds, int(g["n_spin"]))
            # same summation order; libm vs Numba's cos may differ in the last ulp per term
            assert np.allclose(got, g["avg_d%d_s%g" % (dim, ds)], rtol=0, atol=1e-14)


@pytest.mark.parametrize("tag", ["between", "dense"])
def test_fit_rows_match_reference_other_protocols(ukbb, tag):
    """Oracle _fit_voxel on a between-shell protocol (mfu:1921-1956) and on the 271-row dense
    protocol against MFModel.fit of the unmodified reference (tests/golden/fit_protocols.npz,
    oracle/make_golden.py fit_protocol_cases)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fit_protocols.npz"))
    sch = g["sch_" + tag]
    tab = orc.init_table(ukbb["dictionary"], ukbb["sch_mat"], ukbb["orientation"])
    plan = orc.plan_scheme(tab, sch)
    mask = g["mask"] > 0
    sig_csf, sig_ear = orc.iso_signals(sch, float(ukbb["T2_csf"]), float(ukbb["DIFF_csf"]),
                                       float(ukbb["T2_ear"]), ukbb["DIFF_ear"])
    Kv = g["numfasc"][mask].astype(int)
    rows = orc.fit_rows(tab, plan, g["data_" + tag][mask], Kv, g["csf"][mask], np.zeros(Kv.size),
                        g["peaks"][mask][:, :6], sig_csf, sig_ear)

    def ref(name):
        return g["fit_%s_%s" % (tag, name)][mask]
    assert np.allclose(rows[:, 0], ref("M0"), rtol=1e-9, atol=0)
    for k in range(2):
        assert np.allclose(rows[:, 1 + k], ref("frac_f%d" % k), rtol=1e-9, atol=1e-300)
        ids = rows[:, 3 + k].astype(int)
        assert np.array_equal(ukbb["rad"][ids] * (rows[:, 1 + k] > 0), ref("rad_f%d" % k))
    assert np.allclose(rows[:, 5], ref("frac_csf"), rtol=1e-9, atol=1e-300)
    ysq = np.sum(g["data_" + tag][mask] ** 2, axis=1) / sch.shape[0]
    assert np.all(np.abs(rows[:, -2] - ref("MSE")) <= 1e-12 * ysq + 1e-9 * ref("MSE"))
    assert np.allclose(rows[:, -1], ref("R2"), rtol=1e-9, atol=1e-12)


def test_solve2_gram_matches_reference_order_solver():
    """oracle.solve2_gram (BLAS Gram + vectorised branches, used for index checks at shapes
    where the strided Gram takes minutes) against the reference-order C restatement."""
    rng = np.random.default_rng(77)
    for trial in range(20):
        M, n1, n2 = int(rng.integers(8, 60)), int(rng.integers(1, 70)), int(rng.integers(1, 70))
        signed = trial % 3 == 0
        A = rng.standard_normal((M, n1 + n2)) if signed else rng.random((M, n1 + n2)) + 0.05
        y = A[:, [rng.integers(0, n1), n1 + rng.integers(0, n2)]] @ rng.random(2) + 0.05 * rng.standard_normal(M)
        if trial % 5 == 4:
            y = -np.abs(y)
        w, sub, tot, obj, yrec = orc.solve(A, y, [n1, n2])
        wg, subg, objg = orc.solve2_gram(A, y, [n1, n2])
        assert np.array_equal(sub, subg), trial
        assert np.allclose(w, wg, rtol=1e-9, atol=1e-12) and abs(obj - objg) <= 1e-10 * float(y @ y)
