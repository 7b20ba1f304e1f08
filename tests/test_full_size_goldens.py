"""Full-size parity against the UNMODIFIED reference (tests/golden/full_*.npz, written by
oracle/make_golden_full.py in the build container): BASELINE configs 1-5 at their real
shapes and the reference's own test_hcp_dict at its full 782 atoms.

  -m "not gpu": the CPU oracle against the same goldens where it finishes in seconds
  -m gpu      : the CUDA path through the public API / C ABI
Tolerances (BASELINE.json north_star): atom indices exact, weights / fractions / derived maps
1e-9 relative, MSE with floor 1e-12 |y|^2 / M.
"""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN
from tests import phantom
from oracle import oracle as orc


def _load(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip("%s not generated" % name)
    return np.load(path)


def ukbb_dictionary():
    d = _load("ukbb_dictionary.npz")
    dic = {k: d[k] for k in d.files}
    for k in ("num_atom", "num_ear"):
        dic[k] = int(dic[k])
    for k in ("T2_csf", "DIFF_csf", "T2_ear"):
        dic[k] = float(dic[k])
    dic["fasc_propnames"] = ["rad", "fin"]
    return dic


def _dictionary(tag):
    return ukbb_dictionary() if tag == "ukbb986" else phantom.make_dictionary(1000)


def _compare_maps(fit_maps, names, g, Y):
    ysq = np.sum(Y ** 2, axis=-1) / Y.shape[-1]
    assert list(names) == [str(s) for s in g["param_names"]]
    for p in names:
        got, ref = fit_maps[p], g["fit_" + p]
        assert got.shape == ref.shape, p
        if p == "MSE":
            assert np.all(np.abs(got - ref) <= 1e-12 * ysq.reshape(ref.shape) + 1e-9 * np.abs(ref)), p
        elif p == "R2":
            assert np.allclose(got, ref, rtol=1e-9, atol=1e-12), p
        elif p.startswith(("rad_f", "fin_f", "fvf_f", "dperp_in_f")):
            assert np.array_equal(got, ref), p          # dictionary lookups by atom index: exact
        else:
            assert np.allclose(got, ref, rtol=1e-9, atol=1e-300), p


def _oracle_maps(dic, sch, g, sel):
    """params rows of the CPU oracle for voxels `sel` -> the columns the maps are made of."""
    tab = orc.init_table(dic["dictionary"], dic["sch_mat"], dic["orientation"])
    plan = orc.plan_scheme(tab, sch)
    sig_csf, _ = orc.iso_signals(sch, dic["T2_csf"], dic["DIFF_csf"], dic["T2_ear"], dic["DIFF_ear"])
    Y = g["Y"].astype(np.float64)
    maxfasc = int(g["K"].max())
    return np.stack([orc.fit_voxel(tab, plan, Y[v], int(g["K"][v]), int(g["csf"][v]), 0, g["peaks"][v, :3 * maxfasc],
                                   maxfasc, True, False, sig_csf, None) for v in sel])


# ------------------------------------------------------------------------------------------
# CPU: the oracle against the reference at full size
# ------------------------------------------------------------------------------------------
def test_oracle_config1_full_dictionary():
    """BASELINE config 1 (16 x 16 x 4, numfasc = 1 + CSF, real 986-atom dictionary): every voxel."""
    g = _load("full_config1.npz")
    dic = ukbb_dictionary()
    sch = phantom.load_schemes()[1]
    rows = _oracle_maps(dic, sch, g, range(1024))
    M0, nu, ID, nucsf = rows[:, 0], rows[:, 1], rows[:, 2].astype(int), rows[:, 3]
    assert np.allclose(M0, g["fit_M0"].ravel(), rtol=1e-9)
    assert np.allclose(nu, g["fit_frac_f0"].ravel(), rtol=1e-9, atol=1e-300)
    assert np.allclose(nucsf, g["fit_frac_csf"].ravel(), rtol=1e-9, atol=1e-300)
    assert np.array_equal(dic["rad"][ID] * (nu > 0), g["fit_rad_f0"].ravel())


@pytest.mark.parametrize("tag", ["analytic1000", "ukbb986"])
def test_oracle_config3_full_size(tag):
    """BASELINE config 3 shape (numfasc = 2, CSF 30 %, N = 1000 / the real 986 atoms): a
    24-voxel subsample of the 256 reference voxels (the oracle needs ~0.1 s per voxel)."""
    g = _load("full_config3_%s.npz" % tag)
    dic = _dictionary(tag)
    sch = phantom.load_schemes()[1]
    sel = np.arange(0, 256, 11)
    rows = _oracle_maps(dic, sch, g, sel)
    prop = "rad" if tag == "ukbb986" else "fvf"
    assert np.allclose(rows[:, 0], g["fit_M0"][sel], rtol=1e-9)
    for k in range(2):
        assert np.allclose(rows[:, 1 + k], g["fit_frac_f%d" % k][sel], rtol=1e-9, atol=1e-300)
        assert np.array_equal(dic[prop][rows[:, 3 + k].astype(int)] * (rows[:, 1 + k] > 0),
                              g["fit_%s_f%d" % (prop, k)][sel])
    assert np.allclose(rows[:, 5], g["fit_frac_csf"][sel], rtol=1e-9, atol=1e-300)


def _config2_dictionaries(g, voxels):
    """Explicit dictionaries of the config-2 voxels from the CPU oracle's rotation."""
    dic = phantom.make_dictionary(800)
    sch = phantom.load_schemes()[1]
    tab = orc.init_table(dic["dictionary"], dic["sch_mat"], dic["orientation"])
    plan = orc.plan_scheme(tab, sch)
    return [np.hstack([orc.rotate(tab, plan, g["peaks"][v, :3]), orc.rotate(tab, plan, g["peaks"][v, 3:]),
                       g["sig_csf"][:, None]]) for v in voxels]


def test_oracle_config2_full_size():
    g = _load("full_config2.npz")
    vox = [0, 9, 21]
    Y = g["Y"].astype(np.float64)
    for v, A in zip(vox, _config2_dictionaries(g, vox)):
        w, sub, tot, obj, _ = orc.solve(A, Y[v], [800, 800, 1])
        assert np.array_equal(sub, g["sub3"][v]) and np.allclose(w, g["w3"][v], rtol=1e-9, atol=1e-300)
        assert abs(obj - g["obj3"][v]) <= 1e-12 * float(Y[v] @ Y[v]) + 1e-9 * abs(obj)
        w, sub, tot, obj, _ = orc.solve(np.ascontiguousarray(A[:, :1600]), Y[v], [800, 800])
        assert np.array_equal(sub, g["sub2"][v]) and np.allclose(w, g["w2"][v], rtol=1e-9, atol=1e-300)


def config4_problem(V=4, N=300, M=100, seed=507):
    """Same arithmetic-only generator as oracle/make_golden_full.py (bit-reproducible)."""
    rng = np.random.default_rng(seed)
    nt = 3 * N
    decay = 1.0 / (1.0 + 3.0 * rng.random((1, nt)) * np.linspace(0, 1, M)[:, None]) ** 2
    base = rng.random((M, nt)) * decay
    A = base[None] * (1.0 + 0.05 * rng.standard_normal((V, M, nt)))
    idx = rng.integers(0, N, (V, 3))
    wts = 0.2 + 0.8 * rng.random((V, 3))
    Y = np.stack([A[v][:, idx[v] + N * np.arange(3)] @ wts[v] for v in range(V)])
    Y = Y + 0.02 * rng.standard_normal(Y.shape)
    return A, Y, idx


def test_oracle_config4_one_voxel():
    g = _load("full_config4.npz")
    A, Y, idx = config4_problem()
    assert float(np.sum(A)) + float(np.sum(Y)) == float(g["checksum"]), "problem generator drifted"
    w, sub, tot, obj, _ = orc.solve(A[0], Y[0], [300, 300, 300])
    assert np.array_equal(sub, g["sub"][0]) and np.array_equal(w, g["w"][0]) and obj == float(g["obj"][0])


# ------------------------------------------------------------------------------------------
# GPU: the CUDA path against the reference at full size
# ------------------------------------------------------------------------------------------
def _fit_maps(model, g, sch, shape):
    Y = g["Y"].astype(np.float64)
    fit = model.fit(Y.reshape(shape + (-1,)), np.ones(shape), g["K"].astype(float).reshape(shape),
                    peaks=g["peaks"].reshape(shape + (6,)), pgse_scheme=sch, csf_mask=g["csf"].reshape(shape),
                    verbose=0)
    return {p: getattr(fit, p) for p in fit.param_names}, fit.param_names, Y.reshape(shape + (-1,))


@pytest.mark.gpu
def test_gpu_config1_matches_reference():
    from microstructure_fingerprinting_b200 import MFModel
    g = _load("full_config1.npz")
    model = MFModel(ukbb_dictionary())
    maps, names, Y = _fit_maps(model, g, phantom.load_schemes()[1], (16, 16, 4))
    _compare_maps(maps, names, g, Y)
    model.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["analytic1000", "ukbb986"])
def test_gpu_config3_matches_reference(tag):
    """256 voxels at the benchmark shape through MFModel.fit (fast tier + exact re-evaluation)
    against the unmodified reference, on the analytic and on the real Monte-Carlo dictionary
    (atoms correlated to 0.999999997); also exact=True, and the hand-over share."""
    from microstructure_fingerprinting_b200 import MFModel
    g = _load("full_config3_%s.npz" % tag)
    dic = _dictionary(tag)
    model = MFModel(dic)
    maps, names, Y = _fit_maps(model, g, phantom.load_schemes()[1], (256,))
    _compare_maps(maps, names, g, Y)
    fit = model.fit(Y, np.ones(256), 2, peaks=g["peaks"], pgse_scheme=phantom.load_schemes()[1],
                    csf_mask=g["csf"], verbose=0, exact=True)
    for p in names:
        assert np.array_equal(getattr(fit, p), maps[p]), p
    model.close()


@pytest.mark.gpu
def test_gpu_config2_matches_reference():
    """[800, 800, 1] and [800, 800] on explicit per-voxel dictionaries (mfb_rotate_multishell +
    mfb_solve_batch) against the reference's solve_exhaustive_posweights, 32 voxels."""
    import torch
    from microstructure_fingerprinting_b200 import mf_utils as mfu
    g = _load("full_config2.npz")
    dic = phantom.make_dictionary(800)
    sch = phantom.load_schemes()[1]
    msi = mfu.init_PGSE_multishell_interp(dic["dictionary"], dic["sch_mat"], dic["orientation"])
    plan = mfu.GpuPlan(msi, mfu.SchemePlan(msi, sch))
    V, M, N = 32, sch.shape[0], 800
    A = torch.empty((V, M, 2 * N + 1), dtype=torch.float64, device="cuda")
    A[:, :, :N] = plan.rotate(g["peaks"][:, :3])
    A[:, :, N:2 * N] = plan.rotate(g["peaks"][:, 3:])
    A[:, :, 2 * N] = torch.from_numpy(g["sig_csf"]).cuda()[None, :]
    Y = g["Y"].astype(np.float64)
    ysq = np.sum(Y ** 2, axis=1)
    for tag, sizes, Av in (("3", [N, N, 1], A), ("2", [N, N], A[:, :, :2 * N].contiguous())):
        w, sub, tot, obj, yrec = mfu.solve_exhaustive_posweights_batch(Av, Y, np.asarray(sizes))
        assert np.array_equal(sub, g["sub" + tag])
        assert np.allclose(w, g["w" + tag], rtol=1e-9, atol=1e-300)
        assert np.all(np.abs(obj - g["obj" + tag]) <= 1e-12 * ysq + 1e-9 * np.abs(obj))
    plan.close()


@pytest.mark.gpu
def test_gpu_config4_matches_reference():
    """[300, 300, 300] (2.7e7 tuples per voxel): triple scan + exact re-evaluation, bit-identical
    weights and objective (1-3 blocks reproduce the reference's arithmetic)."""
    from microstructure_fingerprinting_b200 import mf_utils as mfu
    g = _load("full_config4.npz")
    A, Y, idx = config4_problem()
    assert float(np.sum(A)) + float(np.sum(Y)) == float(g["checksum"]), "problem generator drifted"
    w, sub, tot, obj, yrec = mfu.solve_exhaustive_posweights_batch(A, Y, np.array([300, 300, 300]))
    assert np.array_equal(sub, g["sub"]) and np.array_equal(w, g["w"]) and np.array_equal(obj, g["obj"])


def config5_dictionary(sch, N=2000):
    from microstructure_fingerprinting_b200 import mf_utils as mfu
    gam = mfu.get_gyromagnetic_ratio("H")
    b = (gam * sch[:, 5] * sch[:, 3]) ** 2 * (sch[:, 4] - sch[:, 5] / 3)
    n_d = int(np.ceil(np.sqrt(N * 1.25)))
    n_f = int(np.ceil(N / n_d))
    DP, FI = np.meshgrid(np.geomspace(0.02e-9, 1.2e-9, n_d), np.linspace(0.2, 0.9, n_f), indexing="ij")
    dperp, f_in = DP.ravel()[:N], FI.ravel()[:N]
    return f_in[None, :] * np.exp(-b[:, None] * dperp[None, :]) + (1 - f_in[None, :]) * np.exp(-b[:, None] * 1.5e-9)


@pytest.mark.gpu
def test_gpu_config5_matches_reference():
    """AxCaliber 2D protocol, M = 1776, N = 2000 atoms per fascicle: rotate_atom_2Dprotocol per
    fascicle + [2000, 2000] search through the chunked pipeline, against the reference."""
    from microstructure_fingerprinting_b200 import mf_utils as mfu
    g = _load("full_config5.npz")
    sch = np.load(os.path.join(GOLDEN, "lowlevel_rotation.npz"))["ax_sch"]
    sig = config5_dictionary(sch)
    assert abs(float(np.sum(sig)) - float(g["dict_checksum"])) <= 1e-9 * float(g["dict_checksum"])
    Y = g["Y"].astype(np.float64)
    w, sub, obj, ok = mfu.solve_rotated_2Dprotocol_batch(sig, sch, np.array([0.0, 0.0, 1.0]), g["peaks"], Y, 2.0e-9)
    assert np.all(ok)
    assert np.array_equal(sub, g["sub"])
    assert np.allclose(w, g["w"], rtol=1e-9, atol=1e-300)
    assert np.all(np.abs(obj - g["obj"]) <= 1e-12 * np.sum(Y ** 2, axis=1) + 1e-9 * np.abs(obj))


@pytest.mark.gpu
def test_gpu_hcp_dict_full_782_atoms():
    """The reference's test_hcp_dict at its full size: rotate_atom of the 552 x 782 HCP
    dictionary along two directions, CSF column, [782, 782, 1] search (noiseless: recovers the
    planted atoms and fractions; noisy: same tuple and weights as the reference)."""
    from microstructure_fingerprinting_b200 import mf_utils as mfu
    h = _load("hcp_dictionary.npz")
    g = _load("full_hcp.npz")
    dic, sch_b0 = h["dic"], h["sch_b0"]
    N = dic.shape[1]
    S0 = np.repeat(h["S0_col"][:, None], N, axis=1)
    refdir = np.array([0.0, 0.0, 1.0])
    for case in range(2):
        fd = g["fd%d" % case]
        D = np.zeros((sch_b0.shape[0], 2 * N + 1))
        for k in range(2):
            D[:, k * N:(k + 1) * N] = mfu.rotate_atom(dic, sch_b0, refdir, fd[:, k], float(h["WM_DIFF"]), S0)
        D[:, -1] = h["sig_csf"]
        assert np.allclose(D[:, 0:N:16], g["rot%d_cols" % case], rtol=1e-12, atol=1e-13)
        y = g["y%d" % case]
        w, sub, tot, obj, yrec = mfu.solve_exhaustive_posweights(D, y.copy(), np.array([N, N, 1]))
        assert np.array_equal(sub, g["sub%d" % case])
        assert np.allclose(w, g["w%d" % case], rtol=1e-9, atol=1e-9 * np.max(g["w%d" % case]))
        if case == 0:
            assert sub[0] == 86 and np.allclose(w / w.sum(), g["nu0"])
