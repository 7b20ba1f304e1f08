"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the
C ABI of libmfb200.so.  Comparisons: committed reference goldens (tests/golden), the CPU
oracle on the same seeded inputs, and size-independent properties at larger sizes."""
import os

import numpy as np
import pytest

from tests.conftest import SOLVER_CASE_NAMES
from microstructure_fingerprinting_b200 import MFModel, _lib, mf_utils as mfu
from oracle import oracle as orc
from tests.phantom import compare_rows, make_dictionary, make_phantom, oracle_rows

pytestmark = pytest.mark.gpu


def test_boundary_cases_1d():
    # reference tests/integration/test_exhaustive_fingerprinting.py:38-59, verbatim data
    s2 = np.sqrt(2.0)
    A = np.array([[0.0], [1.0], [0.0]])
    Y = np.array([[1, 0, s2 / 2, 0, s2 / 2], [0, 0, -s2 / 2, 2, s2 / 2], [0, 1, 0, 0, 0]])
    w_exp, obj_exp = [0, 0, 0, 2, s2 / 2], [1, 1, 1, 0, 0.5]
    for i in range(5):
        w, sub, tot, obj, yrec = mfu.solve_exhaustive_posweights(A, Y[:, i].copy(), np.array([1]))
        assert np.isclose(w[0], w_exp[i]) and np.isclose(obj, obj_exp[i])
        assert sub.dtype == np.int32 and w.shape == (1,) and yrec.shape == (3,)


def test_boundary_cases_2d():
    # reference tests/integration/test_exhaustive_fingerprinting.py:62-89, verbatim data
    s2, s3 = np.sqrt(2.0), np.sqrt(3.0)
    A = np.array([[0.5, s3 * 0.5], [s3 * 0.5, 0.5]])
    Y = np.array([[-s3 / 2, 0.5, -1, -s3 / 2, 0.5001, 0.5, s3 / 2, s2 / 2, -s2 / 2.0],
                  [0.5, -s3 / 2, 0, 0.5001, -s3 / 2, s3 / 2, 0.5, s2 / 2, -s2 / 2.0]])
    w_exp = np.array([[0, 0], [0, 0], [0, 0], [8.66025404e-05, 0], [0, 8.66025404e-05],
                      [1, 0], [0, 1], [0.51763809, 0.51763809], [0, 0]])
    obj_exp = np.array([1, 1, 1, 1.0001000025, 1.0001000025, 0, 0, 0, 1])
    for i in range(9):
        w, sub, tot, obj, _ = mfu.solve_exhaustive_posweights(A, Y[:, i].copy(), np.array([1, 1]))
        assert np.all(np.isclose(w, w_exp[i])), i
        assert np.isclose(obj, obj_exp[i]), i


@pytest.mark.parametrize("name", [n for n in SOLVER_CASE_NAMES if n[1] in "123"])
def test_solver_bit_identical_to_reference(solver_cases, name):
    """1-3 blocks: same summation order, no FMA -> bit-identical to the Numba reference."""
    A, Y, sizes = solver_cases[name + "_A"], solver_cases[name + "_Y"], solver_cases[name + "_sizes"]
    for v in range(Y.shape[0]):
        w, sub, tot, obj, yrec = mfu.solve_exhaustive_posweights(A, Y[v].copy(), sizes)
        assert np.array_equal(sub, solver_cases[name + "_sub"][v]), (name, v)
        assert np.array_equal(w, solver_cases[name + "_w"][v]), (name, v)
        assert obj == solver_cases[name + "_obj"][v], (name, v)
        assert np.allclose(yrec, solver_cases[name + "_yrec"][v], rtol=1e-13, atol=1e-15)
    # batched call, shared dictionary: same answers
    w, sub, tot, obj, yrec = mfu.solve_exhaustive_posweights_batch(A, Y, sizes)
    assert np.array_equal(sub, solver_cases[name + "_sub"])
    assert np.array_equal(w, solver_cases[name + "_w"])
    assert np.array_equal(obj, solver_cases[name + "_obj"])


def test_reference_synthetic(ref_synth):
    # reference test_synthetic_data (:94-153): recovers the ground truth, beats the noise
    A, Y, sizes = ref_synth["A"], ref_synth["Y"], ref_synth["sizes"]
    w, sub, tot, obj, _ = mfu.solve_exhaustive_posweights_batch(A, Y, sizes)
    assert np.array_equal(tot, ref_synth["tot"])
    assert np.array_equal(tot.T, ref_synth["ID"])
    assert np.array_equal(w, ref_synth["w"])
    assert np.array_equal(obj, ref_synth["obj"])


@pytest.mark.parametrize("sizes,M,signed", [([200, 170], 100, False), ([130, 140, 1], 105, False),
                                            ([40, 30, 20], 64, True), ([300], 90, True),
                                            ([65, 1], 33, False), ([1, 7], 20, False),
                                            ([64, 64, 5], 40, False)])
def test_batch_solve_vs_oracle(sizes, M, signed):
    """Per-voxel dictionaries at sizes the oracle finishes in seconds: bit-exact."""
    rng = np.random.default_rng(sum(sizes) + M)
    V, nt = 5, int(np.sum(sizes))
    A = rng.standard_normal((V, M, nt)) if signed else rng.random((V, M, nt)) + 0.01
    st = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    Y = np.zeros((V, M))
    for v in range(V):
        gt = st + np.array([rng.integers(0, n) for n in sizes])
        Y[v] = A[v][:, gt] @ rng.random(len(sizes)) + 0.02 * rng.standard_normal(M)
    w, sub, tot, obj, yrec = mfu.solve_exhaustive_posweights_batch(A, Y, np.asarray(sizes))
    for v in range(V):
        ow, osub, otot, oobj, oyrec = orc.solve(A[v], Y[v], np.asarray(sizes))
        assert np.array_equal(sub[v], osub) and np.array_equal(tot[v], otot)
        assert np.array_equal(w[v], ow) and obj[v] == oobj
        assert np.array_equal(yrec[v], oyrec)


def test_solver_edge_cases():
    rng = np.random.default_rng(5)
    A = rng.random((12, 9)) + 0.1
    # y = 0: nothing beats the zero solution -> w = 0, indices 0, objective 0
    w, sub, tot, obj, yrec = mfu.solve_exhaustive_posweights(A, np.zeros(12), np.array([4, 5]))
    assert np.all(w == 0) and np.all(sub == 0) and obj == 0 and np.all(yrec == 0)
    # y anti-correlated with every atom
    w, sub, tot, obj, _ = mfu.solve_exhaustive_posweights(A, -np.ones(12), np.array([4, 4, 1]))
    assert np.all(w == 0) and np.all(sub == 0) and obj == 12.0
    # duplicate columns everywhere: first index in loop order wins (exact ties)
    A2 = np.tile(A[:, :1], (1, 6))
    y = 2.0 * A[:, 0]
    for sizes in ([6], [3, 3], [2, 2, 2]):
        g = mfu.solve_exhaustive_posweights(A2, y, np.array(sizes))
        o = orc.solve(A2, y, np.array(sizes))
        assert np.array_equal(g[1], o[1]) and np.array_equal(g[0], o[0]) and g[3] == o[3]
    with pytest.raises(AssertionError):
        mfu.solve_exhaustive_posweights(np.zeros((3, 2)), np.ones(3), np.array([1, 1]))
    with pytest.raises(AssertionError):
        mfu.solve_exhaustive_posweights(A, np.ones(11), np.array([4, 5]))


@pytest.mark.parametrize("mode", ["exact", "between"])
def test_rotation(ukbb, mode):
    msi = mfu.init_PGSE_multishell_interp(ukbb["dictionary"], ukbb["sch_mat"], ukbb["orientation"])
    tab = orc.init_table(ukbb["dictionary"], ukbb["sch_mat"], ukbb["orientation"])
    plan = orc.plan_scheme(tab, ukbb["sch_" + mode])
    for d, ref in zip(ukbb["dirs"], ukbb["rot_" + mode]):
        D = mfu.interp_PGSE_from_multishell(ukbb["sch_" + mode], d, msinterp=msi)
        assert D.shape == ref.shape
        assert np.allclose(D, ref, rtol=1e-12, atol=1e-15)      # unmodified reference
        assert np.array_equal(D, orc.rotate(tab, plan, d))       # oracle: bit-exact
    # batched directions + non-initialised mode
    Db = mfu.interp_PGSE_from_multishell(ukbb["sch_" + mode], ukbb["dirs"], ukbb["dictionary"],
                                         ukbb["sch_mat"], ukbb["orientation"])
    assert np.allclose(Db, ukbb["rot_" + mode], rtol=1e-12, atol=1e-15)
    with pytest.raises(ValueError, match="unit norm"):
        mfu.interp_PGSE_from_multishell(ukbb["sch_" + mode], np.array([0, 0, 2.0]), msinterp=msi)


def _ukbb_model(ukbb):
    dic = {"dictionary": ukbb["dictionary"], "sch_mat": ukbb["sch_mat"],
           "orientation": ukbb["orientation"], "num_atom": ukbb["dictionary"].shape[1],
           "num_ear": ukbb["DIFF_ear"].size, "T2_csf": float(ukbb["T2_csf"]),
           "DIFF_csf": float(ukbb["DIFF_csf"]), "T2_ear": float(ukbb["T2_ear"]),
           "DIFF_ear": ukbb["DIFF_ear"], "fasc_propnames": ["rad", "fin"],
           "rad": ukbb["rad"], "fin": ukbb["fin"]}
    return MFModel(dic)


@pytest.mark.parametrize("tag", ["A", "B", "C", "D"])
def test_mfmodel_fit_matches_reference_maps(ukbb, tag):
    """MFModel.fit end to end against maps produced by the unmodified reference."""
    numfasc, csf, ear = ukbb["numfasc"], ukbb["csf"], ukbb["ear"]
    if tag == "B":
        csf = ear = None
    elif tag == "A":
        ear = None
    elif tag == "C":
        numfasc = np.minimum(numfasc, 1)
    model = _ukbb_model(ukbb)
    fit = model.fit(ukbb["data"], ukbb["mask"], numfasc, peaks=ukbb["peaks"], bvals=ukbb["bvals"],
                    bvecs=ukbb["bvecs"], csf_mask=csf, ear_mask=ear, verbose=0)
    assert list(fit.param_names) == [str(s) for s in ukbb["fit%s_param_names" % tag]]
    ysq = np.sum(ukbb["data"] ** 2, axis=-1) / ukbb["data"].shape[-1]
    for p in fit.param_names:
        got, ref = getattr(fit, p), ukbb["fit%s_%s" % (tag, p)]
        assert got.shape == ref.shape, p
        if p == "MSE":
            assert np.all(np.abs(got - ref) <= 1e-12 * ysq + 1e-9 * np.abs(ref)), p
        elif p == "R2":
            assert np.allclose(got, ref, rtol=1e-9, atol=1e-12), p
        else:
            assert np.allclose(got, ref, rtol=1e-9, atol=1e-300), p


@pytest.mark.parametrize("seed,ear", [(1, False), (2, True)])
def test_fit_rows_bit_exact_vs_oracle(seed, ear):
    ph = make_phantom(n_atoms=80, n_vox=120, seed=seed, ear=ear)
    n0 = _lib.launch_count()
    rows = ph.gpu_rows()
    assert _lib.launch_count() > n0
    ref = oracle_rows(ph)
    compare_rows(rows, ref, ph, exact_bits=True)


def test_fit_input_modes_and_errors(ukbb):
    model = _ukbb_model(ukbb)
    mask, data = ukbb["mask"], ukbb["data"]
    nf = np.minimum(ukbb["numfasc"], 1)
    base = model.fit(data, mask, nf, peaks=ukbb["peaks"], pgse_scheme=ukbb["sch_exact"], verbose=0)
    # colat / longit input gives the same maps
    u = ukbb["peaks"][..., :3]
    cl = np.stack([np.arccos(np.clip(u[..., 2], -1, 1)), np.arctan2(u[..., 1], u[..., 0])], axis=-1)
    alt = model.fit(data, mask, nf, colat_longit=cl, pgse_scheme=ukbb["sch_exact"], verbose=0)
    assert np.array_equal(alt.rad_f0, base.rad_f0)
    assert np.allclose(alt.M0, base.M0, rtol=1e-9)
    # tensors input: principal eigenvector +-u
    T = np.zeros(mask.shape + (6,))
    lam1, lam2 = 2e-3, 1e-4
    DT = lam2 * np.eye(3) + (lam1 - lam2) * u[..., :, None] * u[..., None, :]
    T[..., 0], T[..., 1], T[..., 2] = DT[..., 0, 0], DT[..., 0, 1], DT[..., 1, 1]
    T[..., 3], T[..., 4], T[..., 5] = DT[..., 0, 2], DT[..., 1, 2], DT[..., 2, 2]
    alt = model.fit(data, mask, nf, tensors=T, pgse_scheme=ukbb["sch_exact"], verbose=0)
    assert np.array_equal(alt.rad_f0, base.rad_f0)
    with pytest.raises(ValueError, match="non-empty mask"):
        model.fit(data, np.zeros_like(mask), nf, peaks=ukbb["peaks"], pgse_scheme=ukbb["sch_exact"])
    with pytest.raises(RuntimeError):
        model.fit(data, mask, nf, pgse_scheme=ukbb["sch_exact"])
    with pytest.raises(TypeError):
        model.fit(data, mask, nf, peaks=ukbb["peaks"])
    with pytest.raises(ValueError, match="greater than"):
        model.fit(data, mask, 3, peaks=ukbb["peaks"], pgse_scheme=ukbb["sch_exact"])
    zp = ukbb["peaks"].copy()
    zp[0, 0, 0, :3] = 0
    with pytest.raises(ValueError, match="zero vector"):
        model.fit(data, mask, 1, peaks=zp, pgse_scheme=ukbb["sch_exact"])
    with pytest.raises(ValueError, match="must be explicitely passed"):
        base.write_nifti("/tmp/should_not_exist")


def test_write_nifti_roundtrip(ukbb, tmp_path):
    from microstructure_fingerprinting_b200 import nifti
    model = _ukbb_model(ukbb)
    fit = model.fit(ukbb["data"], ukbb["mask"], np.minimum(ukbb["numfasc"], 1), peaks=ukbb["peaks"],
                    pgse_scheme=ukbb["sch_exact"], csf_mask=1, verbose=0)
    names = fit.write_nifti(str(tmp_path / "sub.nii.gz"), affine=np.eye(4))
    assert len(names) == len(fit.param_names)
    for p, fn in zip(fit.param_names, names):
        assert fn.endswith("sub_%s.nii.gz" % p)
        vol, aff = nifti.load(fn)
        assert np.array_equal(vol, getattr(fit, p))
    # file-path inputs (data / mask as NIfTI) reproduce the array-input fit
    nifti.save(ukbb["data"], np.eye(4), str(tmp_path / "dwi.nii"))
    nifti.save(ukbb["mask"], np.eye(4), str(tmp_path / "mask.nii"))
    fit2 = model.fit(str(tmp_path / "dwi.nii"), str(tmp_path / "mask.nii"),
                     np.minimum(ukbb["numfasc"], 1), peaks=ukbb["peaks"],
                     pgse_scheme=ukbb["sch_exact"], csf_mask=1, verbose=0)
    assert fit2.affine is not None and np.array_equal(fit2.M0, fit.M0)


def test_full_size_properties():
    """BASELINE-size shapes (N = 1000 atoms, M = 105) on a voxel subset: properties that do
    not need the oracle at full size -- planted noiseless atoms are recovered, weights are
    non-negative, the objective equals |y - y_rec|^2, results are independent of batching."""
    ph = make_phantom(n_atoms=1000, n_vox=96, seed=9, frac_k=(0.0, 0.3, 0.7), snr=1e9)
    rows = ph.gpu_rows()
    mf = ph.maxfasc
    assert np.all(rows[:, 1:1 + mf] >= 0) and np.all(rows[:, 0] > 0)
    k2 = (ph.K == 2) & (ph.csf == 0)
    assert np.array_equal(rows[k2, 1 + mf:1 + 2 * mf], ph.atoms[k2].astype(float))
    assert np.allclose(rows[k2, 0], 800.0, rtol=1e-6)
    assert np.all(rows[:, -2] < 1e-6)
    again = ph.gpu_rows()
    assert np.array_equal(rows, again)                    # deterministic
    sub = np.arange(10, 40)
    ref = oracle_rows(ph, sub)                            # oracle on a subsample
    compare_rows(rows[sub], ref, ph, idx=sub)


@pytest.mark.parametrize("n_atoms,n_vox,snr", [(200, 1500, 30.0), (1000, 600, 30.0), (333, 800, 1e9),
                                                (64, 800, 10.0)])
def test_fast_tier_equals_exact_tier(n_atoms, n_vox, snr):
    """The DMMA screening tier only selects; the winner is re-evaluated in the reference's
    arithmetic, so rows must be bit-identical to the exact (reference-order) tier."""
    ph = make_phantom(n_atoms=n_atoms, n_vox=n_vox, seed=n_atoms, frac_k=(0.0, 0.1, 0.9),
                      csf_frac=0.4, snr=snr)
    fast = ph.gpu_rows(flags=0)
    exact = ph.gpu_rows(flags=1)
    assert np.array_equal(fast, exact)
    sub = np.arange(0, 24)
    compare_rows(fast[sub], oracle_rows(ph, sub), ph, idx=sub, exact_bits=True)


def test_fast_tier_degenerate_voxels():
    """Identical / nearly parallel peaks, single-fascicle data in a 2-fascicle voxel, zero
    signal: the screening tier must hand these to the exact tier or agree with it."""
    ph = make_phantom(n_atoms=120, n_vox=64, seed=77, frac_k=(0.0, 0.0, 1.0), csf_frac=0.5)
    ph.peaks[0:8, 3:6] = ph.peaks[0:8, 0:3]                       # identical peaks
    ph.peaks[8:16, 3:6] = ph.peaks[8:16, 0:3] + 1e-9              # nearly identical
    ph.peaks[8:16, 3:6] /= np.linalg.norm(ph.peaks[8:16, 3:6], axis=1, keepdims=True)
    ph.Y[16:20] = 0.0                                             # zero signal
    ph.Y[20:24] = -np.abs(ph.Y[20:24])                            # nothing beats w = 0
    from tests.phantom import rotate_columns
    ph.Y[24:32] = 500.0 * rotate_columns(ph.dic, ph.sch, ph.peaks[24:32, :3], ph.atoms[24:32, 0])
    fast = ph.gpu_rows(flags=0)
    exact = ph.gpu_rows(flags=1)
    assert np.array_equal(fast, exact)
    compare_rows(fast, oracle_rows(ph), ph, exact_bits=True)


def test_fast_tier_large_validation():
    """20k voxels at the benchmark shape (N = 1000, M = 105): the screening tier and the
    reference-order tier must produce bit-identical rows; only a small fraction of voxels
    may need the exact tier."""
    from microstructure_fingerprinting_b200 import mf_utils as mfu
    ph = make_phantom(n_atoms=1000, n_vox=20000, seed=2024, frac_k=(0.0, 0.0, 1.0), csf_frac=0.3)
    msi = mfu.init_PGSE_multishell_interp(ph.dic["dictionary"], ph.dic["sch_mat"], ph.dic["orientation"])
    plan = mfu.GpuPlan(msi, mfu.SchemePlan(msi, ph.sch), ph.sig_csf, None)
    fast = plan.fit_host(ph.Y, ph.peaks, ph.K, ph.csf, None, 2, True, False, flags=0)
    st = plan.stats()
    exact = plan.fit_host(ph.Y, ph.peaks, ph.K, ph.csf, None, 2, True, False, flags=1)
    plan.close()
    assert np.array_equal(fast, exact)
    assert st[0] + st[1] == 20000 and st[1] < 0.02 * 20000, st


@pytest.mark.parametrize("name", [n for n in SOLVER_CASE_NAMES if n[1] in "45"])
def test_solver_4up_matches_reference(solver_cases, name):
    """4-5 blocks: the reference runs scipy.optimize.nnls per tuple; the GPU enumerates the
    supports in closed form -> same tuple, weights and objective to rounding (1e-9)."""
    A, Y, sizes = solver_cases[name + "_A"], solver_cases[name + "_Y"], solver_cases[name + "_sizes"]
    ysq = np.sum(Y ** 2, axis=1)
    for v in range(Y.shape[0]):
        w, sub, tot, obj, yrec = mfu.solve_exhaustive_posweights(A, Y[v].copy(), sizes)
        assert sub.dtype == np.int64
        rsub, rw = solver_cases[name + "_sub"][v], solver_cases[name + "_w"][v]
        # a block whose optimal weight is 0 leaves its index undetermined (every atom of that
        # block ties to rounding in the reference): indices must agree wherever weight > 0
        active = (rw > 1e-9 * rw.max()) | (w > 1e-9 * max(w.max(), 1e-300))
        assert np.array_equal(sub[active], rsub[active]), (name, v)
        assert np.allclose(w, rw, rtol=1e-9, atol=1e-9 * rw.max()), (name, v)
        assert abs(obj - solver_cases[name + "_obj"][v]) <= 1e-12 * ysq[v] + 1e-9 * abs(obj)
        assert np.allclose(yrec, solver_cases[name + "_yrec"][v], rtol=1e-9, atol=1e-9 * np.sqrt(ysq[v]))


def test_fit_parallel_multi_gpu(ukbb):
    """parallel=True shards the ROI over all visible GPUs (contiguous spans, no collective):
    the maps must be identical to the single-GPU fit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    ph = make_phantom(n_atoms=200, n_vox=3001, seed=31, csf_frac=0.3)
    dic = ph.dic
    model = MFModel(dic)
    shape = (3001,)
    one = model.fit(ph.Y, np.ones(shape), ph.K.astype(float), peaks=ph.peaks, pgse_scheme=ph.sch,
                    csf_mask=ph.csf.astype(float), verbose=0, parallel=False)
    many = model.fit(ph.Y, np.ones(shape), ph.K.astype(float), peaks=ph.peaks, pgse_scheme=ph.sch,
                     csf_mask=ph.csf.astype(float), verbose=0, parallel=True)
    for p in one.param_names:
        assert np.array_equal(getattr(one, p), getattr(many, p)), p


def test_fit_sharded_two_plans_one_gpu():
    """The multi-GPU path of MFModel.fit (one plan and one host thread per shard,
    cost-weighted contiguous bounds, disjoint writes into the params array) exercised on ONE
    device: devices=[0, 0, 0] runs three plans concurrently on GPU 0.  Mixed K / CSF / EAR
    voxels make the cost-weighted shards uneven.  Rows must equal the single-plan fit."""
    ph = make_phantom(n_atoms=120, n_vox=2700, seed=37, csf_frac=0.3, ear=True)
    model = MFModel(ph.dic)
    rng = np.random.default_rng(1)
    mask = (rng.random((20, 25, 10)) < 0.5).astype(float)
    V = int(mask.sum())
    in_mask = mask > 0

    def vol(a, trailing=()):
        out = np.zeros(mask.shape + trailing)
        out[in_mask] = a[:V]
        return out
    kw = dict(peaks=vol(ph.peaks, (ph.peaks.shape[1],)), pgse_scheme=ph.sch, csf_mask=vol(ph.csf.astype(float)),
              ear_mask=vol(ph.ear.astype(float)), verbose=0)
    data = vol(ph.Y, (ph.Y.shape[1],))
    one = model.fit(data, mask, vol(ph.K.astype(float)), devices=[0], **kw)
    many = model.fit(data, mask, vol(ph.K.astype(float)), devices=[0, 0, 0], **kw)
    exact = model.fit(data, mask, vol(ph.K.astype(float)), devices=[0, 0], exact=True, **kw)
    for p in one.param_names:
        assert np.array_equal(getattr(one, p), getattr(many, p)), p
        assert np.array_equal(getattr(one, p), getattr(exact, p)), p
    # and against the oracle on a subsample of the ROI
    idx = np.arange(0, V, 97)
    ref = oracle_rows(ph, idx)
    assert np.allclose(one.M0[in_mask][idx], ref[:, 0], rtol=1e-9)
    assert np.array_equal(one.frac_f0[in_mask][idx] > 0, ref[:, 1] > 0)
    model.close()


@pytest.mark.parametrize("layout", ["c64", "c32", "fortran", "strided", "int16"])
def test_fit_volume_layouts(layout):
    """mfb_fit_volume gathers the ROI voxels from the caller's volume whatever its layout:
    C-contiguous float64 / float32, Fortran-ordered (what a NIfTI file yields), a strided view,
    an integer volume (converted once on the host).  Maps equal those of a contiguous float64
    copy of the same values."""
    ph = make_phantom(n_atoms=64, n_vox=1000, seed=41, csf_frac=0.3)
    model = MFModel(ph.dic)
    rng = np.random.default_rng(2)
    mask = (rng.random((12, 10, 15)) < 0.5).astype(float)
    V = int(mask.sum())
    in_mask = mask > 0
    M = ph.Y.shape[1]
    base = np.zeros(mask.shape + (M,))
    Y = ph.Y[:V]
    if layout == "c32":
        Y = Y.astype(np.float32).astype(np.float64)
    if layout == "int16":
        Y = np.round(Y)
    base[in_mask] = Y
    if layout == "c64":
        data = base
    elif layout == "c32":
        data = base.astype(np.float32)
    elif layout == "fortran":
        data = np.asfortranarray(base)
    elif layout == "strided":
        big = np.zeros((12, 10, 15, 2 * M + 3))
        big[..., 1:2 * M:2] = base
        data = big[..., 1:2 * M:2]
        assert not data.flags.c_contiguous
    else:
        data = base.astype(np.int16)

    def vol(a, trailing=()):
        out = np.zeros(mask.shape + trailing)
        out[in_mask] = a[:V]
        return out
    kw = dict(peaks=vol(ph.peaks, (ph.peaks.shape[1],)), pgse_scheme=ph.sch, csf_mask=vol(ph.csf.astype(float)),
              verbose=0)
    ref = model.fit(np.ascontiguousarray(base), mask, vol(ph.K.astype(float)), **kw)
    got = model.fit(data, mask, vol(ph.K.astype(float)), **kw)
    for p in ref.param_names:
        assert np.array_equal(getattr(ref, p), getattr(got, p)), p
    # ROI rows of the plain C-ABI call with a dense (V, M) host array
    rows = ph.gpu_rows()[:V] if layout == "c64" else None
    if rows is not None:
        assert np.array_equal(ref.M0[in_mask], rows[:, 0])
    model.close()


def test_entry_points_restore_the_current_device():
    """Every C-ABI entry point runs on its plan's device and puts the caller's current device
    back (a library call must not move the host thread to another GPU)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    ph = make_phantom(n_atoms=64, n_vox=64, seed=43)
    torch.cuda.set_device(0)
    rows1 = ph.gpu_rows(device=1)
    assert torch.cuda.current_device() == 0
    assert np.array_equal(rows1, ph.gpu_rows(device=0))


@pytest.fixture(scope="module")
def lowlevel():
    import os
    from tests.conftest import GOLDEN
    return np.load(os.path.join(GOLDEN, "lowlevel_rotation.npz"))


def test_lerp_rows_kernel_vs_oracle():
    rng = np.random.default_rng(3)
    R, N, V, M = 57, 33, 4, 19
    table = rng.standard_normal((R, N))
    rl, rh = rng.integers(0, R, (V, M)), rng.integers(0, R, (V, M))
    wl, wh, sc = rng.standard_normal((V, M)), rng.standard_normal((V, M)), rng.random((V, M))
    got = mfu._lerp_rows(table, rl, rh, wl, wh, sc)
    assert np.array_equal(got, orc.lerp_rows(table, rl, rh, wl, wh, sc))
    got = mfu._lerp_rows(table, rl, rh, wl, wh)
    assert np.array_equal(got, orc.lerp_rows(table, rl, rh, wl, wh))


def test_rotate_atom_matches_reference(lowlevel):
    """mfu.rotate_atom on a subset of the reference's HCP Monte-Carlo dictionary."""
    g = lowlevel
    ref = np.array([0.0, 0.0, 1.0])
    for d, want in zip(g["hcp_dirs"], g["hcp_rot"]):
        got = mfu.rotate_atom(g["hcp_sig"], g["hcp_sch"], ref, d, float(g["hcp_DIFF"]), g["hcp_S0"],
                              warnings=False)
        assert got.shape == want.shape
        assert np.allclose(got, want, rtol=1e-13, atol=0)
    got1 = mfu.rotate_atom(g["hcp_sig"][:, 2], g["hcp_sch"], ref, g["hcp_dirs"][1], float(g["hcp_DIFF"]),
                           g["hcp_S0"][:, 2], warnings=False)
    assert got1.shape == g["hcp_rot1d"].shape and np.allclose(got1, g["hcp_rot1d"], rtol=1e-13, atol=0)
    batch = mfu.rotate_atom(g["hcp_sig"], g["hcp_sch"], ref, g["hcp_dirs"], float(g["hcp_DIFF"]),
                            g["hcp_S0"], warnings=False)
    assert np.allclose(batch, g["hcp_rot"], rtol=1e-13, atol=0)
    with pytest.raises(ValueError):
        mfu.rotate_atom(g["hcp_sig"], g["hcp_sch"][:-1], ref, ref, 2e-9, g["hcp_S0"])


def test_hcp_pipeline(lowlevel):
    """The reference's test_hcp_dict (tests/integration/test_exhaustive_fingerprinting.py:163-249)
    on a dictionary subset: rotate_atom x2 + CSF column + 3-block exhaustive solve, noiseless."""
    g = lowlevel
    ref = np.array([0.0, 0.0, 1.0])
    n = g["hcp_sig"].shape[1]
    D = np.zeros((g["hcp_sch"].shape[0], 2 * n + 1))
    for k in range(2):
        D[:, k * n:(k + 1) * n] = mfu.rotate_atom(g["hcp_sig"], g["hcp_sch"], ref, g["hcp_dirs"][1 + k],
                                                  float(g["hcp_DIFF"]), g["hcp_S0"], warnings=False)
    D[:, -1] = g["hcp_sig_csf"]
    w, sub, tot, obj, yrec = mfu.solve_exhaustive_posweights(D, g["hcp_y"].copy(), np.array([n, n, 1]))
    i_gt = int(g["hcp_igt"])
    assert np.all(sub[:2] == i_gt) and np.array_equal(sub, g["hcp_sub"])
    assert np.allclose(w / w.sum(), [0.5, 0.3, 0.2])
    assert np.allclose(w, g["hcp_w"], rtol=1e-9)


def test_rotate_atom_2d_matches_reference(lowlevel):
    g = lowlevel
    ref = np.array([0.0, 0.0, 1.0])
    for d, want in zip(g["ax_dirs"], g["ax_rot"]):
        got = mfu.rotate_atom_2Dprotocol(g["ax_sig"], g["ax_sch"], ref, d, float(g["ax_DIFF"]))
        assert got.shape == want.shape
        assert np.allclose(got, want, rtol=1e-12, atol=1e-300)
    got1 = mfu.rotate_atom_2Dprotocol(g["ax_sig"][:, 1], g["ax_sch"], ref, g["ax_dirs"][2],
                                      float(g["ax_DIFF"]))
    assert got1.shape == g["ax_rot1d"].shape and np.allclose(got1, g["ax_rot1d"], rtol=1e-12)
    bad = g["ax_sch"].copy()
    bad[5, 2] = 0.1
    with pytest.raises(ValueError, match="zeros for gz"):
        mfu.rotate_atom_2Dprotocol(g["ax_sig"], bad, ref, ref, 2e-9)


@pytest.mark.parametrize("sizes", [[800, 800], [800, 800, 1], [300, 520]])
def test_solve_batch_fast_equals_exact(sizes):
    """BASELINE config 2 shape (M = 100, ~800 atoms per fascicle, explicit per-voxel
    dictionaries): the DMMA screening path of mfb_solve_batch must return exactly what the
    reference-order search returns."""
    rng = np.random.default_rng(sum(sizes))
    V, M, nt = 48, 100, int(np.sum(sizes))
    base = rng.random((M, nt)) * np.exp(-3.0 * rng.random((1, nt)) * np.linspace(0, 1, M)[:, None])
    A = base[None] * (1.0 + 0.05 * rng.standard_normal((V, M, nt)))
    st = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    Y = np.stack([A[v][:, st + np.array([rng.integers(0, n) for n in sizes])] @ rng.random(len(sizes))
                  for v in range(V)])
    Y += 0.02 * rng.standard_normal(Y.shape)
    fast = mfu.solve_exhaustive_posweights_batch(A, Y, np.asarray(sizes))
    exact = mfu.solve_exhaustive_posweights_batch(A, Y, np.asarray(sizes), exact=True)
    for f, e in zip(fast, exact):
        assert np.array_equal(f, e)
    # shared dictionary (strideA = 0)
    fast = mfu.solve_exhaustive_posweights_batch(A[0], Y, np.asarray(sizes))
    exact = mfu.solve_exhaustive_posweights_batch(A[0], Y, np.asarray(sizes), exact=True)
    for f, e in zip(fast, exact):
        assert np.array_equal(f, e)


@pytest.mark.parametrize("sizes,M", [([300, 260], 150), ([200, 200, 1], 130), ([129, 65], 552),
                                     ([64, 300, 1], 257)])
def test_solve_batch_general_M_equals_exact(sizes, M):
    """M > 112 (the i1 tile no longer fits in shared memory): the TMA-fed k-chunked DMMA
    screening path (k_normalize + k_gemm_pairs) must return exactly what the
    reference-order search returns, for per-voxel and for shared dictionaries."""
    rng = np.random.default_rng(sum(sizes) + M)
    V, nt = 24, int(np.sum(sizes))
    base = rng.random((M, nt)) * np.exp(-3.0 * rng.random((1, nt)) * np.linspace(0, 1, M)[:, None])
    A = base[None] * (1.0 + 0.05 * rng.standard_normal((V, M, nt)))
    st = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    Y = np.stack([A[v][:, st + np.array([rng.integers(0, n) for n in sizes])] @ rng.random(len(sizes))
                  for v in range(V)])
    Y += 0.02 * rng.standard_normal(Y.shape)
    for dic in (A, A[0]):
        fast = mfu.solve_exhaustive_posweights_batch(dic, Y, np.asarray(sizes))
        exact = mfu.solve_exhaustive_posweights_batch(dic, Y, np.asarray(sizes), exact=True)
        for f, e in zip(fast, exact):
            assert np.array_equal(f, e)
    # and the single-voxel wrapper against the CPU oracle
    for v in range(3):
        w, sub, tot, obj, yrec = mfu.solve_exhaustive_posweights(A[v], Y[v].copy(), np.asarray(sizes))
        wo, subo, toto, objo, _ = orc.solve(A[v], Y[v], sizes)
        assert np.array_equal(sub, subo) and np.array_equal(w, wo) and obj == objo


@pytest.mark.parametrize("scheme,n_atoms,n_vox", [("between", 200, 1200), ("dense", 150, 500),
                                                  ("between", 1000, 300)])
def test_fit_materialised_fast_tier_equals_exact_tier(scheme, n_atoms, n_vox):
    """Between-shell protocols (gradient strengths that match no dense shell,
    mf_utils.py:1921-1956) and protocols with M > 112 run the screening tier on materialised
    dictionaries: rows must be bit-identical to the reference-order tier and to the oracle."""
    ph = make_phantom(n_atoms=n_atoms, n_vox=n_vox, seed=n_atoms + 5, frac_k=(0.05, 0.15, 0.8),
                      csf_frac=0.4, scheme=scheme)
    msi = mfu.init_PGSE_multishell_interp(ph.dic["dictionary"], ph.dic["sch_mat"], ph.dic["orientation"])
    plan = mfu.GpuPlan(msi, mfu.SchemePlan(msi, ph.sch), ph.sig_csf, None)
    fast = plan.fit_host(ph.Y, ph.peaks, ph.K, ph.csf, None, 2, True, False, flags=0)
    st = plan.stats()
    exact = plan.fit_host(ph.Y, ph.peaks, ph.K, ph.csf, None, 2, True, False, flags=1)
    plan.close()
    assert np.array_equal(fast, exact)
    n2 = int(np.sum(ph.K == 2))
    assert st[0] > 0.9 * n2, st          # the screening tier decided almost every 2-fascicle voxel
    sub = np.arange(0, 16)
    compare_rows(fast[sub], oracle_rows(ph, sub), ph, idx=sub, exact_bits=True)


def _triple_problem(sizes, M, V, seed, planted=3, signed=False):
    rng = np.random.default_rng(seed)
    nt = int(np.sum(sizes))
    base = rng.random((M, nt)) * np.exp(-3.0 * rng.random((1, nt)) * np.linspace(0, 1, M)[:, None])
    if signed:
        base = base * rng.choice([-1.0, 1.0], size=(M, nt))
    A = base[None] * (1.0 + 0.05 * rng.standard_normal((V, M, nt)))
    st = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    wts = rng.uniform(0.2, 1.0, (V, 3))
    wts[:, planted:] = 0.0
    Y = np.stack([A[v][:, st + np.array([rng.integers(0, n) for n in sizes])] @ wts[v] for v in range(V)])
    Y += 0.02 * rng.standard_normal(Y.shape)
    return A, Y


@pytest.mark.parametrize("sizes,M", [([40, 36, 1], 60), ([30, 28, 26], 60), ([140, 120, 1], 130)])
@pytest.mark.parametrize("scale", [1e-4, 1e-3])
def test_solve_batch_tiny_magnitude_follows_reference_tolerance(sizes, M, scale):
    """The reference's three-block solver accepts Cramer numerators D_i >= -tol with an ABSOLUTE
    tol = 2.2204e-14 (mf_utils.py:562): when A and y are scaled down so that |a|^4 |a.y|
    approaches tol it returns solutions with negative weights.  The screening tiers must hand
    such voxels to the reference-order tier: results equal the exact tier and the CPU oracle."""
    A, Y = _triple_problem(sizes, M, 10, 1000 + sum(sizes), planted=2, signed=True)
    A, Y = A * scale, Y * scale
    fast = mfu.solve_exhaustive_posweights_batch(A, Y, np.asarray(sizes))
    exact = mfu.solve_exhaustive_posweights_batch(A, Y, np.asarray(sizes), exact=True)
    for f, e in zip(fast, exact):
        assert np.array_equal(f, e)
    for v in range(4):
        wo, subo, toto, objo, _ = orc.solve(A[v], Y[v], sizes)
        assert np.array_equal(fast[1][v], subo) and np.array_equal(fast[0][v], wo) and fast[3][v] == objo


@pytest.mark.parametrize("sizes,M,planted,signed", [([60, 70, 50], 100, 3, False), ([300, 300, 300], 100, 3, False),
                                                    ([90, 90, 6], 105, 3, False), ([33, 47, 129], 150, 3, False),
                                                    ([64, 64, 64], 60, 2, False), ([50, 40, 30], 80, 3, True),
                                                    # the scan streams its largest block: every position of a short one
                                                    ([6, 90, 90], 105, 3, False), ([90, 6, 90], 105, 3, False),
                                                    ([5, 200, 37], 64, 3, False), ([150, 9, 11], 100, 2, False)])
def test_solve_batch_triple_scan_equals_exact(sizes, M, planted, signed):
    """Three searched blocks (reference `_3`, mf_utils.py:470-607; BASELINE config 4 shape
    [300, 300, 300]): the DMMA + FP64 triple scan must return exactly what the
    reference-order search returns, and decide most 3-compartment voxels itself."""
    V = 12 if sizes[0] >= 300 else 40
    A, Y = _triple_problem(sizes, M, V, sum(sizes) + M, planted, signed)
    _lib.solve_stats(reset=True)
    fast = mfu.solve_exhaustive_posweights_batch(A, Y, np.asarray(sizes))
    stats = _lib.solve_stats(reset=True)
    exact = mfu.solve_exhaustive_posweights_batch(A, Y, np.asarray(sizes), exact=True)
    for f, e in zip(fast, exact):
        assert np.array_equal(f, e)
    assert stats[0] + stats[1] == V
    if planted == 3 and not signed:
        assert stats[0] >= 0.8 * V, stats
    fast = mfu.solve_exhaustive_posweights_batch(A[0], Y, np.asarray(sizes))     # shared dictionary
    exact = mfu.solve_exhaustive_posweights_batch(A[0], Y, np.asarray(sizes), exact=True)
    for f, e in zip(fast, exact):
        assert np.array_equal(f, e)
    if sizes[0] < 100:
        for v in range(3):
            w, sub, tot, obj, yrec = mfu.solve_exhaustive_posweights(A[v], Y[v].copy(), np.asarray(sizes))
            wo, subo, toto, objo, _ = orc.solve(A[v], Y[v], sizes)
            assert np.array_equal(sub, subo) and np.array_equal(w, wo) and obj == objo


def test_fit_two_fascicles_plus_ear_triple_scan():
    """[N, N, E] voxels (two fascicles + the extra-axonal restricted block, mf.py:398-408)
    run the triple scan on materialised dictionaries; [N, N, 1, E] voxels stay on the
    support-enumeration search.  Rows must equal the reference-order tier's and the oracle's."""
    ph = make_phantom(n_atoms=90, n_vox=400, seed=31, frac_k=(0.05, 0.15, 0.8), csf_frac=0.25,
                      ear=True, n_ear=6, ear_frac=0.6, ear_max_k=2)
    msi = mfu.init_PGSE_multishell_interp(ph.dic["dictionary"], ph.dic["sch_mat"], ph.dic["orientation"])
    plan = mfu.GpuPlan(msi, mfu.SchemePlan(msi, ph.sch), ph.sig_csf, ph.sig_ear)
    fast = plan.fit_host(ph.Y, ph.peaks, ph.K, ph.csf, ph.ear, 2, True, True, flags=0)
    st = plan.stats()
    exact = plan.fit_host(ph.Y, ph.peaks, ph.K, ph.csf, ph.ear, 2, True, True, flags=1)
    plan.close()
    assert np.array_equal(fast, exact)
    n_nne = int(np.sum((ph.K == 2) & (ph.ear == 1) & (ph.csf == 0)))
    assert n_nne > 50 and st[0] > 0, (n_nne, st)
    sel = np.where((ph.K == 2) & (ph.ear == 1) & (ph.csf == 0))[0][:12]
    compare_rows(fast[sel], oracle_rows(ph, sel), ph, idx=sel, exact_bits=True)


@pytest.mark.parametrize("n_atoms,n_ear,seed", [(90, 6, 33), (200, 10, 35)])
def test_fit_two_fascicles_csf_ear_projected_triple_scan(n_atoms, n_ear, seed):
    """[N, N, 1, E] voxels (two fascicles + CSF + EAR, mf.py:398-419 -> reference `_4up`): the
    triple scan runs on the blocks projected off the CSF column, tuples that pass the
    unconstrained-gain test are solved exactly by support enumeration.  Rows must equal the
    reference-order tier's bit for bit and agree with the oracle (scipy.optimize.nnls per tuple)
    to 1e-9; most voxels whose EAR compartment is active must be decided by the fast tier."""
    ph = make_phantom(n_atoms=n_atoms, n_vox=240, seed=seed, frac_k=(0.0, 0.0, 1.0), csf_frac=1.0,
                      ear=True, n_ear=n_ear, ear_frac=1.0, ear_max_k=2)
    msi = mfu.init_PGSE_multishell_interp(ph.dic["dictionary"], ph.dic["sch_mat"], ph.dic["orientation"])
    plan = mfu.GpuPlan(msi, mfu.SchemePlan(msi, ph.sch), ph.sig_csf, ph.sig_ear)
    fast = plan.fit_host(ph.Y, ph.peaks, ph.K, ph.csf, ph.ear, 2, True, True, flags=0)
    st = plan.stats()
    exact = plan.fit_host(ph.Y, ph.peaks, ph.K, ph.csf, ph.ear, 2, True, True, flags=1)
    plan.close()
    assert np.array_equal(fast, exact)
    assert st[0] + st[1] == 240 and st[0] >= 120, st      # fast-tier share (EAR weight > 0 in every voxel)
    sel = np.arange(0, 240, 24)
    ref = oracle_rows(ph, sel)
    got = fast[sel]
    # 4 blocks: to rounding, indices wherever the weight is positive (the reference's per-tuple
    # scipy NNLS leaves the index of a zero-weight block undetermined)
    assert np.allclose(got[:, 0], ref[:, 0], rtol=1e-9)
    for c in (1, 2, 5, 6):
        assert np.allclose(got[:, c], ref[:, c], rtol=1e-9, atol=1e-12), c
    for c_w, c_id in ((1, 3), (2, 4), (6, 7)):
        act = (ref[:, c_w] > 1e-9) | (got[:, c_w] > 1e-9)
        assert np.array_equal(got[act, c_id], ref[act, c_id]), c_id


@pytest.mark.parametrize("csf_frac", [0.0, 1.0])
def test_fit_ear_flagged_but_inactive(csf_frac):
    """Voxels flagged for the EAR compartment whose signal holds none: the non-negative optimum
    puts no weight on the EAR block, the winner is the best pair of fascicle atoms (+ CSF).  The
    triple scan certifies it from its pair job (third index 0, the first tuple of the reference's
    loop order that contains the pair) instead of handing the voxel to the reference-order tier.
    Rows must equal the reference-order tier's; for [N, N, E] also the oracle's, bit for bit."""
    ph = make_phantom(n_atoms=120, n_vox=200, seed=51, frac_k=(0.0, 0.0, 1.0), csf_frac=csf_frac,
                      ear=True, n_ear=5, ear_frac=0.5, ear_max_k=2)
    ear_all = np.ones_like(ph.ear)                   # half of the voxels have no EAR signal
    msi = mfu.init_PGSE_multishell_interp(ph.dic["dictionary"], ph.dic["sch_mat"], ph.dic["orientation"])
    plan = mfu.GpuPlan(msi, mfu.SchemePlan(msi, ph.sch), ph.sig_csf, ph.sig_ear)
    csf_on = csf_frac > 0
    fast = plan.fit_host(ph.Y, ph.peaks, ph.K, ph.csf, ear_all, 2, csf_on, True, flags=0)
    st = plan.stats()
    exact = plan.fit_host(ph.Y, ph.peaks, ph.K, ph.csf, ear_all, 2, csf_on, True, flags=1)
    plan.close()
    assert np.array_equal(fast, exact)
    c_ear = 2 * 2 + int(csf_on) + 1
    n_inactive = int(np.sum(fast[:, c_ear] == 0))
    assert n_inactive >= 40, n_inactive
    assert st[0] >= 150, st                          # most voxels, active or not, decided by the fast tier
    if not csf_on:
        ph.ear, ph.ear_on, ph.csf_on = ear_all, True, False
        sel = np.where(fast[:, c_ear] == 0)[0][:8]
        compare_rows(fast[sel], oracle_rows(ph, sel), ph, idx=sel, exact_bits=True)


@pytest.fixture(scope="module")
def mc_cases():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mc_cases.npz"))


def test_monte_carlo_average_matches_reference(mc_cases):
    """mfb_mc_average against the unmodified Numba kernel (mf_utils.py:2758-2812) and the
    CPU oracle.  Tolerance 1e-12 absolute on signals in [-1, 1]: the GPU sums the spins in
    a tree and its cos() differs from libm's in the last ulp."""
    g = mc_cases
    n_spin = int(g["n_spin"])
    for dim in (2, 3):
        for ds in (1.0, 0.73):
            got = mfu.monte_carlo_average(np.ascontiguousarray(g["phases"][:, :dim]), g["pick"].astype(np.int64),
                                          np.ascontiguousarray(g["gsc"][:, :dim]), ds, n_spin)
            assert got.shape == (g["pick"].size,) and got.dtype == np.float64
            assert np.allclose(got, g["avg_d%d_s%g" % (dim, ds)], rtol=0, atol=1e-12)
            assert np.allclose(got, orc.mc_average(g["phases"][:, :dim], g["pick"], g["gsc"][:, :dim], ds, n_spin),
                               rtol=0, atol=1e-12)
    # larger problem, single component: exact value for a constant phase, average of cos
    ph = np.full((3 * 50000, 1), 0.5)
    got = mfu.monte_carlo_average(ph, np.array([0, 2, 1], dtype=np.int64), np.array([[0.0], [1.0], [2.0]]), 1.0, 50000)
    assert np.allclose(got, np.cos([0.0, 0.5, 1.0]), rtol=0, atol=1e-13)
    with pytest.raises(IndexError):
        mfu.monte_carlo_average(ph, np.array([3], dtype=np.int64), np.array([[1.0]]), 1.0, 50000)


def test_get_PGSE_from_phases_matches_reference(mc_cases, tmp_path):
    """File interface (mf_utils.py:2815-3015): phase files written like the simulator does
    (big-endian double, little-endian single), (Delta, delta) mapping, gradient scaling,
    diffusivity rescaling; error behaviour of the reference."""
    g = mc_cases
    for i, nm in enumerate("xyz"):
        g["phases"][:, i].astype(">f8").tofile(str(tmp_path / ("sub_phase_%s.bdouble" % nm)))
        g["phases"][:, i].astype("<f4").tofile(str(tmp_path / ("sub_phase_%s.lfloat" % nm)))
    fx = str(tmp_path / "sub_phase_x.bdouble")
    assert np.allclose(mfu.get_PGSE_from_phases(fx, g["sim"], g["new"]), g["file_bdouble_d3"], rtol=0, atol=1e-12)
    assert np.allclose(mfu.get_PGSE_from_phases(fx, g["sim"], g["new"], dim=3, D_sim=2.0e-9, D=1.1e-9),
                       g["file_bdouble_d3_D"], rtol=0, atol=1e-12)
    assert np.allclose(mfu.get_PGSE_from_phases(str(tmp_path / "sub_phase_x.lfloat"), g["sim"], g["new2"], dim=2),
                       g["file_lfloat_d2"], rtol=0, atol=1e-12)
    with pytest.raises(NameError):
        mfu.get_PGSE_from_phases(fx, g["sim"], g["new"], D=1e-9)
    with pytest.raises(ValueError, match="dim should be"):
        mfu.get_PGSE_from_phases(fx, g["sim"], g["new"], dim=4)
    bad = g["new"].copy()
    bad[3, 4] = 0.0333          # a Delta no simulated sequence used (TE stays >= Delta + delta)
    with pytest.raises(ValueError, match="not used to simulate"):
        mfu.get_PGSE_from_phases(fx, g["sim"], bad)
    with pytest.raises(RuntimeError, match="does not exist"):
        mfu.get_PGSE_from_phases(str(tmp_path / "nope_phase_x.bdouble"), g["sim"], g["new"])
    with pytest.raises(ValueError, match="not supported"):
        (tmp_path / "sub_phase_x.bint").write_bytes(b"0" * 24)
        mfu.get_PGSE_from_phases(str(tmp_path / "sub_phase_x.bint"), g["sim"], g["new"])


def test_rotate_atom_2d_batched_and_pipeline(lowlevel):
    """Batched rotate_atom_2Dprotocol (one vectorised host plan, one launch) equals the
    per-direction calls, and the chunked AxCaliber pipeline (per-direction decisions on the host in a
    worker thread, plans expanded, dictionaries assembled and searched on the GPU) agrees with the
    hand-written per-voxel sequence rotate_atom_2Dprotocol x 2 + solve_exhaustive_posweights:
    indices exact, weights to 1e-12."""
    g = lowlevel
    ref = np.array([0.0, 0.0, 1.0])
    sig, sch, DIFF = g["ax_sig"], g["ax_sch"], float(g["ax_DIFF"])
    batch = mfu.rotate_atom_2Dprotocol(sig, sch, ref, g["ax_dirs"], DIFF)
    assert batch.shape == (g["ax_dirs"].shape[0],) + sig.shape
    for v, want in enumerate(g["ax_rot"]):
        assert np.allclose(batch[v], want, rtol=1e-12, atol=1e-300)
        assert np.array_equal(batch[v], mfu.rotate_atom_2Dprotocol(sig, sch, ref, g["ax_dirs"][v], DIFF))
    # a fascicle in the gradient plane projects both gradient lines onto one: reference AssertionError
    with pytest.raises(AssertionError, match="pairs of opposite directions"):
        mfu.rotate_atom_2Dprotocol(sig, sch, ref, np.array([np.sqrt(0.5), np.sqrt(0.5), 0.0]), DIFF)

    # pipeline on a richer dictionary: scaled / mixed copies of the fixture atoms
    rng = np.random.default_rng(12)
    N = 24
    mix = rng.random((sig.shape[1], N)) + 0.05
    dic = sig @ (mix / mix.sum(axis=0, keepdims=True))
    V = 20
    peaks = rng.standard_normal((V, 2, 3))
    peaks[:, :, 2] += np.sign(peaks[:, :, 2]) * 1.0            # away from the gradient plane
    peaks /= np.linalg.norm(peaks, axis=2, keepdims=True)
    peaks[3, 1] = [np.sqrt(0.5), np.sqrt(0.5), 0.0]            # breaks the protocol's assumptions
    Y = np.zeros((V, sig.shape[0]))
    truth = rng.integers(0, N, (V, 2))
    for v in range(V):
        if v == 3:
            continue
        D = [mfu.rotate_atom_2Dprotocol(dic, sch, ref, peaks[v, k], DIFF) for k in range(2)]
        Y[v] = 0.6 * D[0][:, truth[v, 0]] + 0.4 * D[1][:, truth[v, 1]] + 1e-3 * rng.standard_normal(sig.shape[0])
    w, sub, obj, ok = mfu.solve_rotated_2Dprotocol_batch(dic, sch, ref, peaks, Y, DIFF, chunk=7)
    assert not ok[3] and ok.sum() == V - 1
    for v in (0, 1, 5, 19):
        D = np.hstack([mfu.rotate_atom_2Dprotocol(dic, sch, ref, peaks[v, k], DIFF) for k in range(2)])
        w1, sub1, tot1, obj1, _ = mfu.solve_exhaustive_posweights(D, Y[v].copy(), np.array([N, N]))
        # (the pipeline expands its plans on the GPU: the parallel-signal factor goes through the device's
        # exp(), the dictionaries can differ from the host path's in the last bit)
        assert np.array_equal(sub[v], sub1) and np.allclose(w[v], w1, rtol=1e-12, atol=0.0)
        assert abs(obj[v] - obj1) <= 1e-12 * float(Y[v] @ Y[v])
        assert np.array_equal(sub[v], truth[v])


def test_solve_batch_fuzz_fast_equals_exact():
    """Random shapes (pair scan M <= 112, general-M pair scan, triple scan; signed data, shared
    dictionaries, zero signals, zero planted weights): screening tier == reference-order tier."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "fuzz_solve_batch.py")
    spec = importlib.util.spec_from_file_location("fuzz_solve_batch", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run(ncases=40, seed=77, verbose=False) == 0


def test_fit_fuzz_fast_equals_exact():
    """Random MFModel.fit-path problems (dictionary size, exact-G / between-shell / M = 271
    protocols, SNR, CSF / EAR, fascicle-count mix, planted degenerate voxels): rows of the
    screening tiers == rows of the reference-order tier, bit for bit."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "fuzz_fit.py")
    spec = importlib.util.spec_from_file_location("fuzz_fit", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run(ncases=20, seed=5, verbose=False) == 0


def test_solve_fuzz_vs_oracle():
    """Random small shapes (1-3 blocks, signed and non-negative data): whatever tier decides,
    weights, indices and objective are bit-identical to the CPU oracle (= the reference)."""
    rng = np.random.default_rng(99)
    for case in range(36):
        nb = int(rng.integers(1, 4))
        M = int(rng.choice([4, 9, 33, 100, 120]))
        sizes = [int(rng.integers(1, 60)) for _ in range(nb)]
        if nb >= 2 and rng.random() < 0.5:
            sizes[0], sizes[1] = int(rng.integers(8, 150)), int(rng.integers(8, 150))
        if nb == 3 and rng.random() < 0.4:
            sizes[2] = 1
        nt = int(np.sum(sizes))
        A = rng.random((M, nt)) + 0.02
        if rng.random() < 0.3:
            A *= rng.choice([-1.0, 1.0], size=A.shape)
        st = np.concatenate(([0], np.cumsum(sizes)[:-1]))
        V = 5
        Y = np.stack([A[:, st + np.array([rng.integers(0, n) for n in sizes])] @ (rng.random(nb) * (rng.random(nb) < 0.8))
                      for _ in range(V)]) + rng.choice([0.0, 0.03]) * rng.standard_normal((V, M))
        w, sub, tot, obj, yrec = mfu.solve_exhaustive_posweights_batch(A, Y, np.asarray(sizes))
        for v in range(V):
            wo, subo, toto, objo, _ = orc.solve(A, Y[v], sizes)
            assert np.array_equal(sub[v], subo) and np.array_equal(w[v], wo) and obj[v] == objo, (case, sizes, M, v)


def test_hcp_pipeline_batched(lowlevel):
    """solve_rotated_batch (rotate_atom x 2 + CSF column + 3-block search for many voxels, plans
    on the host, dictionaries assembled and searched on the GPU) equals the per-voxel sequence of
    the reference's test_hcp_dict."""
    g = lowlevel
    ref = np.array([0.0, 0.0, 1.0])
    sig, sch, DIFF, S0 = g["hcp_sig"], g["hcp_sch"], float(g["hcp_DIFF"]), g["hcp_S0"]
    n = sig.shape[1]
    rng = np.random.default_rng(3)
    V = 9
    peaks = rng.standard_normal((V, 2, 3))
    peaks /= np.linalg.norm(peaks, axis=2, keepdims=True)
    peaks[0, 0], peaks[0, 1] = g["hcp_dirs"][1], g["hcp_dirs"][2]
    Y = np.zeros((V, sig.shape[0]))
    truth = rng.integers(0, n, (V, 2))
    Ds = []
    for v in range(V):
        D = np.zeros((sig.shape[0], 2 * n + 1))
        for k in range(2):
            D[:, k * n:(k + 1) * n] = mfu.rotate_atom(sig, sch, ref, peaks[v, k], DIFF, S0, warnings=False)
        D[:, -1] = g["hcp_sig_csf"]
        Ds.append(D)
        Y[v] = 0.5 * D[:, truth[v, 0]] + 0.3 * D[:, n + truth[v, 1]] + 0.2 * D[:, -1]
    Y[0] = g["hcp_y"]
    w, sub, obj, ok = mfu.solve_rotated_batch(sig, sch, ref, peaks, Y, DIFF, S0, sig_iso=g["hcp_sig_csf"], chunk=4)
    assert ok.all()
    assert np.array_equal(sub[0], g["hcp_sub"]) and np.allclose(w[0], g["hcp_w"], rtol=1e-9)
    for v in range(V):
        w1, sub1, tot1, obj1, _ = mfu.solve_exhaustive_posweights(Ds[v], Y[v].copy(), np.array([n, n, 1]))
        assert np.array_equal(sub[v], sub1) and np.array_equal(w[v], w1) and obj[v] == obj1


def test_workspace_trim_and_reuse():
    """mfb_solve_batch keeps its workspace between calls; mfb_trim frees it and the next call
    allocates it again with identical results."""
    rng = np.random.default_rng(8)
    A = rng.random((40, 61)) + 0.05
    Y = rng.random((6, 40))
    first = mfu.solve_exhaustive_posweights_batch(A, Y, np.array([30, 30, 1]))
    _lib.trim(0)
    again = mfu.solve_exhaustive_posweights_batch(A, Y, np.array([30, 30, 1]))
    for f, g_ in zip(first, again):
        assert np.array_equal(f, g_)
    with pytest.raises(ValueError):
        _lib.trim(-1)


@pytest.mark.parametrize("tag", ["between", "dense"])
def test_mfmodel_fit_other_protocols_match_reference_maps(ukbb, tag):
    """MFModel.fit on a between-shell protocol and on the 271-row dense protocol (both screened on
    materialised dictionaries) against maps produced by the unmodified reference."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fit_protocols.npz"))
    model = _ukbb_model(ukbb)
    fit = model.fit(g["data_" + tag], g["mask"], g["numfasc"], peaks=g["peaks"], pgse_scheme=g["sch_" + tag],
                    csf_mask=g["csf"], verbose=0)
    assert list(fit.param_names) == [str(s) for s in g["fit_%s_param_names" % tag]]
    ysq = np.sum(g["data_" + tag] ** 2, axis=-1) / g["data_" + tag].shape[-1]
    for p in fit.param_names:
        got, ref = getattr(fit, p), g["fit_%s_%s" % (tag, p)]
        assert got.shape == ref.shape, p
        if p == "MSE":
            assert np.all(np.abs(got - ref) <= 1e-12 * ysq + 1e-9 * np.abs(ref)), p
        elif p == "R2":
            assert np.allclose(got, ref, rtol=1e-9, atol=1e-12), p
        else:
            assert np.allclose(got, ref, rtol=1e-9, atol=1e-300), p


def test_c_abi_rejects_bad_arguments():
    """The C ABI never throws: bad arguments come back as MFB_EINVAL with a message, which the
    Python mirror turns into ValueError (include/mfb200.h conventions)."""
    import ctypes
    import torch
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    A = torch.rand((10, 6), dtype=torch.float64, device=dev)
    y = torch.rand((2, 10), dtype=torch.float64, device=dev)
    w = torch.zeros((2, 2), dtype=torch.float64, device=dev)
    sub = torch.zeros((2, 2), dtype=torch.int32, device=dev)
    obj = torch.zeros(2, dtype=torch.float64, device=dev)
    sizes = np.array([3, 3], dtype=np.int64)
    p = sizes.ctypes.data_as(ctypes.c_void_p)

    def call(nb, lda, sz=p, Aptr=A.data_ptr()):
        return lib.mfb_solve_batch(0, 2, 10, nb, sz, Aptr, lda, 0, y.data_ptr(), w.data_ptr(), sub.data_ptr(),
                                   obj.data_ptr(), None, 0, None)
    assert call(2, 6) == 0
    assert call(2, 5) == _lib.MFB_EINVAL and b"lda" in lib.mfb_last_error()
    assert call(6, 6) == _lib.MFB_EINVAL
    assert call(2, 6, Aptr=None) == _lib.MFB_EINVAL
    bad = np.array([3, 0], dtype=np.int64)
    assert call(2, 6, sz=bad.ctypes.data_as(ctypes.c_void_p)) == _lib.MFB_EINVAL
    assert lib.mfb_fit(None, 1, None, None, None, None, None, 2, 0, 0, None, 0, None) == _lib.MFB_EINVAL
    assert lib.mfb_mc_average(0, 10, 9, A.data_ptr(), 1, A.data_ptr(), A.data_ptr(), 1.0, 5, A.data_ptr(), None) == _lib.MFB_EINVAL
    with pytest.raises(ValueError):
        _lib.check(_lib.MFB_EINVAL, "x")
    ph = make_phantom(n_atoms=16, n_vox=4, seed=1)
    msi = mfu.init_PGSE_multishell_interp(ph.dic["dictionary"], ph.dic["sch_mat"], ph.dic["orientation"])
    plan = mfu.GpuPlan(msi, mfu.SchemePlan(msi, ph.sch), None, None)
    try:
        # a CSF voxel on a plan without a CSF column, and more than two fascicles
        with pytest.raises(ValueError):
            plan.fit_host(ph.Y, ph.peaks, ph.K, np.ones(4, np.uint8), None, ph.maxfasc, True, False)
        with pytest.raises(ValueError):
            plan.fit_host(ph.Y, np.zeros((4, 9)), np.full(4, 3, np.int32), None, None, 3, False, False)
        # mfb_fit_volume: unknown element type, missing data; an empty ROI is not an error; the
        # plan still works after the failed calls (the pipeline threads and streams are left clean)
        out = np.zeros((4, 1 + 2 * ph.maxfasc + 2))
        vp = ctypes.c_void_p
        args = lambda data, dt, V=4: (plan.handle, V, data, dt, None, ph.Y.shape[1], 1, ph.peaks.ctypes.data_as(vp),  # noqa: E731
                                      ph.K.ctypes.data_as(vp), None, None, ph.maxfasc, 0, 0, out.ctypes.data_as(vp), 0)
        assert lib.mfb_fit_volume(*args(ph.Y.ctypes.data_as(vp), 7)) == _lib.MFB_EINVAL
        assert lib.mfb_fit_volume(*args(None, _lib.MFB_F64)) == _lib.MFB_EINVAL
        assert lib.mfb_fit_volume(*args(ph.Y.ctypes.data_as(vp), _lib.MFB_F64, V=0)) == 0
        assert lib.mfb_fit_volume(*args(ph.Y.ctypes.data_as(vp), _lib.MFB_F64)) == 0
        assert np.array_equal(out, plan.fit_host(ph.Y, ph.peaks, ph.K, None, None, ph.maxfasc, False, False))
    finally:
        plan.close()


@pytest.mark.gpu
def test_plan2d_device_equals_host_plan():
    """rotate_atom_2Dprotocol's interpolation plan (reference mf_utils.py:1440-1690): the GPU
    expansion of the host's per-direction decisions (mfb_plan2d) must give the rows and the two
    lerp weights of the host plan bit for bit, and the parallel-signal factor to the last bits of
    exp(); directions that break the protocol's assumptions get a zero plan in both."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "lowlevel_rotation.npz"))
    proto = mfu._Protocol2D(g["ax_sch"], np.array([0.0, 0.0, 1.0]), 2.0e-9)
    rng = np.random.default_rng(20)
    d = rng.standard_normal((700, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[3] = [1.0, 0.0, 0.0]                     # in the gradient plane: breaks the protocol
    d[4] = [0.0, 0.0, 1.0]                     # the reference direction itself
    d[5] = [0.0, 0.0, -1.0]
    d[6] = [np.sqrt(0.5), np.sqrt(0.5), 0.0]   # along a gradient line
    host = proto.plan(d, strict=False)
    devp = proto.plan_device(d)
    assert np.array_equal(host[5], devp[5]) and not host[5].all() and host[5].sum() > 600
    for k in range(4):
        assert np.array_equal(host[k], devp[k].cpu().numpy()), k
    sc = devp[4].cpu().numpy()
    assert np.allclose(sc, host[4], rtol=1e-15, atol=0.0)      # exp(): 1 ulp on the device + < 1 ulp in NumPy
    assert np.array_equal(sc == 0.0, host[4] == 0.0)
