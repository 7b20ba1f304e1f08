"""CPU-side checks: the C-ABI library loads and exports every symbol include/mfb200.h
declares; host marshalling logic (tables, scheme plans, NIfTI I/O, sharding)."""
import os
import re

import numpy as np
import pytest

from microstructure_fingerprinting_b200 import _lib, mf_utils as mfu, nifti
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mfb200.h")).read()
    declared = set(re.findall(r"\b(mfb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    m = re.search(r"#define MFB_ABI_VERSION (\d+)", hdr)
    assert lib.mfb_version() == int(m.group(1)) == _lib.MFB_ABI_VERSION
    assert lib.mfb_launch_count() >= 0


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.MFBError):
        mfu.solve_exhaustive_posweights(np.ones((3, 2)), np.ones(3), np.array([1, 1]))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "microstructure_fingerprinting_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f == "__init__.py" and False, f


def test_table_matches_oracle_and_reference(ukbb):
    msi = mfu.init_PGSE_multishell_interp(ukbb["dictionary"], ukbb["sch_mat"], ukbb["orientation"])
    assert np.array_equal(msi["nodes"], ukbb["ref_nodes"])
    assert np.array_equal(msi["table"], ukbb["ref_table"])
    assert np.array_equal(msi["off"], ukbb["ref_off"])
    assert np.array_equal(msi["Gms_un"], ukbb["ref_Gms_un"])
    assert msi["num_subs"] == ukbb["dictionary"].shape[1]
    assert len(msi["interpolators"]) == msi["Gms_un"].size


@pytest.mark.parametrize("mode", ["exact", "between"])
def test_scheme_plan_matches_oracle(ukbb, mode):
    msi = mfu.init_PGSE_multishell_interp(ukbb["dictionary"], ukbb["sch_mat"], ukbb["orientation"])
    sp = mfu.SchemePlan(msi, ukbb["sch_" + mode])
    tab = orc.init_table(ukbb["dictionary"], ukbb["sch_mat"], ukbb["orientation"])
    op = orc.plan_scheme(tab, ukbb["sch_" + mode])
    for k in ("shell_lo", "shell_hi", "gw_lo", "gw_hi", "gdir"):
        assert np.array_equal(getattr(sp, k), op[k]), k


def test_scheme_from_bvals_matches_reference(ukbb):
    sch = mfu.get_PGSE_scheme_from_bval_bvec_dense(ukbb["sch_mat"], ukbb["bvals"], ukbb["bvecs"], 1e-3)
    assert np.array_equal(sch, ukbb["sch_exact"])


def test_scheme_errors(ukbb):
    msi = mfu.init_PGSE_multishell_interp(ukbb["dictionary"], ukbb["sch_mat"], ukbb["orientation"])
    bad = ukbb["sch_exact"].copy()
    bad[:, 3] *= 10.0
    with pytest.raises(ValueError, match="Extrapolation not supported"):
        mfu.SchemePlan(msi, bad)
    bad = ukbb["sch_exact"].copy()
    bad[0, 4] *= 2
    with pytest.raises(ValueError, match="Delta, delta and TE"):
        mfu.SchemePlan(msi, bad)
    bad = ukbb["sch_exact"].copy()
    bad[5, :3] *= 1.5
    with pytest.raises(ValueError, match="unit norm"):
        mfu.SchemePlan(msi, bad)
    with pytest.raises(RuntimeError):
        mfu.import_PGSE_scheme(np.zeros((3, 6)))


def test_dt_vec_to_2darray():
    v = np.arange(1.0, 7.0)
    row = mfu.DT_vec_to_2Darray(v, "row")
    assert np.array_equal(row, [[1, 2, 3], [2, 4, 5], [3, 5, 6]])
    col = mfu.DT_vec_to_2Darray(v, "column")
    assert np.array_equal(col, [[1, 2, 4], [2, 3, 5], [4, 5, 6]])
    dia = mfu.DT_vec_to_2Darray(v, "diagonal")
    assert np.array_equal(dia, [[1, 4, 6], [4, 2, 5], [6, 5, 3]])


@pytest.mark.parametrize("ext", [".nii", ".nii.gz"])
def test_nifti_roundtrip(tmp_path, ext):
    rng = np.random.default_rng(0)
    vol = rng.standard_normal((4, 3, 2, 5))
    aff = np.array([[2.0, 0, 0, -10], [0, 2.5, 0, 7], [0, 0, 3.0, 1], [0, 0, 0, 1]])
    p = str(tmp_path / ("vol" + ext))
    nifti.save(vol, aff, p)
    back, aff2 = nifti.load(p)
    assert np.array_equal(back, vol)
    assert np.allclose(aff2, aff)


def test_reads_reference_style_nifti(tmp_path):
    # float32 volume with qform only, as produced by common neuroimaging tools
    import struct
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, 3, 2, 2, 2, 1, 1, 1, 1)
    struct.pack_into("<h", hdr, 70, 16)
    struct.pack_into("<h", hdr, 72, 32)
    struct.pack_into("<8f", hdr, 76, 1, 2, 2, 2, 1, 1, 1, 1)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<2h", hdr, 252, 1, 0)
    struct.pack_into("<6f", hdr, 256, 0, 0, 0, 5, 6, 7)
    hdr[344:348] = b"n+1\x00"
    data = np.arange(8, dtype=np.float32)
    p = str(tmp_path / "q.nii")
    with open(p, "wb") as f:
        f.write(bytes(hdr) + b"\0\0\0\0" + data.tobytes())
    vol, aff = nifti.load(p)
    assert vol.shape == (2, 2, 2) and vol[1, 0, 0] == 1.0 and vol[0, 1, 0] == 2.0
    assert np.allclose(aff, [[2, 0, 0, 5], [0, 2, 0, 6], [0, 0, 2, 7], [0, 0, 0, 1]])


def test_cleanup_2fascicles_matches_reference():
    from microstructure_fingerprinting_b200 import cleanup_2fascicles
    from tests.conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "cleanup_cases.npz"))
    pk, nf = cleanup_2fascicles(g["f1"], g["f2"], "peaks", g["u1"], g["u2"], g["mask"])
    assert np.array_equal(pk, g["peaks_pk"]) and np.array_equal(nf, g["peaks_nf"])
    pk, nf = cleanup_2fascicles(None, None, "colat_longit", g["cl1"], g["cl2"], g["mask"],
                                frac12=np.stack([g["f1"], g["f2"]], -1))
    assert np.array_equal(pk, g["colat_pk"]) and np.array_equal(nf, g["colat_nf"])
    pk, nf = cleanup_2fascicles(g["f1"], g["f2"], "tensor", g["t1"], g["t2"], g["mask"])
    assert np.array_equal(pk, g["tensor_pk"]) and np.array_equal(nf, g["tensor_nf"])
    assert set(np.unique(nf)) <= {0.0, 1.0, 2.0}
    with pytest.raises(ValueError):
        cleanup_2fascicles(None, None, "peaks", g["u1"], g["u2"], g["mask"])
    with pytest.raises(ValueError, match="Unknown peak mode"):
        cleanup_2fascicles(g["f1"], g["f2"], "bogus", g["u1"], g["u2"], g["mask"])


def test_2d_protocol_batched_plan_matches_reference_goldens():
    """Host logic of rotate_atom_2Dprotocol (mf_utils.py:1440-1690): the vectorised plan of
    _Protocol2D, applied with the NumPy lerp of the oracle, reproduces the outputs of the
    unmodified reference for the fixture directions, and a fascicle in the gradient plane raises
    the reference's AssertionError (strict) / is flagged (non-strict)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "lowlevel_rotation.npz"))
    ref = np.array([0.0, 0.0, 1.0])
    proto = mfu._Protocol2D(g["ax_sch"], ref, float(g["ax_DIFF"]))
    rl, rh, wl, wh, sc = proto.plan(g["ax_dirs"])
    got = orc.lerp_rows(proto.table(g["ax_sig"]), rl, rh, wl, wh, sc)
    assert np.allclose(got, g["ax_rot"], rtol=1e-12, atol=1e-300)
    bad = np.array([[np.sqrt(0.5), np.sqrt(0.5), 0.0]])
    with pytest.raises(AssertionError, match="pairs of opposite directions"):
        proto.plan(bad)
    out = proto.plan(np.vstack([g["ax_dirs"][:1], bad]), strict=False)
    assert out[5].tolist() == [True, False] and np.all(out[4][1] == 0.0)
    assert np.array_equal(out[0][0], rl[0]) and np.array_equal(out[3][0], wh[0])


def test_rotate_atom_plan_matches_reference_goldens():
    """Host plan of rotate_atom (mf_utils.py:1205-1437) + NumPy lerp == reference outputs."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "lowlevel_rotation.npz"))
    ref = np.array([0.0, 0.0, 1.0])
    S0 = g["hcp_S0"]
    table, rl, rh, wl, wh = mfu._rotate_atom_plan(g["hcp_sig"], g["hcp_sch"], ref, g["hcp_dirs"],
                                                  np.array([[float(g["hcp_DIFF"])]]), S0, warnings=False)
    got = orc.lerp_rows(table, rl, rh, wl, wh)
    assert np.allclose(got, g["hcp_rot"], rtol=1e-13, atol=0)


@pytest.mark.parametrize("full", [True, False])
@pytest.mark.parametrize("csf_on,ear_on", [(False, False), (True, False), (True, True)])
def test_mfmodelfit_maps_match_per_property_restatement(full, csf_on, ear_on):
    """MFModelFit builds its maps span by span in threads; compare every map with a plain
    per-property restatement of reference mf.py:1055-1175 on a (partial or full) 3-D mask."""
    from microstructure_fingerprinting_b200.mf import MFModelFit
    rng = np.random.default_rng(5)
    shape = (37, 41, 53)
    mask = np.ones(shape) if full else (rng.random(shape) < 0.6).astype(float)
    in_mask = mask > 0
    V, nf, N, E = int(in_mask.sum()), 2, 50, 4
    P = 1 + 2 * nf + csf_on + 2 * ear_on + 2
    rows = rng.random((V, P))
    rows[:, 1:1 + nf] *= rng.random((V, nf)) < 0.8            # some zero fractions
    rows[:, 1 + nf:1 + 2 * nf] = rng.integers(0, N, (V, nf))
    if ear_on:
        rows[:, 2 * nf + csf_on + 2] = rng.integers(0, E, V)
        rows[:, 2 * nf + csf_on + 1] *= rng.random(V) < 0.5
    peaks = rng.standard_normal((V, 3 * nf))
    info = {'maxfasc': nf, 'csf_on': csf_on, 'ear_on': ear_on, 'affine': np.eye(4), 'mask': mask,
            'fasc_propnames': ['fvf', 'rad'], 'peaks_roi': peaks, '_dict_fvf': rng.random(N),
            '_dict_rad': rng.random(N), 'DIFF_ear': rng.random(E)}
    fit = MFModelFit(info, rows)

    def vol(values, trailing=()):
        out = np.zeros(shape + trailing)
        out[in_mask] = values
        return out
    expect = {'M0': vol(rows[:, 0]), 'MSE': vol(rows[:, -2]), 'R2': vol(rows[:, -1])}
    for k in range(nf):
        expect['frac_f%d' % k] = vol(rows[:, k + 1])
        expect['peak_f%d' % k] = vol(peaks[:, 3 * k:3 * k + 3], (3,))
    for prop in ('fvf', 'rad'):
        tot = np.zeros(V)
        for k in range(nf):
            pk = info['_dict_' + prop][rows[:, 1 + nf + k].astype(int)] * (rows[:, k + 1] > 0)
            tot += rows[:, k + 1] * pk
            expect['%s_f%d' % (prop, k)] = vol(pk)
        expect[prop + '_tot'] = vol(tot)
    if csf_on:
        expect['frac_csf'] = vol(rows[:, 2 * nf + 1])
    if ear_on:
        c = 2 * nf + csf_on + 1
        expect['frac_ear'] = vol(rows[:, c])
        expect['D_ear'] = vol(info['DIFF_ear'][rows[:, c + 1].astype(int)] * (rows[:, c] > 0))
    assert set(fit.param_names) == set(expect)
    order = ['M0', 'frac_f0', 'peak_f0', 'frac_f1', 'peak_f1', 'fvf_f0', 'fvf_f1', 'fvf_tot', 'rad_f0',
             'rad_f1', 'rad_tot']
    assert fit.param_names[:len(order)] == order and fit.param_names[-2:] == ['MSE', 'R2']
    for name, e in expect.items():
        got = getattr(fit, name)
        assert got.shape == e.shape and np.array_equal(got, e), name
