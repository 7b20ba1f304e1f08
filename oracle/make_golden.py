"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, numba, scipy):

    NUMBA_CACHE_DIR=/tmp/numba_cache PYTHONDONTWRITEBYTECODE=1 \
        python oracle/make_golden.py

The GPU box has no /root/reference, so the vectors written here are committed
and are what tests/ read.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np

REF = os.environ.get("MF_REFERENCE", "/root/reference")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
from microstructure_fingerprinting import mf_utils as mfu  # noqa: E402
from microstructure_fingerprinting import mf as refmf  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
FIX = os.path.join(REF, "tests", "integration", "fixtures")
os.makedirs(OUT, exist_ok=True)


def solver_cases():
    """Seeded problems for every block count / branch of the solvers."""
    rng = np.random.default_rng(20260101)
    cases = {}
    specs = [
        ("s1_pos", 24, [37], "pos"), ("s1_signed", 24, [37], "signed"),
        ("s2_pos", 30, [23, 19], "pos"), ("s2_signed", 30, [23, 19], "signed"),
        ("s2_iso", 30, [41, 1], "pos"), ("s2_dup", 30, [12, 12], "dup"),
        ("s3_pos_iso", 28, [17, 15, 1], "pos"), ("s3_signed", 28, [9, 11, 7], "signed"),
        ("s3_pos_ear", 28, [13, 1, 6], "pos"), ("s3_dup", 28, [8, 8, 1], "dup"),
        ("s4_pos", 26, [6, 5, 1, 4], "pos"), ("s4_signed", 26, [4, 3, 2, 3], "signed"),
        ("s5_pos", 26, [3, 3, 3, 1, 2], "pos"),
    ]
    for name, M, sizes, kind in specs:
        nt = int(np.sum(sizes))
        nvox = 6
        if kind == "signed":
            A = rng.standard_normal((M, nt))
        else:
            A = rng.random((M, nt)) + 0.05
        if kind == "dup":           # identical sub-dictionaries -> exact ties / Det == 0
            A[:, sizes[0]:2 * sizes[0]] = A[:, :sizes[0]]
        st = np.concatenate(([0], np.cumsum(sizes)[:-1]))
        Y = np.zeros((nvox, M))
        for v in range(nvox):
            gt = st + np.array([rng.integers(0, n) for n in sizes])
            wgt = rng.random(len(sizes))
            if v % 3 == 1:
                wgt[rng.integers(0, len(sizes))] = 0.0      # inactive compartment
            Y[v] = A[:, gt] @ wgt
            if v % 2 == 0:
                Y[v] += 0.05 * rng.standard_normal(M)       # noisy
            if kind == "signed" and v == 5:
                Y[v] = -np.abs(Y[v])                         # pushes towards w = 0
        W = np.zeros((nvox, len(sizes)))
        SUB = np.zeros((nvox, len(sizes)), dtype=np.int64)
        OBJ = np.zeros(nvox)
        YREC = np.zeros((nvox, M))
        for v in range(nvox):
            w, sub, tot, obj, yrec = mfu.solve_exhaustive_posweights(
                A, Y[v].copy(), np.asarray(sizes))
            W[v], SUB[v], OBJ[v], YREC[v] = w, sub, obj, yrec
        cases[name + "_A"] = A
        cases[name + "_Y"] = Y
        cases[name + "_sizes"] = np.asarray(sizes)
        cases[name + "_w"] = W
        cases[name + "_sub"] = SUB
        cases[name + "_obj"] = OBJ
        cases[name + "_yrec"] = YREC
    np.savez_compressed(os.path.join(OUT, "solver_cases.npz"), **cases)
    print("solver_cases:", len(specs), "cases")


def ukbb_subset(n_atoms=48, n_ear=3):
    d = mfu.loadmat(os.path.join(FIX, "ukbb_90_dirs_dictionary_hcp_deltas.mat"))
    rng = np.random.default_rng(7)
    cols = np.sort(rng.choice(d["dictionary"].shape[1], n_atoms, replace=False))
    dic = {
        "dictionary": np.ascontiguousarray(d["dictionary"][:, cols]),
        "sch_mat": d["sch_mat"],
        "orientation": d["orientation"],
        "num_atom": n_atoms,
        "num_ear": n_ear,
        "T2_csf": d["T2_csf"], "DIFF_csf": d["CSF_DIFF"],
        "T2_ear": d["T2_ear"], "DIFF_ear": d["Dear"][:n_ear].copy(),
        "fasc_propnames": ["rad", "fin"],
        "rad": d["rad"][cols].copy(), "fin": d["fin"][cols].copy(),
    }
    return dic


def rotation_and_fit_cases():
    dic = ukbb_subset()
    bvals = np.loadtxt(os.path.join(FIX, "1000521_bvals.txt"))
    bvecs = np.loadtxt(os.path.join(FIX, "1000521_bvecs.txt"))
    sch_dense = dic["sch_mat"]
    # exact-G scheme (what MFModel.fit builds from bvals/bvecs, mf:838)
    sch_exact = mfu.get_PGSE_scheme_from_bval_bvec_dense(sch_dense, bvals, bvecs, 1e-3)
    # between-shell scheme: true (unsnapped) G values, as in the reference's
    # tests/integration/test_PGSE_from_multishell.py UKBB half
    gam = mfu.get_gyromagnetic_ratio("H")
    Del, dlt = sch_dense[0, 4], sch_dense[0, 5]
    sch_between = sch_exact.copy()
    sch_between[:, 3] = np.sqrt(bvals * 1e6 / (Del - dlt / 3)) / (gam * dlt)
    Gmax = np.max(sch_dense[:, 3])
    sch_between[:, 3] = np.minimum(sch_between[:, 3], Gmax)  # stay inside dense range

    msi = mfu.init_PGSE_multishell_interp(dic["dictionary"], sch_dense, dic["orientation"])
    rng = np.random.default_rng(11)
    dirs = rng.standard_normal((6, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dirs[0] = [0.0, 0.0, 1.0]
    dirs[1] = [1.0, 0.0, 0.0]
    rot_exact = np.stack([mfu.interp_PGSE_from_multishell(sch_exact, d, msinterp=msi)
                          for d in dirs])
    rot_between = np.stack([mfu.interp_PGSE_from_multishell(sch_between, d, msinterp=msi)
                            for d in dirs])
    # flattened interpolator of the reference itself (nodes/rows per shell)
    nodes = np.concatenate([f.x for f in msi["interpolators"]])
    table = np.vstack([f._y for f in msi["interpolators"]])
    off = np.cumsum([0] + [f.x.size for f in msi["interpolators"]])

    # ---- MFModel.fit on a small phantom: mixed numfasc / csf / ear ----
    model = refmf.MFModel(dic)
    shape = (4, 3, 2)
    V = int(np.prod(shape))
    N = dic["num_atom"]
    numfasc = rng.integers(0, 3, size=shape)
    numfasc.flat[0] = 2
    numfasc.flat[1] = 0
    csf = (rng.random(shape) < 0.5).astype(float)
    ear = np.zeros(shape)
    ear.flat[[2, 5, 7]] = 1                # K-dependent: gives 3-, and 4-block voxels
    mask = np.ones(shape)
    mask.flat[3] = 0
    peaks = rng.standard_normal(shape + (6,))
    peaks[..., :3] /= np.linalg.norm(peaks[..., :3], axis=-1, keepdims=True)
    peaks[..., 3:] /= np.linalg.norm(peaks[..., 3:], axis=-1, keepdims=True)
    sig_csf = np.exp(-sch_exact[:, 6] / dic["T2_csf"]) * np.exp(
        -(gam * sch_exact[:, 3] * sch_exact[:, 5]) ** 2
        * (sch_exact[:, 4] - sch_exact[:, 5] / 3) * dic["DIFF_csf"])
    data = np.zeros(shape + (sch_exact.shape[0],))
    for idx in np.ndindex(shape):
        K = numfasc[idx]
        y = np.zeros(sch_exact.shape[0])
        for k in range(K):
            a = mfu.interp_PGSE_from_multishell(sch_exact, peaks[idx][3 * k:3 * k + 3],
                                                msinterp=msi)
            y += rng.uniform(0.2, 0.6) * a[:, rng.integers(0, N)]
        y += csf[idx] * rng.uniform(0.05, 0.3) * sig_csf
        if K == 0 and csf[idx] == 0:
            y += 0.3 * sig_csf
        y *= 700.0
        y += 700.0 / 30.0 * rng.standard_normal(y.size)
        data[idx] = y
    fit_maps = {}
    for tag, kw in [
        ("A", dict(numfasc=numfasc, csf_mask=csf, ear_mask=None)),
        ("B", dict(numfasc=numfasc, csf_mask=None, ear_mask=None)),
        ("C", dict(numfasc=np.minimum(numfasc, 1), csf_mask=csf, ear_mask=ear)),
        ("D", dict(numfasc=numfasc, csf_mask=csf, ear_mask=ear)),
    ]:
        ft = model.fit(data, mask, kw["numfasc"], peaks=peaks, bvals=bvals, bvecs=bvecs,
                       csf_mask=kw["csf_mask"], ear_mask=kw["ear_mask"], verbose=0,
                       parallel=False)
        for p in ft.param_names:
            fit_maps["fit%s_%s" % (tag, p)] = getattr(ft, p)
        fit_maps["fit%s_param_names" % tag] = np.array(ft.param_names)
        print("fit", tag, ft.param_names)
    np.savez_compressed(
        os.path.join(OUT, "ukbb_subset.npz"),
        dictionary=dic["dictionary"], sch_mat=dic["sch_mat"],
        orientation=np.asarray(dic["orientation"], dtype=np.float64),
        T2_csf=dic["T2_csf"], DIFF_csf=dic["DIFF_csf"], T2_ear=dic["T2_ear"],
        DIFF_ear=dic["DIFF_ear"], rad=dic["rad"], fin=dic["fin"],
        bvals=bvals, bvecs=bvecs, sch_exact=sch_exact, sch_between=sch_between,
        dirs=dirs, rot_exact=rot_exact, rot_between=rot_between,
        ref_nodes=nodes, ref_table=table, ref_off=off, ref_Gms_un=msi["Gms_un"],
        data=data, mask=mask, numfasc=numfasc, csf=csf, ear=ear, peaks=peaks,
        **fit_maps)
    print("ukbb_subset written")


def fit_protocol_cases():
    """MFModel.fit of the unmodified reference on a between-shell protocol (gradient strengths
    that match no dense shell, mfu:1921-1956) and on the 271-row dense protocol (M > 112): the
    two protocol kinds that the GPU build screens on materialised dictionaries."""
    dic = ukbb_subset()
    bvals = np.loadtxt(os.path.join(FIX, "1000521_bvals.txt"))
    bvecs = np.loadtxt(os.path.join(FIX, "1000521_bvecs.txt"))
    sch_dense = dic["sch_mat"]
    sch_exact = mfu.get_PGSE_scheme_from_bval_bvec_dense(sch_dense, bvals, bvecs, 1e-3)
    gam = mfu.get_gyromagnetic_ratio("H")
    Del, dlt = sch_dense[0, 4], sch_dense[0, 5]
    sch_between = sch_exact.copy()
    sch_between[:, 3] = np.minimum(np.sqrt(bvals * 1e6 / (Del - dlt / 3)) / (gam * dlt), np.max(sch_dense[:, 3]))
    msi = mfu.init_PGSE_multishell_interp(dic["dictionary"], sch_dense, dic["orientation"])
    model = refmf.MFModel(dic)
    rng = np.random.default_rng(77)
    shape = (5, 4, 1)
    N = dic["num_atom"]
    numfasc = rng.integers(1, 3, size=shape)
    csf = (rng.random(shape) < 0.5).astype(float)
    mask = np.ones(shape)
    peaks = rng.standard_normal(shape + (6,))
    peaks[..., :3] /= np.linalg.norm(peaks[..., :3], axis=-1, keepdims=True)
    peaks[..., 3:] /= np.linalg.norm(peaks[..., 3:], axis=-1, keepdims=True)
    out = {"numfasc": numfasc, "csf": csf, "mask": mask, "peaks": peaks, "sch_between": sch_between,
           "sch_dense": sch_dense}
    for tag, sch in (("between", sch_between), ("dense", sch_dense)):
        b = (gam * sch[:, 3] * sch[:, 5]) ** 2 * (sch[:, 4] - sch[:, 5] / 3)
        sig_csf = np.exp(-sch[:, 6] / dic["T2_csf"]) * np.exp(-b * dic["DIFF_csf"])
        data = np.zeros(shape + (sch.shape[0],))
        for idx in np.ndindex(shape):
            y = np.zeros(sch.shape[0])
            for k in range(numfasc[idx]):
                a = mfu.interp_PGSE_from_multishell(sch, peaks[idx][3 * k:3 * k + 3], msinterp=msi)
                y += rng.uniform(0.2, 0.6) * a[:, rng.integers(0, N)]
            y += csf[idx] * rng.uniform(0.05, 0.3) * sig_csf
            data[idx] = 700.0 * y + 700.0 / 30.0 * rng.standard_normal(y.size)
        ft = model.fit(data, mask, numfasc, peaks=peaks, pgse_scheme=sch, csf_mask=csf, verbose=0, parallel=False)
        out["data_" + tag] = data
        for pn in ft.param_names:
            out["fit_%s_%s" % (tag, pn)] = getattr(ft, pn)
        out["fit_%s_param_names" % tag] = np.array(ft.param_names)
        print("fit protocol", tag, sch.shape[0], "rows done")
    np.savez_compressed(os.path.join(OUT, "fit_protocols.npz"), **out)


def reference_test_vectors():
    """Outputs of the reference on its own seeded test (test_synthetic_data,
    tests/integration/test_exhaustive_fingerprinting.py:94-153, shrunk)."""
    np.random.seed(141414)
    Nfasc, Natoms, M, Nvox = 2, 60, 100, 5
    A = np.random.randn(M * (Nfasc * Natoms + 1)).reshape((M, Nfasc * Natoms + 1), order="F")
    ID = np.zeros((3, Nvox), dtype=int)
    ID[0] = np.random.randint(0, Natoms, Nvox)
    ID[1] = np.random.randint(0, Natoms, Nvox) + Natoms
    ID[2] = 2 * Natoms
    w_gt = np.random.rand(3, Nvox)
    Y = np.stack([A[:, ID[:, i]] @ w_gt[:, i] for i in range(Nvox)])
    Y += 0.1 * (2.0 * np.random.rand(Nvox, M) - 1.0)
    sizes = np.array([Natoms, Natoms, 1])
    res = [mfu.solve_exhaustive_posweights(A, Y[i].copy(), sizes) for i in range(Nvox)]
    np.savez_compressed(os.path.join(OUT, "ref_synthetic.npz"), A=A, Y=Y, sizes=sizes,
                        ID=ID, w=np.stack([r[0] for r in res]),
                        tot=np.stack([r[2] for r in res]),
                        obj=np.array([r[3] for r in res]))
    print("ref_synthetic written")


def lowlevel_rotation_cases():
    """rotate_atom on a subset of the HCP Monte-Carlo dictionary (the reference's
    test_hcp_dict pipeline, tests/integration/test_exhaustive_fingerprinting.py:163-249) and
    rotate_atom_2Dprotocol on the AxCaliber fixture scheme with analytic atoms."""
    rng = np.random.default_rng(99)
    ld = mfu.loadmat(os.path.join(FIX, "MC_dictionary_hcp.mat"))
    dic = ld["dic_fascicle_refdir"]
    cols = np.sort(rng.choice(dic.shape[1], 16, replace=False))
    cols[3] = 86                                   # the atom test_hcp_dict plants
    cols = np.unique(cols)
    sig, S0 = np.ascontiguousarray(dic[:, cols]), np.ascontiguousarray(ld["S0_fascicle"][:, cols])
    sch = mfu.import_PGSE_scheme(os.path.join(FIX, "hcp_mgh_1003.scheme1"))
    sch_b0 = np.vstack((np.zeros((40, sch.shape[1])), sch))
    sch_b0[:40, 4:] = sch[0, 4:]
    refdir = np.array([0.0, 0.0, 1.0])
    dirs = rng.standard_normal((4, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dirs[0] = [0.0, 0.0, 1.0]
    rot = np.stack([mfu.rotate_atom(sig, sch_b0, refdir, d, ld["WM_DIFF"], S0) for d in dirs])
    rot1d = mfu.rotate_atom(sig[:, 2], sch_b0, refdir, dirs[1], ld["WM_DIFF"], S0[:, 2])
    # the joint pipeline of test_hcp_dict on the subset: 2 fascicles + CSF, noiseless
    i_gt = int(np.where(cols == 86)[0][0])
    fd = dirs[1:3]
    nu_gt = np.array([0.5, 0.3, 0.2])
    D = np.zeros((sch_b0.shape[0], 2 * cols.size + 1))
    y = np.zeros(sch_b0.shape[0])
    for k in range(2):
        D[:, k * cols.size:(k + 1) * cols.size] = mfu.rotate_atom(sig, sch_b0, refdir, fd[k],
                                                                  ld["WM_DIFF"], S0)
        y += 500 * nu_gt[k] * D[:, k * cols.size + i_gt]
    D[:, -1] = ld["sig_csf"]
    y += 500 * nu_gt[2] * ld["sig_csf"]
    w, sub, tot, obj, yrec = mfu.solve_exhaustive_posweights(D, y, np.array([cols.size, cols.size, 1]))

    # ---- 2D protocol ----
    sch2 = mfu.import_PGSE_scheme(os.path.join(FIX, "2D_qspace_clean_rot_xy.scheme"))
    gam = mfu.get_gyromagnetic_ratio("H")
    DIFF2 = 2.0e-9
    G, Del, dl = sch2[:, 3], sch2[:, 4], sch2[:, 5]
    ref2 = np.array([0.0, 0.0, 1.0])
    eff = mfu.rotate_scheme_mat(sch2.copy(), np.array([0, 0, 1.0]), ref2)
    gperp = np.sqrt(np.sum(eff[:, :2] ** 2, axis=1))
    bperp = (gam * dl * G * gperp) ** 2 * (Del - dl / 3)
    bpar = (gam * dl * np.abs(eff[:, 2]) * G) ** 2 * (Del - dl / 3)
    dperp = np.array([0.05e-9, 0.2e-9, 0.45e-9, 0.8e-9, 1.3e-9])
    sig2 = np.exp(-bpar[:, None] * DIFF2) * np.exp(-bperp[:, None] * dperp[None, :])
    sig2 *= (1.0 + 0.01 * rng.standard_normal(sig2.shape))       # MC-like variability
    dirs2 = rng.standard_normal((4, 3))
    dirs2 /= np.linalg.norm(dirs2, axis=1, keepdims=True)
    dirs2[0] = [0.0, 0.0, 1.0]
    # the reference normalises its sch_mat argument in place when refdir is the z axis: hand it a
    # fresh copy per call so that every golden comes from the pristine fixture scheme
    rot2 = np.stack([mfu.rotate_atom_2Dprotocol(sig2, sch2.copy(), ref2, d, DIFF2) for d in dirs2])
    rot2_1d = mfu.rotate_atom_2Dprotocol(sig2[:, 1], sch2.copy(), ref2, dirs2[2], DIFF2)
    np.savez_compressed(
        os.path.join(OUT, "lowlevel_rotation.npz"),
        hcp_sig=sig, hcp_S0=S0, hcp_sch=sch_b0, hcp_dirs=dirs, hcp_rot=rot, hcp_rot1d=rot1d,
        hcp_DIFF=ld["WM_DIFF"], hcp_sig_csf=ld["sig_csf"], hcp_y=y, hcp_w=w, hcp_sub=sub,
        hcp_obj=obj, hcp_igt=i_gt,
        ax_sig=sig2, ax_sch=sch2, ax_dirs=dirs2, ax_rot=rot2, ax_rot1d=rot2_1d, ax_DIFF=DIFF2)
    print("lowlevel_rotation written; hcp solve:", sub, w / w.sum())


def cleanup_cases():
    """cleanup_2fascicles (mf.py:36-335) on random inputs in its three peak modes."""
    rng = np.random.default_rng(404)
    shape = (6, 5, 4)
    mask = (rng.random(shape) < 0.85).astype(float)
    f1 = rng.random(shape) * 0.8
    f2 = rng.random(shape) * 0.6
    f1.flat[:6] = [0.5, 0.05, 0.3, 0.1, 0.0, 0.19]
    f2.flat[:6] = [0.5, 0.05, 0.1, 0.3, 0.0, 0.5]
    u1 = rng.standard_normal(shape + (3,)); u1 /= np.linalg.norm(u1, axis=-1, keepdims=True)
    u2 = rng.standard_normal(shape + (3,)); u2 /= np.linalg.norm(u2, axis=-1, keepdims=True)
    near = rng.random(shape) < 0.25                    # nearly parallel pairs (merge branch)
    u2[near] = u1[near] * rng.choice([-1.0, 1.0], size=near.sum())[:, None] + 0.05 * rng.standard_normal((near.sum(), 3))
    u2 /= np.linalg.norm(u2, axis=-1, keepdims=True)
    out = {"mask": mask, "f1": f1, "f2": f2, "u1": u1, "u2": u2}
    pk, nf = refmf.cleanup_2fascicles(f1, f2, "peaks", u1, u2, mask)
    out["peaks_pk"], out["peaks_nf"] = pk, nf
    cl1 = np.stack([np.arccos(np.clip(u1[..., 2], -1, 1)), np.arctan2(u1[..., 1], u1[..., 0])], -1)
    cl2 = np.stack([np.arccos(np.clip(u2[..., 2], -1, 1)), np.arctan2(u2[..., 1], u2[..., 0])], -1)
    pk, nf = refmf.cleanup_2fascicles(None, None, "colat_longit", cl1, cl2, mask,
                                      frac12=np.stack([f1, f2], -1))
    out["cl1"], out["cl2"], out["colat_pk"], out["colat_nf"] = cl1, cl2, pk, nf

    def tens(u):
        DT = 1e-4 * np.eye(3) + 1.9e-3 * u[..., :, None] * u[..., None, :]
        T = np.zeros(u.shape[:-1] + (1, 6))
        T[..., 0, 0], T[..., 0, 1], T[..., 0, 2] = DT[..., 0, 0], DT[..., 0, 1], DT[..., 1, 1]
        T[..., 0, 3], T[..., 0, 4], T[..., 0, 5] = DT[..., 0, 2], DT[..., 1, 2], DT[..., 2, 2]
        return T
    t1, t2 = tens(u1), tens(u2)
    t2[0, 0, 0] = 0
    pk, nf = refmf.cleanup_2fascicles(f1, f2, "tensor", t1, t2, mask)
    out["t1"], out["t2"], out["tensor_pk"], out["tensor_nf"] = t1, t2, pk, nf
    np.savez_compressed(os.path.join(OUT, "cleanup_cases.npz"), **out)
    print("cleanup_cases written", np.unique(out["peaks_nf"], return_counts=True))


def mc_cases():
    """monte_carlo_average / get_PGSE_from_phases (mfu:2758-3015) on synthetic spin phases:
    3 reference (Delta, delta) pairs, 6000 spins, phases ~ N(0, 1.5) per component, stored
    as big-endian double and little-endian single phase files like the simulator writes."""
    import tempfile
    rng = np.random.default_rng(424242)
    n_ref, n_spin = 3, 6000
    sim = np.zeros((n_ref, 7))
    sim[:, 0] = 1.0 / np.sqrt(3); sim[:, 1] = 1.0 / np.sqrt(3); sim[:, 2] = 1.0 / np.sqrt(3)
    sim[:, 3] = [0.04, 0.05, 0.06]
    sim[:, 4] = [0.020, 0.035, 0.050]
    sim[:, 5] = [0.008, 0.010, 0.012]
    sim[:, 6] = 0.08
    phases = 1.5 * rng.standard_normal((n_ref * n_spin, 3))
    n_seq = 40
    pick = rng.integers(0, n_ref, n_seq)
    new = np.zeros((n_seq, 7))
    g = rng.standard_normal((n_seq, 3)); g /= np.linalg.norm(g, axis=1, keepdims=True)
    new[:, :3] = g
    new[:, 3] = rng.uniform(0.0, 0.08, n_seq)
    new[:5, 3] = 0.0
    new[:, 4:6] = sim[pick, 4:6]
    new[:, 6] = 0.08
    out = {"sim": sim, "new": new, "phases": phases, "n_spin": n_spin}
    # direct calls of the Numba kernel
    gsc = (new[:, :3] * new[:, 3:4]) / (sim[pick, :3] * sim[pick, 3:4])
    for dim in (2, 3):
        for ds in (1.0, 0.73):
            out["avg_d%d_s%g" % (dim, ds)] = mfu.monte_carlo_average(
                np.ascontiguousarray(phases[:, :dim]), pick.astype(np.int64),
                np.ascontiguousarray(gsc[:, :dim]), float(ds), n_spin)
    out["pick"], out["gsc"] = pick, gsc
    # through the file interface
    with tempfile.TemporaryDirectory() as tmp:
        for i, nm in enumerate("xyz"):
            phases[:, i].astype(">f8").tofile(os.path.join(tmp, "sub_phase_%s.bdouble" % nm))
            phases[:, i].astype("<f4").tofile(os.path.join(tmp, "sub_phase_%s.lfloat" % nm))
        out["file_bdouble_d3"] = mfu.get_PGSE_from_phases(os.path.join(tmp, "sub_phase_x.bdouble"), sim, new)
        out["file_bdouble_d3_D"] = mfu.get_PGSE_from_phases(os.path.join(tmp, "sub_phase_x.bdouble"), sim, new,
                                                              dim=3, D_sim=2.0e-9, D=1.1e-9)
        new2 = new.copy()
        ang = rng.uniform(0, 2 * np.pi, n_seq)
        new2[:, 0], new2[:, 1], new2[:, 2] = np.cos(ang), np.sin(ang), 0.0
        out["new2"] = new2
        out["file_lfloat_d2"] = mfu.get_PGSE_from_phases(os.path.join(tmp, "sub_phase_x.lfloat"), sim, new2, dim=2)
    np.savez_compressed(os.path.join(OUT, "mc_cases.npz"), **out)
    print("mc_cases written", out["file_bdouble_d3"][:4])


if __name__ == "__main__":
    if "--only-protocols" in sys.argv:
        fit_protocol_cases()
        sys.exit(0)
    mc_cases()
    if "--only-mc" in sys.argv:
        sys.exit(0)
    fit_protocol_cases()
    cleanup_cases()
    lowlevel_rotation_cases()
    solver_cases()
    reference_test_vectors()
    rotation_and_fit_cases()
