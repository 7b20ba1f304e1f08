"""Synthetic multi-shell phantoms for tests, smoke() and bench.py (NumPy only).

Dense sampling: the 271-row UKBB-like scheme (3 shells x 90 directions + b0) stored in
tests/golden/ukbb_subset.npz; subject protocol: the 105 b-values / b-vectors of the same
file snapped to the dense shells (what MFModel.fit builds from bvals/bvecs).  Atoms: an
analytic cylinder-symmetric two-compartment family, smooth in x = |g.u|:
    S = exp(-TE/T2) [ f exp(-b (Dpar x^2 + Din (1-x^2))) + (1-f) exp(-b (Dpar x^2 + Dex (1-x^2))) ]
"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GAMMA = 2 * np.pi * 42.577480e6


def _bvalues(sch):
    return (GAMMA * sch[:, 3] * sch[:, 5]) ** 2 * (sch[:, 4] - sch[:, 5] / 3)


def load_schemes():
    g = np.load(os.path.join(GOLDEN, "ukbb_subset.npz"))
    return g["sch_mat"], g["sch_exact"], g["sch_between"], g["bvals"], g["bvecs"]


def make_dictionary(n_atoms, n_ear=4):
    """Analytic dictionary in the key layout MFModel expects (reference mf.py:506-513)."""
    sch_dense = load_schemes()[0]
    b = _bvalues(sch_dense)
    x2 = sch_dense[:, 2] ** 2                       # orientation = z
    n_f = int(np.ceil(n_atoms ** (1 / 3)))
    n_in = int(np.ceil(np.sqrt(n_atoms / n_f)))
    n_ex = int(np.ceil(n_atoms / (n_f * n_in)))
    f = np.linspace(0.25, 0.85, n_f)
    d_in = np.linspace(0.02e-9, 0.45e-9, n_in)
    d_ex = np.linspace(0.5e-9, 1.4e-9, n_ex)
    F, DI, DE = [a.ravel()[:n_atoms] for a in np.meshgrid(f, d_in, d_ex, indexing="ij")]
    dpar, T2, TE = 2.0e-9, 0.07, sch_dense[:, 6]
    intra = np.exp(-b[:, None] * (dpar * x2[:, None] + DI[None, :] * (1 - x2[:, None])))
    extra = np.exp(-b[:, None] * (dpar * x2[:, None] + DE[None, :] * (1 - x2[:, None])))
    D = np.exp(-TE / T2)[:, None] * (F[None, :] * intra + (1 - F[None, :]) * extra)
    return {"dictionary": np.ascontiguousarray(D), "sch_mat": sch_dense,
            "orientation": np.array([0.0, 0.0, 1.0]), "num_atom": n_atoms, "num_ear": n_ear,
            "T2_csf": 2.03, "DIFF_csf": 3.0e-9, "T2_ear": 2.03,
            "DIFF_ear": np.linspace(0.1e-9, 1.0e-9, n_ear),
            "fasc_propnames": ["fvf", "dperp_in"], "fvf": F.copy(), "dperp_in": DI.copy()}


def host_table(dic):
    """Lookup table of the dense sampling (NumPy restatement used only to SYNTHESISE
    phantom signals; same construction as reference mf_utils.py:2008-2080)."""
    sch = dic["sch_mat"]
    x_all = np.abs(sch[:, :3] @ np.asarray(dic["orientation"], dtype=float))
    Gun, inv = np.unique(sch[:, 3], return_inverse=True)
    shells = []
    for s, G in enumerate(Gun):
        rows = np.where(inv == s)[0]
        if G == 0:
            shells.append((np.array([0.0, 1.0]), np.repeat([dic["dictionary"][rows[0]]], 2, axis=0)))
            continue
        xs, first = np.unique(x_all[rows], return_index=True)
        ys = dic["dictionary"][rows][first]
        near = np.abs(xs - xs[0]) < 1e-3
        c = int(near.sum())
        if c > 1:
            xs = np.append(xs[near].mean(), xs[c:])
            ys = np.vstack([ys[near].mean(axis=0, keepdims=True), ys[c:]])
        shells.append((xs, ys))
    return Gun, shells


def _rotate_shell(shells, s, xv, atoms):
    xs, ys = shells[s]
    j = np.clip(np.searchsorted(xs, xv), 1, xs.size - 1)
    w_hi = (xv - xs[j - 1]) / (xs[j] - xs[j - 1])
    w_lo = (xs[j] - xv) / (xs[j] - xs[j - 1])
    a = atoms[:, None]
    return w_hi * ys[j, a] + w_lo * ys[j - 1, a]


def rotate_columns(dic, sch, dirs, atoms, table=None):
    """Signals of atom `atoms[v]` rotated to `dirs[v]` for every voxel: (V, M).  Gradient
    strengths between two dense shells are blended linearly in G (reference
    mf_utils.py:1921-1956).  Large batches are split over threads by voxel blocks (the
    result does not depend on the split)."""
    Gun, shells = table if table is not None else host_table(dic)
    V, M = dirs.shape[0], sch.shape[0]
    blk = 1 << 15
    if V > 2 * blk:
        import os
        from concurrent.futures import ThreadPoolExecutor
        cuts = list(range(0, V, blk)) + [V]
        with ThreadPoolExecutor(min(16, os.cpu_count() or 1)) as ex:
            parts = list(ex.map(lambda i: rotate_columns(dic, sch, dirs[cuts[i]:cuts[i + 1]],
                                                         atoms[cuts[i]:cuts[i + 1]], (Gun, shells)),
                                range(len(cuts) - 1)))
        return np.concatenate(parts, axis=0)
    out = np.zeros((V, M))
    x = np.abs(dirs @ sch[:, :3].T)                  # (V, M)
    for G in np.unique(sch[:, 3]):
        cols = np.where(sch[:, 3] == G)[0]
        xv = x[:, cols]
        hit = np.where(Gun == G)[0]
        if hit.size:
            out[:, cols] = _rotate_shell(shells, int(hit[0]), xv, atoms)
        else:
            h = int(np.argmax(Gun > G))
            lo, hi = Gun[h - 1], Gun[h]
            out[:, cols] = ((G - lo) / (hi - lo)) * _rotate_shell(shells, h, xv, atoms) + \
                ((hi - G) / (hi - lo)) * _rotate_shell(shells, h - 1, xv, atoms)
    return out


class Phantom(object):
    pass


def make_phantom(n_atoms=96, n_vox=64, seed=0, frac_k=(0.1, 0.4, 0.5), csf_frac=0.3, ear=False,
                 snr=30.0, n_ear=4, dic=None, scheme="exact", ear_frac=0.1, ear_max_k=1):
    """Seeded phantom: numfasc in {0,1,2} with probabilities frac_k, CSF on csf_frac of the
    voxels, optional EAR on ~10%, crossing angle U(15,90) deg, M0 = 800, Gaussian noise."""
    rng = np.random.default_rng(seed)
    ph = Phantom()
    ph.dic = dic if dic is not None else make_dictionary(n_atoms, n_ear)
    # subject protocol: 105 rows snapped to the dense shells ("exact"), the same rows with
    # their original gradient strengths ("between"), or the 271-row dense scheme itself
    ph.sch = {"exact": load_schemes()[1], "between": load_schemes()[2], "dense": load_schemes()[0]}[scheme]
    ph.bvals, ph.bvecs = load_schemes()[3], load_schemes()[4]
    M = ph.sch.shape[0]
    V = n_vox
    ph.K = rng.choice(3, size=V, p=np.asarray(frac_k) / np.sum(frac_k)).astype(np.int32)
    ph.csf = (rng.random(V) < csf_frac).astype(np.uint8)
    ph.ear = ((rng.random(V) < ear_frac) & (ph.K <= ear_max_k)).astype(np.uint8) if ear else np.zeros(V, np.uint8)
    u1 = rng.standard_normal((V, 3))
    u1 /= np.linalg.norm(u1, axis=1, keepdims=True)
    t = rng.standard_normal((V, 3))
    t -= np.sum(t * u1, axis=1, keepdims=True) * u1
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    ang = np.deg2rad(rng.uniform(15, 90, V))[:, None]
    u2 = np.cos(ang) * u1 + np.sin(ang) * t
    ph.peaks = np.ascontiguousarray(np.hstack([u1, u2]))
    ph.atoms = rng.integers(0, ph.dic["num_atom"], size=(V, 2))
    b = _bvalues(ph.sch)
    ph.sig_csf = np.exp(-ph.sch[:, 6] / ph.dic["T2_csf"]) * np.exp(-b * ph.dic["DIFF_csf"])
    ph.sig_ear = np.stack([np.exp(-ph.sch[:, 6] / ph.dic["T2_ear"]) * np.exp(-b * d)
                           for d in np.atleast_1d(ph.dic["DIFF_ear"])], axis=1)
    nu1 = rng.uniform(0.3, 0.7, V)
    nu_csf = rng.uniform(0.0, 0.3, V) * ph.csf
    nu_ear = rng.uniform(0.05, 0.2, V) * ph.ear
    rest = 1.0 - nu_csf - nu_ear
    w1 = np.where(ph.K == 2, nu1, 1.0) * rest * (ph.K >= 1)
    w2 = (1.0 - nu1) * rest * (ph.K == 2)
    Y = w1[:, None] * rotate_columns(ph.dic, ph.sch, u1, ph.atoms[:, 0])
    Y += w2[:, None] * rotate_columns(ph.dic, ph.sch, u2, ph.atoms[:, 1])
    Y += nu_csf[:, None] * ph.sig_csf[None, :]
    Y += nu_ear[:, None] * ph.sig_ear[:, rng.integers(0, ph.sig_ear.shape[1])][None, :]
    empty = (ph.K + ph.csf + ph.ear) == 0
    Y[empty] = 0.3 * ph.sig_csf[None, :]
    M0 = 800.0
    ph.Y = np.ascontiguousarray(M0 * Y + (M0 / snr) * rng.standard_normal((V, M)))
    ph.maxfasc = int(ph.K.max())
    ph.csf_on, ph.ear_on = bool(ph.csf.any()), bool(ph.ear.any())
    ph.peaks = np.ascontiguousarray(ph.peaks[:, :3 * ph.maxfasc])

    def gpu_rows(flags=0, device=0):
        from microstructure_fingerprinting_b200 import mf_utils as mfu
        msi = mfu.init_PGSE_multishell_interp(ph.dic["dictionary"], ph.dic["sch_mat"],
                                              ph.dic["orientation"])
        plan = mfu.GpuPlan(msi, mfu.SchemePlan(msi, ph.sch), ph.sig_csf if ph.csf_on else None,
                           ph.sig_ear if ph.ear_on else None, device=device)
        try:
            return plan.fit_host(ph.Y, ph.peaks, ph.K, ph.csf, ph.ear, ph.maxfasc, ph.csf_on,
                                 ph.ear_on, flags=flags)
        finally:
            plan.close()
    ph.gpu_rows = gpu_rows
    return ph


def oracle_rows(ph, idx=None):
    """params rows from the CPU oracle (tests / smoke / cpu_baseline only)."""
    from oracle import oracle as orc
    tab = orc.init_table(ph.dic["dictionary"], ph.dic["sch_mat"], ph.dic["orientation"])
    plan = orc.plan_scheme(tab, ph.sch)
    sel = slice(None) if idx is None else idx
    K, csf, ear = ph.K[sel], ph.csf[sel], ph.ear[sel]
    P = 1 + 2 * ph.maxfasc + ph.csf_on + 2 * ph.ear_on + 2
    out = np.zeros((K.size, P))
    Y, peaks = ph.Y[sel], ph.peaks[sel]
    for i in range(K.size):
        out[i] = orc.fit_voxel(tab, plan, Y[i], K[i], csf[i], ear[i], peaks[i], ph.maxfasc,
                               ph.csf_on, ph.ear_on, ph.sig_csf, ph.sig_ear)
    return out


def compare_rows(rows, ref, ph, idx=None, exact_bits=False):
    """The parity contract (BASELINE.json north_star): atom indices exact; M0, fractions
    within 1e-9 relative; MSE with an absolute floor 1e-12*|y|^2/M; R2 within 1e-9."""
    sel = slice(None) if idx is None else idx
    mf = ph.maxfasc
    Y = ph.Y[sel]
    ysq = np.sum(Y ** 2, axis=1) / Y.shape[1]
    ids = slice(1 + mf, 1 + 2 * mf)
    assert np.array_equal(rows[:, ids], ref[:, ids]), "fascicle atom indices differ"
    if ph.ear_on:
        j = 2 * mf + ph.csf_on + 2
        assert np.array_equal(rows[:, j], ref[:, j]), "EAR atom indices differ"
    if exact_bits:
        assert np.array_equal(rows[:, :-1], ref[:, :-1]), "rows are not bit-identical"
    frac_cols = [0] + list(range(1, 1 + mf)) + ([2 * mf + 1] if ph.csf_on else []) + \
        ([2 * mf + ph.csf_on + 1] if ph.ear_on else [])
    for c in frac_cols:
        assert np.allclose(rows[:, c], ref[:, c], rtol=1e-9, atol=1e-300), "column %d" % c
    assert np.all(np.abs(rows[:, -2] - ref[:, -2]) <= 1e-12 * ysq + 1e-9 * np.abs(ref[:, -2]))
    assert np.allclose(rows[:, -1], ref[:, -1], rtol=1e-9, atol=1e-12)
