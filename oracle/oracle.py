"""CPU oracle for the per-voxel exhaustive dictionary fit -- TEST INFRASTRUCTURE ONLY.

Python face of ``mf_oracle.c`` (plain C, -ffp-contract=off) plus NumPy
restatements of the host-side pieces of the reference.  Nothing under
``microstructure_fingerprinting_b200/`` imports this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs do.

Parity status: PINNED against the unmodified reference (see
``oracle/make_golden.py`` and ``tests/test_oracle_golden.py``).

Citations are path:line in rensonnetg/microstructure_fingerprinting
(mfu = microstructure_fingerprinting/mf_utils.py, mf = .../mf.py).
"""
import ctypes
import itertools
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    """Compile libmf_oracle.so with gcc (building the checker is not using it)."""
    so = os.path.join(_HERE, "libmf_oracle.so")
    src = os.path.join(_HERE, "mf_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmf_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


# --------------------------------------------------------------------------
# Rotation (multi-shell interpolation)
# --------------------------------------------------------------------------

def init_table(sig_ms, sch_mat_ms, ordir):
    """init_PGSE_multishell_interp (mfu:1959-2085) flattened to one lookup table.

    Returns dict(Gms_un, off (n_shells+1), nodes (R,), table (R,N), DeldelTE, num_subs).
    """
    sig_ms = np.asarray(sig_ms, dtype=np.float64)
    sch_mat_ms = np.asarray(sch_mat_ms, dtype=np.float64)
    ordir = np.squeeze(np.asarray(ordir)).astype(np.float64)
    if sig_ms.ndim == 1:
        sig_ms = sig_ms.reshape((sig_ms.size, 1))
    x_all = np.abs(np.dot(sch_mat_ms[:, 0:3], ordir))          # mfu:2006
    Gms_un, i_G = np.unique(sch_mat_ms[:, 3], return_inverse=True)  # mfu:2008
    nodes, rows, off = [], [], [0]
    for s in range(Gms_un.size):
        ind = np.where(i_G == s)[0]
        if Gms_un[s] == 0:                                     # mfu:2019-2046
            xs = np.array([0.0, 1.0])
            ys = np.repeat([sig_ms[ind[0], :]], 2, axis=0)
        else:
            xs, first = np.unique(x_all[ind], return_index=True)   # mfu:2048-2053
            ys = sig_ms[ind, :][first, :]
            near = np.abs(xs - xs[0]) < 1e-3                    # mfu:2059
            c = int(np.sum(near))
            if c > 1:                                           # mfu:2065-2072
                xs = np.append(np.mean(xs[near]), xs[c:])
                ys = np.append(np.mean(ys[near, :], axis=0, keepdims=True),
                               ys[c:, :], axis=0)
        nodes.append(xs)
        rows.append(ys)
        off.append(off[-1] + xs.size)
    return {"Gms_un": Gms_un,
            "off": np.asarray(off, dtype=np.int32),
            "nodes": np.ascontiguousarray(np.concatenate(nodes)),
            "table": np.ascontiguousarray(np.vstack(rows)),
            "DeldelTE": sch_mat_ms[0, 4:7].copy(),
            "num_subs": sig_ms.shape[1]}


def plan_scheme(tab, sch_mat):
    """Map each subject measurement to dense shell(s): mfu:1786-1839."""
    sch_mat = np.asarray(sch_mat, dtype=np.float64)
    if not np.all(np.isclose(tab["DeldelTE"], sch_mat[:, 4:7])):
        raise ValueError("Delta, delta and TE values should all be "
                         "identical to those in the multi-shell sampling.")
    gn = np.sqrt(np.sum(sch_mat[:, 0:3] ** 2, axis=1))
    if np.any(np.abs(1 - gn[gn > 0]) > 1e-3):
        raise ValueError("Gradient directions in multi-shell scheme matrix"
                         " should all either have zero or unit norm.")
    Gms = tab["Gms_un"]
    M = sch_mat.shape[0]
    lo = np.zeros(M, dtype=np.int32)
    hi = np.zeros(M, dtype=np.int32)
    gw_lo = np.ones(M)
    gw_hi = np.zeros(M)
    for m in range(M):
        G = sch_mat[m, 3]
        i = np.where(G == Gms)[0]
        if i.size > 0:
            lo[m] = hi[m] = i[0]
        else:
            ih = int(np.argmax(Gms > G))
            if ih == 0:
                raise ValueError("Gradient intensity %g is not in the [%g, %g]"
                                 " range spanned by the multi-shell sampling."
                                 " Extrapolation not supported." % (G, Gms[0], Gms[-1]))
            lo[m], hi[m] = ih - 1, ih
            gw_hi[m] = (G - Gms[ih - 1]) / (Gms[ih] - Gms[ih - 1])
            gw_lo[m] = (Gms[ih] - G) / (Gms[ih] - Gms[ih - 1])
    return {"gdir": np.ascontiguousarray(sch_mat[:, 0:3]), "shell_lo": lo,
            "shell_hi": hi, "gw_lo": gw_lo, "gw_hi": gw_hi}


def rotate(tab, plan, newdir, out=None):
    """interp_PGSE_from_multishell, fast mode (mfu:1693): returns (M, N)."""
    u = np.ascontiguousarray(np.squeeze(np.asarray(newdir, dtype=np.float64)))
    nrm = np.sqrt((u ** 2).sum())
    if np.abs(1 - nrm) > 1e-3:                                   # mfu:1798-1802
        raise ValueError("Orientation vector of the new signal must have unit norm."
                         " Detected %g." % (nrm,))
    M = plan["gdir"].shape[0]
    N = tab["table"].shape[1]
    if out is None:
        out = np.empty((M, N))
    ldd = out.strides[0] // 8
    _lib().orc_rotate_multishell(
        M, N, tab["off"].size - 1, _i(tab["off"]), _d(tab["nodes"]), _d(tab["table"]),
        _d(plan["gdir"]), _i(plan["shell_lo"]), _i(plan["shell_hi"]),
        _d(plan["gw_lo"]), _d(plan["gw_hi"]), _d(u), _d(out), ldd)
    return out


# --------------------------------------------------------------------------
# Solvers
# --------------------------------------------------------------------------

def _solve_4up(A, y, sizes):
    """solve_exhaustive_posweights_4up (mfu:612-657): scipy.optimize.nnls per tuple."""
    import scipy.optimize
    st = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.int64)
    w_best = np.zeros(sizes.size)
    idx_best = np.zeros(sizes.size, dtype=np.int64)
    min_obj = np.sum(y ** 2)
    for idx in itertools.product(*[range(int(n)) for n in sizes]):
        w, r = scipy.optimize.nnls(A[:, st + np.asarray(idx)], y)
        obj = r * r
        if obj < min_obj:
            w_best, min_obj, idx_best = w, obj, np.atleast_1d(idx).astype(np.int64)
    return w_best, idx_best, st + idx_best, min_obj


def solve(A, y, dicsizes):
    """solve_exhaustive_posweights (mfu:115-214). Same 5-tuple as the reference."""
    assert isinstance(A, np.ndarray) and A.ndim == 2
    assert not np.any(np.all(A == 0, axis=0)), "All-zero columns detected in A"
    A = np.ascontiguousarray(A, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    assert A.size > 0 and y.size > 0
    assert A.shape[0] == y.size
    sizes = np.asarray(dicsizes).astype(np.int64)
    assert np.all(sizes > 0) and A.shape[1] == np.sum(sizes)
    M, lda = A.shape
    nb = sizes.size
    st = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    if nb >= 4:
        w, sub, tot, obj = _solve_4up(A, y, sizes)
        return w, sub, tot, obj, np.dot(A[:, tot], w)
    w = np.zeros(nb)
    sub = np.zeros(nb, dtype=np.int32)
    obj = ctypes.c_double(0.0)
    L = _lib()
    if nb == 1:
        L.orc_solve_1(M, int(sizes[0]), _d(A), lda, _d(y), _d(w), _i(sub), ctypes.byref(obj))
    elif nb == 2:
        rc = L.orc_solve_2(M, int(sizes[0]), int(sizes[1]), _d(A), lda, _d(y),
                           _d(w), _i(sub), ctypes.byref(obj))
        assert rc == 0
    else:
        rc = L.orc_solve_3(M, int(sizes[0]), int(sizes[1]), int(sizes[2]), _d(A), lda,
                           _d(y), _d(w), _i(sub), ctypes.byref(obj))
        assert rc == 0
    tot = (st + sub).astype(np.int32)
    y_rec = np.empty(M)
    L.orc_reconstruct(M, nb, _d(A), lda, _i(tot), _d(w), _d(y_rec))
    return w, sub, tot, obj.value, y_rec


# --------------------------------------------------------------------------
# _fit_voxel / fit loop
# --------------------------------------------------------------------------

def solve2_gram(A, y, dicsizes):
    """Two-block solve_exhaustive_posweights (`_2`, mfu:288-392) with the Gram terms formed by
    BLAS and the four sign branches evaluated for all (i1, i2) at once in NumPy.  Same branch
    logic, same residual formula, first minimum in (i1, i2) loop order; only the summation
    order of the dot products differs from the reference, so indices agree except where two
    residuals tie to rounding.  For shapes where the strided C / Numba Gram takes minutes
    (M = 1776, N = 2000: BASELINE config 5).  Returns (w (2,), idx_sub (2,), min_obj)."""
    A = np.asarray(A, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    N1, N2 = int(dicsizes[0]), int(dicsizes[1])
    A1, A2 = A[:, :N1], A[:, N1:N1 + N2]
    A11 = np.einsum('mi,mi->i', A1, A1)[:, None]
    A22 = np.einsum('mi,mi->i', A2, A2)[None, :]
    A12 = A1.T @ A2
    Y1 = (A1.T @ y)[:, None]
    Y2 = (A2.T @ y)[None, :]
    y_sq = float(y @ y)
    w1d = A22 * Y1 - A12 * Y2                                   # mfu:331-332
    w2d = A11 * Y2 - A12 * Y1
    Det = A11 * A22 - A12 ** 2
    with np.errstate(divide='ignore', invalid='ignore'):
        w0, w1 = w1d / Det, w2d / Det
        both = y_sq + w0 ** 2 * A11 + w1 ** 2 * A22 + 2 * (w0 * w1 * A12 - w0 * Y1 - w1 * Y2)   # mfu:341-345
        only1 = np.broadcast_to(y_sq - Y1 ** 2 / A11, A12.shape)    # w = Y1/A11
        only2 = np.broadcast_to(y_sq - Y2 ** 2 / A22, A12.shape)
    res = np.full(A12.shape, np.inf)
    pp = (w1d > 0) & (w2d > 0)
    res[pp] = both[pp]
    c1 = (w1d >= 0) & (w2d <= 0) & ~pp & (Y1 >= 0)              # mfu:348-357
    res[c1] = only1[c1]
    c2 = (w1d <= 0) & (w2d >= 0) & ~pp & ~((w1d >= 0) & (w2d <= 0)) & (Y2 >= 0)   # mfu:358-367
    res[c2] = only2[c2]
    nn = (w1d < 0) & (w2d < 0)                                  # mfu:368-381
    c3 = nn & (Y1 > 0)
    res[c3] = only1[c3]
    c4 = nn & ~(Y1 > 0) & (Y2 > 0)
    res[c4] = only2[c4]
    flat = int(np.argmin(res))                                  # first minimum in (i1, i2) order
    i1, i2 = divmod(flat, N2)
    if not res[i1, i2] < y_sq:                                  # strict < from min_obj = y_sq
        return np.zeros(2), np.zeros(2, dtype=np.int64), y_sq
    if pp[i1, i2]:
        w = np.array([w0[i1, i2], w1[i1, i2]])
    elif c1[i1, i2] or c3[i1, i2]:
        w = np.array([Y1[i1, 0] / A11[i1, 0], 0.0])
    else:
        w = np.array([0.0, Y2[0, i2] / A22[0, i2]])
    return w, np.array([i1, i2], dtype=np.int64), float(res[i1, i2])


def fit_voxel(tab, plan, y, K, csf_i, ear_i, peaks_i, maxfasc, csf_on, ear_on,
              sig_csf=None, sig_ear=None, D=None):
    """_fit_voxel (mf:340-461): returns the params row."""
    M = plan["gdir"].shape[0]
    N = tab["table"].shape[1]
    E = 0 if sig_ear is None else sig_ear.shape[1]
    csf_on, ear_on = int(bool(csf_on)), int(bool(ear_on))
    P = 1 + 2 * maxfasc + csf_on + 2 * ear_on + 2
    row = np.zeros(P)
    K = int(K)
    csf_i, ear_i = bool(csf_i), bool(ear_i)
    if K + csf_i + ear_i == 0:                                   # mf:387-388
        return row
    dicsize = K * N + csf_i + ear_i * E
    if D is None:
        D = np.zeros((M, maxfasc * N + csf_on + ear_on * E))
    sizes = []
    for k in range(K):                                           # mf:391-398
        rotate(tab, plan, peaks_i[3 * k:3 * k + 3], out=D[:, k * N:(k + 1) * N])
        sizes.append(N)
    if csf_i:                                                    # mf:401-403
        D[:, K * N] = sig_csf
        sizes.append(1)
    if ear_i:                                                    # mf:404-408
        st = K * N + csf_i
        D[:, st:st + E] = sig_ear
        sizes.append(E)
    w, sub, _tot, sos, y_rec = solve(D[:, :dicsize], y, np.asarray(sizes))
    M0 = np.sum(w)                                               # mf:420-425
    nu = w / M0 if np.abs(M0) > 0 else w
    row[0] = M0
    row[1:K + 1] = nu[:K]
    row[1 + maxfasc:1 + maxfasc + K] = sub[:K]
    if csf_i:
        row[2 * maxfasc + 1] = nu[K]
    if ear_i:
        i_ear = 2 * maxfasc + csf_on + 1
        row[i_ear] = nu[K + csf_i]
        row[i_ear + 1] = sub[K + csf_i]
    row[P - 2] = sos / M                                         # mf:446
    if M > 1 and np.std(y_rec) > 0 and np.std(y) > 0:            # mf:449-450
        row[P - 1] = np.corrcoef(y, y_rec)[0, 1] ** 2
    return row


def iso_signals(sch_mat, T2_csf, DIFF_csf, T2_ear, DIFF_ear):
    """CSF / EAR columns: mf:841-846, 918-925."""
    gam = 2 * np.pi * 42.577480e6                                # mfu:1138-1150 ('H')
    G, Delta, delta, TE = (sch_mat[:, 3], sch_mat[:, 4], sch_mat[:, 5], sch_mat[:, 6])
    b = (gam * G * delta) ** 2 * (Delta - delta / 3)
    sig_csf = np.exp(-TE / T2_csf) * np.exp(-b * DIFF_csf)
    DIFF_ear = np.atleast_1d(DIFF_ear)
    sig_ear = np.zeros((sch_mat.shape[0], DIFF_ear.size))
    for i in range(DIFF_ear.size):
        sig_ear[:, i] = np.exp(-TE / T2_ear) * np.exp(-b * DIFF_ear[i])
    return sig_csf, sig_ear


def fit_rows(tab, plan, Y, Kv, csf, ear, peaks, sig_csf=None, sig_ear=None):
    """The voxel loop of MFModel.fit (mf:1018-1028) -> params_in_mask (V,P)."""
    Kv = np.asarray(Kv).astype(int)
    csf = np.asarray(csf) > 0
    ear = np.asarray(ear) > 0
    maxfasc = int(Kv.max())
    csf_on, ear_on = bool(csf.any()), bool(ear.any())
    V = Y.shape[0]
    P = 1 + 2 * maxfasc + csf_on + 2 * ear_on + 2
    out = np.zeros((V, P))
    for i in range(V):
        out[i] = fit_voxel(tab, plan, Y[i], Kv[i], csf[i], ear[i], peaks[i], maxfasc,
                           csf_on, ear_on, sig_csf, sig_ear)
    return out


def lerp_rows(table, row_lo, row_hi, w_lo, w_hi, scale=None):
    """The scipy interp1d two-weight lerp over explicit rows, as used by rotate_atom
    (mfu:1423-1426) and rotate_atom_2Dprotocol (mfu:1678-1684): NumPy restatement of the
    mfb_lerp_rows kernel, (V, M) plan arrays -> (V, M, N)."""
    out = w_hi[..., None] * table[row_hi] + w_lo[..., None] * table[row_lo]
    if scale is not None:
        out = scale[..., None] * out
    return out


def mc_average(sim_phases, delta_mapping, gscaling, Dscaling, num_spins):
    """monte_carlo_average (mfu:2758-2812), sequential C restatement."""
    ph = np.ascontiguousarray(sim_phases, dtype=np.float64)
    dm = np.ascontiguousarray(delta_mapping, dtype=np.int64)
    gs = np.ascontiguousarray(gscaling, dtype=np.float64)
    out = np.zeros(dm.size)
    L = _lib()
    L.orc_mc_average.argtypes = [ctypes.c_longlong, ctypes.c_int, _dp, ctypes.POINTER(ctypes.c_longlong), _dp,
                                 ctypes.c_double, ctypes.c_longlong, _dp]
    L.orc_mc_average.restype = None
    L.orc_mc_average(dm.size, ph.shape[1], _d(ph), dm.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)),
                     _d(gs), float(Dscaling), int(num_spins), _d(out))
    return out
