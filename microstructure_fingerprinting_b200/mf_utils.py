"""Low-level API of the reference's mf_utils.py for the fit path, B200-backed.

Mirrors (same names, argument meaning, return types and error behaviour):
  solve_exhaustive_posweights       reference mf_utils.py:115-214
  init_PGSE_multishell_interp       reference mf_utils.py:1959-2085
  interp_PGSE_from_multishell       reference mf_utils.py:1693-1956
  import_PGSE_scheme                reference mf_utils.py:2128-2192
  get_PGSE_scheme_from_bval_bvec_dense  reference mf_utils.py:2197-2300
  get_gyromagnetic_ratio            reference mf_utils.py:1138-1150
  DT_vec_to_2Darray                 reference mf_utils.py:901-957
  loadmat                           reference mf_utils.py:3026-3087
  rotate_atom                       reference mf_utils.py:1205-1437
  rotate_atom_2Dprotocol            reference mf_utils.py:1440-1690
  rotate_scheme_mat, vrrotvec2mat   reference mf_utils.py:1153-1202, 842-858
  monte_carlo_average               reference mf_utils.py:2758-2812
  get_PGSE_from_phases              reference mf_utils.py:2815-3015
The numerical work (rotation, exhaustive search) runs in libmfb200.so on the GPU; this
module only validates, marshals and keeps the small per-study tables on the host.
There is no CPU fallback.
"""
import ctypes

import numpy as np

from . import _lib

__all__ = ["solve_exhaustive_posweights", "solve_exhaustive_posweights_batch",
           "init_PGSE_multishell_interp", "interp_PGSE_from_multishell",
           "import_PGSE_scheme", "get_PGSE_scheme_from_bval_bvec_dense",
           "get_gyromagnetic_ratio", "DT_vec_to_2Darray", "loadmat", "from_ipython",
           "rotate_atom", "rotate_atom_2Dprotocol", "rotate_scheme_mat", "vrrotvec2mat",
           "monte_carlo_average", "get_PGSE_from_phases", "solve_rotated_2Dprotocol_batch",
           "solve_rotated_batch",
           "MultiShellTable", "SchemePlan", "GpuPlan"]


# ----------------------------------------------------------------------------------
# small host utilities
# ----------------------------------------------------------------------------------

def get_gyromagnetic_ratio(element='H'):
    """Gyromagnetic ratio [rad/(s T)] (reference mf_utils.py:1138-1150)."""
    mhz_per_tesla = {'hydrogen': 42.577480e6, 'H': 42.577480e6, 'proton': 42.577480e6,
                     'carbon': 10.7084e6, 'C': 10.7084e6,
                     'phosphorus': 17.235e6, 'P': 17.235e6}
    if element not in mhz_per_tesla:
        raise ValueError('Gyromagnetic ratio for nucleus of element %s'
                         'unknown.' % element)
    return 2 * np.pi * mhz_per_tesla[element]


def from_ipython():
    """True when running under IPython (reference mf_utils.py:3090-3100)."""
    try:
        __IPYTHON__  # noqa: F821
        return True
    except NameError:
        return False


def loadmat(filename):
    """scipy.io.loadmat with nested MATLAB structs turned into dicts
    (behaviour of reference mf_utils.py:3026-3087)."""
    import scipy.io

    def is_struct(obj):
        return type(obj).__name__ == 'mat_struct'

    def to_dict(obj):
        return {k: (to_dict(v) if is_struct(v) else v) for k, v in obj.__dict__.items()}

    raw = scipy.io.loadmat(filename, struct_as_record=False, squeeze_me=True)
    return {k: (to_dict(v) if is_struct(v) else v) for k, v in raw.items()}


def DT_vec_to_2Darray(DT_vec, order):
    """(..., 6) tensor coefficients -> (..., 3, 3) symmetric arrays
    (reference mf_utils.py:901-957)."""
    if DT_vec.shape[-1] != 6:
        raise ValueError("Last dimension of input should have size 6,"
                         " detected %d." % DT_vec.shape[-1])
    # position in the 6-vector of [xx, xy, xz, yy, yz, zz]
    layouts = {'row': (0, 1, 2, 3, 4, 5), 'column': (0, 1, 3, 2, 4, 5),
               'diagonal': (0, 3, 5, 1, 4, 2)}
    if order not in layouts:
        raise ValueError("Unknown order option \"%s\"." % order)
    xx, xy, xz, yy, yz, zz = (DT_vec[..., i] for i in layouts[order])
    out = np.zeros(DT_vec.shape[:-1] + (3, 3))
    out[..., 0, 0], out[..., 1, 1], out[..., 2, 2] = xx, yy, zz
    out[..., 0, 1] = out[..., 1, 0] = xy
    out[..., 0, 2] = out[..., 2, 0] = xz
    out[..., 1, 2] = out[..., 2, 1] = yz
    return out


def import_PGSE_scheme(scheme):
    """Load / validate a PGSE scheme [gx gy gz G Delta delta TE] per row
    (reference mf_utils.py:2128-2192). Always returns a 2-D array."""
    if isinstance(scheme, str):
        with open(scheme, 'r') as f:
            header = f.readline()
        sch_mat = np.loadtxt(scheme, skiprows=1 if 'version' in header.lower() else 0)
    elif isinstance(scheme, np.ndarray):
        sch_mat = scheme
    else:
        raise TypeError("Unable to import a PGSE scheme matrix from input")
    if sch_mat.ndim == 1:
        sch_mat = sch_mat[np.newaxis, :]
    if sch_mat.shape[1] != 7:
        raise RuntimeError("Detected %s instead of expected 7 colums in"
                           " PGSE scheme matrix." % sch_mat.shape[1])
    gnorm = np.sqrt(np.sum(sch_mat[:, :3] ** 2, axis=1))
    n_bad = np.sum(np.abs(1 - gnorm[gnorm > 0]) > 1e-4)
    if n_bad > 0:
        raise ValueError("Detected %d non-zero gradients which did not have"
                         " unit norm. Please normalize." % n_bad)
    G, Delta, delta, TE = (sch_mat[:, i] for i in (3, 4, 5, 6))
    checks = [(G < 0, 'negative gradient intensity (4th column).'),
              (Delta < 0, 'negative gradient separation Delta (5th column).'),
              (delta < 0, 'negative gradient duration delta (6th column).'),
              (TE < 0, 'negative echo time TE (7th column).')]
    for bad, what in checks:
        if np.any(bad):
            raise ValueError('Detected %d sequence(s) with %s' % (np.sum(bad), what))
    if np.any(delta > Delta):
        raise ValueError('Detected %d sequence(s) in which delta (6th column)'
                         ' was greater than Delta (5th column).' % np.sum(delta > Delta))
    if np.any(TE < (Delta + delta) * 0.999):
        raise ValueError('Detected %d sequence(s) in which TE (7th column)'
                         ' was lower than Delta+delta.' % np.sum(TE < (Delta + delta)))
    return sch_mat


def get_PGSE_scheme_from_bval_bvec_dense(sch_mat_dense, bvals, bvecs, Gtol=1e-3):
    """Scheme matrix from b-values [s/mm^2] / b-vectors, snapping each gradient
    intensity to the dense sampling's shells (reference mf_utils.py:2197-2300)."""
    sch_ref = import_PGSE_scheme(sch_mat_dense)
    if isinstance(bvals, str):
        bvals = np.loadtxt(bvals)
    if isinstance(bvecs, str):
        bvecs = np.atleast_2d(np.loadtxt(bvecs))
    bvals = bvals * 1e6  # s/mm^2 -> s/m^2
    if np.ndim(bvecs) != 2:
        raise ValueError("bvecs array should have 2 dimensions,"
                         " detected %d." % np.ndim(bvecs))
    if bvecs.shape[0] != bvals.size and bvecs.shape[1] != bvals.size:
        raise ValueError("Number of b-vectors does not match number"
                         " of b-values (%d)" % bvals.size)
    if not np.all(sch_ref[0, 4:6] == sch_ref[:, 4:6]):
        raise ValueError('Detected different pairs of (Delta, delta) values'
                         ' in reference scheme matrix (note that zeros '
                         'count as values),'
                         ' which is currently not supported.')
    sch_mat = np.zeros((bvals.size, 7))
    if bvecs.shape[0] == 3:
        sch_mat[:, :3] = bvecs.transpose()
    elif bvecs.shape[1] == 3:
        sch_mat[:, :3] = bvecs
    else:
        raise ValueError("Vectors in bvecs should be 3-dimensional."
                         " However, detected no dimension with size 3.")
    gnorm = np.sqrt(np.sum(sch_mat[:, :3] ** 2, axis=1))
    nz = gnorm > 0
    sch_mat[nz, :3] = sch_mat[nz, :3] / gnorm[nz][:, np.newaxis]

    gam = get_gyromagnetic_ratio('H')
    Del, dlt, TE = sch_ref[0, 4], sch_ref[0, 5], sch_ref[0, 6]
    G = np.sqrt(bvals / (Del - dlt / 3)) / (gam * dlt)
    G_shells = np.unique(sch_ref[:, 3])
    Geff = np.zeros(bvals.shape[0])
    n_mapped = 0
    for Gs in G_shells:
        hit = np.where(np.abs(Gs - G) < Gtol)[0]
        n_mapped += hit.size
        Geff[hit] = Gs
    if n_mapped != G.size:
        raise ValueError('Mismatch between reference scheme matrix and bvals. '
                         ' Could only map %d/%d b-values (equivalently, gradient'
                         ' intensities G) from the specified bvals to the b-values'
                         ' contained in the reference scheme matrix. You may want to'
                         ' change the tolerance on gradient intensity G (currently '
                         '%g T/m).' % (n_mapped, G.size, Gtol))
    sch_mat[:, 3] = Geff
    sch_mat[:, 4:7] = np.array([Del, dlt, TE])
    return sch_mat


# ----------------------------------------------------------------------------------
# multi-shell interpolation tables
# ----------------------------------------------------------------------------------

class _ShellNodes(object):
    """Data view of one shell of the lookup table (the reference stores a
    scipy interp1d here; only its node data `.x`, `.y` is kept)."""

    def __init__(self, x, y):
        self.x = x
        self.y = y
        self._y = y


class MultiShellTable(dict):
    """Return type of init_PGSE_multishell_interp: a dict with the reference's keys
    ('scheme_DeldelTE', 'num_subs', 'Gms_un', 'interpolators') plus the flattened
    lookup table ('nodes' (R,), 'table' (R, N), 'off' (n_shells+1,)) the GPU reads."""


def _check_unit_or_zero_gradients(sch_mat):
    gnorm = np.sqrt(np.sum(sch_mat[:, 0:3] ** 2, axis=1))
    if np.any(np.abs(1 - gnorm[gnorm > 0]) > 1e-3):
        raise ValueError("Gradient directions in multi-shell scheme matrix"
                         " should all either have zero or unit norm.")


def init_PGSE_multishell_interp(sig_ms, sch_mat_ms, ordir):
    """Initialises the multi-shell lookup table (reference mf_utils.py:1959-2085).

    Per dense shell: sorted unique nodes x = |g.ordir| (first-occurrence rows), the
    near-perpendicular cluster |x - x[0]| < 1e-3 replaced by its mean, b0 shell ->
    nodes [0, 1] with identical rows.
    """
    ordir = np.asarray(ordir)
    if ordir.size != 3:
        raise ValueError("Direction of dictionary computed with dense"
                         " sampling (ordir) should have 3 entries.")
    ordir = np.squeeze(ordir).astype(np.float64)
    sch_mat_ms = np.asarray(sch_mat_ms, dtype=np.float64)
    if not np.all(np.isclose(sch_mat_ms[0, 4:7], sch_mat_ms[:, 4:7])):
        raise ValueError("Delta, delta and TE values should all be "
                         "identical in multi-shell sampling.")
    sig_ms = np.asarray(sig_ms, dtype=np.float64)
    if sig_ms.ndim == 1:
        sig_ms = sig_ms.reshape((sig_ms.size, 1))
    if sch_mat_ms.shape[0] != sig_ms.shape[0]:
        raise ValueError("Number of lines in dense multishell scheme"
                         " (%d) does not match number of signal values"
                         " per substrate (%d)." % (sch_mat_ms.shape[0], sig_ms.shape[0]))
    ordirnorm = np.sqrt((ordir ** 2).sum())
    if np.abs(1 - ordirnorm) > 1e-3:
        raise ValueError("Orientation vector of the multi-shell signal "
                         "must have unit norm. Detected %g." % (ordirnorm,))
    _check_unit_or_zero_gradients(sch_mat_ms)

    x_all = np.abs(np.dot(sch_mat_ms[:, 0:3], ordir))
    Gms_un, shell_of = np.unique(sch_mat_ms[:, 3], return_inverse=True)
    shells = []
    for s, G in enumerate(Gms_un):
        rows = np.where(shell_of == s)[0]
        if G == 0:
            same = np.all(np.isclose(sig_ms[rows, :], sig_ms[rows[0], :]), axis=0)
            if np.any(~same):
                bad = np.where(~same)[0]
                raise ValueError('Distinct signal values in provided multi-'
                                 'shell sampling for zero gradients '
                                 '(b0 acquistions), for '
                                 '%d substrate(s) [%s]' %
                                 (bad.shape[0], " ".join("{:d}".format(b) for b in bad)))
            shells.append(_ShellNodes(np.array([0.0, 1.0]),
                                      np.repeat([sig_ms[rows[0], :]], 2, axis=0)))
            continue
        xs, first = np.unique(x_all[rows], return_index=True)
        ys = sig_ms[rows, :][first, :]
        cluster = np.abs(xs - xs[0]) < 1e-3
        c = int(np.sum(cluster))
        if c > 1:
            xs = np.append(np.mean(xs[cluster]), xs[c:])
            ys = np.append(np.mean(ys[cluster, :], axis=0, keepdims=True), ys[c:, :], axis=0)
        if xs.size < 2:
            raise ValueError("Shell G=%g of the dense sampling has fewer than 2 distinct "
                             "|g.ordir| nodes; cannot interpolate." % G)
        shells.append(_ShellNodes(xs, ys))
    out = MultiShellTable()
    out['scheme_DeldelTE'] = sch_mat_ms[0, 4:7]
    out['num_subs'] = sig_ms.shape[1]
    out['Gms_un'] = Gms_un
    out['interpolators'] = shells
    out['nodes'] = np.ascontiguousarray(np.concatenate([s.x for s in shells]))
    out['table'] = np.ascontiguousarray(np.vstack([s.y for s in shells]))
    out['off'] = np.cumsum([0] + [s.x.size for s in shells]).astype(np.int32)
    return out


class SchemePlan(object):
    """Subject scheme mapped onto the dense shells (reference mf_utils.py:1786-1839)."""

    def __init__(self, msinterp, sch_mat):
        sch_mat = np.asarray(sch_mat, dtype=np.float64)
        if sch_mat.ndim != 2 or sch_mat.shape[1] < 7:
            raise ValueError("sch_mat should have shape (Nseq, 7).")
        if not np.all(np.isclose(msinterp['scheme_DeldelTE'], sch_mat[:, 4:7])):
            raise ValueError("Delta, delta and TE values should all be "
                             "identical to those in the multi-shell sampling.")
        _check_unit_or_zero_gradients(sch_mat)
        Gms = msinterp['Gms_un']
        M = sch_mat.shape[0]
        self.M = M
        self.gdir = np.ascontiguousarray(sch_mat[:, 0:3])
        self.shell_lo = np.zeros(M, dtype=np.int32)
        self.shell_hi = np.zeros(M, dtype=np.int32)
        self.gw_lo = np.ones(M)
        self.gw_hi = np.zeros(M)
        for G in np.unique(sch_mat[:, 3]):
            rows = sch_mat[:, 3] == G
            same = np.where(G == Gms)[0]
            if same.size > 0:
                self.shell_lo[rows] = self.shell_hi[rows] = same[0]
                continue
            ih = int(np.argmax(Gms > G))
            if ih == 0:
                raise ValueError("Gradient intensity %g is not in the [%g, %g]"
                                 " range spanned by the multi-shell sampling."
                                 " Extrapolation not supported." % (G, Gms[0], Gms[-1]))
            self.shell_lo[rows], self.shell_hi[rows] = ih - 1, ih
            self.gw_hi[rows] = (G - Gms[ih - 1]) / (Gms[ih] - Gms[ih - 1])
            self.gw_lo[rows] = (Gms[ih] - G) / (Gms[ih] - Gms[ih - 1])


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class GpuPlan(object):
    """Owner of one mfb_plan (lookup table + subject scheme + iso columns on one GPU)."""

    def __init__(self, msinterp, scheme_plan, sig_csf=None, sig_ear=None, device=0):
        torch = _lib.require_cuda()
        lib = _lib.load()
        self.device = int(device)
        self.M = scheme_plan.M
        self.N = int(msinterp['table'].shape[1])
        self.E = 0
        nodes = np.ascontiguousarray(msinterp['nodes'], dtype=np.float64)
        table = np.ascontiguousarray(msinterp['table'], dtype=np.float64)
        off = np.ascontiguousarray(msinterp['off'], dtype=np.int32)
        if sig_csf is not None:
            sig_csf = np.ascontiguousarray(sig_csf, dtype=np.float64)
        if sig_ear is not None:
            sig_ear = np.ascontiguousarray(sig_ear, dtype=np.float64)
            self.E = int(sig_ear.shape[1])
        with torch.cuda.device(self.device):
            self.handle = lib.mfb_plan_create(
                self.device, self.M, self.N, int(nodes.size), int(off.size - 1), _ptr(off),
                _ptr(nodes), _ptr(table), _ptr(scheme_plan.gdir), _ptr(scheme_plan.shell_lo),
                _ptr(scheme_plan.shell_hi), _ptr(scheme_plan.gw_lo), _ptr(scheme_plan.gw_hi),
                _ptr(sig_csf), _ptr(sig_ear), self.E)
        if not self.handle:
            raise _lib.MFBError("mfb_plan_create failed: " + _lib.last_error())

    def close(self):
        if getattr(self, 'handle', None):
            _lib.load().mfb_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def rotate(self, dirs):
        """dirs (V, 3) -> torch.cuda tensor (V, M, N)."""
        torch = _lib.require_cuda()
        dirs = np.ascontiguousarray(np.atleast_2d(dirs), dtype=np.float64)
        V = dirs.shape[0]
        dev = torch.device('cuda', self.device)
        d_dirs = torch.from_numpy(dirs).to(dev)
        out = torch.empty((V, self.M, self.N), dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        rc = _lib.load().mfb_rotate_multishell(self.handle, V, d_dirs.data_ptr(), out.data_ptr(),
                                               self.N, st)
        _lib.check(rc, "mfb_rotate_multishell")
        return out

    def fit_host(self, y, peaks, K, csf, ear, maxfasc, csf_on, ear_on, flags=0):
        """Host arrays in, params rows (V, P) out (mfb_fit_host)."""
        y = np.ascontiguousarray(y, dtype=np.float64)
        V = y.shape[0]
        K = np.ascontiguousarray(K, dtype=np.int32)
        peaks = None if maxfasc == 0 else np.ascontiguousarray(peaks, dtype=np.float64)
        csf = None if csf is None else np.ascontiguousarray(csf, dtype=np.uint8)
        ear = None if ear is None else np.ascontiguousarray(ear, dtype=np.uint8)
        P = 1 + 2 * maxfasc + int(csf_on) + 2 * int(ear_on) + 2
        out = np.zeros((V, P))
        rc = _lib.load().mfb_fit_host(self.handle, V, _ptr(y), _ptr(peaks), _ptr(K), _ptr(csf),
                                      _ptr(ear), int(maxfasc), int(csf_on), int(ear_on), _ptr(out),
                                      int(flags))
        _lib.check(rc, "mfb_fit_host")
        return out

    def fit_volume(self, data, voxel_offset, meas_stride, peaks, K, csf, ear, maxfasc, csf_on,
                   ear_on, out, flags=0):
        """ROI voxels scattered in the host volume `data` (float64 / float32, any strides):
        voxel v's signal starts at element voxel_offset[v] of data's buffer, measurements are
        meas_stride elements apart.  The library gathers, uploads, fits and writes the params
        rows into out (V, P) (mfb_fit_volume); nothing is copied on the Python side."""
        V = int(voxel_offset.shape[0])
        P = 1 + 2 * maxfasc + int(csf_on) + 2 * int(ear_on) + 2
        assert out.shape == (V, P) and out.dtype == np.float64 and out.flags.c_contiguous
        assert voxel_offset.dtype == np.int64 and voxel_offset.flags.c_contiguous
        dtype = {np.dtype(np.float64): _lib.MFB_F64, np.dtype(np.float32): _lib.MFB_F32}[data.dtype]
        K = np.ascontiguousarray(K, dtype=np.int32)
        peaks = None if maxfasc == 0 else np.ascontiguousarray(peaks, dtype=np.float64)
        csf = None if csf is None else np.ascontiguousarray(csf, dtype=np.uint8)
        ear = None if ear is None else np.ascontiguousarray(ear, dtype=np.uint8)
        rc = _lib.load().mfb_fit_volume(self.handle, V, ctypes.c_void_p(data.ctypes.data), dtype,
                                        _ptr(voxel_offset), 0, int(meas_stride), _ptr(peaks), _ptr(K),
                                        _ptr(csf), _ptr(ear), int(maxfasc), int(csf_on), int(ear_on),
                                        _ptr(out), int(flags))
        _lib.check(rc, "mfb_fit_volume")
        return out

    def fit_device(self, y, peaks, K, csf, ear, maxfasc, csf_on, ear_on, flags=0, out=None):
        """torch.cuda tensors in / out (mfb_fit); inputs must live on this plan's GPU."""
        torch = _lib.require_cuda()
        V = y.shape[0]
        P = 1 + 2 * maxfasc + int(csf_on) + 2 * int(ear_on) + 2
        if out is None:
            out = torch.empty((V, P), dtype=torch.float64, device=y.device)
        st = torch.cuda.current_stream(y.device).cuda_stream

        def p(t):
            return None if t is None else t.data_ptr()
        rc = _lib.load().mfb_fit(self.handle, V, p(y), p(peaks), p(K), p(csf), p(ear), int(maxfasc),
                                 int(csf_on), int(ear_on), p(out), int(flags), st)
        _lib.check(rc, "mfb_fit")
        return out

    def stats(self):
        buf = np.zeros(8)
        _lib.check(_lib.load().mfb_fit_stats(self.handle, _ptr(buf), 8), "mfb_fit_stats")
        return buf


def _plan_for(msinterp, sch_mat, device=0):
    """GpuPlan cached on the table object, keyed by the scheme's bytes."""
    sch_mat = np.ascontiguousarray(sch_mat, dtype=np.float64)
    cache = msinterp.__dict__.setdefault('_gpu_plans', {}) if isinstance(msinterp, MultiShellTable) \
        else {}
    key = (device, sch_mat.shape, sch_mat.tobytes())
    if key not in cache:
        if len(cache) >= 4:
            cache.clear()
        cache[key] = GpuPlan(msinterp, SchemePlan(msinterp, sch_mat), device=device)
    return cache[key]


def interp_PGSE_from_multishell(sch_mat, newdir, sig_ms=None, sch_mat_ms=None, ordir=None,
                                msinterp=None):
    """Single-fascicle PGSE signal rotated to `newdir` by interpolation in the dense
    multi-shell sampling (reference mf_utils.py:1693-1956).

    Returns an (Nseq, Nsub) array passed through numpy.squeeze, like the reference.
    Extension: `newdir` of shape (V, 3) returns (V, Nseq, Nsub).
    """
    if msinterp is None:
        if sig_ms is None or sch_mat_ms is None or ordir is None:
            raise ValueError("If msinterp is not specified, sig_ms, "
                             "sch_mat_ms and ordir must all be specified.")
        msinterp = init_PGSE_multishell_interp(sig_ms, sch_mat_ms, ordir)
    else:
        if msinterp['Gms_un'].size != len(msinterp['interpolators']):
            raise ValueError("msinterp['Gms_un'] has size %d vs expected %d to match "
                             "len(msinterp['interpolators'])"
                             % (msinterp['Gms_un'].size, len(msinterp['interpolators'])))
        if msinterp['interpolators'][0].y.shape[1] != msinterp['num_subs']:
            raise ValueError("Inconsistency in msinterp regarding number of substrates. "
                             "Make sure the interpolator was initialized"
                             " on the right dictionary.")
    newdir = np.asarray(newdir, dtype=np.float64)
    batched = newdir.ndim == 2 and newdir.shape[0] != 1 and newdir.shape[1] == 3 and newdir.size > 3
    if not batched:
        if newdir.size != 3:
            raise ValueError("Direction of fascicle for new signal (newdir)"
                             " should have 3 entries.")
        newdir = newdir.reshape(1, 3)
    norms = np.sqrt((newdir ** 2).sum(axis=1))
    if np.any(np.abs(1 - norms) > 1e-3):
        raise ValueError("Orientation vector of the new signal must have unit norm. Detected"
                         " %g." % (norms[np.argmax(np.abs(1 - norms))],))
    plan = _plan_for(msinterp, np.asarray(sch_mat, dtype=np.float64))
    out = plan.rotate(newdir).cpu().numpy()
    if batched:
        return out
    return np.squeeze(out[0])


# ----------------------------------------------------------------------------------
# exhaustive combinatorial NNLS
# ----------------------------------------------------------------------------------

def solve_exhaustive_posweights_batch(A, Y, dicsizes, device=0, return_device=False, exact=False):
    """Batched form of solve_exhaustive_posweights: Y is (V, M); A is (M, Ntot) (shared by
    all voxels) or (V, M, Ntot).  Returns (w (V,K), ind_subdic (V,K), ind_totdic (V,K),
    min_obj (V,), y_recons (V,M)).  `exact=True` forces the reference-order tier for every
    voxel (verification; the results are the same by construction)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device('cuda', device)
    sizes = np.ascontiguousarray(np.asarray(dicsizes).astype(np.int64))
    nb = int(sizes.size)

    def to_dev(x):
        if isinstance(x, torch.Tensor):
            return x.to(device=dev, dtype=torch.float64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
    dA, dY = to_dev(A), to_dev(Y)
    V, M = dY.shape
    ntot = int(sizes.sum())
    if dA.dim() == 2:
        strideA = 0
        assert dA.shape == (M, ntot)
    else:
        assert dA.shape == (V, M, ntot)
        strideA = M * ntot
    w = torch.zeros((V, nb), dtype=torch.float64, device=dev)
    sub = torch.zeros((V, nb), dtype=torch.int32, device=dev)
    obj = torch.zeros((V,), dtype=torch.float64, device=dev)
    yrec = torch.zeros((V, M), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        rc = lib.mfb_solve_batch(device, V, M, nb, _ptr(sizes), dA.data_ptr(), ntot, strideA,
                                 dY.data_ptr(), w.data_ptr(), sub.data_ptr(), obj.data_ptr(),
                                 yrec.data_ptr(), 1 if exact else 0, st)
    _lib.check(rc, "mfb_solve_batch")
    starts = torch.from_numpy(np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.int32)).to(dev)
    tot = sub + starts[None, :]
    if return_device:
        return w, sub, tot, obj, yrec
    return (w.cpu().numpy(), sub.cpu().numpy(), tot.cpu().numpy(), obj.cpu().numpy(),
            yrec.cpu().numpy())


def solve_exhaustive_posweights(A, y, dicsizes, printmsg=None):
    """Solves NNLS with 1-sparsity constraints combinatorially
    (reference mf_utils.py:115-214): min_{w>=0} ||A w - y||^2 with exactly one active
    column per sub-dictionary.

    Returns (w_nneg (K,), ind_atoms_subdic (K,) int32, ind_atoms_totdic (K,) int32,
    min_obj float, y_recons (M,)); index arrays are int64 for 4 or more blocks, as in
    the reference.
    """
    if printmsg is not None:
        print(printmsg, end="")
    assert isinstance(A, np.ndarray), "A should be a NumPy ndarray"
    assert A.ndim == 2, "A should be a 2D array"
    assert not np.any(np.all(A == 0, axis=0)), "All-zero columns detected in A"
    assert isinstance(y, np.ndarray), "y should be a NumPy ndarray"
    assert A.size > 0 and y.size > 0, "A and y should not be empty arrays"
    msg = ("Number of rows in A (%d) should match number of elements in y (%d)"
           % (A.shape[0], y.size))
    assert A.shape[0] == y.size, msg
    assert isinstance(dicsizes, np.ndarray), "dicsizes should be a NumPy ndarray"
    assert np.all(dicsizes > 0), "All entries of dicsizes should be > 0"
    msg = ("Number of columns of A (%d) does not equal sum of size of "
           "sub-matrices in diclengths array (%d)" % (A.shape[1], np.sum(dicsizes)))
    assert A.shape[1] == np.sum(dicsizes), msg

    w, sub, tot, obj, yrec = solve_exhaustive_posweights_batch(
        A.astype(np.float64), y.astype(np.float64).reshape(1, -1), dicsizes)
    idt = np.int64 if dicsizes.size >= 4 else np.int32
    return (w[0], sub[0].astype(idt), tot[0].astype(idt), float(obj[0]), yrec[0])


# ----------------------------------------------------------------------------------
# HARDI / AxCaliber rotations of an M-row dictionary (low-level API)
# ----------------------------------------------------------------------------------

def _lerp_rows(table, row_lo, row_hi, w_lo, w_hi, scale=None, device=0, return_device=False):
    """out[v, m, :] = scale * (w_hi * table[row_hi] + w_lo * table[row_lo]) on the GPU
    (mfb_lerp_rows).  Plan arrays are (V, M); returns a NumPy array (V, M, N)."""
    torch = _lib.require_cuda()
    dev = torch.device('cuda', device)
    table = np.ascontiguousarray(table, dtype=np.float64)
    V, M = row_lo.shape
    N = table.shape[1]

    def up(x, dt):
        return torch.from_numpy(np.ascontiguousarray(x, dtype=dt)).to(dev)
    d_t = up(table, np.float64)
    d_rl, d_rh = up(row_lo, np.int32), up(row_hi, np.int32)
    d_wl, d_wh = up(w_lo, np.float64), up(w_hi, np.float64)
    d_sc = None if scale is None else up(scale, np.float64)
    out = torch.empty((V, M, N), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        rc = _lib.load().mfb_lerp_rows(device, V, M, N, d_t.data_ptr(), d_rl.data_ptr(), d_rh.data_ptr(),
                                       d_wl.data_ptr(), d_wh.data_ptr(),
                                       None if d_sc is None else d_sc.data_ptr(), out.data_ptr(), N, st)
    _lib.check(rc, "mfb_lerp_rows")
    return out if return_device else out.cpu().numpy()


def _lerp_plan(xs, x_new):
    """scipy interp1d._call_linear bookkeeping for sorted nodes xs: interval index clipped
    to [1, n-1] (linear extrapolation outside) and the two separately rounded weights."""
    j = np.clip(np.searchsorted(xs, x_new), 1, xs.size - 1)
    x_lo, x_hi = xs[j - 1], xs[j]
    return j - 1, j, (x_hi - x_new) / (x_hi - x_lo), (x_new - x_lo) / (x_hi - x_lo)


def _rotate_atom_plan(sig, sch_mat, ordir, dirs, DIFF, S0, warnings=True):
    """Interpolation plan of rotate_atom (reference mf_utils.py:1205-1437) for a batch of
    directions dirs (V, 3): lookup table (raw rows, then per shell the nodes' rows including
    the free-diffusion anchor and the merged near-perpendicular cluster) and (row_lo, row_hi,
    w_lo, w_hi), each (V, M)."""
    V, M = dirs.shape[0], sig.shape[0]
    gam = get_gyromagnetic_ratio('H')
    gnorm = np.sqrt((sch_mat[:, 0:3] ** 2).sum(axis=1, keepdims=True))
    gnorm[gnorm == 0] = np.inf
    gunit = sch_mat[:, 0:3] / gnorm
    x_or = np.abs(np.dot(gunit, ordir / np.sqrt((ordir ** 2).sum())))
    x_new = np.stack([np.abs(np.dot(gunit, d / np.sqrt((d ** 2).sum()))) for d in dirs])  # (V, M)
    bvals = (gam * sch_mat[:, 3] * sch_mat[:, 5]) ** 2 * (sch_mat[:, 4] - sch_mat[:, 5] / 3)
    shells, shell_of = np.unique(sch_mat[:, 3:6], return_inverse=True, axis=0)
    shell_of = np.ravel(shell_of)

    rows = [sig]                      # table: the raw rows first (b0 rows are copied), then shells
    n_rows = M
    row_lo = np.zeros((V, M), dtype=np.int32)
    row_hi = np.zeros((V, M), dtype=np.int32)
    w_lo = np.ones((V, M))
    w_hi = np.zeros((V, M))
    for s in range(shells.shape[0]):
        ind = np.where(shell_of == s)[0]
        bval = bvals[ind[0]]
        if bval == 0:
            row_lo[:, ind] = row_hi[:, ind] = ind
            continue
        if ind.size < 2:
            raise ValueError("Fewer than 2 identical (G, Del, del) triplets "
                             "detected for triplet %d/%d (%g, %g, %g), b=%g"
                             " s/mm^2, probably not a HARDI shell." %
                             (s + 1, shells.shape[0], shells[s, 0], shells[s, 1], shells[s, 2], bval / 1e6))
        if ind.size < 10 and warnings:
            print("WARNING: rotate_atom: fewer than 10 data points detected"
                  " for acquisition parameters (G, Del, del) %d/%d "
                  "(%g, %g, %g), b=%g s/mm^2.\nQuality of approximation may be poor."
                  % (s + 1, shells.shape[0], shells[s, 0], shells[s, 1], shells[s, 2], bval / 1e6))
        same_S0 = np.all(np.isclose(S0[ind, :], S0[ind[0], :]), axis=0)
        if np.any(~same_S0):
            bad = np.where(~same_S0)[0]
            raise ValueError('Distinct values in provided S0 image '
                             'for shell  %d/%d (b=%g s/mm^2) for %d substrate(s) [%s]' %
                             (s + 1, shells.shape[0], bval / 1e6, bad.shape[0],
                              " ".join("{:d}".format(b) for b in bad)))
        xs, first = np.unique(x_or[ind], return_index=True)
        ys = sig[ind, :][first, :]
        if not np.any(xs == 1):
            xs = np.append(xs, [1])
            ys = np.append(ys, np.exp(-bval * DIFF) * S0[ind[0], :][np.newaxis, :], axis=0)
        cluster = np.abs(xs - xs[0]) < 1e-3
        c = int(np.sum(cluster))
        if c > 1:
            xs = np.append(np.mean(xs[cluster]), xs[c:])
            ys = np.append(np.mean(ys[cluster, :], axis=0, keepdims=True), ys[c:, :], axis=0)
        lo, hi, wl, wh = _lerp_plan(xs, x_new[:, ind])
        row_lo[:, ind], row_hi[:, ind] = n_rows + lo, n_rows + hi
        w_lo[:, ind], w_hi[:, ind] = wl, wh
        rows.append(ys)
        n_rows += ys.shape[0]
    return np.vstack(rows), row_lo, row_hi, w_lo, w_hi


def rotate_atom(sig, sch_mat, ordir, newdir, DIFF, S0, warnings=True):
    """Rotate HARDI signals of single fascicles from `ordir` to `newdir`
    (reference mf_utils.py:1205-1437).

    Per (G, Delta, delta) shell of `sch_mat`: nodes = sorted unique |g.ordir| (first
    occurrences), the free-diffusion point (1, exp(-b*DIFF)*S0) appended unless a node
    equals 1, the near-perpendicular cluster merged into its mean, linear interpolation /
    extrapolation at |g.newdir|; b0 rows are copied.  Returns an array shaped like `sig`.
    Extension: `newdir` of shape (V, 3) returns (V,) + sig.shape.
    """
    assert isinstance(sig, np.ndarray), "Input sig should be a NumPy ndarray"
    assert isinstance(sch_mat, np.ndarray), "Input sch_mat should be a NumPy ndarray"
    assert isinstance(ordir, np.ndarray), "Input ordir should be a NumPy ndarray"
    assert isinstance(newdir, np.ndarray), "Input newdir should be a NumPy ndarray"
    sig_shape = sig.shape
    if sig.ndim == 1:
        sig = sig.reshape((sig.size, 1))
    if not isinstance(DIFF, np.ndarray):
        DIFF = np.array([[DIFF]])
    assert isinstance(S0, np.ndarray), "Input S0 should be a NumPy ndarray"
    if S0.ndim == 1:
        S0 = S0[:, np.newaxis]
    if sch_mat.shape[1] < 6:
        raise ValueError('sch_mat must be a N-by-6 or7 matrix')
    if sch_mat.shape[0] != sig.shape[0]:
        raise ValueError('sch_mat and sig must have the same number of rows')
    assert sig.shape == S0.shape, "The S0 matrix should have the same size as the signal matrix"

    batched = newdir.ndim == 2 and newdir.shape[1] == 3 and newdir.size > 3
    dirs = newdir.reshape(-1, 3) if batched else newdir.reshape(1, 3)
    V = dirs.shape[0]
    table, row_lo, row_hi, w_lo, w_hi = _rotate_atom_plan(sig, sch_mat, ordir, dirs, DIFF, S0, warnings)
    out = _lerp_rows(table, row_lo, row_hi, w_lo, w_hi)
    if np.any(np.isnan(out)):
        raise ValueError('Nan detected after rotation of substrate(s).')
    if batched:
        return out.reshape((V,) + sig_shape)
    return np.reshape(out[0], sig_shape)


def vrrotvec2mat(rotax, theta):
    """Rotation matrix of angle theta about the unit axis rotax (reference mf_utils.py:842-858)."""
    if rotax.size != 3:
        raise ValueError("rotation axis should be a 3-element NumPy array")
    if ~np.isclose(np.sum(rotax ** 2), 1):
        raise ValueError("rotation axis should have unit norm")
    s, c = np.sin(theta), np.cos(theta)
    t = 1 - c
    x, y, z = rotax[0], rotax[1], rotax[2]
    return np.array([[t * x * x + c, t * x * y - s * z, t * x * z + s * y],
                     [t * x * y + s * z, t * y * y + c, t * y * z - s * x],
                     [t * x * z - s * y, t * y * z + s * x, t * z * z + c]])


def rotate_scheme_mat(sch_mat, cyldir1, cyldir2):
    """Scheme matrix seen from a fascicle rotated from cyldir1 to cyldir2
    (reference mf_utils.py:1153-1202): DWI(fasc(dir2); sch) = DWI(fasc(dir1); sch_eff)."""
    if cyldir1.size != 3 or cyldir2.size != 3:
        raise ValueError("cyldir1 and cyldir2 should be 3-elements NumPy arrays.")
    if (~np.isclose(np.sum(cyldir1 ** 2), 1) or ~np.isclose(np.sum(cyldir2 ** 2), 1)):
        raise ValueError("cyldir1 and cyldir2 should have unit norm.")
    axis = np.cross(cyldir1, cyldir2)
    n2 = np.sum(axis ** 2)
    if not n2 > 0:
        return sch_mat
    axis = axis / np.sqrt(n2)
    ang = np.arccos(np.dot(cyldir1, cyldir2))
    g = sch_mat[:, :3] @ vrrotvec2mat(axis, -ang).T
    g[np.abs(g) <= np.finfo(float).eps] = 0
    gn = np.sqrt(np.sum(g ** 2, axis=1, keepdims=True))
    nz = np.squeeze(gn > 0)
    g[nz, :] = g[nz, :] / gn[nz, :]
    return np.hstack((g, sch_mat[:, 3:])) if sch_mat.shape[1] > 3 else g


def _perp_frame(sch_mat, fascdir):
    """Unit in-plane directions, perpendicular and parallel gradient intensities of every
    sequence in the frame of a fascicle along fascdir (reference mf_utils.py:1502-1517)."""
    eff = rotate_scheme_mat(sch_mat, np.array([0, 0, 1]), fascdir)
    # NB: like the reference, the in-plane directions are normalised IN PLACE in `eff`; when
    # fascdir is the z axis `eff` is `sch_mat` itself, so the (private) scheme copy is
    # normalised and a later frame computed from it sees unit in-plane norms
    # (reference mf_utils.py:1503-1508, 1530-1535).  Reproduced for parity.
    gp = eff[:, 0:2]
    nrm = np.sqrt(np.sum(gp ** 2, axis=1))
    nz = nrm > 0
    gp[nz, :] = eff[nz, 0:2] / nrm[nz][:, np.newaxis]
    G = sch_mat[:, 3]
    return gp, nz, G * nrm, np.abs(eff[:, 2]) * G


def _opposite_pairs(dirs_un):
    """Index pairs (i, j) of unique in-plane directions with d_i . d_j ~ -1."""
    return np.where(np.isclose(dirs_un @ dirs_un.T, -1))


class _Protocol2D(object):
    """Direction-independent part of rotate_atom_2Dprotocol (reference
    mf_utils.py:1440-1690): reference frame, lookup table of perpendicular signals, and per
    (Delta, delta) pair the sorted signed-G nodes of every reference line."""

    def __init__(self, sch_mat, refdir, DIFF):
        if np.any(sch_mat[:, 2] != 0):
            raise ValueError("Use the original schemefile with zeros for gz.\n"
                             "Specify the reference and new orientations separately.")
        sch = np.array(sch_mat, dtype=np.float64, copy=True)      # private: never touch the caller's
        self.gam = get_gyromagnetic_ratio('H')
        self.G, self.Delta, self.delta = sch[:, 3].copy(), sch[:, 4].copy(), sch[:, 5].copy()
        self.is_b0, self.is_b = (self.G == 0), (self.G != 0)
        self.M = M = sch.shape[0]
        self.DIFF = DIFF
        g_ref, nz_ref, Gperp_ref, Gpar_ref = _perp_frame(sch, refdir)
        # NB: when refdir is the z axis _perp_frame normalised the in-plane directions of `sch`
        # in place (like the reference); every new frame is computed from that state
        assert np.all(np.isclose(self.G ** 2, Gperp_ref ** 2 + Gpar_ref ** 2)), \
            "Inconsistency in parallel and perpendicular gradient components for reference fasicle."
        self.S_par_ref = np.exp(-(self.gam * self.delta * Gpar_ref) ** 2 * (self.Delta - self.delta / 3) * DIFF)
        assert np.all(np.isclose(self.S_par_ref[self.is_b0], 1)), \
            "Reference fascicle: parallel signal should  be one in b0 sequences."
        self.sch = sch
        # unique laboratory gradient directions: sequences sharing one behave identically
        self.lab_un, self.lab_id = np.unique(sch[:, 0:3], return_inverse=True, axis=0)
        self.lab_id = np.ravel(self.lab_id)
        pairs, pair_of = np.unique(sch[:, 4:6], return_inverse=True, axis=0)
        self.pair_of = np.ravel(pair_of)
        self.n_pairs = pairs.shape[0]
        self.pairs = []
        self.b0_row = np.full(self.n_pairs, -1, dtype=np.int64)    # table row of the shell's b0 signal
        self.extra = []                                            # (pair, rows to average)
        for ip in range(self.n_pairs):
            ind = np.where(self.pair_of == ip)[0]
            ref_un, ref_id = np.unique(g_ref[ind, :], return_inverse=True, axis=0)
            ref_id = np.ravel(ref_id)
            assert ref_un.shape[0] in (3, 5), (
                "Problem at delta pair %d/%d: found %d unique gradient directions in plane perpendicular"
                " to reference fascicle (including b0 zero dirs)." % (ip + 1, self.n_pairs, ref_un.shape[0]))
            ig, ig_op = _opposite_pairs(ref_un)
            assert ig.size in (2, 4), (
                "Problem at delta pair %d/%d: found %d instead of 4 (2x2, redundant) pairs of opposite "
                "directions in plane perpendicular to reference fascicle." % (ip + 1, self.n_pairs, ig.size))
            lines = {}
            for k in range(ig.size):          # nodes of the line through ref_un[ig[k]], signed along it
                sel_ref = ind[(ref_id == ig[k]) | (ref_id == ig_op[k])]
                Gs_ref = Gperp_ref[sel_ref] * np.sign(g_ref[sel_ref, :] @ ref_un[ig[k], :])
                order = np.argsort(Gs_ref, kind="mergesort")        # interp1d(assume_sorted=False)
                lines[int(ig[k])] = (Gs_ref[order], sel_ref[order])
            b0s = np.where(self.is_b0 & (self.pair_of == ip))[0]
            if b0s.size == 1:
                self.b0_row[ip] = b0s[0]
            elif b0s.size > 1:
                self.b0_row[ip] = M + len(self.extra)
                self.extra.append(b0s)
            self.pairs.append({"ind": ind, "ref_un": ref_un, "lines": lines,
                               "lab": np.unique(self.lab_id[ind])})

    def table(self, sig):
        """Lookup-table rows: perpendicular reference signals, then the mean b0 signal of the
        (Delta, delta) pairs with several b0 sequences."""
        S_perp_ref = sig / self.S_par_ref[:, np.newaxis]
        if not self.extra:
            return S_perp_ref
        return np.vstack([S_perp_ref] + [np.mean(sig[rows, :], axis=0) for rows in self.extra])

    def frames(self, newdirs):
        """In-plane unit directions gp (V, U, 2), in-plane norms nrm (V, U) and |g_z| (V, U) of
        the U unique laboratory directions seen from fascicles along newdirs (V, 3)
        (rotate_scheme_mat + the in-plane normalisation of the reference, mfu:1153-1202,
        1530-1535)."""
        d = np.asarray(newdirs, dtype=np.float64)
        V = d.shape[0]
        if np.any(~np.isclose(np.sum(d ** 2, axis=1), 1)):
            raise ValueError("cyldir1 and cyldir2 should have unit norm.")
        U = self.lab_un.shape[0]
        ax = np.stack([-d[:, 1], d[:, 0], np.zeros(V)], axis=1)          # cross(z, newdir)
        n2 = np.sum(ax ** 2, axis=1)
        rot = n2 > 0
        g = np.broadcast_to(self.lab_un[np.newaxis], (V, U, 3)).copy()
        if np.any(rot):
            a = ax[rot] / np.sqrt(n2[rot])[:, np.newaxis]
            ang = -np.arccos(d[rot, 2])
            s_, c_ = np.sin(ang), np.cos(ang)
            t_ = 1 - c_
            x, y, z = a[:, 0], a[:, 1], a[:, 2]
            R = np.empty((a.shape[0], 3, 3))
            R[:, 0, 0], R[:, 0, 1], R[:, 0, 2] = t_ * x * x + c_, t_ * x * y - s_ * z, t_ * x * z + s_ * y
            R[:, 1, 0], R[:, 1, 1], R[:, 1, 2] = t_ * x * y + s_ * z, t_ * y * y + c_, t_ * y * z - s_ * x
            R[:, 2, 0], R[:, 2, 1], R[:, 2, 2] = t_ * x * z - s_ * y, t_ * y * z + s_ * x, t_ * z * z + c_
            gr = np.matmul(self.lab_un[np.newaxis], np.transpose(R, (0, 2, 1)))   # (Vr, U, 3)
            gr[np.abs(gr) <= np.finfo(float).eps] = 0
            gn = np.sqrt(np.sum(gr ** 2, axis=2, keepdims=True))
            np.divide(gr, gn, out=gr, where=gn > 0)
            g[rot] = gr
        gp = g[:, :, 0:2].copy()
        nrm = np.sqrt(np.sum(gp ** 2, axis=2))
        np.divide(gp, nrm[:, :, np.newaxis], out=gp, where=(nrm > 0)[:, :, np.newaxis])
        return gp, nrm, np.abs(g[:, :, 2])

    def _static_tables(self):
        """Per-measurement constants of the plan (they do not depend on the new directions): class
        (pair, laboratory direction) of every b > 0 sequence, and the reference lines of all pairs
        concatenated (sorted signed-G nodes and their table rows)."""
        if getattr(self, "_static", None) is not None:
            return self._static
        M = self.M
        m_class = np.full(M, -1, dtype=np.int32)          # -1: b0 sequence (keeps its own row)
        m_b0row = np.full(M, -1, dtype=np.int32)
        classes = []                                       # (pair, position in the pair's lab list, measurements)
        line_id, line_off, line_nodes, line_rows = {}, [0], [], []
        for ip, pr in enumerate(self.pairs):
            ind, lab = pr["ind"], pr["lab"]
            for j in range(lab.size):
                m_rows = ind[self.lab_id[ind] == lab[j]]
                if not np.any(self.is_b[m_rows]):
                    continue                               # the b0 "direction"
                mb = m_rows[self.is_b[m_rows]]
                m_class[mb] = len(classes)
                m_b0row[mb] = self.b0_row[ip]
                classes.append((ip, j, mb))
            for im in sorted(pr["lines"]):
                nodes, node_rows = pr["lines"][im]
                line_id[(ip, im)] = len(line_off) - 1
                line_nodes.append(nodes)
                line_rows.append(node_rows)
                line_off.append(line_off[-1] + nodes.size)
        self._static = {
            "m_class": m_class, "m_b0row": m_b0row, "classes": classes, "line_id": line_id,
            "line_off": np.asarray(line_off, dtype=np.int32),
            "line_nodes": np.ascontiguousarray(np.concatenate(line_nodes)),
            "line_rows": np.ascontiguousarray(np.concatenate(line_rows).astype(np.int32)),
            "m_lab": np.ascontiguousarray(self.lab_id.astype(np.int32)),
            "m_isb0": np.ascontiguousarray(self.is_b0.astype(np.uint8)),
            "m_gd": np.ascontiguousarray(self.gam * self.delta), "m_tt": np.ascontiguousarray(self.Delta - self.delta / 3)}
        return self._static

    def classify(self, newdirs, strict=True):
        """Direction-dependent decisions of the plan, per direction and per class (pair, laboratory
        direction) instead of per sequence: kind (0 = no rule reaches the class, 2 = the gradient became
        parallel to the new fascicle: mean b0 signal of the shell, 3 = interpolate along a reference
        line), the reference line and the sign of the perpendicular gradient along it; plus the in-plane
        norm and |g_z| of every unique laboratory direction and the directions that break the
        protocol's assumptions (`bad`; with strict=True they raise the reference's AssertionError).
        Everything of size (V, M) is left to the expansion (`plan` on the host, mfb_plan2d on the GPU)."""
        st = self._static_tables()
        gp, nrm, gz = self.frames(newdirs)
        V = gp.shape[0]
        C = len(st["classes"])
        kind = np.zeros((V, C), dtype=np.uint8)
        line = np.zeros((V, C), dtype=np.int32)
        sgn_c = np.zeros((V, C))
        bad = np.zeros(V, dtype=bool)
        c = 0
        for ip, pr in enumerate(self.pairs):
            ind, lab, ref_un = pr["ind"], pr["lab"], pr["ref_un"]
            P = lab.size
            rows = gp[:, lab, :]                                         # (V, P, 2)
            # np.unique(axis=0) of each voxel's rows: lexicographic sort + exact duplicates
            order = np.lexsort((rows[:, :, 1], rows[:, :, 0]), axis=-1)   # (V, P)
            srt = np.take_along_axis(rows, order[:, :, np.newaxis], axis=1)
            first = np.ones((V, P), dtype=bool)
            first[:, 1:] = np.any(srt[:, 1:, :] != srt[:, :-1, :], axis=2)
            uid_sorted = np.cumsum(first, axis=1) - 1                    # unique index of sorted row
            n_un = uid_sorted[:, -1] + 1
            bad_here = (n_un != 3) & (n_un != 5)
            bad |= bad_here
            if strict and np.any(bad_here):
                v = int(np.where(bad_here)[0][0])
                raise AssertionError(
                    "Problem at delta pair %d/%d: found %d unique gradient directions in plane perpendicular to "
                    "new fascicle (including b0 zero dirs)." % (ip + 1, self.n_pairs, n_un[v]))
            new_id = np.empty((V, P), dtype=np.int64)                    # unique index of lab dir lab[j]
            np.put_along_axis(new_id, order, uid_sorted, axis=1)
            # opposite pairs among the unique rows; a line is named by its lower unique index
            dots = np.einsum('vik,vjk->vij', srt, srt)
            opp = np.isclose(dots, -1) & first[:, :, np.newaxis] & first[:, np.newaxis, :]
            opp &= np.triu(np.ones((P, P), dtype=bool), 1)[np.newaxis]
            n_lines = opp.sum(axis=(1, 2))
            bad_here = (n_lines != 1) & (n_lines != 2)
            bad |= bad_here
            if strict and np.any(bad_here):
                v = int(np.where(bad_here)[0][0])
                raise AssertionError(
                    "Problem at delta pair %d/%d: found %d instead of 2 pairs of opposite directions, in plane "
                    " perpendicular to new fascicle." % (ip + 1, self.n_pairs, n_lines[v]))
            # per sorted row: is it the lower / upper member of a line?
            lower = opp.any(axis=2)                                      # (V, P) sorted position i of (i, j)
            upper = opp.any(axis=1)
            partner_of_upper = np.argmax(opp, axis=1)                    # for sorted j: its i
            ar = np.arange(V)
            for j in range(P):
                m_rows = ind[self.lab_id[ind] == lab[j]]
                if not np.any(self.is_b[m_rows]):
                    continue                                             # the b0 "direction"
                assert st["classes"][c][0] == ip and st["classes"][c][1] == j
                mb = st["classes"][c][2]
                # gradients that became parallel to the new fascicle: mean b0 signal of the shell
                van = ~(nrm[:, lab[j]] > 0)                              # (V,)
                if np.any(van):
                    assert self.b0_row[ip] >= 0, (
                        "Shell %d/%d: some new line directions are completely parallel to new fascicle, "
                        "implying free diffusion. However, no b0 measurements in the reference signal are "
                        "available for this shell. We therefore can't properly scale the new signal."
                        % (ip + 1, self.n_pairs))
                    kind[van, c] = 2
                # the first sorted row of this direction's unique class carries the line flags
                rep = np.argmax(uid_sorted == new_id[:, j][:, np.newaxis], axis=1)
                is_lo, is_up = lower[ar, rep], upper[ar, rep]
                on_line = (is_lo | is_up) & ~van & ~bad
                if np.any(on_line):
                    assert np.all(self.is_b[mb]), (
                        "Problem at delta pair %d/%d: trying to interpolate b0 sequences." % (ip + 1, self.n_pairs))
                    line_pos = np.where(is_lo, rep, partner_of_upper[ar, rep])   # sorted position of line_new
                    line_new = srt[ar, line_pos, :]                          # (V, 2)
                    sgn = np.sign(np.sum(rows[:, j, :] * line_new, axis=1))  # (V,)
                    i_max = np.argmax(line_new @ ref_un.T, axis=1)           # closest reference line
                    for im in np.unique(i_max[on_line]):
                        if int(im) not in pr["lines"]:
                            raise ValueError("rotate_atom_2Dprotocol: no reference line matches a new line "
                                             "direction at delta pair %d/%d" % (ip + 1, self.n_pairs))
                        vs = on_line & (i_max == im)
                        kind[vs, c] = 3
                        line[vs, c] = st["line_id"][(ip, int(im))]
                        sgn_c[vs, c] = sgn[vs]
                c += 1
        return {"nrm": nrm, "gz": gz, "kind": kind, "line": line, "sgn": sgn_c, "bad": bad}

    def plan(self, newdirs, strict=True):
        """Interpolation plan of every sequence for every direction: (row_lo, row_hi, w_lo,
        w_hi, scale), each (V, M).  strict=True raises the reference's AssertionError when a
        direction breaks the protocol's assumptions (e.g. a fascicle in the gradient plane,
        which projects both gradient lines onto one); strict=False returns a sixth array
        ok (V,) instead and gives those directions an all-zero plan.  (Host expansion of
        `classify`; the batched pipeline expands on the GPU, mfb_plan2d.)"""
        cl = self.classify(newdirs, strict)
        st = self._static_tables()
        nrm, gz, bad = cl["nrm"], cl["gz"], cl["bad"]
        V, M = nrm.shape[0], self.M
        G = self.G
        Gperp = G[np.newaxis, :] * nrm[:, self.lab_id]
        Gpar = gz[:, self.lab_id] * G[np.newaxis, :]
        assert np.all(np.isclose(G[np.newaxis, :] ** 2, Gperp ** 2 + Gpar ** 2)), \
            "Inconsistency in parallel and perpendicular gradient components for new fascicle."
        S_par = np.exp(-(self.gam * self.delta[np.newaxis, :] * Gpar) ** 2 *
                       (self.Delta - self.delta / 3)[np.newaxis, :] * self.DIFF)
        assert np.all(np.isclose(S_par[:, self.is_b0], 1)), \
            "New fascicle: parallel signal should  be equal to 1 in b0 sequences."
        row_lo = np.broadcast_to(np.arange(M, dtype=np.int32), (V, M)).copy()   # b0: own row
        row_hi = row_lo.copy()
        w_lo, w_hi = np.ones((V, M)), np.zeros((V, M))
        covered = np.broadcast_to(self.is_b0, (V, M)).copy()
        for c, (ip, j, mb) in enumerate(st["classes"]):
            van = cl["kind"][:, c] == 2
            if np.any(van):
                vv = np.where(van)[0][:, np.newaxis]
                row_lo[vv, mb[np.newaxis, :]] = row_hi[vv, mb[np.newaxis, :]] = self.b0_row[ip]
                w_lo[vv, mb[np.newaxis, :]], w_hi[vv, mb[np.newaxis, :]] = 1.0, 0.0
                covered[vv, mb[np.newaxis, :]] = True
            on = cl["kind"][:, c] == 3
            for li in np.unique(cl["line"][on, c]):
                vs = np.where(on & (cl["line"][:, c] == li))[0]
                nodes = st["line_nodes"][st["line_off"][li]:st["line_off"][li + 1]]
                node_rows = st["line_rows"][st["line_off"][li]:st["line_off"][li + 1]]
                x = Gperp[vs[:, np.newaxis], mb[np.newaxis, :]] * cl["sgn"][vs, c][:, np.newaxis]
                lo, hi, wl, wh = _lerp_plan(nodes, x)
                vv = vs[:, np.newaxis]
                row_lo[vv, mb[np.newaxis, :]], row_hi[vv, mb[np.newaxis, :]] = node_rows[lo], node_rows[hi]
                w_lo[vv, mb[np.newaxis, :]], w_hi[vv, mb[np.newaxis, :]] = wl, wh
                covered[vv, mb[np.newaxis, :]] = True
        # sequences no rule reached keep a zero perpendicular signal, like the reference
        scale = np.where(covered, S_par, 0.0)
        if strict:
            return row_lo, row_hi, w_lo, w_hi, scale
        scale[bad, :] = 0.0
        return row_lo, row_hi, w_lo, w_hi, scale, ~bad


    def plan_device(self, newdirs, device=0):
        """The same plan with the per-sequence expansion on the GPU (mfb_plan2d): the host only takes
        the per-direction decisions (`classify`, ~1/7 of the work of `plan`), uploads V x (2 U + 3 C)
        values and leaves the V x M x 5 plan entries to one kernel on the current CUDA stream.
        Returns (row_lo, row_hi, w_lo, w_hi, scale) as CUDA tensors (V, M) and ok (V,) as a NumPy
        array.  Rows and weights are the host plan's bit for bit; `scale` goes through the device's
        exp() and can differ from NumPy's in the last bit."""
        torch = _lib.require_cuda()
        dev = torch.device('cuda', device)
        st = self._static_tables()
        cache = getattr(self, "_dev_static", None)
        if cache is None or cache[0] != device:
            up = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
            cache = (device, {k: up(st[k]) for k in ("m_class", "m_lab", "m_isb0", "m_b0row", "m_gd", "m_tt",
                                                      "line_off", "line_nodes", "line_rows")},
                     up(self.G))
            self._dev_static = cache
        ds, d_G = cache[1], cache[2]
        cl = self.classify(newdirs, strict=False)
        V, M, U, C = cl["nrm"].shape[0], self.M, cl["nrm"].shape[1], cl["kind"].shape[1]
        ok = ~cl["bad"]
        d = {k: torch.from_numpy(np.ascontiguousarray(cl[k])).to(dev, non_blocking=True)
             for k in ("nrm", "gz", "kind", "line", "sgn")}
        d_ok = torch.from_numpy(np.ascontiguousarray(ok.astype(np.uint8))).to(dev, non_blocking=True)
        row_lo = torch.empty((V, M), dtype=torch.int32, device=dev)
        row_hi = torch.empty((V, M), dtype=torch.int32, device=dev)
        w_lo = torch.empty((V, M), dtype=torch.float64, device=dev)
        w_hi = torch.empty((V, M), dtype=torch.float64, device=dev)
        scale = torch.empty((V, M), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.load().mfb_plan2d(
                device, V, M, U, C, ds["m_class"].data_ptr(), ds["m_lab"].data_ptr(), ds["m_isb0"].data_ptr(),
                ds["m_b0row"].data_ptr(), d_G.data_ptr(), ds["m_gd"].data_ptr(), ds["m_tt"].data_ptr(), float(self.DIFF),
                d["nrm"].data_ptr(), d["gz"].data_ptr(), d["kind"].data_ptr(), d["line"].data_ptr(), d["sgn"].data_ptr(),
                d_ok.data_ptr(), ds["line_off"].data_ptr(), ds["line_nodes"].data_ptr(), ds["line_rows"].data_ptr(),
                row_lo.data_ptr(), row_hi.data_ptr(), w_lo.data_ptr(), w_hi.data_ptr(), scale.data_ptr(),
                torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "mfb_plan2d")
        for t in list(d.values()) + [d_ok]:
            t.record_stream(torch.cuda.current_stream(dev))
        return row_lo, row_hi, w_lo, w_hi, scale, ok


def rotate_atom_2Dprotocol(sig, sch_mat, refdir, newdir, DIFF, return_device=False):
    """Rotate signals of a 2D AxCaliber-like protocol (gradients in the xy plane, pairs of
    opposite polarities along one or two lines) from a fascicle along `refdir` to `newdir`
    (reference mf_utils.py:1440-1690): signal = parallel free-diffusion factor times the
    perpendicular signal, the latter interpolated linearly in the signed perpendicular
    gradient intensity along the closest reference line, per (Delta, delta) pair.

    Extension: `newdir` of shape (V, 3) rotates to V directions at once (one vectorised host
    plan, one GPU launch) and returns (V,) + sig.shape; with return_device=True the result
    stays on the GPU as a torch tensor (V, M, N)."""
    sig_shape = sig.shape
    if sig.ndim == 1:
        sig = sig[:, np.newaxis]
    if np.any(sch_mat[:, 2] != 0):
        raise ValueError("Use the original schemefile with zeros for gz.\n"
                         "Specify the reference and new orientations separately.")
    if sig_shape[0] != sch_mat.shape[0]:
        raise ValueError("Signal and scheme matrix must have the same "
                         "number of elements (sequences) along their first"
                         " dimension. Detected %d and %d." % (sig_shape[0], sch_mat.shape[0]))
    newdir = np.asarray(newdir, dtype=np.float64)
    batched = newdir.ndim == 2
    proto = _Protocol2D(sch_mat, refdir, DIFF)
    row_lo, row_hi, w_lo, w_hi, scale = proto.plan(newdir.reshape(-1, 3))
    out = _lerp_rows(proto.table(sig), row_lo, row_hi, w_lo, w_hi, scale, return_device=return_device)
    if return_device:
        return out
    if batched:
        return out.reshape((newdir.shape[0],) + sig_shape)
    return np.reshape(out[0], sig_shape)


def _solve_rotated_batch(table, plan_fn, N, peaks, Y, sig_iso, chunk, device):
    """Chunked rotate + search pipeline shared by solve_rotated_batch (HARDI, rotate_atom) and
    solve_rotated_2Dprotocol_batch: plan_fn(dirs (D, 3)) -> (row_lo, row_hi, w_lo, w_hi, scale or
    None, ok (D,)).  A worker thread prepares chunk c + 1 -- host plan, upload, dictionary assembly
    on the GPU (mfb_lerp_rows) on its own CUDA stream -- while the main thread searches chunk c
    (mfb_solve_batch); the rotated dictionaries never leave the GPU."""
    import threading
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device('cuda', device)
    peaks = np.ascontiguousarray(peaks, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    if peaks.ndim != 3 or peaks.shape[1] != 2 or peaks.shape[2] != 3:
        raise ValueError("peaks should have shape (V, 2, 3)")
    V, K = peaks.shape[0], peaks.shape[1]
    M = Y.shape[1]
    if Y.shape[0] != V:
        raise ValueError("Y should have %d rows" % V)
    iso = 0 if sig_iso is None else 1
    ntot = K * N + iso
    sizes = np.ascontiguousarray(np.array([N] * K + [1] * iso, dtype=np.int64))
    nb = sizes.size
    d_table = torch.from_numpy(np.ascontiguousarray(table, dtype=np.float64)).to(dev)
    d_iso = None if sig_iso is None else torch.from_numpy(np.ascontiguousarray(sig_iso, dtype=np.float64)).to(dev)
    w_out = np.zeros((V, nb))
    sub_out = np.zeros((V, nb), dtype=np.int32)
    obj_out = np.zeros(V)
    ok_out = np.zeros(V, dtype=bool)
    chunks = [(s0, min(V, s0 + chunk)) for s0 in range(0, V, chunk)]
    side = torch.cuda.Stream(device=dev)          # assembly stream of the worker
    ready = {}

    def prepare(c):
        """plan (host) -> upload -> assemble A (GPU, side stream); leaves (A, Yg, good, event)."""
        try:
            s0, s1 = chunks[c]
            nv = s1 - s0
            with torch.cuda.device(dev), torch.cuda.stream(side):
                # (a plan function may expand its plan on the GPU: it then runs on the side stream and returns tensors)
                rl, rh, wl, wh, sc, ok = plan_fn(peaks[s0:s1].reshape(-1, 3))
                good = np.where(ok.reshape(nv, K).all(axis=1))[0]
                d_pl = [None if x is None else (x if isinstance(x, torch.Tensor) else
                                                torch.from_numpy(np.ascontiguousarray(x)).to(dev, non_blocking=True))
                        for x in (rl, rh, wl, wh, sc)]
                A = torch.empty((nv, M, ntot), dtype=torch.float64, device=dev)
                if iso:
                    A[:, :, K * N] = d_iso[None, :]
                for k in range(K):
                    # directions are ordered (voxel, fascicle): fascicle k of every voxel is a strided view
                    pk = [None if x is None else x.view(nv, K, M)[:, k, :].contiguous() for x in d_pl]
                    rc = lib.mfb_lerp_rows(device, nv, M, N, d_table.data_ptr(), pk[0].data_ptr(), pk[1].data_ptr(),
                                           pk[2].data_ptr(), pk[3].data_ptr(),
                                           None if pk[4] is None else pk[4].data_ptr(),
                                           A.data_ptr() + 8 * k * N, ntot, side.cuda_stream)
                    _lib.check(rc, "mfb_lerp_rows")
                if good.size != nv:
                    A = A.index_select(0, torch.from_numpy(good).to(dev)).contiguous()
                Yg = torch.from_numpy(Y[s0:s1][good]).to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
            ready[c] = (A, Yg, good, ev)
        except BaseException as exc:       # re-raised in the main thread
            ready[c] = exc

    if chunks:
        prepare(0)
    main_stream = torch.cuda.current_stream(dev)
    for c, (s0, s1) in enumerate(chunks):
        item = ready.pop(c)
        if isinstance(item, BaseException):
            raise item
        th = None
        if c + 1 < len(chunks):
            th = threading.Thread(target=prepare, args=(c + 1,))
            th.start()
        A, Yg, good, ev = item
        main_stream.wait_event(ev)
        if good.size:
            A.record_stream(main_stream)
            Yg.record_stream(main_stream)
            w, sub, tot, obj, _ = solve_exhaustive_posweights_batch(A, Yg, sizes, device=device, return_device=True)
            w_out[s0 + good], sub_out[s0 + good], obj_out[s0 + good] = w.cpu().numpy(), sub.cpu().numpy(), obj.cpu().numpy()
            ok_out[s0 + good] = True
        del A, Yg
        if th is not None:
            th.join()
    return w_out, sub_out, obj_out, ok_out


def solve_rotated_2Dprotocol_batch(sig, sch_mat, refdir, peaks, Y, DIFF, sig_iso=None, chunk=48, device=0):
    """Low-level AxCaliber-like pipeline for many voxels (extension; per voxel it is what the
    reference's users write by hand, cf. tests/integration/test_exhaustive_fingerprinting.py:163-249):

        D_k = rotate_atom_2Dprotocol(sig, sch_mat, refdir, peaks[v, k], DIFF)      k = 0, 1
        solve_exhaustive_posweights([D_0, D_1 (, sig_iso)], Y[v], [N, N (, 1)])

    sig (M, N) single-fascicle dictionary along refdir, peaks (V, 2, 3) unit vectors, Y (V, M),
    optional isotropic column sig_iso (M,).  Returns (w (V, nb), ind_subdic (V, nb) int32,
    min_obj (V,), ok (V,) bool); voxels whose peak breaks the protocol's assumptions (reference
    AssertionError, e.g. a fascicle lying in the gradient plane) have ok = False and zero outputs."""
    sig = np.ascontiguousarray(sig, dtype=np.float64)
    if sig.ndim != 2 or np.asarray(Y).shape[1] != sig.shape[0]:
        raise ValueError("sig should be (M, N) and Y (V, M)")
    proto = _Protocol2D(sch_mat, refdir, DIFF)
    return _solve_rotated_batch(proto.table(sig), lambda d: proto.plan_device(d, device), sig.shape[1], peaks, Y,
                                sig_iso, chunk, device)


def solve_rotated_batch(sig, sch_mat, ordir, peaks, Y, DIFF, S0, sig_iso=None, chunk=256, device=0):
    """The same pipeline for HARDI-like protocols, i.e. the reference's test_hcp_dict sequence
    (tests/integration/test_exhaustive_fingerprinting.py:163-249) for many voxels:

        D_k = rotate_atom(sig, sch_mat, ordir, peaks[v, k], DIFF, S0)               k = 0, 1
        solve_exhaustive_posweights([D_0, D_1 (, sig_iso)], Y[v], [N, N (, 1)])

    Returns (w, ind_subdic, min_obj, ok) like solve_rotated_2Dprotocol_batch (ok is all True:
    rotate_atom has no per-direction failure mode)."""
    sig = np.ascontiguousarray(sig, dtype=np.float64)
    S0 = np.asarray(S0, dtype=np.float64)
    if S0.ndim == 1:
        S0 = S0[:, np.newaxis]
    if sig.ndim != 2 or np.asarray(Y).shape[1] != sig.shape[0]:
        raise ValueError("sig should be (M, N) and Y (V, M)")
    if not isinstance(DIFF, np.ndarray):
        DIFF = np.array([[DIFF]])
    ordir = np.asarray(ordir, dtype=np.float64)
    sch_mat = np.asarray(sch_mat, dtype=np.float64)
    # the table does not depend on the directions: build it once from a dummy direction
    table = _rotate_atom_plan(sig, sch_mat, ordir, ordir.reshape(1, 3), DIFF, S0, warnings=False)[0]

    def plan_fn(dirs):
        _, rl, rh, wl, wh = _rotate_atom_plan(sig, sch_mat, ordir, dirs, DIFF, S0, warnings=False)
        return rl, rh, wl, wh, None, np.ones(dirs.shape[0], dtype=bool)
    return _solve_rotated_batch(table, plan_fn, sig.shape[1], peaks, Y, sig_iso, chunk, device)


# ----------------------------------------------------------------------------------
# Monte-Carlo dictionary generation
# ----------------------------------------------------------------------------------

def monte_carlo_average(sim_phases, delta_mapping, gscaling, Dscaling, num_spins, device=0):
    """Monte-Carlo DW-MRI signal as the spin average of the dephasing
    (reference mf_utils.py:2758-2812):

        S_i = (1/n_spin) sum_l cos(Dscaling * sum_d gscaling[i, d] * sim_phases[map(i)*n_spin + l, d])

    sim_phases (n_ref*n_spin, n_dim) float64, delta_mapping (n_seq,) int64, gscaling
    (n_seq, n_dim) float64.  Returns the (n_seq,) float64 signal.  Runs on the GPU
    (mfb_mc_average); NumPy arrays or CUDA tensors are accepted."""
    torch = _lib.require_cuda()
    dev = torch.device('cuda', device)

    def up(x, dt_np, dt_t):
        if isinstance(x, torch.Tensor):
            return x.to(device=dev, dtype=dt_t).contiguous()
        return torch.from_numpy(np.ascontiguousarray(x, dtype=dt_np)).to(dev)
    d_ph = up(sim_phases, np.float64, torch.float64)
    d_map = up(delta_mapping, np.int64, torch.int64)
    d_gs = up(gscaling, np.float64, torch.float64)
    if d_ph.dim() != 2 or d_gs.dim() != 2 or d_map.dim() != 1:
        raise ValueError("sim_phases and gscaling should be 2D arrays, delta_mapping a 1D array")
    n_entries, dim = int(d_ph.shape[0]), int(d_ph.shape[1])
    n_seq = int(d_map.shape[0])
    if tuple(d_gs.shape) != (n_seq, dim):
        raise ValueError("gscaling should have shape (%d, %d), detected %s"
                         % (n_seq, dim, tuple(d_gs.shape)))
    num_spins = int(num_spins)
    if n_seq > 0:
        lo, hi = int(d_map.min()), int(d_map.max())
        if lo < 0 or (hi + 1) * num_spins > n_entries:
            raise IndexError("delta_mapping refers to reference sequences outside sim_phases "
                             "(%d entries, %d spins per sequence)" % (n_entries, num_spins))
    out = torch.zeros((n_seq,), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        rc = _lib.load().mfb_mc_average(device, n_entries, dim, d_ph.data_ptr(), n_seq, d_map.data_ptr(),
                                        d_gs.data_ptr(), float(Dscaling), num_spins, out.data_ptr(), st)
    _lib.check(rc, "mfb_mc_average")
    return out.cpu().numpy()


def _phase_file_encoding(ext):
    """(byte order, dtype code, bytes per item) from a phase-file extension such as
    '.bdouble' or '.lfloat' (reference mf_utils.py:2905-2930)."""
    if not ext:
        raise ValueError("Phase file extension not found.\nAborting as there is no way to tell "
                         "which level of precision was used to store the phase values (e.g., "
                         "float, double, ...).")
    order = {'b': '>', 'l': '<'}.get(ext[1].lower())
    if order is None:
        raise ValueError("Phase file extension (after the dot) should start with a b for big "
                         "endian or with a l for little endian. Detected: \"%s\"." % ext[1])
    kind = ext[2:]
    if kind in ('single', 'float'):
        return order, 'f4', 4
    if kind == 'double':
        return order, 'f8', 8
    raise ValueError("Data type of phase file specified in file extension (\"%s\") not supported."
                     % kind)


def get_PGSE_from_phases(phasefile, sch_mat_sim, sch_mat, dim=None, D_sim=None, D=None):
    """PGSE signal of the protocol `sch_mat` from the spins' phases accumulated in a
    Monte-Carlo simulation run with protocol `sch_mat_sim` (reference
    mf_utils.py:2815-3015).  `phasefile` is 'path/to/base_phase_x.bdouble'; its siblings
    '_phase_y', '_phase_z' are read for dim > 1.  The extension gives byte order (b/l) and
    precision (double / single / float).  Every sequence of `sch_mat` must use a
    (Delta, delta) pair present in `sch_mat_sim`; gradients are rescaled component-wise.
    The spin average runs on the GPU (monte_carlo_average)."""
    import os
    names, maxdim = ['x', 'y', 'z'], 3
    d_ratio_sqrt = 1.0
    if D is not None:
        if D_sim is None:
            raise NameError("Simulation diffusivity should be specified if new signal "
                            "diffusivity is set.")
        d_ratio_sqrt = float(np.sqrt(D / D_sim))
    if dim is None:
        dim = maxdim
    elif dim > maxdim:
        raise ValueError("dim should be less than or equal to %d." % maxdim)
    sim = import_PGSE_scheme(sch_mat_sim)
    new = import_PGSE_scheme(sch_mat)
    if np.any(new[:, dim:maxdim] != 0):
        print("WARNING get_PGSE_from_phases: detected non-zero entries in gradient components "
              "after dimension %d.\nThose components will be ignored but make sure the right "
              "acquisition protocol was provided.\nIt is common for such protocols to contain "
              "zeros in those gradient components, for instance after projection into the "
              "xy-plane of a 3D protocol.\n" % dim)
    n_seq, n_ref = new.shape[0], sim.shape[0]
    g_sim = sim[:, :3] * sim[:, 3][:, np.newaxis]
    g_new = new[:, :3] * new[:, 3][:, np.newaxis]
    # reference sequence of each new sequence: LAST simulated row with the same (Delta, delta)
    same = np.all(new[:, np.newaxis, 4:6] == sim[np.newaxis, :, 4:6], axis=2)      # (n_seq, n_ref)
    mapping = np.where(same.any(axis=1), n_ref - 1 - np.argmax(same[:, ::-1], axis=1), -1).astype(np.int64)
    bad = np.where(mapping < 0)[0]
    if bad.size:
        listing = '\n'.join('\t%4d -- %5g -- %5g' % (b, new[b, 4] * 1e3, new[b, 5] * 1e3) for b in bad)
        raise ValueError('Acquisition protocol contains %d (Delta,delta) pair(s) (out of %d) not used '
                         'to simulate the directional phases in the Monte Carlo simulation. List of '
                         'unmatched sequences:\nSequ. no. -- Delta [ms] -- delta [ms]\n%s'
                         % (bad.size, n_seq, listing))
    gscaling = g_new[:, :dim] / g_sim[mapping, :dim]
    if not os.path.isfile(phasefile):
        raise RuntimeError("File %s does not exist." % phasefile)
    nbytes = os.path.getsize(phasefile)
    folder, tail = os.path.split(phasefile)
    base, ext = os.path.splitext(tail)
    order, code, width = _phase_file_encoding(ext)
    if nbytes % (n_ref * width) != 0:
        raise RuntimeError("Phase file %s is either corrupted or inconsistently named. Storage "
                           "precision of items (%d bytes) times number of reference simulation "
                           "sequences (%d) does not divide total size (%d bytes)."
                           % (phasefile, width, n_ref, nbytes))
    n_entries = nbytes // width
    n_spins = n_entries // n_ref
    phases = np.zeros((n_entries, dim))
    for i in range(dim):
        path_i = os.path.join(folder, base[:-len(names[i])] + names[i] + ext)
        if not os.path.isfile(path_i):
            raise RuntimeError("Phase file %s not found." % path_i)
        phases[:, i] = np.fromfile(path_i, dtype=order + code, count=n_entries, sep="")
    return monte_carlo_average(phases, mapping, gscaling, d_ratio_sqrt, n_spins)
