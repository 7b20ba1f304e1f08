"""cleanup_2fascicles: peak pre-processing in front of MFModel.fit (reference mf.py:36-335).

Host-side NumPy (elementwise thresholds and a 2-way sort per voxel; not on the fit path).
Selects 0, 1 or 2 of two detected fascicle orientations from their weights and crossing
angle: merge directions closer than 15 degrees, drop a fascicle 2.5x lighter than the other
unless it weighs more than 0.20, drop absolute weights under 0.075, sort by weight.
"""
import numpy as np

from . import mf_utils as mfu
from . import nifti

RATIO = 2.5        # dominant / secondary weight ratio above which the secondary is dropped ...
W_KEEP = 0.20      # ... unless its weight exceeds this
W_SMALL = 0.075    # absolute weight under which a fascicle is dropped
ANG_MIN = 15       # crossing angle [deg] under which the two orientations are merged


def _load(x):
    return nifti.load(x)[0] if isinstance(x, str) else x


def _principal_dirs(tensors6):
    d, eigv = np.linalg.eigh(mfu.DT_vec_to_2Darray(tensors6, order='column'))
    return eigv[..., -1] * (np.abs(d)[..., -1] > 0)[:, np.newaxis]


def cleanup_2fascicles(frac1, frac2, peakmode, mu1, mu2, mask, frac12=None):
    """Cleans up two detected fascicle orientations per voxel.

    Arguments as in the reference: `frac1`, `frac2` weights (arrays or NIfTI paths), `peakmode`
    one of 'colat_longit' (last dim 2), 'peaks' (3) or 'tensor' (6, NIfTI lower-triangular
    order), `mu1`, `mu2` the orientations, `mask`, optional `frac12` holding both weights.

    Returns (peaks_out mask.shape + (6,), num_fasc_out mask.shape); in one-fascicle voxels the
    fascicle is population 0.
    """
    if (frac1 is None or frac2 is None) and frac12 is None:
        raise ValueError("If fractions of first and second fascicles set to None,"
                         " argument frac12 is required to specify both fractions"
                         " simultanously. A total of 6 arguments should be passed, not 5.")
    mask, frac1, frac2 = _load(mask), _load(frac1), _load(frac2)
    if frac12 is not None:
        frac12 = _load(frac12)
        if frac12.shape[-1] < 2:
            raise ValueError("Last dimension of frac12 should have size at least 2.")
        if frac12.shape[mask.ndim] == 1:          # (..., 1, 2)
            frac1, frac2 = frac12[..., 0, 0], frac12[..., 0, 1]
        else:
            frac1, frac2 = frac12[..., 0], frac12[..., 1]
    if frac1.shape != mask.shape:
        raise ValueError("frac1 should have the same shape as mask")
    if frac2.shape != mask.shape:
        raise ValueError("frac2 should have the same shape as mask")
    mu1, mu2 = _load(mu1), _load(mu2)
    sizes = {'colat_longit': 2, 'peaks': 3, 'tensor': 6}
    if peakmode not in sizes:
        raise ValueError('Unknown peak mode %s' % peakmode)
    if peakmode == 'tensor':
        if mu1.shape[mask.ndim] == 1:
            mu1 = mu1[..., 0, :]
        if mu2.shape[mask.ndim] == 1:
            mu2 = mu2[..., 0, :]
    if mu1.shape[-1] != sizes[peakmode] or mu2.shape[-1] != sizes[peakmode]:
        raise ValueError("In '%s' peak mode, last dimension of mu1 and mu2 should have size %d. "
                         "Detected %d and %d." % (peakmode, sizes[peakmode], mu1.shape[-1], mu2.shape[-1]))

    roi = mask > 0
    n = int(np.sum(roi))
    f = np.stack([frac1[roi], frac2[roi]], axis=1).astype(np.float64)     # cleaned weights
    f_in = f.copy()
    m1, m2 = mu1[roi], mu2[roi]
    peaks = np.zeros((n, 6))
    if peakmode == 'colat_longit':
        for k, m in enumerate((m1, m2)):
            peaks[:, 3 * k + 0] = np.sin(m[..., 0]) * np.cos(m[..., 1])
            peaks[:, 3 * k + 1] = np.sin(m[..., 0]) * np.sin(m[..., 1])
            peaks[:, 3 * k + 2] = np.cos(m[..., 0])
    elif peakmode == 'peaks':
        peaks[:, :3], peaks[:, 3:] = m1, m2
    else:
        peaks[:, :3], peaks[:, 3:] = _principal_dirs(m1), _principal_dirs(m2)
    num = np.full(n, 2.0)

    # merge directions closer than ANG_MIN into population 0 (sign-aware)
    dp = np.sum(peaks[:, :3] * peaks[:, 3:], axis=-1)
    merge = np.abs(np.clip(dp, -1, 1)) > np.cos(ANG_MIN * np.pi / 180)
    if np.any(merge):
        summed = peaks[merge, :3] + peaks[merge, 3:] * np.sign(dp[merge])[:, np.newaxis]
        peaks[merge, :3] = summed / np.sqrt(np.sum(summed ** 2, axis=1))[:, np.newaxis]
        peaks[merge, 3:] = 0
        f[merge, 0] = f_in[merge, 0] + f_in[merge, 1]
        f[merge, 1] = 0
        num[merge] = 1

    # relatively small fascicles: population 0 too small -> population 1 takes its place
    small0 = (f[:, 1] > RATIO * f[:, 0]) & (f[:, 0] < W_KEEP)
    if np.any(small0):
        peaks[small0, :3] = peaks[small0, 3:]
        peaks[small0, 3:] = 0
        f[small0, 0] = f[small0, 1]
        f[small0, 1] = 0
        num[small0] = (f[small0, 0] > 0) * 1
    small1 = (f[:, 0] > RATIO * f[:, 1]) & (f[:, 1] < W_KEEP)
    if np.any(small1):
        peaks[small1, 3:] = 0
        f[small1, 1] = 0
        num[small1] = (f[small1, 0] > 0) * 1

    # absolutely small weights
    tiny0 = f[:, 0] < W_SMALL
    if np.any(tiny0):
        peaks[tiny0, :3] = 0
        f[tiny0, 0] = 0
        num[tiny0] = num[tiny0] - 1
    tiny1 = f[:, 1] < W_SMALL
    if np.any(tiny1):
        peaks[tiny1, 3:] = 0
        f[tiny1, 1] = 0
        num[tiny1] = (f[tiny1, 0] > 0) * 1

    # heaviest fascicle first (same ordering rule as the reference: reversed ascending argsort)
    order = np.argsort(f, axis=-1)[:, ::-1]
    cols = (np.kron(order, 3 * np.ones((1, 3), dtype=int)) + np.tile(np.array([[0, 1, 2]]), [n, 2]))
    peaks = peaks[np.arange(n)[:, np.newaxis], cols]

    peaks_out = np.zeros(mask.shape + (6,))
    peaks_out[roi] = peaks
    num_out = np.zeros(mask.shape)
    num_out[roi] = num
    return peaks_out, num_out
