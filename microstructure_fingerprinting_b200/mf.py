"""MFModel / MFModelFit: the reference's DIPY-style front end (reference mf.py:464-1229)
with the voxel loop (mf.py:978-1028) and _fit_voxel (mf.py:340-461) replaced by one
batched call into libmfb200.so per GPU.

Input marshalling and the output maps keep the reference's semantics: same arguments,
same ROI ordering (np.where(mask > 0)), same params-row layout, same exception types.
`parallel=True` keeps its keyword and now means "shard the ROI over all visible GPUs"
(contiguous chunks, no collective) instead of a multiprocessing.Pool.
"""
import contextlib
import os
import threading
import time

import numpy as np

from . import _lib
from . import mf_utils as mfu
from . import nifti


def shard_bounds(n_items, n_shards, cost=None):
    """Contiguous split of range(n_items) into n_shards spans: returns the n_shards + 1
    boundaries.  Used for the per-GPU ROI shards (no data-path collective: voxels are
    independent given the replicated plan).  Without `cost` the spans hold equal counts; with
    a per-item cost (n_items,) they hold (nearly) equal total cost, so that GPUs finish together
    when the fascicle count varies over the ROI."""
    if cost is None or n_shards <= 1 or n_items == 0:
        return np.linspace(0, n_items, n_shards + 1).astype(np.int64)
    csum = np.concatenate(([0.0], np.cumsum(np.asarray(cost, dtype=np.float64))))
    if not csum[-1] > 0:
        return np.linspace(0, n_items, n_shards + 1).astype(np.int64)
    targets = csum[-1] * np.arange(1, n_shards) / n_shards
    inner = np.searchsorted(csum, targets, side="left")
    b = np.concatenate(([0], inner, [n_items])).astype(np.int64)
    return np.maximum.accumulate(b)


def voxel_cost(numfasc_roi, csf_roi, ear_roi, n_atoms, n_ear):
    """Relative cost of fitting a voxel: the number of atom tuples searched (the dominant term
    of SURVEY 8d's flop count), floored at one unit."""
    K = np.asarray(numfasc_roi)
    c = np.where(K >= 2, float(n_atoms) ** 2, np.where(K == 1, float(n_atoms), 1.0))
    c = c * np.where(np.asarray(ear_roi) > 0, float(max(n_ear, 1)), 1.0)
    return np.maximum(c * np.where(np.asarray(csf_roi) > 0, 1.3, 1.0), 1.0)


def _as_array(x):
    """Array or path to a NIfTI file -> (ndarray, affine or None)."""
    if isinstance(x, str):
        return nifti.load(x)
    return x, None


def _shape_str(shape):
    return " ".join("%d" % s for s in shape)


def _bool_compartment(arg, name, img_shape, in_mask, ROI_size):
    """csf_mask / ear_mask argument -> (bool vector over the ROI, affine or None)
    (reference mf.py:852-894)."""
    if arg is None:
        return np.zeros(ROI_size, dtype=bool), None
    if np.isscalar(arg) and not isinstance(arg, str):
        return np.full(ROI_size, arg > 0, dtype=bool), None
    arr, aff = _as_array(arg)
    if arr.shape != img_shape:
        raise ValueError("Arg. %s incomptabible. Based on data,"
                         " it should have shape (%s), detected (%s)"
                         " instead." % (name, _shape_str(img_shape), _shape_str(arr.shape)))
    return arr[in_mask] > 0, aff


class MFModel():
    r"""Microstructure Fingerprinting model (reference mf.py:464-1051)."""
    MAX_FASC = 2  # max number fascicles in a voxel
    MAX_PROG_LINES = 100
    DFT_DISP_ITVL = 5

    def __init__(self, dictionary):
        """dictionary: path to a MATLAB .mat file or a dict with the keys the reference
        uses: 'dictionary' (Nms, N), 'sch_mat' (Nms, 7), 'orientation' (3,), 'num_atom',
        'num_ear', 'T2_csf', 'DIFF_csf', 'T2_ear', 'DIFF_ear', 'fasc_propnames' and one
        array per property name."""
        if isinstance(dictionary, str):
            self.dic = mfu.loadmat(dictionary)
        elif isinstance(dictionary, dict):
            self.dic = dictionary
        else:
            raise ValueError("Dictionary should either be a valid path to a"
                             " Matlab-like mat file or a Python dictionary.")
        self.ms_interpolator = mfu.init_PGSE_multishell_interp(
            self.dic['dictionary'], self.dic['sch_mat'], self.dic['orientation'])
        self._plans = {}                      # (shard, device) -> {'lock', 'key', 'plan'}
        self._plans_lock = threading.Lock()
        print("Initiated model based on dictionary with %d single-fascicle"
              " fingerprint(s) and %d fingerprint(s) for the extra-axonal"
              " restricted (EAR) compartment." %
              (self.dic['num_atom'], self.dic['num_ear']))

    # ------------------------------------------------------------------
    @contextlib.contextmanager
    def _plan(self, slot, dev, key, scheme_plan, sig_csf, sig_ear):
        """The mfb_plan of shard `slot` on device `dev` for this protocol / CSF / EAR columns,
        kept between fit() calls (lookup table, workspace and page-locked staging slots stay
        allocated); one fit at a time per plan."""
        with self._plans_lock:
            entry = self._plans.get((slot, dev))
            if entry is None:
                entry = self._plans[(slot, dev)] = {'lock': threading.Lock(), 'key': None, 'plan': None}
        with entry['lock']:
            if entry['key'] != key:
                if entry['plan'] is not None:
                    entry['plan'].close()
                    entry['plan'] = None
                entry['plan'] = mfu.GpuPlan(self.ms_interpolator, scheme_plan, sig_csf, sig_ear, device=dev)
                entry['key'] = key
            yield entry['plan']

    def close(self):
        """Release the GPU plans (device memory, page-locked staging) this model holds."""
        with self._plans_lock:
            for entry in self._plans.values():
                with entry['lock']:
                    if entry['plan'] is not None:
                        entry['plan'].close()
                        entry['plan'], entry['key'] = None, None

    # ------------------------------------------------------------------
    def _peaks_in_roi(self, peaks, colat_longit, tensors, img_shape, in_mask, ROI_size,
                      maxfasc, ndim_mask, VRB):
        """One of the three orientation inputs -> (ROI_size, 3*k) array
        (reference mf.py:693-800)."""
        affine = None
        if peaks is not None:
            arr, affine = _as_array(peaks)
            if arr.shape[:-1] != img_shape:
                raise ValueError("Arg. peaks not compatible. Based on data,"
                                 " it should have shape (%s x), with x a "
                                 "multiple of 3. Got (%s) instead." %
                                 (_shape_str(img_shape), _shape_str(arr.shape)))
            if arr.shape[-1] % 3 != 0:
                raise ValueError("Size of last dimension of arg. peaks should"
                                 " be a multiple of 3, got %d instead." % arr.shape[-1])
            if arr.shape[-1] > maxfasc * 3 and VRB >= 1:
                print("Ignoring last %d value(s) along last dimension of"
                      " peaks, as max number of axon populations in mask"
                      " is %d." % (arr.shape[-1] - maxfasc * 3, maxfasc))
            return np.asarray(arr[in_mask, :3 * maxfasc], dtype=np.float64), affine
        if colat_longit is not None:
            args, allowed, mode = colat_longit, ((2,),), 'angles'
        elif tensors is not None:
            args, allowed, mode = tensors, ((6,), (1, 6)), 'tensors'
        else:
            raise RuntimeError("At least one of peaks, colat_longit and"
                               " tensors must be specified.")
        if not isinstance(args, list):
            args = [args]
        peaks_roi = np.zeros((ROI_size, 3 * len(args)))
        if len(args) > maxfasc and VRB >= 1:
            print("Ignoring %d peak orientation argument(s) because"
                  " max number of axon populations in mask is %d." %
                  (len(args) - maxfasc, maxfasc))
        for i in range(min(len(args), maxfasc)):
            arr, aff = _as_array(args[i])
            if affine is None:
                affine = aff
            if arr.shape not in [img_shape + d for d in allowed]:
                expected = " or ".join("(" + _shape_str(img_shape + d) + ")" for d in allowed)
                raise ValueError("Peak orientation arg. %d of %d seems "
                                 "incompatible. Based on data, it should have"
                                 " shape %s, got (%s) instead." %
                                 (i + 1, len(args), expected, _shape_str(arr.shape)))
            if mode == 'angles':
                th, ph = arr[in_mask, 0], arr[in_mask, 1]
                peaks_roi[:, 3 * i + 0] = np.sin(th) * np.cos(ph)
                peaks_roi[:, 3 * i + 1] = np.sin(th) * np.sin(ph)
                peaks_roi[:, 3 * i + 2] = np.cos(th)
            else:
                if arr.shape[ndim_mask] == 1 and arr.ndim == ndim_mask + 2:
                    arr = arr[(slice(None),) * ndim_mask + (0, slice(None))]
                # principal eigenvector; zero tensors keep a zero peak
                d, eigv = np.linalg.eigh(mfu.DT_vec_to_2Darray(arr[in_mask, :], order='column'))
                nonzero = (np.abs(d)[..., -1] > 0)[:, np.newaxis]
                peaks_roi[:, 3 * i:3 * i + 3] = eigv[..., -1] * nonzero
        return peaks_roi, affine

    # ------------------------------------------------------------------
    def fit(self, data, mask, numfasc, *,
            peaks=None, colat_longit=None, tensors=None,
            pgse_scheme=None, bvals=None, bvecs=None,
            csf_mask=None, ear_mask=None,
            verbose=1, parallel=False, devices=None, exact=False):
        r"""Perform fingerprinting on the pre-computed dictionary (reference mf.py:516-1051).

        Arguments are those of the reference.  `parallel=True` shards the ROI over all
        visible GPUs; `devices` (list of CUDA device indices) overrides that choice.
        `exact=True` forces the reference-order tier for every voxel (slower; results are
        the same by construction, this is a verification knob).

        Returns an MFModelFit with one attribute per estimated parameter map.
        """
        VRB = verbose
        st_0 = time.time()
        data_arr, nii_affine = _as_array(data)
        if isinstance(data, str) and VRB >= 2:
            print("Data loaded from file %s in %g s." % (data, time.time() - st_0))
        mask_arr, aff = _as_array(mask)
        if nii_affine is None:
            nii_affine = aff

        img_shape = mask_arr.shape
        in_mask = mask_arr > 0
        ROI_size = int(np.count_nonzero(in_mask))
        if ROI_size == 0:
            raise ValueError("No voxel detected in mask. Please provide "
                             "a non-empty mask.")
        if data_arr.shape[:-1] != img_shape:
            raise ValueError("Data and mask not compatible. Based on data,"
                             " mask should have shape (%s), "
                             "got (%s) instead." %
                             (_shape_str(data_arr.shape[:-1]), _shape_str(img_shape)))

        # number of fascicles per voxel (reference mf.py:660-687)
        if np.isscalar(numfasc) and not isinstance(numfasc, str):
            numfasc_roi = np.full(ROI_size, numfasc, dtype=int)
        else:
            nf_arr, _ = _as_array(numfasc)
            if nf_arr.shape != img_shape:
                raise ValueError("Data and argument numfasc not compatible. "
                                 " Based on data, numfasc should have "
                                 "shape (%s), got (%s) instead." %
                                 (_shape_str(img_shape), _shape_str(nf_arr.shape)))
            numfasc_roi = nf_arr[in_mask].astype(int)
        maxfasc = int(np.max(numfasc_roi))
        if maxfasc > MFModel.MAX_FASC:
            raise ValueError("Detected %d mask voxel(s) in numfasc with"
                             " number of axon populations greater than"
                             " allowed maximum of %d." %
                             (np.sum(numfasc_roi > MFModel.MAX_FASC), MFModel.MAX_FASC))
        if np.any(numfasc_roi < 0):
            raise ValueError("Detected %d mask voxel(s) with a negative number of axon "
                             "populations in numfasc." % np.sum(numfasc_roi < 0))

        peaks_roi, aff = self._peaks_in_roi(peaks, colat_longit, tensors, img_shape, in_mask,
                                            ROI_size, maxfasc, mask_arr.ndim, VRB)
        if nii_affine is None:
            nii_affine = aff

        # missing or non-unit peak directions (reference mf.py:803-815; the unit-norm
        # check is the per-voxel one of interp_PGSE_from_multishell, mf_utils.py:1798-1802,
        # hoisted in front of the launch)
        for i in range(maxfasc):
            present = numfasc_roi >= i + 1
            pk = peaks_roi[:, 3 * i:3 * i + 3]
            if not present.all():
                pk = pk[present]
            sq = np.einsum('ij,ij->i', pk, pk)
            num_0 = int(np.count_nonzero(sq == 0))
            if num_0 > 0:
                num_0 = int(np.sum(np.sum(np.abs(pk), axis=1) == 0))
            if num_0 > 0:
                raise ValueError("Detected %d voxel(s) in which the main "
                                 "orientation of axon population %d/%d was "
                                 "a zero vector, although numfasc "
                                 "specifies the presence of that "
                                 "population." % (num_0, i + 1, maxfasc))
            nrm = np.sqrt(sq)
            bad = np.abs(1 - nrm) > 1e-3
            if np.any(bad):
                raise ValueError("Orientation vector of the new signal must have unit norm."
                                 " Detected %g." % (nrm[np.argmax(bad)],))

        # subject protocol (reference mf.py:821-846)
        if pgse_scheme is not None:
            if isinstance(pgse_scheme, str):
                pgse_scheme = np.loadtxt(pgse_scheme, skiprows=1)
            if pgse_scheme.shape[1] != 7:
                raise ValueError("pgse_scheme should have 7 columns, "
                                 " detected %d instead." % (pgse_scheme.shape[1],))
        else:
            if bvals is None or bvecs is None:
                raise TypeError("If no schemefile is provided, then both"
                                " bvals and bvecs must be specified.")
            pgse_scheme = mfu.get_PGSE_scheme_from_bval_bvec_dense(
                self.dic['sch_mat'], bvals, bvecs, 1e-3)
        pgse_scheme = np.asarray(pgse_scheme, dtype=np.float64)
        num_seq = pgse_scheme.shape[0]
        if data_arr.shape[-1] != num_seq:
            raise ValueError("Data holds %d measurements per voxel but the protocol "
                             "describes %d." % (data_arr.shape[-1], num_seq))
        gam = mfu.get_gyromagnetic_ratio('H')
        G, Delta, delta, TE = (pgse_scheme[:, i] for i in (3, 4, 5, 6))
        b = (gam * G * delta) ** 2 * (Delta - delta / 3)

        csf_roi, aff = _bool_compartment(csf_mask, 'csf_mask', img_shape, in_mask, ROI_size)
        if nii_affine is None:
            nii_affine = aff
        ear_roi, aff = _bool_compartment(ear_mask, 'ear_mask', img_shape, in_mask, ROI_size)
        if nii_affine is None:
            nii_affine = aff
        csf_on = bool(np.any(csf_roi))
        ear_on = bool(np.any(ear_roi))

        n_empty = 0
        if VRB >= 2:
            n_empty = np.sum((numfasc_roi + csf_roi + ear_roi) == 0)
        if n_empty > 0 and VRB >= 2:
            print("WARNING: detected %d voxel(s) in mask with zero "
                  " axon population, no cerebrospinal fluid (CSF) and no"
                  " extra-axonal restricted (EAR) compartment specified."
                  " No estimation will be performed there." % (n_empty,))

        # patient-specific CSF / EAR columns (reference mf.py:918-925)
        sig_csf = sig_ear = None
        if csf_on:
            sig_csf = np.exp(-TE / self.dic['T2_csf']) * np.exp(-b * self.dic['DIFF_csf'])
        if ear_on:
            DIFF_ear = np.atleast_1d(self.dic['DIFF_ear'])
            sig_ear = np.zeros((num_seq, self.dic['num_ear']))
            for i in range(self.dic['num_ear']):
                sig_ear[:, i] = np.exp(-TE / self.dic['T2_ear']) * np.exp(-b * DIFF_ear[i])
            if np.any(np.all(sig_ear == 0, axis=0)):
                raise AssertionError("All-zero columns detected in A")
        if csf_on and np.all(sig_csf == 0):
            raise AssertionError("All-zero columns detected in A")

        # ---- batched estimation on the GPU(s) ----
        torch = _lib.require_cuda()
        scheme_plan = mfu.SchemePlan(self.ms_interpolator, pgse_scheme)
        if devices is None:
            devices = list(range(torch.cuda.device_count())) if parallel else [0]
        devices = list(devices)[:max(1, min(len(devices), ROI_size))]
        # The ROI signals are never copied on the Python side (the reference's
        # data_arr[mask > 0], mf.py:644): every voxel is described by the element offset of
        # its first measurement in the caller's volume and libmfb200 gathers chunk by chunk
        # in a helper thread while the GPU fits the previous chunks (mfb_fit_volume).
        if data_arr.dtype not in (np.float64, np.float32) or not data_arr.dtype.isnative:
            data_arr = data_arr.astype(np.float64)
        item = data_arr.itemsize
        if any(st % item for st in data_arr.strides):
            data_arr = np.ascontiguousarray(data_arr)
        meas_stride = data_arr.strides[-1] // item
        if data_arr.flags.c_contiguous and ROI_size == in_mask.size:
            vox_off = np.arange(0, ROI_size * num_seq, num_seq, dtype=np.int64)
        elif data_arr.flags.c_contiguous:
            vox_off = np.flatnonzero(in_mask.ravel()).astype(np.int64) * num_seq
        else:
            vox_off = np.zeros(ROI_size, dtype=np.int64)
            for d, c in enumerate(np.nonzero(in_mask)):
                vox_off += c.astype(np.int64) * (data_arr.strides[d] // item)
        K32 = numfasc_roi.astype(np.int32)
        csf_u8 = csf_roi.view(np.uint8) if csf_on else None
        ear_u8 = ear_roi.view(np.uint8) if ear_on else None
        peaks_c = np.ascontiguousarray(peaks_roi[:, :3 * maxfasc], dtype=np.float64)
        num_params = 1 + maxfasc * 2 + csf_on * 1 + ear_on * 2 + 2
        params_in_mask = np.empty((ROI_size, num_params))
        st_est = time.time()
        if VRB >= 2:
            print("Starting estimation in %d voxel(s) on %d GPU(s)." % (ROI_size, len(devices)))
        cost = None
        if len(devices) > 1:
            cost = voxel_cost(numfasc_roi, csf_roi, ear_roi, self.dic['num_atom'], self.dic.get('num_ear', 1))
        bounds = shard_bounds(ROI_size, len(devices), cost)
        errors = []
        plan_key = (pgse_scheme.tobytes(), None if sig_csf is None else sig_csf.tobytes(),
                    None if sig_ear is None else sig_ear.tobytes())

        def work(rank, dev):
            lo, hi = int(bounds[rank]), int(bounds[rank + 1])
            if hi <= lo:
                return
            try:
                with self._plan(rank, dev, plan_key, scheme_plan, sig_csf, sig_ear) as plan:
                    plan.fit_volume(data_arr, vox_off[lo:hi], meas_stride, peaks_c[lo:hi], K32[lo:hi],
                                    None if csf_u8 is None else csf_u8[lo:hi],
                                    None if ear_u8 is None else ear_u8[lo:hi],
                                    maxfasc, csf_on, ear_on, params_in_mask[lo:hi],
                                    flags=1 if exact else 0)
            except BaseException as exc:  # re-raised in the caller's thread
                errors.append(exc)

        if len(devices) == 1:
            work(0, devices[0])
        else:
            threads = [threading.Thread(target=work, args=(r, d)) for r, d in enumerate(devices)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        if errors:
            raise errors[0]
        if VRB >= 2:
            print("Estimation performed in %g second(s)." % (time.time() - st_est))

        fitinfo = {'maxfasc': maxfasc, 'csf_on': csf_on, 'ear_on': ear_on,
                   'affine': nii_affine, 'mask': mask_arr,
                   'fasc_propnames': [x.strip() for x in self.dic['fasc_propnames']],
                   'peaks_roi': peaks_roi}
        for n in fitinfo['fasc_propnames']:
            fitinfo['_dict_' + n] = self.dic[n]
        if ear_on:
            fitinfo['DIFF_ear'] = np.atleast_1d(self.dic['DIFF_ear'])
        return MFModelFit(fitinfo, params_in_mask, verbose=VRB)


class MFModelFit():
    """Fitted maps (reference mf.py:1054-1229): one attribute per name in
    `param_names`, each shaped like the mask (peaks: mask.shape + (3,))."""

    def __init__(self, fitinfo, model_params, verbose=0):
        self.affine = fitinfo['affine']
        numfasc = fitinfo['maxfasc']
        csf_on = fitinfo['csf_on']
        ear_on = fitinfo['ear_on']
        mask = fitinfo['mask']
        in_mask = mask > 0
        ROI_size = model_params.shape[0]
        assert ROI_size == np.count_nonzero(in_mask), 'Inconsistent mask and model parameter array'
        # ROI rows -> volumes.  Every map is described by a function of a span of params rows;
        # spans are filled by a few threads (NumPy releases the GIL in the copies, gathers and
        # products involved), each writing its own part of every volume.  A full mask needs no
        # scatter index at all.
        full = ROI_size == in_mask.size
        roi_flat = None if full else np.flatnonzero(in_mask.ravel())
        peaks_roi = fitinfo['peaks_roi']
        specs = []   # (name, trailing shape, f(rows, lo, hi) -> values of the span)

        def col(c):
            return lambda rows, lo, hi: rows[:, c]

        def prop_of(values, k):
            def f(rows, lo, hi):   # zero where the fascicle got no weight
                return values[rows[:, 1 + numfasc + k].astype(np.intp)] * (rows[:, k + 1] > 0)
            return f

        def total_of(values):
            def f(rows, lo, hi):
                tot = np.zeros(hi - lo)
                for k in range(numfasc):
                    tot += rows[:, k + 1] * prop_of(values, k)(rows, lo, hi)
                return tot
            return f

        specs.append(('M0', (), col(0)))
        for k in range(numfasc):
            specs.append(('frac_f%d' % k, (), col(k + 1)))
            specs.append(('peak_f%d' % k, (3,), lambda rows, lo, hi, k=k: peaks_roi[lo:hi, 3 * k:3 * (k + 1)]))
        # fascicle-specific properties and their fraction-weighted voxel totals
        for prop in fitinfo['fasc_propnames']:
            values = np.asarray(fitinfo['_dict_' + prop])
            for k in range(numfasc):
                specs.append((prop + '_f%d' % k, (), prop_of(values, k)))
            specs.append((prop + '_tot', (), total_of(values)))
        if csf_on:
            specs.append(('frac_csf', (), col(2 * numfasc + 1)))
        if ear_on:
            c_ear = 2 * numfasc + csf_on + 1
            DIFF_ear = fitinfo['DIFF_ear']
            specs.append(('frac_ear', (), col(c_ear)))
            specs.append(('D_ear', (), lambda rows, lo, hi: DIFF_ear[rows[:, c_ear + 1].astype(np.intp)]
                          * (rows[:, c_ear] > 0)))
        specs.append(('MSE', (), col(model_params.shape[1] - 2)))
        specs.append(('R2', (), col(model_params.shape[1] - 1)))

        alloc = np.empty if full else np.zeros
        flat = {}
        for name, trailing, _ in specs:
            vol = alloc(mask.shape + trailing)
            setattr(self, name, vol)
            flat[name] = vol.reshape((-1,) + trailing)

        def fill(lo, hi):
            rows = model_params[lo:hi]
            where = slice(lo, hi) if full else roi_flat[lo:hi]
            for name, _, f in specs:
                flat[name][where] = f(rows, lo, hi)

        span = 1 << 16
        cuts = list(range(0, ROI_size, span)) + [ROI_size]
        n_thr = min(8, os.cpu_count() or 1, len(cuts) - 1)
        if n_thr <= 1:
            fill(0, ROI_size)
        else:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(n_thr) as ex:
                list(ex.map(lambda i: fill(cuts[i], cuts[i + 1]), range(len(cuts) - 1)))
        names = [name for name, _, _ in specs]
        self.param_names = names
        if verbose >= 2:
            print("Microstructure Fingerprinting fit object constructed. "
                  "'param_names' lists the property maps:")
            for p in names:
                print('\t%s' % (p,))
            print("Call 'write_nifti' to write the corresponding NIfTI files.")

    def write_nifti(self, output_basename, affine=None):
        """Exports every map as <basename>_<param>.nii[.gz] (reference mf.py:1177-1229).
        Returns the list of files created."""
        if affine is None:
            affine = self.affine
        if affine is None:
            raise ValueError("Argument affine must be explicitely passed  because "
                             "no affine transform matrix was found during model "
                             "fitting. Expecting NumPy array with shape (4, 4).")
        niigz = '.nii.gz'
        if len(output_basename) > len(niigz) and output_basename.endswith(niigz):
            path, fname = os.path.split(output_basename[:-len(niigz)])
            ext = niigz
        else:
            path, tail = os.path.split(output_basename)
            fname, ext = os.path.splitext(tail)
            if ext not in ['', '.nii']:
                raise ValueError("Unknown NIfTI extension %s in output %s" %
                                 (ext, output_basename))
            ext = '.nii'
        basename = os.path.join(path, fname)
        fnames = []
        for p in self.param_names:
            fn = '%s_%s%s' % (basename, p, ext)
            nifti.save(getattr(self, p), affine, fn)
            fnames.append(fn)
        return fnames
