"""MFModel.fit path (numfasc = 2, CSF on 30 % of the voxels, N = 1000) on the three protocol
kinds: gradient strengths that match the dense shells exactly (fused table-source kernels),
between-shell strengths and M = 271 (screening on materialised dictionaries)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from microstructure_fingerprinting_b200 import mf_utils as mfu  # noqa: E402
from tests.phantom import make_phantom  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
for scheme in ("exact", "between", "dense"):
    ph = make_phantom(n_atoms=1000, n_vox=V, seed=4, frac_k=(0.0, 0.0, 1.0), csf_frac=0.3, scheme=scheme)
    msi = mfu.init_PGSE_multishell_interp(ph.dic["dictionary"], ph.dic["sch_mat"], ph.dic["orientation"])
    plan = mfu.GpuPlan(msi, mfu.SchemePlan(msi, ph.sch), ph.sig_csf, None)
    best = 1e30
    for rep in range(3):
        t0 = time.perf_counter()
        rows = plan.fit_host(ph.Y, ph.peaks, ph.K, ph.csf, None, 2, True, False, flags=0)
        best = min(best, time.perf_counter() - t0)
    st = plan.stats()
    plan.close()
    print("%-7s M %3d: %.0f voxels/s (host buffers), screened %d, reference-order %d" % (scheme, ph.Y.shape[1], V / best, st[0], st[1]))
