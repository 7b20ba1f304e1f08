"""Large randomized validation: the screening (fast) tier must return exactly the rows of
the reference-order (exact) tier.  python tools/validate_fast_vs_exact.py [voxels per case]"""
import sys

import numpy as np

sys.path.insert(0, ".")
from microstructure_fingerprinting_b200 import mf_utils as mfu  # noqa: E402
from tests.phantom import make_phantom  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
cases = [(1000, 30.0, 0.3, "exact"), (1000, 100.0, 0.5, "exact"), (500, 10.0, 0.3, "exact"), (257, 1e6, 0.5, "exact"),
         (300, 30.0, 0.4, "between"),
         (1000, 5.0, 0.3, "exact"), (777, 3.0, 0.5, "exact")]      # low SNR: one-atom winners, per-atom restriction masks
bad = 0
for n_atoms, snr, csf_frac, scheme in cases:
    nv = V if scheme == "exact" else V // 4
    ph = make_phantom(n_atoms=n_atoms, n_vox=nv, seed=int(snr) + n_atoms, frac_k=(0.0, 0.05, 0.95), csf_frac=csf_frac,
                      snr=snr, scheme=scheme)
    msi = mfu.init_PGSE_multishell_interp(ph.dic["dictionary"], ph.dic["sch_mat"], ph.dic["orientation"])
    plan = mfu.GpuPlan(msi, mfu.SchemePlan(msi, ph.sch), ph.sig_csf, None)
    fast = plan.fit_host(ph.Y, ph.peaks, ph.K, ph.csf, None, 2, True, False, flags=0)
    st = plan.stats()
    exact = plan.fit_host(ph.Y, ph.peaks, ph.K, ph.csf, None, 2, True, False, flags=1)
    plan.close()
    diff = int(np.sum(np.any(fast != exact, axis=1)))
    bad += diff
    print("N %d snr %g csf %.1f %s: %d voxels, screened %d, exact %d, rows differing %d" %
          (n_atoms, snr, csf_frac, scheme, nv, st[0], st[1], diff))
print("TOTAL differing rows:", bad)
sys.exit(1 if bad else 0)
