"""Fit-path throughput of voxels with the EAR compartment: [N, N, E] (two fascicles + EAR) and
[N, N, 1, E] (+ CSF), N atoms per fascicle, E EAR atoms, through mfb_fit_host; share of voxels the
fast tier decides; fast tier == exact tier on a subsample.  usage: bench_fit_ear.py [N] [E] [V]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from microstructure_fingerprinting_b200 import mf_utils as mfu  # noqa: E402
from tests.phantom import make_phantom  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
E = int(sys.argv[2]) if len(sys.argv) > 2 else 10
V = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
ONLY = sys.argv[4] if len(sys.argv) > 4 else ""          # e.g. "NNE" or "NN1E": run one composition only
for csf_frac, label in ((0.0, "[N,N,E]"), (1.0, "[N,N,1,E]")):
    if ONLY and ONLY != label.replace("[", "").replace("]", "").replace(",", ""):
        continue
    for ear_frac, what in ((1.0, "EAR active in every voxel"), (0.5, "EAR signal in half of the voxels")):
        ph = make_phantom(n_atoms=N, n_vox=V, seed=61, frac_k=(0, 0, 1), csf_frac=csf_frac, ear=True, n_ear=E,
                          ear_frac=ear_frac, ear_max_k=2)
        ear = np.ones_like(ph.ear)
        msi = mfu.init_PGSE_multishell_interp(ph.dic["dictionary"], ph.dic["sch_mat"], ph.dic["orientation"])
        plan = mfu.GpuPlan(msi, mfu.SchemePlan(msi, ph.sch), ph.sig_csf, ph.sig_ear)
        args = (ph.Y, ph.peaks, ph.K, ph.csf, ear, 2, csf_frac > 0, True)
        plan.fit_host(*[a[:256] if isinstance(a, np.ndarray) else a for a in args])
        best = 1e30
        for rep in range(2):
            t0 = time.perf_counter()
            rows = plan.fit_host(*args)
            best = min(best, time.perf_counter() - t0)
        st = plan.stats()
        n = 48
        ex = plan.fit_host(*[a[:n] if isinstance(a, np.ndarray) else a for a in args], flags=1)
        t0 = time.perf_counter()
        ex = plan.fit_host(*[a[:n] if isinstance(a, np.ndarray) else a for a in args], flags=1)
        t_ex = (time.perf_counter() - t0) / n
        print("%-10s N %d E %d V %d, %s: %.0f voxels/s; fast tier %d, exact tier %d voxels (ill-conditioned %d, near ties / fewer "
              "columns %.6f); exact tier alone %.1f ms per voxel; first %d rows == exact tier: %s"
              % (label, N, E, V, what, V / best, st[0], st[1], st[6], st[7], t_ex * 1e3, n, bool(np.array_equal(rows[:n], ex))))
        plan.close()
