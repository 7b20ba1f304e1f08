"""Timing experiment: steady-state vs per-CTA overhead of k_fast_pairs (kernel-only ms from
mfb_fit_stats).  MFB_FAST_DEBUG=3 removes epilogue and gathers (results invalid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from microstructure_fingerprinting_b200 import mf_utils as mfu  # noqa: E402
from tests.phantom import make_phantom  # noqa: E402

dev = torch.device("cuda")
for N, V in [(1024, 8000), (4096, 600)]:
    ph = make_phantom(n_atoms=N, n_vox=V, seed=5, frac_k=(0, 0, 1), csf_frac=0.0)
    msi = mfu.init_PGSE_multishell_interp(ph.dic["dictionary"], ph.dic["sch_mat"], ph.dic["orientation"])
    plan = mfu.GpuPlan(msi, mfu.SchemePlan(msi, ph.sch), ph.sig_csf, None)
    y, pk = torch.from_numpy(ph.Y).to(dev), torch.from_numpy(ph.peaks).to(dev)
    K, c = torch.from_numpy(ph.K).to(dev), torch.from_numpy(ph.csf).to(dev)
    for i in range(2):
        try:
            plan.fit_device(y, pk, K, c, None, 2, True, False, flags=2)
        except Exception as e:
            print("ERR", str(e)[:100])
    st = plan.stats()
    fl = 2.0 * 108 * N * N * st[4]
    print("debug", os.environ.get("MFB_FAST_DEBUG"), "N", N, "kernel ms %.1f  DMMA %.2f TFLOP/s (%.1f%% of 37.1)"
          % (st[2], fl / st[2] / 1e9, 100 * fl / st[2] / 1e9 / 37.1))
    plan.close()
