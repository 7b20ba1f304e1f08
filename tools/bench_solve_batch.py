"""BASELINE config 2: solve_exhaustive_posweights on explicit per-voxel dictionaries
(numfasc = 2, ~800 atoms per fascicle, M = 105, optional CSF column), device-resident."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from microstructure_fingerprinting_b200 import mf_utils as mfu  # noqa: E402
from tests.phantom import make_phantom  # noqa: E402

N, V = 800, 4096
ph = make_phantom(n_atoms=N, n_vox=V, seed=11, frac_k=(0, 0, 1), csf_frac=0.0)
msi = mfu.init_PGSE_multishell_interp(ph.dic["dictionary"], ph.dic["sch_mat"], ph.dic["orientation"])
plan = mfu.GpuPlan(msi, mfu.SchemePlan(msi, ph.sch), ph.sig_csf, None)
M = ph.Y.shape[1]
dev = torch.device("cuda")
for csf in (0, 1):
    ntot = 2 * N + csf
    A = torch.empty((V, M, ntot), dtype=torch.float64, device=dev)
    A[:, :, :N] = plan.rotate(ph.peaks[:, :3])
    A[:, :, N:2 * N] = plan.rotate(ph.peaks[:, 3:6])
    if csf:
        A[:, :, 2 * N] = torch.from_numpy(ph.sig_csf).to(dev)[None, :]
    Y = torch.from_numpy(ph.Y).to(dev)
    sizes = np.array([N, N] + ([1] if csf else []))
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = mfu.solve_exhaustive_posweights_batch(A, Y, sizes, return_device=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    F = 2.0 * M * N * N + 4.0 * M * ntot + (65.0 if csf else 25.0) * N * N
    print("sizes %s: %.1f voxels/s, %.2f TFLOP/s algorithmic (%.0f%% of DGEMM peak 35.47)" %
          (sizes.tolist(), V / dt, F * V / dt / 1e12, 100 * F * V / dt / 35.47e12))
