"""Isolated runs of the one-fascicle kernel (k_single_fascicle: fused rotation + Gram terms +
closed forms, HBM / L2 bound) and of the reference-order pair kernel k_pairs<3> ([N, N, 1] in the
exact tier), for timing and ncu captures."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from tests.phantom import make_phantom  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "single"
if what == "single":
    ph = make_phantom(n_atoms=1000, n_vox=200000, seed=2, frac_k=(0, 1, 0), csf_frac=0.5)
    flags = 0
else:
    ph = make_phantom(n_atoms=1000, n_vox=256, seed=2, frac_k=(0, 0, 1), csf_frac=1.0)
    flags = 1      # exact tier: k_rotate_assemble + k_colstats + k_cross3 + k_pairs<3>
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rows = ph.gpu_rows(flags=flags)
    dt = time.perf_counter() - t0
print("%s: %d voxels in %.3f s = %.0f voxels/s (host buffers, plan creation included)" % (what, rows.shape[0], dt, rows.shape[0] / dt))
