"""Device-resident fit throughput at the benchmark shape, small and quick (kernel A/B runs:
MFB_LIB selects the library build).  usage: quick_fit_bench.py [voxels] [csf_frac] [flags] [snr]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from microstructure_fingerprinting_b200 import _lib, mf_utils as mfu  # noqa: E402
from tests.phantom import make_phantom  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
csf_frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 2     # 2: kernel bracketed by events (single stream); 0: production path
snr = float(sys.argv[4]) if len(sys.argv) > 4 else 30.0
ph = make_phantom(n_atoms=1000, n_vox=V, seed=100, frac_k=(0, 0, 1), csf_frac=csf_frac, snr=snr)
msi = mfu.init_PGSE_multishell_interp(ph.dic["dictionary"], ph.dic["sch_mat"], ph.dic["orientation"])
plan = mfu.GpuPlan(msi, mfu.SchemePlan(msi, ph.sch), ph.sig_csf, None)
dev = torch.device("cuda")
d = [torch.from_numpy(x).to(dev) for x in (ph.Y, ph.peaks, ph.K, ph.csf)]
out = torch.empty((V, 8), dtype=torch.float64, device=dev)
best, ref = 1e30, None
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plan.fit_device(d[0], d[1], d[2], d[3], None, 2, True, False, flags=flags, out=out)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rep:
        best = min(best, dt)
st = plan.stats()
F = 0.7 * (2.0 * 105 * 1e6 + 4.0 * 105 * 2000 + 210 + 25e6 + 3.0 * 105 * 2000) + \
    0.3 * (2.0 * 105 * (1e6 + 2000) + 4.0 * 105 * 2001 + 210 + 65e6 + 3.0 * 105 * 2000)
print("SNR %g, %s: %.0f voxels/s; pair kernel %.2f TFLOP/s algorithmic (%.1f ms per launch); exact-tier voxels %d; checksum %.6f"
      % (snr, _lib.LIB_PATH.split("/")[-1], V / best, F * st[4] / (st[2] / 1e3) / 1e12 if st[2] else 0, st[2] / max(st[3], 1), st[1],
         float(out[:, 0].sum())))
