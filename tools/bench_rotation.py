"""HBM-bound pieces measured alone: mfb_rotate_multishell (interp_PGSE_from_multishell
for a batch of directions: 8*M*N bytes written per direction, table reads from L2),
mfb_lerp_rows (rotate_atom / rotate_atom_2Dprotocol row lerp) and mfb_mc_average (spin average:
8*dim bytes read per spin and sequence group, one cos per spin and sequence)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from microstructure_fingerprinting_b200 import _lib, mf_utils as mfu  # noqa: E402
from tests.phantom import make_phantom  # noqa: E402

HBM = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if len(sys.argv) < 2 else float(sys.argv[1])
dev = torch.device("cuda")


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for scheme in ("exact", "between"):
    ph = make_phantom(n_atoms=1000, n_vox=8, seed=1, scheme=scheme)
    msi = mfu.init_PGSE_multishell_interp(ph.dic["dictionary"], ph.dic["sch_mat"], ph.dic["orientation"])
    plan = mfu.GpuPlan(msi, mfu.SchemePlan(msi, ph.sch), None, None)
    V, M, N = 8192, plan.M, plan.N
    rng = np.random.default_rng(0)
    d = rng.standard_normal((V, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d_dirs = torch.from_numpy(d).to(dev)
    out = torch.empty((V, M, N), dtype=torch.float64, device=dev)
    lib = _lib.load()
    ms = timed(lambda: _lib.check(lib.mfb_rotate_multishell(plan.handle, V, d_dirs.data_ptr(), out.data_ptr(), N, None)))
    gb = V * M * N * 8 / 1e9
    print("mfb_rotate_multishell %-7s V %d M %d N %d: %.2f ms, %.0f GB/s written (%.0f%% of the measured HBM copy rate %.0f GB/s, "
          "which counts read + write), %.2f M directions/s" % (scheme, V, M, N, ms, gb / ms * 1e3, 100 * gb / ms * 1e3 / HBM, HBM, V / ms / 1e3))
    plan.close()
    del out

# lerp rows at the AxCaliber shape
g = np.load("tests/golden/lowlevel_rotation.npz")
sch = g["ax_sch"]
M, N, V = sch.shape[0], 2000, 96
table = torch.rand((M + 9, N), dtype=torch.float64, device=dev)
rl = torch.randint(0, M, (V, M), dtype=torch.int32, device=dev); rh = torch.randint(0, M, (V, M), dtype=torch.int32, device=dev)
wl = torch.rand((V, M), dtype=torch.float64, device=dev); wh = 1 - wl
sc = torch.rand((V, M), dtype=torch.float64, device=dev)
out = torch.empty((V, M, N), dtype=torch.float64, device=dev)
lib = _lib.load()
ms = timed(lambda: _lib.check(lib.mfb_lerp_rows(0, V, M, N, table.data_ptr(), rl.data_ptr(), rh.data_ptr(), wl.data_ptr(),
                                                wh.data_ptr(), sc.data_ptr(), out.data_ptr(), N, None)))
gb = V * M * N * 8 / 1e9
print("mfb_lerp_rows V %d M %d N %d: %.2f ms, %.0f GB/s written (%.0f%% of %.0f)" % (V, M, N, ms, gb / ms * 1e3, 100 * gb / ms * 1e3 / HBM, HBM))
del out, table

# Monte-Carlo average: 3 reference sequences x 10^6 spins, 300 sequences
n_ref, n_spin, n_seq, dim = 3, 1000000, 300, 3
ph_ = torch.randn((n_ref * n_spin, dim), dtype=torch.float64, device=dev)
mp = torch.randint(0, n_ref, (n_seq,), dtype=torch.int64, device=dev)
gs = torch.rand((n_seq, dim), dtype=torch.float64, device=dev)
sig = torch.empty((n_seq,), dtype=torch.float64, device=dev)
ms = timed(lambda: _lib.check(lib.mfb_mc_average(0, n_ref * n_spin, dim, ph_.data_ptr(), n_seq, mp.data_ptr(), gs.data_ptr(), 0.9,
                                                 n_spin, sig.data_ptr(), None)))
print("mfb_mc_average %d sequences x %d spins (dim %d): %.2f ms, %.1f G spin-terms/s, %.0f GB/s of phase reads (L2-served across sequences)"
      % (n_seq, n_spin, dim, ms, n_seq * n_spin / ms / 1e6, n_seq * n_spin * dim * 8 / ms / 1e6))
