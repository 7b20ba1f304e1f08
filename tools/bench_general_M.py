"""General-M screening path (k_normalize + k_gemm_pairs) on explicit per-voxel dictionaries:
BASELINE config 5 shape (AxCaliber-like, M = 1776, 2000 atoms per fascicle) and the HCP
shape of the reference's test_hcp_dict (M = 552, 782 atoms), device-resident."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from microstructure_fingerprinting_b200 import mf_utils as mfu  # noqa: E402

dev = torch.device("cuda")
cases = [(1776, 2000, 64, 0), (1776, 2000, 64, 1), (552, 782, 512, 1), (271, 1000, 1024, 0)]
if len(sys.argv) > 1:
    cases = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for M, N, V, csf in cases:
    g = torch.Generator(device=dev).manual_seed(M + N)
    nt = 2 * N + csf
    base = torch.rand((M, nt), generator=g, device=dev, dtype=torch.float64) * torch.exp(
        -3.0 * torch.rand((1, nt), generator=g, device=dev, dtype=torch.float64) *
        torch.linspace(0, 1, M, device=dev, dtype=torch.float64)[:, None])
    A = base[None] * (1.0 + 0.05 * torch.randn((V, M, nt), generator=g, device=dev, dtype=torch.float64))
    i1 = torch.randint(0, N, (V,), generator=g, device=dev)
    i2 = torch.randint(0, N, (V,), generator=g, device=dev) + N
    ar = torch.arange(V, device=dev)
    Y = 0.6 * A[ar, :, i1] + 0.4 * A[ar, :, i2]
    Y = Y + 0.02 * torch.randn(Y.shape, generator=g, device=dev, dtype=torch.float64)
    sizes = np.array([N, N] + ([1] if csf else []))
    best = 1e30
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = mfu.solve_exhaustive_posweights_batch(A, Y, sizes, return_device=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    ok = float((out[1][:, 0] == i1).double().mean()), float((out[1][:, 1] == (i2 - N)).double().mean())
    F = 2.0 * M * N * N + 4.0 * M * nt + (65.0 if csf else 25.0) * N * N
    print("M %d sizes %s V %d: %.1f voxels/s, %.2f TFLOP/s algorithmic (%.0f%% of DGEMM peak 35.47); planted atoms recovered %.2f/%.2f" %
          (M, sizes.tolist(), V, V / best, F * V / best / 1e12, 100 * F * V / best / 35.47e12, ok[0], ok[1]))
    del A, Y, out
    torch.cuda.empty_cache()
