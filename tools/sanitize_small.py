"""One small invocation of every hand-synchronised kernel (mbarrier rings, voxel-wide
atomicMax thresholds) for compute-sanitizer:

    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_small.py [case ...]

cases: fit        MFModel.fit path, table source      k_fast_pairs<0,0>, k_fast_pairs<1,0>
       pairs      mfb_solve_batch, M <= 111           k_fast_pairs<0,1>, k_fast_pairs<1,1>
       gemm       mfb_solve_batch, M > 111            k_gemm_pairs<0,0>, k_gemm_pairs<1,0>
       triples    mfb_solve_batch, three blocks       k_gemm_pairs<0,1>, k_triples
Sizes are tiny (the tools slow kernels down 10-100x); results are checked against the exact
tier so that a sanitizer-clean run is also a correct one.
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from microstructure_fingerprinting_b200 import _lib, mf_utils as mfu  # noqa: E402
from tests.phantom import make_phantom  # noqa: E402


def problem(sizes, M, V, seed):
    rng = np.random.default_rng(seed)
    nt = int(np.sum(sizes))
    base = rng.random((M, nt)) * np.exp(-3.0 * rng.random((1, nt)) * np.linspace(0, 1, M)[:, None])
    A = base[None] * (1.0 + 0.05 * rng.standard_normal((V, M, nt)))
    st = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    Y = np.stack([A[v][:, st + np.array([rng.integers(0, n) for n in sizes])] @ rng.random(len(sizes)) for v in range(V)])
    return A, Y + 0.02 * rng.standard_normal(Y.shape)


def check(sizes, M, V, seed):
    A, Y = problem(sizes, M, V, seed)
    n0 = _lib.launch_count()
    fast = mfu.solve_exhaustive_posweights_batch(A, Y, np.asarray(sizes))
    exact = mfu.solve_exhaustive_posweights_batch(A, Y, np.asarray(sizes), exact=True)
    assert all(np.array_equal(f, e) for f, e in zip(fast, exact)), sizes
    print("solve_batch", sizes, "M", M, "V", V, "ok,", _lib.launch_count() - n0, "launches")


cases = sys.argv[1:] or ["fit", "pairs", "gemm", "triples"]
if "fit" in cases:
    ph = make_phantom(n_atoms=160, n_vox=12, seed=5, frac_k=(0, 0, 1), csf_frac=0.5)
    rows = ph.gpu_rows()
    assert np.array_equal(rows, ph.gpu_rows(flags=1))
    print("fit path ok:", rows.shape)
if "pairs" in cases:
    check([160, 130], 100, 6, 1)
    check([150, 140, 1], 100, 6, 2)
if "gemm" in cases:
    check([140, 130], 150, 4, 3)
    check([130, 140, 1], 130, 4, 4)
if "triples" in cases:
    check([40, 36, 24], 60, 4, 5)
print("done")
