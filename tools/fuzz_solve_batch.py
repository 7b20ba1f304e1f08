"""Randomized shapes: the screening tiers of mfb_solve_batch (pair scan M <= 111, general-M
pair scan, triple scan, CSF-projected triple scan of [N1, N2, 1, N4]) must return exactly what
the reference-order search returns.
python tools/fuzz_solve_batch.py [cases] [seed]"""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from microstructure_fingerprinting_b200 import _lib, mf_utils as mfu  # noqa: E402



def run(ncases=60, seed=2026, verbose=True):
    rng = np.random.default_rng(seed)
    bad = 0
    for case in range(ncases):
        bad += one_case(rng, verbose)
    return bad


def one_case(rng, verbose):
    kind = rng.choice(["pair", "pair_iso", "triple", "quad"])
    M = int(rng.choice([5, 17, 60, 105, 112, 113, 140, 300]))
    if kind == "triple":
        sizes = [int(rng.integers(2, 90)) for _ in range(3)]
    elif kind == "quad":        # two blocks + a single column + a small block (reference `_4up`)
        sizes = [int(rng.integers(2, 40)), int(rng.integers(2, 40)), 1, int(rng.integers(2, 8))]
    else:
        sizes = [int(rng.integers(8, 400)), int(rng.integers(8, 400))] + ([1] if kind == "pair_iso" else [])
    V = int(rng.integers(1, 40))
    nt = int(np.sum(sizes))
    signed = rng.random() < 0.25
    shared = rng.random() < 0.3
    base = rng.random((M, nt)) * np.exp(-3.0 * rng.random((1, nt)) * np.linspace(0, 1, M)[:, None])
    if signed:
        base *= rng.choice([-1.0, 1.0], size=(M, nt))
    A = base[None] * (1.0 + 0.05 * rng.standard_normal((V, M, nt)))
    st = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    wts = rng.uniform(0.0, 1.0, (V, len(sizes))) * (rng.random((V, len(sizes))) < 0.8)
    Y = np.stack([A[v][:, st + np.array([rng.integers(0, n) for n in sizes])] @ wts[v] for v in range(V)])
    Y += rng.choice([0.0, 1e-3, 0.05]) * rng.standard_normal(Y.shape)
    if rng.random() < 0.2:
        Y[0] = 0.0
    dic = A[0] if shared else A
    _lib.solve_stats(reset=True)
    fast = mfu.solve_exhaustive_posweights_batch(dic, Y, np.asarray(sizes))
    stats = _lib.solve_stats()
    exact = mfu.solve_exhaustive_posweights_batch(dic, Y, np.asarray(sizes), exact=True)
    ok = all(np.array_equal(f, e) for f, e in zip(fast, exact))
    if verbose or not ok:
        print("%-8s M %3d sizes %-16s V %2d signed %d shared %d: screened %2d redone %2d %s" %
              (kind, M, sizes, V, signed, shared, stats[0], stats[1], "ok" if ok else "MISMATCH"))
    if not ok:
        for v in range(V):
            if not all(np.array_equal(f[v], e[v]) for f, e in zip(fast, exact)):
                print("   voxel %d: fast sub %s w %s obj %.17g | exact sub %s w %s obj %.17g" %
                      (v, fast[1][v], fast[0][v], fast[3][v], exact[1][v], exact[0][v], exact[3][v]))
    return 0 if ok else 1


if __name__ == "__main__":
    n_bad = run(int(sys.argv[1]) if len(sys.argv) > 1 else 60, int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
    print("mismatching cases:", n_bad)
    sys.exit(1 if n_bad else 0)
