"""Randomized MFModel.fit-path problems: rows of the screening tiers (flags = 0) must equal the
rows of the reference-order tier (flags = 1) bit for bit.  Varies dictionary size, protocol
(exact-G, between-shell, M = 271), SNR, CSF / EAR fractions, fascicle-count mix, and plants
degenerate voxels (identical peaks, zero / negative signals, single-fascicle data).
python tools/fuzz_fit.py [cases] [seed]"""
import sys

import numpy as np

sys.path.insert(0, ".")
from microstructure_fingerprinting_b200 import mf_utils as mfu  # noqa: E402
from tests.phantom import make_phantom, rotate_columns  # noqa: E402


def one_case(rng, verbose):
    n_atoms = int(rng.choice([8, 13, 40, 97, 128, 200, 333]))
    scheme = str(rng.choice(["exact", "exact", "between", "dense"]))
    snr = float(rng.choice([5.0, 30.0, 200.0, 1e9]))
    ear = bool(rng.random() < 0.4)
    V = int(rng.integers(50, 400))
    fk = rng.random(3) + 0.05
    fk[2] += 1.0
    ph = make_phantom(n_atoms=n_atoms, n_vox=V, seed=int(rng.integers(1 << 30)), frac_k=tuple(fk),
                      csf_frac=float(rng.choice([0.0, 0.3, 0.8])), ear=ear, n_ear=int(rng.integers(2, 7)),
                      ear_frac=0.4, ear_max_k=2, snr=snr, scheme=scheme)
    if ph.maxfasc == 2:
        k2 = np.where(ph.K == 2)[0]
        if k2.size >= 8:
            ph.peaks[k2[0], 3:6] = ph.peaks[k2[0], 0:3]                    # identical peaks
            ph.peaks[k2[1], 3:6] = -ph.peaks[k2[1], 0:3]                   # antipodal peaks
            ph.Y[k2[2]] = 0.0
            ph.Y[k2[3]] = -np.abs(ph.Y[k2[3]])
            ph.Y[k2[4]] = 300.0 * rotate_columns(ph.dic, ph.sch, ph.peaks[k2[4]:k2[4] + 1, :3], ph.atoms[k2[4]:k2[4] + 1, 0])[0]
            ph.Y[k2[5]] = 300.0 * ph.sig_csf
    msi = mfu.init_PGSE_multishell_interp(ph.dic["dictionary"], ph.dic["sch_mat"], ph.dic["orientation"])
    plan = mfu.GpuPlan(msi, mfu.SchemePlan(msi, ph.sch), ph.sig_csf if ph.csf_on else None,
                       ph.sig_ear if ph.ear_on else None)
    args = (ph.Y, ph.peaks, ph.K, ph.csf, ph.ear if ph.ear_on else None, ph.maxfasc, ph.csf_on, ph.ear_on)
    fast = plan.fit_host(*args, flags=0)
    st = plan.stats()
    exact = plan.fit_host(*args, flags=1)
    plan.close()
    diff = int(np.sum(np.any(fast != exact, axis=1)))
    if verbose or diff:
        print("N %3d %-7s snr %-6g ear %d V %3d maxfasc %d csf %d: screened %3d exact %3d one-fascicle %3d differing %d"
              % (n_atoms, scheme, snr, ear, V, ph.maxfasc, ph.csf_on, st[0], st[1], st[5], diff))
    return diff


def run(ncases=40, seed=1, verbose=True):
    rng = np.random.default_rng(seed)
    return sum(one_case(rng, verbose) for _ in range(ncases))


if __name__ == "__main__":
    bad = run(int(sys.argv[1]) if len(sys.argv) > 1 else 40, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    print("differing rows:", bad)
    sys.exit(1 if bad else 0)
