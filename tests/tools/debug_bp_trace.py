"""Debug: where do the GPU and oracle BP traces part?  python tests/tools/debug_bp_trace.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
import admm_b200
from admm_b200 import _capi as K
from oracle import pyoracle as O

for seed in (9, 10, 11, 12):
    rng = np.random.default_rng(seed)
    n, p, k = 500, 5000, 40
    x = np.asfortranarray(rng.normal(size=(n, p)))
    bt = np.zeros(p)
    bt[rng.choice(p, k, replace=False)] = rng.uniform(0.5, 1.5, size=k)
    y = x @ bt
    with K.trace(which=0, cap=4000) as tr:
        f = admm_b200.admm_bp(x, y).fit()
    o = O.bp(x, y, trace_cap=4000)
    tg, tc = tr.rows, o["trace"][:o["niter"]]
    m = min(len(tg), len(tc))
    rel = np.abs(tg[:m] - tc[:m]) / np.maximum(np.abs(tc[:m]), 1e-300)
    bad = np.where(rel.max(axis=1) > 1e-7)[0]
    print("seed", seed, "niter", f.niter, o["niter"], "first divergent row", (int(bad[0]) if len(bad) else None))
    if len(bad):
        i = int(bad[0])
        for r in range(max(0, i - 3), min(m, i + 3)):
            print("  it %3d gpu %s" % (r, np.array2string(tg[r], precision=12)))
            print("         cpu %s" % (np.array2string(tc[r], precision=12)))
