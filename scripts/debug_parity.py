"""GPU-vs-oracle field diff at increasing step counts (debug aid; run on the GPU box)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from tests import cases, golden_tools, oracle_tools as ot
from weldform_b200 import harness

name = sys.argv[1] if len(sys.argv) > 1 else "contact_100"
case, ref = golden_tools.load(name)
for ns in [int(a) for a in sys.argv[2:]] or [1, 2, 5, 10, 20, 40, 60, 80, 100]:
    c = dict(case); c["nsteps"] = float(ns)
    g, p = harness.run_case(c), ot.run_port(c)
    bad = {}
    for k in cases.FIELDS_F64 + cases.SCALARS + ["ID", "Nb", "pairs", "cont_pairs"]:
        if k in p and k in g:
            a, b = np.asarray(g[k]).reshape(-1), np.asarray(p[k]).reshape(-1)
            if a.shape != b.shape:
                bad[k] = "shape %s vs %s" % (a.shape, b.shape)
            elif not np.array_equal(a, b):
                bad[k] = "%.2e" % ot.max_rel(a, b)
    print(name, "steps", ns, "rebuilds", g["nb_rebuilds"], "maxpl %.3g" % p["pl_strain"].max(), "mismatch:", bad, flush=True)
