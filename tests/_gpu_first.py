import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
from tests import oracle_tools as ot, cases
from weldform_b200 import harness
def cmp(c, tag, pairs=True):
    t=time.time(); g = harness.run_case(c); tg=time.time()-t
    t=time.time(); p = ot.run_port(c); tp=time.time()-t
    bad = []
    for k in cases.FIELDS_F64+cases.SCALARS:
        if k in g and not np.array_equal(g[k], p[k]): bad.append((k, "%.2e"%ot.max_rel(g[k], p[k])))
    print(tag, "gpu %.2fs port %.2fs"%(tg,tp), "N", int(p["n_particles"][0]), "maxpl %.3g"%p["pl_strain"].max(), "rebuilds", g["nb_rebuilds"], "\n   mismatch:", bad,
          "\n   pairs", np.array_equal(g["pairs"], p["pairs"]), g["pairs"].size//2, p["pairs"].size//2, "ID", np.array_equal(g["ID"], p["ID"]), "Nb", np.array_equal(g["Nb"], p["Nb"]),
          "cont", np.array_equal(g["cont_pairs"], p["cont_pairs"]), g["cont_pairs"].size//2, p["cont_pairs"].size//2, flush=True)
c = cases.perturbed(cases.compression(dx=0.03, nsteps=1)); c["mode"]=1.0
cmp(c, "nb-only")
for ns in (1, 3, 60):
    cmp(cases.perturbed(cases.compression(dx=0.03, nsteps=ns, ti=0.0), amp=60.0), "noTI steps=%d"%ns)
for ns in (1, 60):
    cmp(cases.perturbed(cases.compression(dx=0.03, nsteps=ns), amp=60.0), "TI steps=%d"%ns)
for model in (1,2):
    cmp(cases.perturbed(cases.compression(dx=0.03, nsteps=60, model=model), amp=60.0), "model%d"%model)
for ns in (1, 100):
    cmp(cases.contact_compression(dx=0.0024, nsteps=ns, vmax=5.0, tau=1e-5), "contact steps=%d"%ns)
