"""Generates tests/golden/digest/*.npz: DIGESTS of runs of the UNMODIFIED reference solver (oracle/_ref/wf_ref, Nproc = 1) at the sizes
BASELINE.json's configurations are stated on -- too large to commit as full dumps:

    compression_50k_40     examples/Compression.cpp's scale (dx = 0.0094 -> ~49 k particles), 40 Fraser steps through yield
    compression_1m_3       Compression_thermal at 1 M particles (dx = 0.00362), 3 steps
    compression_4m_2       Compression_thermal at 4 M particles (dx = 0.00228), 2 steps: the largest the reference runs in this container's memory
    contact_1m_10 / 2m_10  the cylinder between two rigid planes at 1 M / 2 M particles (Contact_Compression_1010's scale), 10 steps with contact forces

Run in the build container only:   python tests/golden/make_digest_golden.py [name ...]
(At millions of particles the reference spends far longer tearing its Domain down than running: compression_4m_2 wrote its dump after 11 minutes and was
still freeing memory 85 minutes later; its digest was built from the finished dump with digest_of() and the process stopped.)

A digest holds: the arguments that regenerate the case (tests/cases.compression_fast), the pair count, sum(Nb), for every fp64 field of the
dump its sum / min / max / sum of squares over all particles, and the full rows of SAMPLE_ROWS particles (fixed seed) for the fields in
SAMPLE_FIELDS. The GPU test (tests/test_parity_gpu.py::test_baseline_scale_digest) regenerates the case, runs it and compares element-wise.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import cases, oracle_tools as ot  # noqa: E402

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "digest")
SAMPLE_ROWS = 10000
SAMPLE_FIELDS = ["x", "v", "a", "Displacement", "Density", "Pressure", "pl_strain", "T", "Sigmay", "dTdt", "Sigma", "StrainRate", "contforce", "delta_cont", "normal"]
SPECS = {
    "compression_50k_40": dict(dx=0.0094, nsteps=40, perturb_amp=150.0),
    "compression_1m_3": dict(dx=0.00362, nsteps=3, perturb_amp=60.0),
    # 4 M of BASELINE's ten million particles: the largest the reference runs in the build container's 62 GB (5 M ends in std::bad_alloc: per-pair tables)
    "compression_4m_2": dict(dx=0.00228, nsteps=2, perturb_amp=60.0),
    # BASELINE configs[2] (contact against rigid planes): the cylinder between two rigid planes (penalty contact, friction 0.2), the upper one driven
    # down fast enough that a few hundred surface particles are in contact within the ten steps (1 M and 2 M particles)
    "contact_1m_10": dict(dx=0.00362, nsteps=10, perturb_amp=60.0, contact_planes=1.0, vmax=500.0, tau=1.0e-9),
    "contact_2m_10": dict(dx=0.0028744663089468095, nsteps=10, perturb_amp=60.0, contact_planes=1.0, vmax=500.0, tau=1.0e-9),
}


def sample_index(n):
    return np.sort(np.random.default_rng(20261018).choice(n, size=min(SAMPLE_ROWS, n), replace=False))


def digest_of(dump, n):
    """dump: state dict of a run (reference, port or CUDA path); n = solid particle count"""
    out = {}
    idx = sample_index(n)
    for k in cases.FIELDS_F64:
        if k not in dump:
            continue
        v = np.asarray(dump[k], dtype=np.float64)
        nc = v.size // int(dump["n_particles"][0])
        v = v.reshape(-1, nc)[:n]
        out["stat/" + k] = np.array([v.sum(), v.min(), v.max(), (v * v).sum()])
        if k in SAMPLE_FIELDS:
            out["rows/" + k] = v[idx].copy()
    out["Nb_sum"] = np.array(int(np.asarray(dump["Nb"])[:n].sum()))
    out["Nb_rows"] = np.asarray(dump["Nb"])[:n][idx].astype(np.int32)
    out["ID_rows"] = np.asarray(dump["ID"])[:n][idx].astype(np.int32)
    out["pairs"] = np.array(dump["pairs"].size // 2 if "pairs" in dump else -1)
    out["cont_pairs"] = np.array(dump["cont_pairs"].size // 2 if "cont_pairs" in dump else 0)
    for k in cases.SCALARS:
        if k in dump:
            out["scalar/" + k] = np.asarray(dump[k], dtype=np.float64)
    return out


def main():
    if not ot.have_ref():
        raise SystemExit("oracle/_ref/wf_ref missing: run `make -C oracle ref` in the container that has /root/reference")
    os.makedirs(HERE, exist_ok=True)
    only = set(sys.argv[1:])
    for name, spec in SPECS.items():
        if only and name not in only:
            continue
        t0 = time.time()
        case = cases.compression_fast(**{k: (bool(v) if k == "contact_planes" else v) for k, v in spec.items()})
        ref = ot.run_ref(case, nproc=1, timeout=6 * 3600)
        n = int(ref["solid_part_count"][0])
        dg = digest_of(ref, n)
        dg["spec_keys"] = np.array(sorted(spec))
        dg["spec_vals"] = np.array([float(spec[k]) for k in sorted(spec)])
        dg["n"] = np.array(n)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **dg)
        print("%-22s N=%d pairs=%d maxpl=%.3g yielding particles=%d  %.0f s  %.0f kB" % (name, n, int(dg["pairs"]), ref["pl_strain"].max(), int((ref["pl_strain"] > 0).sum()),
                                                                                     time.time() - t0, os.path.getsize(path) / 1e3), flush=True)


if __name__ == "__main__":
    main()
