"""Generate tests/golden/drivers/*.npz: digests of what the reference's OWN example drivers write when they are run here.

TEST INFRASTRUCTURE (build container only; needs /root/reference through oracle/Makefile's `drivers` target).

    python tests/golden/drivers/make_driver_goldens.py [driver ...]

Each driver binary (oracle/_ref/drv_<name>, built from /root/reference/examples/<name>.cpp, see oracle/Makefile for which are the
unmodified reference and which carry the one-line m_contact_force fix) is started in a scratch directory with OMP_THREAD_LIMIT=1:
the binary still asks for its Nproc = 4 threads, the OpenMP runtime grants one, every parallel loop runs serially and the pair lists
come out in the reference's serial discovery order -- the reproducible configuration (with more threads CalcTempInc adds into the
particles without locks, tests/test_oracle.py). The run is stopped once the output tables listed in PLAN exist (the full runs of the
contact drivers would take days of CPU time: FSW.cpp asks for ~1e6 steps of 41 405 particles).

Stored per driver: the WriteCSV tables (Source/Output.cpp:520-554) named in PLAN as float32 (the text carries 6 significant digits,
which float32 round-trips), History.csv and Prop.csv rows up to the last kept table, and the step count at each kept table.
"""
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
REF = os.path.join(ROOT, "oracle", "_ref")

# driver -> (file key, output indices kept, row stride of the kept tables)
PLAN = {
    "Compression_thermal": ("test06", [1, 2, 10, 30, 60, 106], 1),           # whole run: 2 9xx steps, 106 tables; the body yields around table 30
    "Compression": ("test06", [1], 1),                                      # Gradient_Approach_Set(2): unstable in this snapshot, heap corruption after ~34 steps
    "Contact_Compression_1010": ("test06", [1, 2, 3, 5, 8], 1),             # 457 steps per table
    "FSW": ("test06", [1, 2, 3], 2),                                        # 922 steps per table
    "Compression_1010": ("test06", [1, 2, 3], 1),                           # Fraser, Hollomon, gradKernelCorr; 92 steps per table
    "Compression_1010_half_ghost": ("test06", [1, 2, 3], 1),                # Fraser, ghosts
    "CompressionVel": ("test06", [1, 2, 3, 4], 1),                          # KickDrift, 70 steps per table
    "Compression_TRIP_vel": ("test06", [1, 2, 3], 1),                       # KickDrift, gradKernelCorr
    "Contact_Compression_half_ghost": ("test06", [1, 20, 40, 60], 1),      # KickDrift, adaptive time step, ghosts, one rigid plane
    # third set (oracle/Makefile MORE_DRIVERS)
    "Contact_Compression_TRIP_2surf": ("test06", [1], 2),                   # the reference explodes right after (plastic strain 739 at table 3)
    "Contact_Compression_1010_1side": ("test06", [1, 2, 3], 2),
    "Contact_Compression_2surf_short": ("test06", [1, 5, 10, 15], 2),
    "Contact_Compression_vel": ("test06", [1, 4, 8], 2),
    "Contact_Compression_1010_unif": ("test06", [1, 2, 3], 2),
    "Contact_Compression_thermal": ("test06", [1, 5, 10, 15], 2),
    "Contact_Compression_Rot": ("test06", [1, 5, 10, 15], 2),
    "FSW_thin": ("test06", [1, 10, 20], 8),                                 # 194 262 particles
    "slice_test": ("test06", [1, 2], 4),
    "Traction": ("test06", [1, 10, 20, 30], 2),
    "Traction_JC": ("test06", [1, 5], 2),                                   # from table 10 on the reference's 1- and 4-thread runs differ by O(100): no information
    "Traction_Hollomon": ("test06", [1, 10, 20, 30], 2),
    "Twisting": ("test06", [1, 100, 200, 400], 2),
}
NEEDS = {"FSW": ["Tool.nas"], "FSW_thin": ["Tool.nas"]}


def read_csv_table(path):
    with open(path) as f:
        head = f.readline()
        rows = [[float(t) for t in line.replace(",", " ").split()] for line in f if line.strip()]
    return head.strip(), np.asarray(rows, dtype=np.float64)


def read_rows(path):
    """History.csv / Prop.csv: header line (History only) + numeric rows; ragged or empty fields are kept as NaN."""
    out = []
    if not os.path.exists(path):
        return np.zeros((0, 0))
    with open(path) as f:
        for line in f:
            toks = [t.strip() for t in line.strip().split(",")]
            try:
                out.append([float(t) if t else np.nan for t in toks])
            except ValueError:
                continue                      # header
    w = max((len(r) for r in out), default=0)
    return np.asarray([r + [np.nan] * (w - len(r)) for r in out], dtype=np.float64).reshape(len(out), w)


def run_driver(name, exe, key, wanted, timeout_s, env_extra=None, workdir=None):
    """Start `exe` in a scratch directory, wait until <key>_<last wanted>.csv and its successor's predecessor are complete, stop it.
    Returns (directory, return code or None if it was stopped)."""
    d = workdir or tempfile.mkdtemp(prefix="drv_%s_" % name)
    for f in NEEDS.get(name, []):
        shutil.copy(os.path.join(REF, f), os.path.join(d, f))
    env = dict(os.environ, OMP_THREAD_LIMIT="1")
    env.update(env_extra or {})
    log = open(os.path.join(d, "log.txt"), "w")
    p = subprocess.Popen([exe], cwd=d, stdout=log, stderr=subprocess.STDOUT, env=env)
    last = os.path.join(d, "%s_%04d.csv" % (key, max(wanted)))
    t0 = time.time()
    rc = None
    while time.time() - t0 < timeout_s:
        rc = p.poll()
        if rc is not None:
            break
        if os.path.exists(last) and time.time() - os.path.getmtime(last) > 3.0:   # the table is written in one go from a string stream
            break
        time.sleep(1.0)
    if p.poll() is None:
        p.terminate()
        try:
            p.wait(timeout=10)
        except subprocess.TimeoutExpired:
            p.kill()
        rc = None
    log.close()
    return d, rc


def table_blocks(logpath):
    """The reference prints `step number: N` ... `Output No. K at T has been generated` in the block it writes with every table
    (SolverFraser.cpp:418-452, SolverKickDrift.cpp:332-361) -- and also whenever 60 s of CPU time passed without one, so only the first block of each K counts.
    K is printed after the increment: the block of table file idx shows K = idx + 1. Returns {idx: (steps, time)}."""
    out, step = {}, None
    with open(logpath, errors="replace") as f:
        for line in f:
            if line.startswith("step number:"):
                step = int(line.split(":")[1])
            elif line.startswith("Output No."):
                toks = line.split()
                k, t = int(toks[2]), float(toks[4])
                out.setdefault(k - 1, (-1 if step is None else step, t))   # the KickDrift / LeapFrog loops do not print the step number
    return out


def digest(d, key, wanted, stride):
    out = {}
    blocks = table_blocks(os.path.join(d, "log.txt"))
    kept = []
    for idx in wanted:
        path = os.path.join(d, "%s_%04d.csv" % (key, idx))
        if not os.path.exists(path) or idx not in blocks:
            continue
        head, tab = read_csv_table(path)
        out["header"] = np.array(head)
        out["table_%04d" % idx] = tab[::stride].astype(np.float32)
        out["rows_%04d" % idx] = np.array(tab.shape[0])
        kept.append(idx)
    tlast = blocks[max(kept)][1]
    hist = read_rows(os.path.join(d, "History.csv"))
    prop = read_rows(os.path.join(d, "Prop.csv"))
    # History / Prop get a row per log block, i.e. also at the wall-clock-triggered ones: keep every row up to the last kept table's time;
    # the mirror writes one row per table and the test matches rows by their time stamp
    out["history"] = hist[hist[:, 0] <= tlast * (1 + 1e-9)] if hist.size else hist
    out["prop"] = prop[prop[:, 0] <= tlast * (1 + 1e-9)] if prop.size else prop
    out["stride"] = np.array(stride)
    out["kept"] = np.asarray(kept, dtype=np.int64)
    out["steps_at_table"] = np.asarray([blocks[k][0] for k in kept], dtype=np.int64)
    out["time_at_table"] = np.asarray([blocks[k][1] for k in kept], dtype=np.float64)
    return out


# drivers whose runs amplify round-off (time step at the stability limit, contact onsets): a SECOND reference run with four OpenMP threads (another
# summation order, nothing else changed) is stored next to the golden as alt_<idx>, so that the test can tell the reference's own reproducibility
# from a real difference:   python tests/golden/drivers/make_driver_goldens.py --alt [driver ...]
ALT = ["Contact_Compression_2surf_short", "Contact_Compression_1010_unif", "Contact_Compression_thermal", "Contact_Compression_Rot",
       "Traction", "Traction_JC", "Traction_Hollomon"]


def add_alt(name):
    key, wanted, stride = PLAN[name]
    path = os.path.join(HERE, name + ".npz")
    g = dict(np.load(path))
    t0 = time.time()
    d, rc = run_driver(name, os.path.join(REF, "drv_" + name), key, g["kept"].tolist(), timeout_s=6 * 3600, env_extra={"OMP_THREAD_LIMIT": "4"})
    for idx in g["kept"].tolist():
        _, tab = read_csv_table(os.path.join(d, "%s_%04d.csv" % (key, idx)))
        g["alt_%04d" % idx] = tab[::int(g["stride"])].astype(np.float32)
    np.savez_compressed(path, **g)
    print("%s: alt run (4 threads) %.0f s" % (name, time.time() - t0), flush=True)
    shutil.rmtree(d, ignore_errors=True)


def main(argv):
    if argv and argv[0] == "--alt":
        for name in (argv[1:] or ALT):
            add_alt(name)
        return
    names = argv or list(PLAN)
    for name in names:
        key, wanted, stride = PLAN[name]
        exe = os.path.join(REF, "drv_" + name)
        t0 = time.time()
        d, rc = run_driver(name, exe, key, wanted, timeout_s=6 * 3600)
        dg = digest(d, key, wanted, stride)
        dg["returncode"] = np.array(-999 if rc is None else rc)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **dg)
        print("%s: %.0f s, rc=%s, tables %s, steps %s -> %s.npz" % (name, time.time() - t0, rc, [k for k in dg if k.startswith("table_")],
                                                                    dg["steps_at_table"].tolist()[-3:], name), flush=True)
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main(sys.argv[1:])
