"""Worker of the multi-GPU parity test (launched by torch.distributed.run, one process per GPU): runs a WFC1 case file on
WORLD_SIZE ranks through the slab decomposition and writes each rank's dump to <outdir>/rank<r>.npz."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import wfc  # noqa: E402
from weldform_b200 import api, harness  # noqa: E402


def main():
    case_path, outdir = sys.argv[1], sys.argv[2]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    dist.init_process_group("gloo")                       # only carries the 128-byte NCCL id; the library owns its NCCL communicator
    torch.cuda.set_device(local)
    box = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    case = wfc.read(case_path)
    case = {k: (float(v[0]) if v.size == 1 and k not in ("ID",) else v) for k, v in case.items()}
    ctx, info = harness.context_from_case(case, device=local, comm=(rank, world, box[0]))
    try:
        ctx.solver_init(float(case.get("dt", 1.0e-7)))
        if int(case.get("mode", 0)) == 1:
            ctx.nb_build()
        elif float(case.get("tf", 0.0)) > 0.0:
            ctx.run_until(float(case["tf"]))
        else:
            ctx.run(int(case.get("nsteps", 1)))
        out = harness.dump(ctx, info, True)
        ci = ctx.comm_info()
        out["comm_info"] = np.array([ci[k] for k in ("n_local", "n_halo", "migrated_out", "migrated_in", "layer_lo", "layer_hi", "halo_bytes_per_step", "layers")], np.int64)
        np.savez(os.path.join(outdir, "rank%d.npz" % rank), **out)
    finally:
        ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
