"""world_size-2 gloo worker for test_two_rank_partition_and_grid_reduce_gloo:
the oracle stands in for the engine (no GPU here); partitioning and the grid
reduce are the product's own flexpart_b200.parallel code."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import flexpart_b200 as fb  # noqa: E402
from flexpart_b200 import parallel  # noqa: E402
import cases  # noqa: E402
from oracle_api import Oracle  # noqa: E402


def run(cb, parts, met, nsteps):
    o = Oracle(cb)
    o.fill_rannumb(50000, -320)
    o.upload_met(1, met[0]); o.upload_met(2, met[1])
    o.set_met_bracket((1, 2), (0, 10800))
    o.push_particles(parts)
    for k in range(nsteps):
        o.conccalc(k * 900, 1.0)
        o.step(k * 900)
    return o.fetch_grids(zero_conc=False)["gridunc"]


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 3001  # odd on purpose: ranks get 1501 / 1500 rows
    cb = cases.config_small(nrel=2, npart_each=2000, turboff=1, ctl=-5.0)
    # homogeneous wind: no interpolation spread -> the mesoscale term vanishes and the
    # trajectories do not depend on the order the ran3 stream is consumed in
    met = (fb.MetFields(cb).homogeneous(12.0, 4.0, 0.0), fb.MetFields(cb).homogeneous(12.0, 4.0, 0.0))
    allp = cases.seeded_particles(cb, n, zmax=9000.0)
    mine = parallel.take_partition(allp, rank, world, fb.Particles)
    assert mine.numpart == len(parallel.partition_rows(n, rank, world))
    cbr = cb.clone(maxpart=max(mine.numpart, 1), part_id_stride=world, part_id_offset=rank)
    g = run(cbr, mine, met, 3)
    t = torch.from_numpy(np.ascontiguousarray(g.ravel(order="F")))
    parallel.reduce_grids_to_root([t], root=0)
    if rank == 0:
        full = run(cb.clone(maxpart=n), allp, met, 3).ravel(order="F")
        tot = t.numpy()
        assert abs(tot.sum() - full.sum()) < 1e-5 * full.sum()
        d = np.linalg.norm(tot.astype(np.float64) - full) / np.linalg.norm(full)
        assert d < 1e-6, d
        print("GLOO_OK", d, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
