"""One rank of the multi-GPU grid-exchange test (tests/test_gpu_multirank.py): a share of the
particles on GPU `rank`, steps + conccalc, the sum of the grids to rank 0 through
fpb_comm_init / fpb_reduce_grids_begin / _end.  The NCCL id travels through a file -- the C ABI does
not depend on torch.distributed.  argv: rank world tmpdir mode"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def problem(mode, rank, world, device):
    import flexpart_b200 as fb
    import cases
    n = 6000
    kw = dict(nrel=4, npart_each=n // 4, rng_mode=fb.RNG_PHILOX_INDEX, sort_interval=1, device=device,
              part_id_stride=world, part_id_offset=rank, lage=(86400 * 20,), nspec=2, drydepspec=(1, 0),
              receptors=[(36.0, 18.0, 1.0e9)], nest=(-60.0, -30.0, 48, 24, 2.5, 2.5))
    if mode == "exact":
        kw.update(scatter_mode=fb.SCATTER_DETERMINISTIC, drydepspec=(0, 0))   # (deposition would change the masses)
    cb = cases.config_small(**kw)
    p = cases.seeded_particles(cb, n, zmax=2500.0, lat_range=(-50.0, 50.0), nspec=2)
    if mode == "exact":
        # unit masses and young particles (no kernel weights): every grid contribution is 1.0 and
        # float addition is exact, so the N-rank sum must equal the one-rank grid bit for bit
        p.xmass1[:n] = 1.0
    else:
        p.itramem[:n // 2] = -20000
        p.xmass1[:n, 1] = 0.37
    mine = fb.Particles(cb.cfg.maxpart, 2)
    idx = np.arange(rank, n, world)
    for f in ("xtra1", "ytra1", "ztra1", "itra1", "itramem", "npoint", "nclass", "idt"):
        getattr(mine, f)[:len(idx)] = getattr(p, f)[idx]
    mine.xmass1[:len(idx)] = p.xmass1[idx]
    mine.numpart = len(idx)
    return cb, mine, idx


def run(rank, world, tmp, mode):
    import flexpart_b200 as fb
    import cases
    cb, mine, idx = problem(mode, rank, world, rank if world > 1 else 0)
    eng = fb.Engine(cb)
    eng.fill_rannumb()
    m0, m1 = cases.met_pair(cb)
    eng.upload_met(1, m0); eng.upload_met(2, m1); eng.set_met_bracket((1, 2), (0, 10800))
    idfile = os.path.join(tmp, f"nccl_id_{mode}")
    if world > 1:
        if rank == 0:
            uid = fb.Engine.comm_unique_id()
            with open(idfile + ".tmp", "wb") as f:
                f.write(uid)
            os.rename(idfile + ".tmp", idfile)
        else:
            t0 = time.time()
            while not os.path.exists(idfile):
                if time.time() - t0 > 120:
                    raise SystemExit("timed out waiting for the NCCL id")
                time.sleep(0.05)
            uid = open(idfile, "rb").read()
    else:
        uid = bytes(128)
    eng.comm_init(uid, rank, world)
    eng.push_particles(mine)
    outs = []
    for k in range(4):
        eng.conccalc(k * 900, 1.0)
        eng.step(k * 900, 450)
        if k % 2 == 1:                      # two output intervals
            eng.reduce_grids_begin()
            if k == 1:
                eng.conccalc((k + 1) * 900, 0.0)   # the engine works on while the exchange is in flight
            outs.append(eng.reduce_grids_end())
    eng.pull_particles(mine)
    np.savez(os.path.join(tmp, f"out_{mode}_{world}_{rank}.npz"), idx=idx,
             **{f: getattr(mine, f)[:len(idx)] for f in ("xtra1", "ytra1", "ztra1", "itra1", "uap", "us")},
             xmass1=mine.xmass1[:len(idx)],
             **{f"{name}_{i}": g for i, o in enumerate(outs) for name, g in o.items()})
    eng.close()


if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4])
