"""The multi-GPU path on hardware (SURVEY.md 8c pin 5, 8e): N ranks, each with every N-th particle on
its own GPU, the grids summed to rank 0 through the C ABI's NCCL exchange, against ONE rank holding
all particles.  Needs >= 2 GPUs (run with `gpurun --gpus 2`); the single-GPU part checks that the
one-rank exchange is exactly fpb_fetch_grids."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import flexpart_b200 as fb
import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "_nccl_worker.py")


def _ngpus():
    import torch
    return torch.cuda.device_count()


def _launch(world, tmp, mode):
    procs = [subprocess.Popen([sys.executable, WORKER, str(r), str(world), tmp, mode]) for r in range(world)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    return [np.load(os.path.join(tmp, f"out_{mode}_{world}_{r}.npz")) for r in range(world)]


@pytest.mark.parametrize("mode", ["exact", "general"])
def test_n_rank_exchange_equals_one_rank(mode):
    world = min(_ngpus(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    with tempfile.TemporaryDirectory() as tmp:
        one = _launch(1, tmp, mode)[0]
        many = _launch(world, tmp, mode)
    # trajectories do not depend on the partition (Philox streams keyed by the global particle id)
    for r, o in enumerate(many):
        for f in ("xtra1", "ytra1", "ztra1", "itra1", "uap", "us", "xmass1"):
            assert np.array_equal(o[f], one[f][o["idx"]]), (r, f)
    # the sums on rank 0
    names = [k for k in one.files if k.rsplit("_", 1)[-1] in ("0", "1") and k.split("_")[0] in
             ("gridunc", "griduncn", "drygridunc", "drygriduncn", "creceptor")]
    assert len(names) >= 8
    for k in names:
        a, b = many[0][k], one[k]
        assert np.abs(b).sum() > 0 or not k.startswith("grid" if mode == "exact" else ("grid", "dry")), k
        if mode == "exact" and k.startswith("grid"):
            assert np.array_equal(a, b), k          # exact float sums: bit-identical for any N
        else:
            d = np.linalg.norm((a.astype(np.float64) - b).ravel()) / max(np.linalg.norm(b.astype(np.float64).ravel()), 1e-300)
            assert d < 1e-6, (k, d)
    # ranks > 0 receive nothing
    assert all(np.abs(many[r]["gridunc_0"]).sum() == 0 for r in range(1, world))


def test_one_rank_exchange_is_fetch_grids():
    """nranks = 1: begin/end = staging copy + zero, no NCCL; equals fpb_fetch_grids(zero_conc=1)."""
    res = []
    for use_comm in (False, True):
        cb = cases.config_small(nrel=4, npart_each=1024, rng_mode=fb.RNG_PHILOX_INDEX, nspec=2, drydepspec=(1, 0),
                                receptors=[(36.0, 18.0, 1.0e9)], nest=(-60.0, -30.0, 48, 24, 2.5, 2.5),
                                scatter_mode=fb.SCATTER_DETERMINISTIC)
        eng = fb.Engine(cb)
        eng.fill_rannumb()
        m0, m1 = cases.met_pair(cb)
        eng.upload_met(1, m0); eng.upload_met(2, m1); eng.set_met_bracket((1, 2), (0, 10800))
        if use_comm:
            eng.comm_init(bytes(128), 0, 1)
        p = cases.seeded_particles(cb, 4096, zmax=2500.0, nspec=2)
        p.itramem[:2048] = -20000
        eng.push_particles(p)
        outs = []
        for k in range(4):
            eng.conccalc(k * 900, 1.0)
            eng.step(k * 900, 450)
            if k % 2 == 1:
                if use_comm:
                    eng.reduce_grids_begin()
                    outs.append(eng.reduce_grids_end())
                else:
                    outs.append(eng.fetch_grids())
        res.append(outs)
        eng.close()
    for a, b in zip(*res):
        for k in ("gridunc", "griduncn", "drygridunc", "drygriduncn", "creceptor"):
            assert np.abs(a[k]).sum() > 0 or k == "creceptor", k
            if k.startswith("grid"):    # the deterministic scatter is bit-reproducible from run to run
                assert np.array_equal(a[k], b[k]), k
            else:                       # deposition / receptor sums are float atomics
                np.testing.assert_allclose(a[k], b[k], rtol=1e-5, atol=1e-30)
