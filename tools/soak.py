"""tools/soak.py WORKLOAD STEPS -- long run of a bench workload on one GPU: every step's statistics
are checked (no particle outside the domain, non-finite count, mass bookkeeping)."""
import sys, types, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import flexpart_b200 as fb, bench
wl, nsteps = sys.argv[1], int(sys.argv[2])
args = types.SimpleNamespace(workload=wl, particles=int(sys.argv[3]) if len(sys.argv) > 3 else 1000000, sort_interval=1)
cb, rel = bench.build_workload(args, 0, 1, 0)
c = cb.cfg
span = (nsteps + 2) * 900
eng = fb.Engine(cb); eng.fill_rannumb()
eng.upload_met(1, fb.MetFields(cb).synth(0)); eng.upload_met(2, fb.MetFields(cb).synth(span))
eng.set_met_bracket((1, 2), (0, span))
eng.set_releases(rel)
n, made = eng.release_particles(0)
q = fb.Particles(c.maxpart, c.nspec); q.numpart = n
tot = dict(n_terminated=0, n_nonfinite=0, n_nan_cbl=0, n_substeps=0)
for k in range(nsteps):
    itime = k * 900
    if c.wetdep and k:
        eng.wetdepo(itime, 900, 450)
    eng.conccalc(itime, 1.0)
    st = eng.step(itime, 0)
    for key in tot: tot[key] += st[key]
    if (k + 1) % 4 == 0: eng.zero_conc_grids()
    if (k + 1) % 25 == 0 or k == nsteps - 1:
        eng.pull_particles(q)
        live = q.itra1[:n] != fb.ITRA_DEAD
        x, y, z = q.xtra1[:n][live], q.ytra1[:n][live], q.ztra1[:n][live]
        ok = np.isfinite(x).all() and np.isfinite(y).all() and np.isfinite(z).all() and x.min() >= 0 and x.max() <= c.nx - 1 \
            and y.min() >= 0 and y.max() <= c.ny - 1 and z.min() >= 0 and z.max() <= cb.height[c.nz - 1] and np.isfinite(q.xmass1[:n]).all()
        print(f"step {k + 1}: live {int(live.sum())}, {tot}, z max {z.max():.0f}, mass {q.xmass1[:n][live].sum(0)}, ok={ok}", flush=True)
        assert ok
print("soak ok")
