import sys, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import flexpart_b200 as fb, cases, oracle_api
def run(tag, **over):
    kw = dict(nx=73, ny=37, nz=138, dx=5., dy=5., xlon0=-180.0, ylat0=-90.0, lsynctime=900, ifine=4, ctl=10.0, cblflag=1,
              outlon0=-180.0, outlat0=-90.0, numxgrid=72, numygrid=36, dxout=5., dyout=5.,
              outheights=(100.0, 250.0, 500.0, 1000.0, 2000.0, 3000.0, 5000.0, 8000.0, 12000.0, 50000.0),
              lage=(86400 * 20,), ioutputforeachrelease=0, npart=(20,) * 100, maxpart=2000,
              rng_mode=fb.RNG_PHILOX_INDEX)
    kw.update(over)
    cb = fb.make_config(**kw)
    rel = cases.releases_boxes(cb, seed=100, zmax=2000.0, lat_range=(-60.0, 60.0), width=10.0)
    out = []
    for E in (fb.Engine, oracle_api.Oracle):
        eng = E(cb); eng.fill_rannumb()
        eng.upload_met(1, fb.MetFields(cb).synth(0)); eng.upload_met(2, fb.MetFields(cb).synth(10800))
        eng.set_met_bracket((1, 2), (0, 10800))
        p = fb.Particles(cb.cfg.maxpart, 1); st = fb.ReleaseState(cb.cfg.numpoint)
        fb.release_particles(cb, rel, st, 0, p)
        eng.push_particles(p)
        s = eng.step(0)
        q = fb.Particles(cb.cfg.maxpart, 1); q.numpart = p.numpart; eng.pull_particles(q)
        out.append((s, q))
    (sg, g), (so, o) = out
    n = g.numpart
    nan = np.isnan(g.xtra1[:n]) | np.isnan(g.ztra1[:n])
    dz = np.abs(g.ztra1[:n] - o.ztra1[:n])
    print(tag, 'ifine', cb.cfg.ifine, 'ctl', cb.cfg.ctl, 'gpu nsub', sg['n_substeps'], 'orc nsub', so['n_substeps'], 'gpu nan', nan.sum(),
          'orc nan', np.isnan(o.ztra1[:n]).sum(), 'max dz(non-nan)', np.nanmax(dz), flush=True)
run('base')
run('ctl5', ctl=5.0)
run('table', rng_mode=fb.RNG_REFERENCE)
run('strict', math_mode=fb.MATH_STRICT)
run('strict_table', math_mode=fb.MATH_STRICT, rng_mode=fb.RNG_REFERENCE)
run('nz40', nz=40, height=fb.synth_heights(138)[::3][:40])
