"""tools/convmix_profile.py [NPART] -- fpb_convmix alone on the bench's C2 particles and synthetic soundings (what
bench.py's next_rows.convmix times): for ncu captures of conv_column_kernel."""
import sys, time, types
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import flexpart_b200 as fb, bench, conv_cases
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
args = types.SimpleNamespace(workload="c2", particles=n, sort_interval=1)
cb, rel = bench.build_workload(args, 0, 1, 0)
c = cb.cfg
eng = fb.Engine(cb); eng.fill_rannumb()
eng.upload_met(1, fb.MetFields(cb).synth(0)); eng.upload_met(2, fb.MetFields(cb).synth(10800))
eng.set_met_bracket((1, 2), (0, 10800))
eng.set_releases(rel)
eng.release_particles(0)
for k in range(3):
    eng.step(k * 900, 0)
akm, bkm, akz, bkz, nconvlev = conv_cases.hybrid_levels(c.nz)
f0 = conv_cases.conv_fields(cb, akz, bkz, c.nz, 1)
eng.set_convection(c.nz, c.nzmax, nconvlev, akz[1:], bkz[1:], akm[1:], bkm[1:])
eng.upload_convmet(1, *f0); eng.upload_convmet(2, *f0)
for k in range(3):
    t0 = time.perf_counter()
    ncol, nconv = eng.convmix(2700)
    print(f"fpb_convmix: {1e3 * (time.perf_counter() - t0):.1f} ms, {ncol} occupied columns, {nconv} convecting", flush=True)
