import sys, types
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import flexpart_b200 as fb, bench
args = types.SimpleNamespace(workload='c2', particles=int(sys.argv[1]) if len(sys.argv) > 1 else 1000000, sort_interval=1)
cb, rel = bench.build_workload(args, 0, 1, 0)
eng = fb.Engine(cb); eng.fill_rannumb()
eng.upload_met(1, fb.MetFields(cb).synth(0)); eng.upload_met(2, fb.MetFields(cb).synth(90000))
eng.set_met_bracket((1, 2), (0, 90000))
p = bench.host_particles(cb, rel, pinned=False)
eng.push_particles(p)
for k in range(6):
    eng.conccalc(k * 900, 1.0)
    st = eng.step(k * 900, 0)
    pre, post, lanes_post, nsub = st['n_nan_cbl'], st['n_nonfinite'], st['n_terminated'], st['n_substeps']
    print(f"step {k}: warp iterations {pre} before + {post} after the counter ran out ({100 * post / (pre + post):.1f} % in the tail); "
          f"lanes active: overall {nsub / (pre + post):.1f}, tail {lanes_post / max(post, 1):.1f}, body {(nsub - lanes_post) / max(pre, 1):.1f}")
