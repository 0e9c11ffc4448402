#!/usr/bin/env python
"""tools/metproc_profile.py -- three fpb_calcpar_verttransform calls on a 0.5 deg x 138 level field (run
under ncu by tools/profile_round.sh for the launch list of the met_* kernels)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import flexpart_b200 as fb  # noqa: E402
import conv_cases  # noqa: E402
import met_cases  # noqa: E402

nuvz = 138
kw = dict(nx=721, ny=361, nz=nuvz, dx=0.5, dy=0.5, xlon0=-180.0, ylat0=-90.0, numxgrid=720, numygrid=360, dxout=0.5,
          dyout=0.5, outlon0=-180.0, outlat0=-90.0, npart=(8,), maxpart=64)
cb0 = fb.make_config(**kw, height=fb.synth_heights(nuvz))
akm, bkm, akz, bkz, _ = conv_cases.hybrid_levels(nuvz)
raw = met_cases.raw_fields(cb0, akz, bkz, nuvz, seed=5)
hh, _ = fb.verttransform_heights(cb0, nuvz, akz[1:], bkz[1:], raw)
eng = fb.Engine(fb.make_config(**kw, height=hh))
eng.set_vertical(nuvz, akm[1:], bkm[1:], akz[1:], bkz[1:])
for _ in range(3):
    ms = eng.calcpar_verttransform(1, raw)
    print(f"fpb_calcpar_verttransform: {ms:.1f} ms (kernels {eng.metproc_kernel_ms:.2f} ms)")
eng.close()
