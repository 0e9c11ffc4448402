"""Generate tests/golden/*.npz from the CPU oracle.  The reference itself cannot
be run in this environment (no Fortran compiler), so these fixtures are
regression pins of the oracle, committed so that an accidental change of the
restatement is caught.  Usage: python tests/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import __graft_entry__ as ge  # noqa: E402

ge.build(verbose=False)
from test_oracle_pins import _golden_case  # noqa: E402

res, outs, p = _golden_case()
n = res.numpart_final
os.makedirs(os.path.join(HERE, "golden"), exist_ok=True)
np.savez_compressed(os.path.join(HERE, "golden", "oracle_small_hanna.npz"),
                    numpart=n, particle_steps=res.particle_steps, substeps=res.substeps,
                    itra1=p.itra1[:n], idt=p.idt[:n], xtra1=p.xtra1[:n], ytra1=p.ytra1[:n],
                    ztra1=p.ztra1[:n], gridunc_last=outs[-1]["gridunc"])
print("wrote golden fixture:", n, "particles,", res.particle_steps, "particle-steps")
