"""normalise + resample on N weights, alone (for an ncu launch list): python tools/norm_profile.py N"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from top_down_renderer_b200 import synth
from top_down_renderer_b200.core import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(1)
w = (1.0 / (0.7 + rng.random(n) * 0.5)).astype(np.float32)
ld = rng.uniform(0, 0.4, n).astype(np.float32)
st = np.zeros(n, dtype=synth.STATE_DTYPE); st["scale"] = 2; st["have_init"] = 1
c = Context(0)
c.pf_set_states(st, ld)
for rep in range(3):
    c.pf_set_weights(w)
    c.pf_normalize_resample(0.37, n)
    c.sync()
c.close()
print("done")
