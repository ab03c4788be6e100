import time, numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from top_down_renderer_b200 import synth
from top_down_renderer_b200.core import Context
cm = synth.make_class_map(4000, 4000, 6, seed=1234)
img, lut = synth.to_cv_image(cm), synth.identity_lut(6)
c = Context(0)
for i in range(4):
    t = time.perf_counter(); c.map_set_class_image(img, lut, 6, 1.0); dt = time.perf_counter() - t
    print(f"map_set_class_image 4000x4000x6 (H2D 16 MB + seeds + EDT): {dt*1e3:.2f} ms")
