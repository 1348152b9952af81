"""GPU diagnostic: run REFERENCE and FAST pipelines independently, compare every image after every stage."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fuzzyblue_b200 as fb
from fuzzyblue_b200 import api

p = fb.Parameters()
R = fb.Atmosphere.allocate(fb.Builder(0, kernels=api.KERNELS_REFERENCE), p)
Fp = fb.Atmosphere.allocate(fb.Builder(0, kernels=api.KERNELS_FAST), p)
IM3 = [2, 4, 5, 6, 7]
names = {0: "T", 1: "E", 2: "S", 3: "dE", 4: "dR", 5: "dM", 6: "dens", 7: "dMS"}
probe = (3, 86, 37)

def err(a, b, f16):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 2.0 ** -14 if f16 else 1e-30)

def both(stage, order, outs):
    R.run_stage(stage, order=order); Fp.run_stage(stage, order=order)
    for im in outs:
        a, b = Fp.download(im), R.download(im)
        e = err(a, b, im in IM3)
        w = np.unravel_index(int(e.argmax()), e.shape)
        extra = f" probe fast={a[probe]} ref={b[probe]}" if im in IM3 else ""
        print(f"stage {stage} order {order} {names[im]:5s}: max {e.max():.3e} at {w} fast={a[w]} ref={b[w]} >5e-4: {(e>5e-4).sum()} >1e-3: {(e>1e-3).sum()} >2e-3: {(e>2e-3).sum()}{extra}")

both(0, 0, [0]); both(1, 0, [3]); both(2, 0, [4, 5, 2]); both(6, 0, [1])
for order in (2, 3, 4):
    both(3, order, [6]); both(4, order - 1, [3, 1]); both(5, 0, [7, 2])
