"""A/B of two builds of the sky evaluation: renders a camera sweep (small frames on reduced tables, 4K frames on the
default tables) with the library FUZZYBLUE_B200_LIB selects and writes one SHA-256 per output to argv[1]; run it once
per build and diff the files.  Also prints the device time of the 4K frames."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fuzzyblue_b200 as fb
from fuzzyblue_b200 import synthetic
out = open(sys.argv[1], "w")
b = fb.Builder(0)

def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]

for dims, (W, H), views, alts in ((dict(scattering_r_size=16, scattering_mu_size=64, scattering_mu_s_size=16, scattering_nu_size=4), (512, 288), 24, None),
                                  (dict(), (3840, 2160), 8, None),
                                  (dict(), (640, 360), 9, (59.0, 59.67, 59.69, 59.99, 60.0, 60.001, 61.0, 1e-4, 3.0))):
    pend = fb.Atmosphere.build(b, None, fb.Parameters(**dims)); torch.cuda.synchronize()
    atm = pend.atmosphere()
    r = fb.Renderer(b)
    draws, extra = synthetic.camera_sweep(views, W, H, seed=4 if alts is None else 12, altitudes_km=alts)
    for k in range(views):
        depth = synthetic.analytic_depth(extra[k][0], extra[k][1], W, H)
        c, t = r.draw_host(atm, draws[k], depth)
        print(W, H, k, digest(c), digest(t), "nan_c=%d nan_t=%d" % (np.isnan(c).sum(), np.isnan(t).sum()), file=out)
    if W == 3840:
        dd = torch.from_numpy(np.stack([synthetic.analytic_depth(extra[k][0], extra[k][1], W, H) for k in range(views)])).cuda()
        color = torch.empty((views, H, W, 4), device="cuda"); transm = torch.empty_like(color)
        s = torch.cuda.Stream()
        ts = []
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); r.draw_sweep(s, atm, draws, dd, color, transm, W, H); e1.record(s); s.synchronize()
            ts.append(e0.elapsed_time(e1) / views)
        print("4K ms/frame over %d views: %.4f (min of 4)" % (views, min(ts)))
out.close()
