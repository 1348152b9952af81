"""End-to-end time of one default-dims precompute with the read-back recorded into the command stream
(fb_pending_set_readback), against replay + three separate read calls.  FUZZYBLUE_B200_RB_SLABS selects the slab count."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fuzzyblue_b200 as fb
from fuzzyblue_b200 import api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
b = fb.Builder(0)
s = torch.cuda.Stream()
P = fb.Parameters()
p = fb.Atmosphere.build(b, s, P)
s.synchronize()
atm = p.atmosphere()
hT = torch.empty((P.transmittance_r_size, P.transmittance_mu_size, 4), dtype=torch.float32).pin_memory()
hE = torch.empty((P.irradiance_r_size, P.irradiance_mu_s_size, 4), dtype=torch.float32).pin_memory()
hS = torch.empty((P.scattering_r_size, P.scattering_mu_size, P.scattering_nu_size * P.scattering_mu_s_size, 4), dtype=torch.float16).pin_memory()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
L, vp = api._lib(), api.c_void_p

def separate():
    p.resubmit(s)
    api._check(L.fb_atmosphere_read_transmittance(atm._h, vp(hT.data_ptr()), hT.numel() * 4, api._stream(s)))
    api._check(L.fb_atmosphere_read_scattering(atm._h, vp(hS.data_ptr()), hS.numel() * 2, api._stream(s)))
    api._check(L.fb_atmosphere_read_irradiance(atm._h, vp(hE.data_ptr()), hE.numel() * 4, api._stream(s)))
    s.synchronize()

def recorded():
    p.resubmit(s)
    s.synchronize()

def device_only():
    p.resubmit(s)
    s.synchronize()

def run(fn):
    ts = []
    for i in range(n + 3):
        with torch.cuda.stream(s):
            flush.fill_(i & 255)
        s.synchronize()
        t0 = time.perf_counter(); fn(); t = time.perf_counter() - t0
        if i >= 3: ts.append(t * 1e3)
    return float(np.median(ts)), float(np.min(ts))

print("device only (wall)   median %.3f ms  min %.3f" % run(device_only))
print("separate read-backs  median %.3f ms  min %.3f" % run(separate))
ref = hS.clone()
p.set_readback(hT.data_ptr(), hS.data_ptr(), hE.data_ptr())
hS.zero_()
print("recorded read-backs  median %.3f ms  min %.3f  (slabs %s)" % (run(recorded) + (os.environ.get("FUZZYBLUE_B200_RB_SLABS", "4"),)))
print("same bytes:", bool(torch.equal(ref.view(torch.int16), hS.view(torch.int16))), " launches/step:", p.launch_count())
