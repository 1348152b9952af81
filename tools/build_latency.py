"""Where does a cold fb_atmosphere_build go?  (allocation vs enqueue vs device time)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fuzzyblue_b200 as fb
b = fb.Builder(0)
p = fb.Parameters()
s = torch.cuda.Stream()
w = fb.Atmosphere.build(b, s, p); s.synchronize(); w.close()
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); a = fb.Atmosphere.allocate(b, p); t1 = time.perf_counter(); a.close(); t2 = time.perf_counter()
    pend = fb.Atmosphere.build(b, s, p); t3 = time.perf_counter(); s.synchronize(); t4 = time.perf_counter()
    atm = pend.assert_ready(check=False); t5 = time.perf_counter(); atm.close(); t6 = time.perf_counter()
    print(f"allocate {1e3*(t1-t0):.2f} ms, free {1e3*(t2-t1):.2f} ms | build call {1e3*(t3-t2):.2f} ms, wait {1e3*(t4-t3):.2f} ms | assert_ready {1e3*(t5-t4):.2f} ms, destroy {1e3*(t6-t5):.2f} ms")
