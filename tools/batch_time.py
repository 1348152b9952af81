"""Times a batch of independent atmospheres (config 4 style) on one GPU: ms per atmosphere vs batch size."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fuzzyblue_b200 as fb
from fuzzyblue_b200 import synthetic
b = fb.Builder(0)
s = torch.cuda.Stream()
for n in (1, 8, 32):
    params = synthetic.random_atmospheres(n, seed=20260)
    pend = fb.build_batch(b, params, s); s.synchronize()
    for p in pend: p.close()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(s)
    pend = fb.build_batch(b, params, s)   # includes cudaMalloc of 8 images per atmosphere on the host side
    t1.record(s); s.synchronize()
    print(f"batch {n:3d}: {t0.elapsed_time(t1)/n:.3f} ms per atmosphere (device time incl. allocation stalls)")
    # replay only: resubmit every atmosphere's graph on its own stream
    streams = [torch.cuda.Stream() for _ in range(min(n, 8))]
    for i, p in enumerate(pend): p.resubmit(streams[i % len(streams)])
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    for rep in range(3):
        for i, p in enumerate(pend): p.resubmit(streams[i % len(streams)])
    torch.cuda.synchronize()
    print(f"batch {n:3d}: {(time.perf_counter()-w0)*1e3/(3*n):.3f} ms per atmosphere (graph replays on {len(streams)} streams, wall)")
    for p in pend: p.close()
