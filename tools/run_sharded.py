"""torchrun driver: r-slab sharded precompute over NCCL vs the single-GPU build (bitwise), plus timing.
   torchrun --nproc-per-node N tools/run_sharded.py [default|dump|hires64]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import fuzzyblue_b200 as fb
from fuzzyblue_b200 import api, sharded

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
case = sys.argv[1] if len(sys.argv) > 1 else "default"
dims = {"default": {}, "dump": dict(scattering_r_size=16, scattering_mu_size=64, scattering_mu_s_size=16, scattering_nu_size=4),
        # a 1/64-volume cut of BASELINE.json config 3 (S 128x512x128x32 -> 64x128x64x16), 8 orders
        "hires64": dict(scattering_r_size=64, scattering_mu_size=128, scattering_mu_s_size=64, scattering_nu_size=16,
                        transmittance_mu_size=1024, transmittance_r_size=256, order=8)}[case]
p = fb.Parameters(**dims)
builder = fb.Builder(local)
stream = torch.cuda.Stream()
pend = fb.Atmosphere.allocate(builder, p)
be = sharded.PendingBackend(pend, stream)
with torch.cuda.stream(stream):
    sp = sharded.ShardedPrecompute(be, p.scattering_r_size, p.order, rank, world)
    sp.run()
    stream.synchronize()
    dist.barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    sp.run()
    t1.record(stream)
    stream.synchronize()
ms = torch.tensor([t0.elapsed_time(t1)], device="cuda")
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
S = pend.download(api.IMAGE_SCATTERING, stream); E = pend.download(api.IMAGE_IRRADIANCE, stream)
stream.synchronize()
# single-GPU build of the same thing on this rank
whole = fb.Atmosphere.build(builder, stream, p)
stream.synchronize()
t0.record(stream); whole.resubmit(stream); t1.record(stream); stream.synchronize()
S1 = whole.download(api.IMAGE_SCATTERING, stream); E1 = whole.download(api.IMAGE_IRRADIANCE, stream)
stream.synchronize()
same = bool(np.array_equal(S, S1) and np.array_equal(E, E1))
flags = torch.tensor([int(same)], device="cuda"); dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"case={case} world={world} sharded_ms={ms.item():.3f} single_gpu_ms={t0.elapsed_time(t1):.3f} gathers={sp.gathers} "
          f"bitwise_equal_on_all_ranks={bool(flags.item())}", flush=True)
dist.destroy_process_group()
