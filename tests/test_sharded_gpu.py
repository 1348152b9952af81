"""The r-slab sharded build through the C ABI (fb_atmosphere_build_sharded / fb_pending_run_sharded, NCCL).

world = 1 runs on any GPU box; world = 2 needs two GPUs (`gpurun --gpus 2`; skipped otherwise).  The sharded tables must
be BIT-IDENTICAL to the single-GPU build: every stage runs the same kernel on the same inputs, only on fewer altitude
levels per device."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

# 16 altitude levels (8 per rank at world 2), rows wide enough for several CTAs; order 4 = three exchanges of every kind
DIMS = dict(scattering_r_size=16, scattering_mu_size=32, scattering_mu_s_size=16, scattering_nu_size=4, order=4)


def _single(kernels=0):
    import torch
    import fuzzyblue_b200 as fb
    from fuzzyblue_b200 import api
    b = fb.Builder(0, kernels=kernels)
    pend = fb.Atmosphere.build(b, None, fb.Parameters(**DIMS))
    torch.cuda.synchronize()
    return {k: pend.download(i) for k, i in (("T", api.IMAGE_TRANSMITTANCE), ("E", api.IMAGE_IRRADIANCE), ("S", api.IMAGE_SCATTERING))}


def test_world_one_sharded_build_equals_build():
    import torch
    import fuzzyblue_b200 as fb
    from fuzzyblue_b200 import api, sharded
    want = _single()
    b = fb.Builder(0)
    pend = sharded.build_sharded(b, fb.Parameters(**DIMS), None, 0, 1)
    pend.wait()
    for k, i in (("T", api.IMAGE_TRANSMITTANCE), ("E", api.IMAGE_IRRADIANCE), ("S", api.IMAGE_SCATTERING)):
        assert np.array_equal(pend.download(i).view(np.uint8), want[k].view(np.uint8)), k
    pend.run_sharded(None, 0, 1)          # the same schedule again on the same images
    pend.wait()
    assert np.array_equal(pend.download(api.IMAGE_SCATTERING).view(np.uint8), want["S"].view(np.uint8))
    assert pend.slow_stages() == 0
    with pytest.raises(fb.FuzzyblueError):
        pend.run_sharded(None, 0, 2)      # world > 1 without a communicator


def _worker(rank, world, port, out_dir, flags, kernels):
    import torch
    import torch.distributed as dist
    import fuzzyblue_b200 as fb
    from fuzzyblue_b200 import api, sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)     # carries only the 128-byte NCCL id
    torch.cuda.set_device(rank)
    comm = sharded.NcclComm(rank, rank, world)
    b = fb.Builder(rank, kernels=kernels)
    s = torch.cuda.Stream(device=rank)
    pend = sharded.build_sharded(b, fb.Parameters(**DIMS), comm, rank, world, flags, s)
    pend.wait()
    out = {k: pend.download(i) for k, i in (("T", api.IMAGE_TRANSMITTANCE), ("E", api.IMAGE_IRRADIANCE), ("S", api.IMAGE_SCATTERING))}
    torch.cuda.synchronize()
    pend.run_sharded(comm, rank, world, flags, s)                     # replay on the same images
    pend.wait()
    out["S2"] = pend.download(api.IMAGE_SCATTERING)
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **out)
    pend.close()
    comm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("kernels", [0, 1], ids=["fast", "reference"])
@pytest.mark.parametrize("flags", [1 | 4, 1 | 2], ids=["pipelined", "whole-slab"])
def test_two_gpu_sharded_build_is_bit_identical(tmp_path, flags, kernels):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    want = _single(kernels)
    port = 29700 + (os.getpid() % 1000) + 4 * flags + kernels
    mp.spawn(_worker, args=(2, port, str(tmp_path), flags, kernels), nprocs=2, join=True)
    for rank in range(2):
        got = np.load(tmp_path / f"rank{rank}.npz")
        for k in ("T", "E", "S"):
            assert np.array_equal(got[k].view(np.uint8), want[k].view(np.uint8)), (rank, k)
        assert np.array_equal(got["S2"].view(np.uint8), want["S"].view(np.uint8)), rank
