"""world_size-2 gloo test of the r-slab sharding host logic (fuzzyblue_b200/sharded.py) on CPU.

The stages run through an oracle-backed stand-in for the CUDA backend (the oracle is test infrastructure; the
product backend is PendingBackend).  The sharded schedule + all-gathers must reproduce the single-process tables
EXACTLY, which proves every cross-slab dependency is covered by an exchange."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fuzzyblue_b200 import api, sharded   # noqa: E402
from oracle import oracle as O            # noqa: E402

DIMS = dict(scattering_r_size=4, scattering_mu_size=8, scattering_mu_s_size=4, scattering_nu_size=2,
            transmittance_mu_size=32, transmittance_r_size=8, irradiance_mu_s_size=8, irradiance_r_size=4)


class OracleBackend:
    """Same interface as sharded.PendingBackend, stages evaluated by the CPU oracle on this rank's slab only."""

    def __init__(self, p: O.Params):
        self.p = p
        shapes = {api.IMAGE_TRANSMITTANCE: p.t_shape, api.IMAGE_IRRADIANCE: p.e_shape, api.IMAGE_DELTA_IRRADIANCE: p.e_shape}
        # slabs this rank never computes stay NaN until an all-gather fills them: a missing exchange poisons the result
        self.img = {i: torch.full(shapes.get(i, p.s_shape), float("nan"), dtype=torch.float64) for i in range(8)}

    def tensor(self, image):
        return self.img[image]

    def _np(self, image):
        return self.img[image].numpy()

    def _slab_idx(self, r0, r1):
        R, M, W, _ = self.p.s_shape
        r1 = r1 or R
        return np.arange(r0 * M * W, r1 * M * W, dtype=np.int64), r0, r1

    def run_stage(self, stage, order=0, r_begin=0, r_end=0):
        p, m = self.p, O.F32
        g = self._np
        idx, r0, r1 = self._slab_idx(r_begin, r_end)
        shp = (r1 - r0,) + tuple(p.s_shape[1:])
        if stage == api.STAGE_TRANSMITTANCE:
            g(api.IMAGE_TRANSMITTANCE)[:] = O.transmittance(p, m)
        elif stage == api.STAGE_DIRECT_IRRADIANCE:
            g(api.IMAGE_DELTA_IRRADIANCE)[:] = O.direct_irradiance(p, m, g(api.IMAGE_TRANSMITTANCE))
        elif stage == api.STAGE_CLEAR_IRRADIANCE:
            g(api.IMAGE_IRRADIANCE)[:] = 0
        elif stage == api.STAGE_SINGLE_SCATTERING:
            dR, dM, S = O.single_scattering(p, m, g(api.IMAGE_TRANSMITTANCE), idx)
            g(api.IMAGE_DELTA_RAYLEIGH)[r0:r1] = dR.reshape(shp)
            g(api.IMAGE_DELTA_MIE)[r0:r1] = dM.reshape(shp)
            g(api.IMAGE_SCATTERING)[r0:r1] = S.reshape(shp)
        elif stage == api.STAGE_SCATTERING_DENSITY:
            # the oracle would happily read NaN slabs of a table an order does not use; hand it zeros there
            dms = g(api.IMAGE_DELTA_MULTIPLE_SCATTERING) if order > 2 else np.zeros(p.s_shape)
            out = O.scattering_density(p, m, order, g(api.IMAGE_TRANSMITTANCE), g(api.IMAGE_DELTA_RAYLEIGH), g(api.IMAGE_DELTA_MIE),
                                       dms, g(api.IMAGE_DELTA_IRRADIANCE), idx)
            g(api.IMAGE_SCATTERING_DENSITY)[r0:r1] = out.reshape(shp)
        elif stage == api.STAGE_INDIRECT_IRRADIANCE:
            dms = g(api.IMAGE_DELTA_MULTIPLE_SCATTERING) if order > 1 else np.zeros(p.s_shape)
            dE, E = O.indirect_irradiance(p, m, order, g(api.IMAGE_DELTA_RAYLEIGH), g(api.IMAGE_DELTA_MIE), dms, g(api.IMAGE_IRRADIANCE))
            g(api.IMAGE_DELTA_IRRADIANCE)[:] = dE
            g(api.IMAGE_IRRADIANCE)[:] = E
        elif stage == api.STAGE_MULTIPLE_SCATTERING:
            S_in = np.nan_to_num(g(api.IMAGE_SCATTERING), nan=0.0)
            dMS, S = O.multiple_scattering(p, m, g(api.IMAGE_TRANSMITTANCE), g(api.IMAGE_SCATTERING_DENSITY), S_in, idx)
            g(api.IMAGE_DELTA_MULTIPLE_SCATTERING)[r0:r1] = dMS.reshape(shp)
            g(api.IMAGE_SCATTERING)[r0:r1] = S.reshape(shp)
        else:
            raise ValueError(stage)


def _worker(rank, world, port, order, out_dir, min_chunk_bytes):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = O.Params(order=order, **DIMS)
    be = OracleBackend(p)
    sp = sharded.ShardedPrecompute(be, p.scattering_r_size, order, rank, world, min_chunk_bytes=min_chunk_bytes).run()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), S=be._np(api.IMAGE_SCATTERING), E=be._np(api.IMAGE_IRRADIANCE),
             T=be._np(api.IMAGE_TRANSMITTANCE), dMS=be._np(api.IMAGE_DELTA_MULTIPLE_SCATTERING), gathers=sp.gathers)
    dist.destroy_process_group()


def test_slab_partition():
    assert sharded.slab_of(0, 8, 32) == (0, 4) and sharded.slab_of(7, 8, 32) == (28, 32)
    assert sharded.slab_of(3, 4, 128) == (96, 128)
    with pytest.raises(ValueError):
        sharded.slab_of(0, 3, 32)


@pytest.mark.parametrize("order,min_chunk_bytes", [(3, 16 << 20), (3, 0)])
def test_two_rank_sharded_precompute_equals_single_process(tmp_path, order, min_chunk_bytes):
    """min_chunk_bytes = 0 forces the sub-slab pipeline (2 chunks per 2-row slab) that large tables use."""
    world, port = 2, 29500 + (os.getpid() % 2000) + (1 if min_chunk_bytes else 0)
    mp.spawn(_worker, args=(world, port, order, str(tmp_path), min_chunk_bytes), nprocs=world, join=True)
    chunks = 1 if min_chunk_bytes else 2
    ref = O.precompute(O.Params(order=order, **DIMS), O.F32)
    for rank in range(world):
        got = np.load(tmp_path / f"rank{rank}.npz")
        assert np.array_equal(got["T"], ref.transmittance)
        assert np.array_equal(got["E"], ref.irradiance)
        assert np.array_equal(got["S"], ref.scattering)           # every rank holds the full, identical final table
        # 2 single-scattering gathers + per order: density (+ delta_multiple except after the last) + the final table
        # (tables this small are exchanged in one piece)
        assert int(got["gathers"]) == 2 + chunks * ((order - 1) + (order - 2)) + 1
    # each rank's own slab of the last delta_multiple_scattering is current; the peer's slab is still the previous
    # order's (the last order's temporaries are not exchanged: nothing reads them)
    r0 = np.load(tmp_path / "rank0.npz")["dMS"]
    prev = O.precompute(O.Params(order=order - 1, **DIMS), O.F32)
    assert np.array_equal(r0[:2], ref.delta_multiple_scattering[:2])
    assert np.array_equal(r0[2:], prev.delta_multiple_scattering[2:])
