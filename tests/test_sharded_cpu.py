"""world_size-2 gloo test of the r-slab sharded schedule on CPU.

The schedule is `fb_sharded_plan` of the C library (fuzzyblue_b200/csrc/fb_sharded.cu: pure host logic, runs without a
GPU); `sharded.GlooExecutor` executes its steps with gloo collectives, the stages themselves through an oracle-backed
stand-in for the CUDA kernels (the oracle is test infrastructure; production runs the same step list inside the library
with NCCL).  Every slab a rank never computes starts as NaN, so a dependency the plan does not cover with an exchange
poisons the result: the sharded tables must equal the single-process tables EXACTLY."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import fuzzyblue_b200 as fb               # noqa: E402
from fuzzyblue_b200 import api, sharded   # noqa: E402
from oracle import oracle as O            # noqa: E402

DIMS = dict(scattering_r_size=4, scattering_mu_size=8, scattering_mu_s_size=4, scattering_nu_size=2,
            transmittance_mu_size=32, transmittance_r_size=8, irradiance_mu_s_size=8, irradiance_r_size=4)


class OracleBackend:
    """What sharded.GlooExecutor drives: run_stage(stage, order, begin, end) evaluated by the CPU oracle, tensor(image) = the whole image."""

    def __init__(self, p: O.Params):
        self.p = p
        shapes = {api.IMAGE_TRANSMITTANCE: p.t_shape, api.IMAGE_IRRADIANCE: p.e_shape, api.IMAGE_DELTA_IRRADIANCE: p.e_shape}
        # slabs this rank never computes stay NaN until an all-gather fills them: a missing exchange poisons the result
        self.img = {i: torch.full(shapes.get(i, p.s_shape), float("nan"), dtype=torch.float64) for i in range(8)}
        self.rows_computed = 0

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
            # rows [r_begin, r_end) of the irradiance table only; the oracle evaluates every row (rows of other ranks read
            # NaN slabs and come out NaN) and the rows this rank owns are kept
            e0, e1 = r_begin, (r_end or p.irradiance_r_size)
            dms = g(api.IMAGE_DELTA_MULTIPLE_SCATTERING) if order > 1 else np.zeros(p.s_shape)
            with np.errstate(all="ignore"):
                dE, E = O.indirect_irradiance(p, m, order, g(api.IMAGE_DELTA_RAYLEIGH), g(api.IMAGE_DELTA_MIE), dms, g(api.IMAGE_IRRADIANCE))
            g(api.IMAGE_DELTA_IRRADIANCE)[e0:e1] = dE[e0:e1]
            g(api.IMAGE_IRRADIANCE)[e0:e1] = E[e0:e1]
            self.rows_computed += e1 - e0
        elif stage == api.STAGE_MULTIPLE_SCATTERING:
            S_in = np.nan_to_num(g(api.IMAGE_SCATTERING), nan=0.0)
            dMS, S = O.multiple_scattering(p, m, g(api.IMAGE_TRANSMITTANCE), g(api.IMAGE_SCATTERING_DENSITY), S_in, idx)
            g(api.IMAGE_DELTA_MULTIPLE_SCATTERING)[r0:r1] = dMS.reshape(shp)
            g(api.IMAGE_SCATTERING)[r0:r1] = S.reshape(shp)
        else:
            raise ValueError(stage)


def _worker(rank, world, port, order, out_dir, flags):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = O.Params(order=order, **DIMS)
    be = OracleBackend(p)
    ex = sharded.GlooExecutor(be, fb.Parameters(order=order, **DIMS), rank, world).run(flags)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), S=be._np(api.IMAGE_SCATTERING), E=be._np(api.IMAGE_IRRADIANCE),
             T=be._np(api.IMAGE_TRANSMITTANCE), dMS=be._np(api.IMAGE_DELTA_MULTIPLE_SCATTERING), exchanges=ex.exchanges,
             rows=be.rows_computed)
    dist.destroy_process_group()


def test_slab_partition():
    assert sharded.slab_of(0, 8, 32) == (0, 4) and sharded.slab_of(7, 8, 32) == (28, 32)
    assert sharded.slab_of(3, 4, 128) == (96, 128)
    with pytest.raises(ValueError):
        sharded.slab_of(0, 3, 32)


def _steps(p, rank, world, flags=sharded.GATHER_RESULT):
    return [(s.op, s.stage, s.image, s.order, s.begin, s.end, s.root) for s in sharded.plan(p, rank, world, flags)]


def test_plan_single_rank_is_the_reference_schedule():
    """world = 1: no exchanges, and the stages are the recorded command stream of src/precompute.rs:1671-2048."""
    p = fb.Parameters(order=4)
    st = _steps(p, 0, 1)
    assert all(op == sharded.SHARD_STAGE for op, *_ in st)
    want = [api.STAGE_TRANSMITTANCE, api.STAGE_DIRECT_IRRADIANCE, api.STAGE_SINGLE_SCATTERING, api.STAGE_CLEAR_IRRADIANCE]
    for _ in (2, 3, 4):
        want += [api.STAGE_SCATTERING_DENSITY, api.STAGE_INDIRECT_IRRADIANCE, api.STAGE_MULTIPLE_SCATTERING]
    assert [s[1] for s in st] == want
    assert [s[3] for s in st if s[1] == api.STAGE_SCATTERING_DENSITY] == [2, 3, 4]         # push constant `order`
    assert [s[3] for s in st if s[1] == api.STAGE_INDIRECT_IRRADIANCE] == [1, 2, 3]        # push constant `order - 1`


@pytest.mark.parametrize("world", [2, 4, 8])
def test_plan_properties(world):
    """Default and high-resolution dims: the slabs tile the r axis, the irradiance rows are partitioned, every rank issues
    the same exchanges in the same order (a collective the ranks disagree on would deadlock), and the bytes a rank
    receives are the one mandatory all-gather per order plus halos."""
    hires = dict(transmittance_mu_size=1024, transmittance_r_size=256, scattering_r_size=128, scattering_mu_size=512,
                 scattering_mu_s_size=128, scattering_nu_size=32)
    for p in (fb.Parameters(order=4), fb.Parameters(order=8, **hires)):
        plans = [_steps(p, r, world) for r in range(world)]
        R = p.scattering_r_size
        slabs = sorted((s[4], s[5]) for pl in plans for s in pl if s[0] == sharded.SHARD_STAGE and s[1] == api.STAGE_SINGLE_SCATTERING)
        assert slabs == [(r * R // world, (r + 1) * R // world) for r in range(world)]
        rows = sorted((s[4], s[5]) for pl in plans for s in pl
                      if s[0] == sharded.SHARD_STAGE and s[1] == api.STAGE_INDIRECT_IRRADIANCE and s[3] == 1)
        assert rows[0][0] == 0 and rows[-1][1] == p.irradiance_r_size
        assert all(rows[i][1] == rows[i + 1][0] for i in range(len(rows) - 1))             # a partition of the rows
        exch = [[s for s in pl if s[0] != sharded.SHARD_STAGE] for pl in plans]
        assert all(e == exch[0] for e in exch)
        got = sharded.bytes_received(p, 1, world)
        table = R * p.scattering_mu_size * p.scattering_nu_size * p.scattering_mu_s_size * 8
        assert got["all_gather"] == (world - 1) * table // world * (p.order - 1 + 1)          # density per order + the result
        assert got["halo"] <= 2 * (table // R) * (2 + p.order - 2)
    # 2 GiB tables are exchanged in sub-slabs behind the density kernels, 8 MiB tables in one piece
    n_gather = lambda p, flags: sum(1 for s in _steps(p, 0, 8, flags) if s[0] == sharded.SHARD_ALLGATHER and s[2] == api.IMAGE_SCATTERING_DENSITY)
    assert n_gather(fb.Parameters(order=8, **hires), 0) == 7 * 4 and n_gather(fb.Parameters(order=8, **hires), sharded.NO_PIPELINE) == 7
    assert n_gather(fb.Parameters(order=4), 0) == 3


def test_plan_rejects_uneven_slabs():
    with pytest.raises(fb.FuzzyblueError):
        sharded.plan(fb.Parameters(), 0, 3)
    with pytest.raises(fb.FuzzyblueError):
        sharded.plan(fb.Parameters(), 2, 2)


@pytest.mark.parametrize("order,pipelined", [(3, False), (4, False), (3, True)])
def test_two_rank_sharded_precompute_equals_single_process(tmp_path, order, pipelined):
    """pipelined: sub-slab exchanges forced (2 sub-slabs per 2-level slab), the form the 2 GiB tables use."""
    world, port = 2, 29500 + (os.getpid() % 2000) + order + (7 if pipelined else 0)
    flags = sharded.GATHER_RESULT | (sharded.PIPELINE_ALWAYS if pipelined else 0)
    mp.spawn(_worker, args=(world, port, order, str(tmp_path), flags), nprocs=world, join=True)
    ref = O.precompute(O.Params(order=order, **DIMS), O.F32)
    rows = 0
    for rank in range(world):
        got = np.load(tmp_path / f"rank{rank}.npz")
        assert np.array_equal(got["T"], ref.transmittance)
        assert np.array_equal(got["E"], ref.irradiance)           # rows computed by their owners, broadcast to everyone
        assert np.array_equal(got["S"], ref.scattering)           # every rank holds the full, identical final table
        rows += int(got["rows"])
        # 2 single-scattering halos; per order: the density all-gather, (before the next order) the two ground rows of
        # delta_irradiance and the halo of delta_multiple_scattering; finally the irradiance rows of both ranks + the table
        chunks = 2 if pipelined else 1      # density per order in `chunks` pieces; the result in `chunks` pieces
        assert int(got["exchanges"]) == 2 + chunks * (order - 1) + (order - 2) * 3 + world + chunks
    assert rows == (order - 1) * DIMS["irradiance_r_size"]          # every irradiance row evaluated exactly once per order
    # each rank's own slab of the last delta_multiple_scattering is current; of the peer's slab only the halo slice was
    # ever received, and that one order earlier (the last order's temporaries are not exchanged: nothing reads them)
    r0 = np.load(tmp_path / "rank0.npz")["dMS"]
    prev = O.precompute(O.Params(order=order - 1, **DIMS), O.F32)
    assert np.array_equal(r0[:2], ref.delta_multiple_scattering[:2])
    assert np.array_equal(r0[2], prev.delta_multiple_scattering[2])
    assert np.all(np.isnan(r0[3]))


def _simulate(world, order, dims, flags):
    """All ranks of a world in ONE process: each rank runs its stage steps up to its next exchange; when every rank waits at
    the same exchange it is carried out between the ranks' images (the semantics of sharded.GlooExecutor / the NCCL
    executor: all-gather of sub-slabs, one-slice halos with the neighbours, rows broadcast from their owner)."""
    p = fb.Parameters(order=order, **dims)
    plans = [sharded.plan(p, r, world, flags) for r in range(world)]
    be = [OracleBackend(O.Params(order=order, **dims)) for _ in range(world)]
    n = p.scattering_r_size // world
    pc = [0] * world
    exchanges = 0
    while True:
        for r in range(world):
            while pc[r] < len(plans[r]) and plans[r][pc[r]].op in (sharded.SHARD_STAGE, sharded.SHARD_JOIN):
                s = plans[r][pc[r]]
                if s.op == sharded.SHARD_STAGE:
                    be[r].run_stage(s.stage, s.order, s.begin, s.end)
                pc[r] += 1
        if all(pc[r] == len(plans[r]) for r in range(world)):
            break
        heads = [plans[r][pc[r]] for r in range(world)]          # a rank that has finished while others wait = deadlock
        key = lambda s: (s.op, s.image, s.begin, s.end, s.root)
        assert all(key(h) == key(heads[0]) for h in heads), [repr(h) for h in heads]
        s = heads[0]
        if s.op == sharded.SHARD_ALLGATHER:
            for q in range(world):
                src = be[q].tensor(s.image)[q * n + s.begin:q * n + s.end].clone()
                for r in range(world):
                    if r != q:
                        be[r].tensor(s.image)[q * n + s.begin:q * n + s.end] = src
        elif s.op == sharded.SHARD_HALO:
            first = [be[r].tensor(s.image)[r * n].clone() for r in range(world)]
            last = [be[r].tensor(s.image)[(r + 1) * n - 1].clone() for r in range(world)]
            for r in range(world):
                if r > 0:
                    be[r].tensor(s.image)[r * n - 1] = last[r - 1]
                if r < world - 1:
                    be[r].tensor(s.image)[(r + 1) * n] = first[r + 1]
        elif s.op == sharded.SHARD_BCAST_ROWS:
            src = be[s.root].tensor(s.image)[s.begin:s.end].clone()
            for r in range(world):
                if r != s.root:
                    be[r].tensor(s.image)[s.begin:s.end] = src
        else:
            raise AssertionError(s.op)
        exchanges += 1
        for r in range(world):
            pc[r] += 1
    return be, exchanges


@pytest.mark.parametrize("world,pipelined,r_size,order", [(4, False, 8, 3), (8, False, 8, 3), (8, True, 8, 3), (4, True, 16, 4), (8, False, 16, 4)])
def test_simulated_worlds_of_4_and_8_equal_single_process(world, pipelined, r_size, order):
    """The plan for 4 and 8 ranks (what the 8-GPU runs execute), with fewer irradiance rows than ranks at world 8 and one
    altitude level per rank: every slab a rank never computes starts as NaN, and every rank must end with the single-process
    tables, exactly.  (World 2 runs through real gloo processes above.)"""
    dims = dict(DIMS, scattering_r_size=r_size)
    flags = sharded.GATHER_RESULT | (sharded.PIPELINE_ALWAYS if pipelined else 0)
    be, exchanges = _simulate(world, order, dims, flags)
    ref = O.precompute(O.Params(order=order, **dims), O.F32)
    for r in range(world):
        assert np.array_equal(be[r]._np(api.IMAGE_TRANSMITTANCE), ref.transmittance)
        assert np.array_equal(be[r]._np(api.IMAGE_IRRADIANCE), ref.irradiance), r
        assert np.array_equal(be[r]._np(api.IMAGE_SCATTERING), ref.scattering), r
    assert sum(b.rows_computed for b in be) == (order - 1) * dims["irradiance_r_size"]     # each row once per order
    assert exchanges > 0
