"""r-slab sharding of ONE atmosphere across ranks (BASELINE.json config 3), one process per GPU.

The schedule lives in the C library (`fb_sharded_plan`, fuzzyblue_b200/csrc/fb_sharded.cu, where the dependency
analysis of the shaders is written down): a list of steps per rank — stages on this rank's slab of the r axis,
one-slice halo exchanges with the neighbouring ranks, the all-gather of scattering_density that the ray march of
multiple_scattering needs, 1 KiB broadcasts of the irradiance rows.  Production executes it inside the library with
NCCL over NVLink (`build_sharded` / `PendingAtmosphere.run_sharded` below are thin callers of
`fb_atmosphere_build_sharded` / `fb_pending_run_sharded`).  `GlooExecutor` runs the SAME step list through
`torch.distributed` point-to-point / broadcast calls on host tensors: it is what tests/test_sharded_cpu.py uses with a
world-size-2 gloo group to prove that every cross-slab dependency is covered by a step.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_uint32, c_void_p
from typing import List, Tuple

from . import api

SHARD_STAGE, SHARD_ALLGATHER, SHARD_HALO, SHARD_BCAST_ROWS, SHARD_JOIN = range(5)
GATHER_RESULT, NO_PIPELINE, PIPELINE_ALWAYS = 1, 2, 4

S3 = (api.IMAGE_SCATTERING, api.IMAGE_DELTA_RAYLEIGH, api.IMAGE_DELTA_MIE, api.IMAGE_SCATTERING_DENSITY,
      api.IMAGE_DELTA_MULTIPLE_SCATTERING)


class FbShardStep(ctypes.Structure):
    _fields_ = [("op", ctypes.c_int32), ("stage", ctypes.c_int32), ("image", ctypes.c_int32), ("order", ctypes.c_uint32),
                ("begin", ctypes.c_uint32), ("end", ctypes.c_uint32), ("root", ctypes.c_int32), ("_pad", ctypes.c_int32)]

    def __repr__(self):
        names = ("STAGE", "ALLGATHER", "HALO", "BCAST_ROWS", "JOIN")
        return f"{names[self.op]}(stage={self.stage}, image={self.image}, order={self.order}, [{self.begin},{self.end}), root={self.root})"


def slab_of(rank: int, world: int, r_size: int) -> Tuple[int, int]:
    """Even split of the r axis; r_size must divide by the world size (32/8 and 128/8 do)."""
    if r_size % world:
        raise ValueError(f"scattering_r_size={r_size} does not divide by world size {world}")
    n = r_size // world
    return rank * n, (rank + 1) * n


def plan(params: api.Parameters, rank: int, world: int, flags: int = GATHER_RESULT) -> List[FbShardStep]:
    """The steps rank `rank` of `world` executes for `params` (pure host logic in the C library; no GPU needed)."""
    raw = params.raw()
    n = c_uint32()
    api._check(api._lib().fb_sharded_plan(byref(raw), params.order, rank, world, flags, None, 0, byref(n)))
    steps = (FbShardStep * n.value)()
    api._check(api._lib().fb_sharded_plan(byref(raw), params.order, rank, world, flags, steps, n.value, byref(n)))
    return list(steps)


def bytes_received(params: api.Parameters, rank: int, world: int, flags: int = GATHER_RESULT) -> dict:
    """Bytes this rank receives per precompute, by kind of exchange (for the bench's `collective` record)."""
    R = params.scattering_r_size
    slice_b = params.scattering_mu_size * params.scattering_nu_size * params.scattering_mu_s_size * 8
    row_b = params.irradiance_mu_s_size * 16
    out = {"all_gather": 0, "halo": 0, "rows": 0, "exchanges": 0}
    for s in plan(params, rank, world, flags):
        if s.op == SHARD_ALLGATHER:
            out["all_gather"] += (world - 1) * (s.end - s.begin) * slice_b
        elif s.op == SHARD_HALO:
            out["halo"] += ((rank > 0) + (rank < world - 1)) * slice_b
        elif s.op == SHARD_BCAST_ROWS:
            out["rows"] += (s.end - s.begin) * row_b if s.root != rank else 0
        if s.op in (SHARD_ALLGATHER, SHARD_HALO, SHARD_BCAST_ROWS):
            out["exchanges"] += 1
    out["total"] = out["all_gather"] + out["halo"] + out["rows"]
    return out


class NcclComm:
    """An NCCL communicator created through the library's own thin wrappers (`fb_nccl_*`): what a caller without NCCL
    bindings uses.  The 128-byte unique id travels from rank 0 to the others through `torch.distributed` (any backend)."""

    def __init__(self, device: int, rank: int, world: int, group=None):
        import torch
        import torch.distributed as dist
        self.rank, self.world = rank, world
        ident = (ctypes.c_char * 128)()
        if rank == 0:
            api._check(api._lib().fb_nccl_unique_id(ctypes.cast(ident, c_void_p)))
        dev = torch.device("cuda", device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        t = torch.frombuffer(bytearray(bytes(ident)), dtype=torch.uint8).clone().to(dev)
        dist.broadcast(t, 0, group=group)
        buf = bytes(t.cpu().numpy().tobytes())
        h = c_void_p()
        api._check(api._lib().fb_nccl_comm_create(device, world, rank, ctypes.c_char_p(buf), byref(h)))
        self.handle = h.value

    def close(self):
        if getattr(self, "handle", None):
            api._lib().fb_nccl_comm_destroy(c_void_p(self.handle))
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def build_sharded(builder: api.Builder, params: api.Parameters, comm, rank: int, world: int, flags: int = GATHER_RESULT,
                  stream=None) -> api.PendingAtmosphere:
    """Atmosphere::build across `world` GPUs (src/precompute.rs:1077-1081): `fb_atmosphere_build_sharded`."""
    h = c_void_p()
    raw = params.raw()
    api._check(api._lib().fb_atmosphere_build_sharded(builder._h, byref(raw), params.order, c_void_p(comm.handle if comm is not None else 0),
                                                      rank, world, flags, api._stream(stream), byref(h)))
    return api.PendingAtmosphere(h, builder, params)


class GlooExecutor:
    """Runs a plan with `torch.distributed` collectives on host tensors (gloo): the CPU stand-in for the library's NCCL
    executor.  `backend.run_stage(stage, order, begin, end)` evaluates a stage, `backend.tensor(image)` returns the whole
    image as a tensor whose first dimension is r (3-D images) or the irradiance row (2-D images)."""

    def __init__(self, backend, params: api.Parameters, rank: int, world: int, group=None):
        self.b, self.p, self.rank, self.world, self.group = backend, params, rank, world, group
        self.exchanges = 0

    def run(self, flags: int = GATHER_RESULT):
        import torch.distributed as dist
        n = self.p.scattering_r_size // self.world
        a, b = self.rank * n, (self.rank + 1) * n
        for s in plan(self.p, self.rank, self.world, flags):
            if s.op == SHARD_STAGE:
                self.b.run_stage(s.stage, s.order, s.begin, s.end)
            elif s.op == SHARD_ALLGATHER:
                full = self.b.tensor(s.image)
                for q in range(self.world):          # one in-place broadcast per owner, as the NCCL executor does
                    dist.broadcast(full[q * n + s.begin:q * n + s.end], q, group=self.group)
                self.exchanges += 1
            elif s.op == SHARD_HALO:
                full = self.b.tensor(s.image)
                ops = []
                if self.rank > 0:
                    ops += [dist.P2POp(dist.isend, full[a].contiguous(), self.rank - 1, self.group),
                            dist.P2POp(dist.irecv, full[a - 1], self.rank - 1, self.group)]
                if self.rank < self.world - 1:
                    ops += [dist.P2POp(dist.isend, full[b - 1].contiguous(), self.rank + 1, self.group),
                            dist.P2POp(dist.irecv, full[b], self.rank + 1, self.group)]
                for w in dist.batch_isend_irecv(ops) if ops else []:
                    w.wait()
                self.exchanges += 1
            elif s.op == SHARD_BCAST_ROWS:
                dist.broadcast(self.b.tensor(s.image)[s.begin:s.end], s.root, group=self.group)
                self.exchanges += 1
            elif s.op == SHARD_JOIN:
                pass                                 # host collectives are synchronous
            else:
                raise ValueError(s.op)
        return self
