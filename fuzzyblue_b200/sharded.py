"""r-slab sharding of ONE atmosphere across ranks (BASELINE.json config 3), one process per GPU.

Which stage needs which exchange (dependency analysis of the shaders):
  transmittance, direct/indirect irradiance   tiny 2-D tables: computed replicated on every rank
  single_scattering                           r-local (needs only the transmittance table)
  scattering_density at r                     reads the previous order's 3-D tables at the SAME r +- one slice (u_r lands
                                              within 1e-2 texels of the texel centre) and row 0 of delta_irradiance
  multiple_scattering at r                    marches along the ray through ALL r of scattering_density
                                              (multiple_scattering.comp:35-44)  ->  all-gather before K6
  indirect_irradiance                         rows are linear in r, unrelated to the scattering r grid -> needs ALL r of
                                              the previous order's tables

A slab [r0, r1) is one contiguous byte range of the linear [r][mu][nu*mu_s][4] layout, so every exchange is an in-place
all-gather of slab views (NCCL over NVLink on GPUs; gloo in the CPU test).  This module is pure host logic: the stages
themselves run through a backend (the CUDA library in production, see `PendingBackend`).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

from . import api

S3 = (api.IMAGE_SCATTERING, api.IMAGE_DELTA_RAYLEIGH, api.IMAGE_DELTA_MIE, api.IMAGE_SCATTERING_DENSITY,
      api.IMAGE_DELTA_MULTIPLE_SCATTERING)


def slab_of(rank: int, world: int, r_size: int) -> Tuple[int, int]:
    """Even split of the r axis; r_size must divide by the world size (32/8 and 128/8 do)."""
    if r_size % world:
        raise ValueError(f"scattering_r_size={r_size} does not divide by world size {world}")
    n = r_size // world
    return rank * n, (rank + 1) * n


class PendingBackend:
    """Production backend: stages are kernels of libfuzzyblue_b200.so on this rank's GPU, images are its device memory."""

    def __init__(self, pending: api.PendingAtmosphere, stream=None):
        import torch
        self.pending, self.stream, self._torch = pending, stream, torch

    def run_stage(self, stage: int, order: int = 0, r_begin: int = 0, r_end: int = 0):
        self.pending.run_stage(stage, order, r_begin, r_end, self.stream)

    def tensor(self, image: int):
        """The whole image as a torch CUDA tensor aliasing the library's allocation (first dim = r for 3-D images)."""
        torch = self._torch
        ptr, nbytes = self.pending.image(image)
        shape = self.pending._shape(image)
        f16 = image in S3

        class _Raw:
            __cuda_array_interface__ = {"shape": shape, "typestr": "<f2" if f16 else "<f4", "data": (ptr, False), "version": 2}

        return torch.as_tensor(_Raw(), device="cuda")


class ShardedPrecompute:
    """The command stream of Atmosphere::build (src/precompute.rs:1671-2048) with the 3-D stages restricted to this
    rank's r-slab and the exchanges listed above.

    The two big producers (scattering_density before K6, multiple_scattering before the next order) are launched in
    `chunks` sub-slabs; as soon as a sub-slab has been enqueued its all-gather goes onto a communication stream, so the
    NVLink transfer of sub-slab c overlaps the computation of sub-slab c+1 (GPU backends; the CPU test backend has no
    streams and gathers synchronously)."""

    def __init__(self, backend, r_size: int, order: int, rank: int, world: int, group=None, chunks: int = 4,
                 min_chunk_bytes: int = 16 << 20):
        self.b, self.order, self.rank, self.world, self.group = backend, order, rank, world, group
        self.r0, self.r1 = slab_of(rank, world, r_size)
        self.r_size = r_size
        n = self.r1 - self.r0
        self.chunks = max(c for c in range(1, max(1, min(chunks, n)) + 1) if n % c == 0)
        if world > 1:   # splitting only pays when a sub-slab is a real transfer (>= 16 MiB); small tables go in one piece
            t = backend.tensor(api.IMAGE_SCATTERING_DENSITY)
            while self.chunks > 1 and (t.numel() * t.element_size() // world // self.chunks < min_chunk_bytes or n % self.chunks):
                self.chunks -= 1
        self.gathers = 0
        self.bytes_received = 0          # per rank, summed over all exchanges
        self._comm = None
        if world > 1 and getattr(backend, "stream", None) is not None:
            import torch
            self._comm = torch.cuda.Stream()

    def _all_gather(self, image: int, lo: int = 0, hi: int = 0):
        """All-gather rows [lo, hi) (relative to each rank's slab; default: the whole slab) of `image`."""
        if self.world == 1:
            return
        import torch.distributed as dist
        full = self.b.tensor(image)
        n = self.r_size // self.world
        hi = hi or n
        views = [full[i * n + lo:i * n + hi] for i in range(self.world)]
        if self._comm is None:
            dist.all_gather(views, views[self.rank], group=self.group)
        else:
            import torch
            ev = torch.cuda.Event()
            ev.record(self.b.stream)                 # the sub-slab's producer kernel has been enqueued on the compute stream
            self._comm.wait_event(ev)
            with torch.cuda.stream(self._comm):
                dist.all_gather(views, views[self.rank], group=self.group)
        self.gathers += 1
        self.bytes_received += (self.world - 1) * views[0].numel() * views[0].element_size()

    def _join(self):
        """Dependent stages on the compute stream wait for every exchange issued so far."""
        if self._comm is not None:
            self.b.stream.wait_stream(self._comm)

    def _produce_and_gather(self, stage: int, order: int, image: int, gather: bool = True):
        m = (self.r1 - self.r0) // self.chunks
        for c in range(self.chunks):
            self.b.run_stage(stage, order, self.r0 + c * m, self.r0 + (c + 1) * m)
            if gather:
                self._all_gather(image, c * m, (c + 1) * m)

    def run(self, gather_result: bool = True):
        b, r0, r1 = self.b, self.r0, self.r1
        b.run_stage(api.STAGE_TRANSMITTANCE)
        b.run_stage(api.STAGE_DIRECT_IRRADIANCE)
        b.run_stage(api.STAGE_SINGLE_SCATTERING, 0, r0, r1)
        b.run_stage(api.STAGE_CLEAR_IRRADIANCE)
        if self.order >= 2:
            self._all_gather(api.IMAGE_DELTA_RAYLEIGH)      # order-2 density halo + indirect irradiance of order 1
            self._all_gather(api.IMAGE_DELTA_MIE)
            self._join()
        for order in range(2, self.order + 1):
            # K4 in sub-slabs, each gathered behind the next one's computation: the ray march of K6 crosses every r
            self._produce_and_gather(api.STAGE_SCATTERING_DENSITY, order, api.IMAGE_SCATTERING_DENSITY)
            b.run_stage(api.STAGE_INDIRECT_IRRADIANCE, order - 1)          # replicated: 2-D, reads all r (previous order)
            self._join()
            # K6 likewise; its output feeds the next order's density halo and indirect irradiance
            self._produce_and_gather(api.STAGE_MULTIPLE_SCATTERING, 0, api.IMAGE_DELTA_MULTIPLE_SCATTERING,
                                     gather=order < self.order)
            self._join()
        if gather_result:
            self._all_gather(api.IMAGE_SCATTERING)
            self._join()
        return self
