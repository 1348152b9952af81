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
    rank's r-slab and the exchanges listed above."""

    def __init__(self, backend, r_size: int, order: int, rank: int, world: int, group=None):
        self.b, self.order, self.rank, self.world, self.group = backend, order, rank, world, group
        self.r0, self.r1 = slab_of(rank, world, r_size)
        self.r_size = r_size
        self.gathers = 0

    def _all_gather(self, image: int):
        if self.world == 1:
            return
        import torch.distributed as dist
        full = self.b.tensor(image)
        n = self.r_size // self.world
        views = [full[i * n:(i + 1) * n] for i in range(self.world)]
        dist.all_gather(views, views[self.rank], group=self.group)
        self.gathers += 1

    def run(self, gather_result: bool = True):
        b, r0, r1 = self.b, self.r0, self.r1
        b.run_stage(api.STAGE_TRANSMITTANCE)
        b.run_stage(api.STAGE_DIRECT_IRRADIANCE)
        b.run_stage(api.STAGE_SINGLE_SCATTERING, 0, r0, r1)
        b.run_stage(api.STAGE_CLEAR_IRRADIANCE)
        if self.order >= 2:
            self._all_gather(api.IMAGE_DELTA_RAYLEIGH)      # order-2 density halo + indirect irradiance of order 1
            self._all_gather(api.IMAGE_DELTA_MIE)
        for order in range(2, self.order + 1):
            b.run_stage(api.STAGE_SCATTERING_DENSITY, order, r0, r1)
            b.run_stage(api.STAGE_INDIRECT_IRRADIANCE, order - 1)          # replicated: 2-D, reads all r
            self._all_gather(api.IMAGE_SCATTERING_DENSITY)                 # the ray march of K6 crosses every r
            b.run_stage(api.STAGE_MULTIPLE_SCATTERING, 0, r0, r1)
            if order < self.order:
                self._all_gather(api.IMAGE_DELTA_MULTIPLE_SCATTERING)      # next density halo + next indirect irradiance
        if gather_result:
            self._all_gather(api.IMAGE_SCATTERING)
        return self
