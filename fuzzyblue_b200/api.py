"""Host-side mirror of the reference crate's public API over the C ABI of ``include/fuzzyblue.h``.

The reference (a Rust crate) re-exports six names — ``Atmosphere, Builder, Parameters,
PendingAtmosphere, DrawParameters, Renderer`` (/root/reference/src/lib.rs:8-12).  This module
keeps those names, field names (including the ``absorbtion_*`` spelling of
src/precompute.rs:756,761), argument meaning and call order, with a CUDA stream standing where the
reference takes a ``vk::CommandBuffer``: calls *enqueue* and return, the caller synchronises.

Only ctypes is used here — no torch, no numpy-side arithmetic, and no fallback of any kind: if
``libfuzzyblue_b200.so`` is missing or no sm_100 device is present the import / ``Builder`` fails.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_float, c_int, c_int32, c_size_t, c_uint32, c_uint64, c_void_p
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FUZZYBLUE_B200_LIB") or os.path.join(_HERE, "csrc", "libfuzzyblue_b200.so")   # override: kernel experiments

# enums of include/fuzzyblue.h
FB_OK = 0
(IMAGE_TRANSMITTANCE, IMAGE_IRRADIANCE, IMAGE_SCATTERING, IMAGE_DELTA_IRRADIANCE, IMAGE_DELTA_RAYLEIGH, IMAGE_DELTA_MIE,
 IMAGE_SCATTERING_DENSITY, IMAGE_DELTA_MULTIPLE_SCATTERING) = range(8)
(STAGE_TRANSMITTANCE, STAGE_DIRECT_IRRADIANCE, STAGE_SINGLE_SCATTERING, STAGE_SCATTERING_DENSITY,
 STAGE_INDIRECT_IRRADIANCE, STAGE_MULTIPLE_SCATTERING, STAGE_CLEAR_IRRADIANCE) = range(7)
KERNELS_FAST, KERNELS_REFERENCE = 0, 1


class FuzzyblueError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{_lib().fb_status_string(status).decode()}: {message}")
        self.status = status


# ---------------------------------------------------------------------------------------------
# C structs
# ---------------------------------------------------------------------------------------------
class FbDensityProfileLayer(ctypes.Structure):
    _fields_ = [("width", c_float), ("exp_term", c_float), ("exp_scale", c_float), ("linear_term", c_float),
                ("constant_term", c_float), ("_pad", c_float * 3)]


class FbDensityProfile(ctypes.Structure):
    _fields_ = [("layers", FbDensityProfileLayer * 2)]


class FbParams(ctypes.Structure):
    _fields_ = [("solar_irradiance", c_float * 3), ("sun_angular_radius", c_float),
                ("rayleigh_scattering", c_float * 3), ("bottom_radius", c_float),
                ("mie_scattering", c_float * 3), ("top_radius", c_float),
                ("mie_extinction", c_float * 3), ("mie_phase_function_g", c_float),
                ("ground_albedo", c_float * 3), ("mu_s_min", c_float),
                ("absorption_extinction", c_float * 3),
                ("transmittance_mu_size", c_int32), ("transmittance_r_size", c_int32),
                ("scattering_r_size", c_int32), ("scattering_mu_size", c_int32),
                ("scattering_mu_s_size", c_int32), ("scattering_nu_size", c_int32),
                ("irradiance_mu_s_size", c_int32), ("irradiance_r_size", c_int32), ("_pad", c_int32),
                ("rayleigh_density", FbDensityProfile), ("mie_density", FbDensityProfile),
                ("absorption_density", FbDensityProfile)]


class FbDrawParams(ctypes.Structure):
    _fields_ = [("inverse_viewproj", (c_float * 4) * 4), ("camera_position", c_float * 3), ("_pad", c_uint32),
                ("sun_direction", c_float * 3)]


class FbExtent2D(ctypes.Structure):
    _fields_ = [("width", c_uint32), ("height", c_uint32)]


class FbExtent3D(ctypes.Structure):
    _fields_ = [("width", c_uint32), ("height", c_uint32), ("depth", c_uint32)]


assert ctypes.sizeof(FbParams) == 320 and ctypes.sizeof(FbDrawParams) == 92

# every symbol include/fuzzyblue.h declares: name -> (restype, argtypes)
_P = POINTER
class FbExportLayout(ctypes.Structure):
    """Where the three tables sit in an exported allocation (include/fuzzyblue.h: FbExportLayout)."""
    _fields_ = [("allocation_bytes", c_size_t), ("scattering_offset", c_size_t), ("scattering_bytes", c_size_t),
                ("transmittance_offset", c_size_t), ("transmittance_bytes", c_size_t),
                ("irradiance_offset", c_size_t), ("irradiance_bytes", c_size_t)]


ABI = {
    "fb_status_string": (c_char_p, [c_int]),
    "fb_last_error": (c_char_p, []),
    "fb_version": (c_char_p, []),
    "fb_params_default": (c_int, [_P(FbParams)]),
    "fb_params_default_order": (c_uint32, []),
    "fb_params_transmittance_extent": (c_int, [_P(FbParams), _P(FbExtent2D)]),
    "fb_params_irradiance_extent": (c_int, [_P(FbParams), _P(FbExtent2D)]),
    "fb_params_scattering_extent": (c_int, [_P(FbParams), _P(FbExtent3D)]),
    "fb_params_validate": (c_int, [_P(FbParams)]),
    "fb_builder_create": (c_int, [c_int, _P(c_void_p)]),
    "fb_builder_destroy": (None, [c_void_p]),
    "fb_builder_set_kernels": (c_int, [c_void_p, c_int]),
    "fb_builder_trim": (c_int, [c_void_p]),
    "fb_builder_device": (c_int, [c_void_p]),
    "fb_builder_sm_count": (c_int, [c_void_p]),
    "fb_builder_measure_peaks": (c_int, [c_void_p, _P(ctypes.c_double), _P(ctypes.c_double)]),
    "fb_atmosphere_build": (c_int, [c_void_p, _P(FbParams), c_uint32, c_void_p, _P(c_void_p)]),
    "fb_atmosphere_allocate": (c_int, [c_void_p, _P(FbParams), c_uint32, _P(c_void_p)]),
    "fb_pending_resubmit": (c_int, [c_void_p, c_void_p]),
    "fb_pending_set_readback": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "fb_pending_launch_count": (c_int, [c_void_p]),
    "fb_pending_run_stage": (c_int, [c_void_p, c_int, c_uint32, c_uint32, c_uint32, c_void_p]),
    "fb_pending_image": (c_int, [c_void_p, c_int, _P(c_void_p), _P(c_size_t)]),
    "fb_pending_upload": (c_int, [c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "fb_pending_download": (c_int, [c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "fb_pending_atmosphere": (c_int, [c_void_p, _P(c_void_p)]),
    "fb_pending_assert_ready": (c_int, [c_void_p, c_int, _P(c_void_p)]),
    "fb_pending_destroy": (None, [c_void_p]),
    "fb_atmosphere_transmittance": (c_int, [c_void_p, _P(c_void_p), _P(FbExtent2D)]),
    "fb_atmosphere_scattering": (c_int, [c_void_p, _P(c_void_p), _P(FbExtent3D)]),
    "fb_atmosphere_irradiance": (c_int, [c_void_p, _P(c_void_p), _P(FbExtent2D)]),
    "fb_atmosphere_params": (c_int, [c_void_p, _P(FbParams)]),
    "fb_atmosphere_read_transmittance": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "fb_atmosphere_read_scattering": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "fb_atmosphere_read_irradiance": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "fb_atmosphere_destroy": (None, [c_void_p]),
    "fb_precompute_host": (c_int, [c_void_p, _P(FbParams), c_uint32, c_void_p, c_void_p, c_void_p]),
    "fb_renderer_create": (c_int, [c_void_p, _P(c_void_p)]),
    "fb_renderer_destroy": (None, [c_void_p]),
    "fb_renderer_draw": (c_int, [c_void_p, c_void_p, _P(FbDrawParams), c_void_p, c_void_p, c_void_p, c_uint32, c_uint32, c_void_p]),
    "fb_renderer_draw_blend": (c_int, [c_void_p, c_void_p, _P(FbDrawParams), c_void_p, c_void_p, c_uint32, c_uint32, c_void_p]),
    "fb_renderer_draw_sweep": (c_int, [c_void_p, c_void_p, _P(FbDrawParams), c_uint32, c_void_p, c_void_p, c_void_p, c_uint32, c_uint32, c_void_p]),
    "fb_renderer_draw_host": (c_int, [c_void_p, c_void_p, _P(FbDrawParams), c_void_p, c_void_p, c_void_p, c_uint32, c_uint32]),
    "fb_sky_radiance": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p, c_void_p, c_void_p]),
    "fb_sun_and_sky_irradiance": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p, c_void_p, c_void_p]),
    "fb_atmosphere_build_batch": (c_int, [c_void_p, _P(FbParams), c_uint32, c_uint32, c_void_p, _P(c_void_p)]),
    "fb_builder_set_exportable": (c_int, [c_void_p, c_int]),
    "fb_atmosphere_export_fd": (c_int, [c_void_p, _P(c_int), _P(FbExportLayout)]),
    "fb_external_memory_read_fd": (c_int, [c_int, c_int, c_size_t, c_size_t, c_void_p, c_size_t]),
    "fb_external_semaphore_import_fd": (c_int, [c_int, c_int, c_int, _P(c_void_p)]),
    "fb_external_semaphore_signal": (c_int, [c_void_p, c_uint64, c_void_p]),
    "fb_external_semaphore_wait": (c_int, [c_void_p, c_uint64, c_void_p]),
    "fb_external_semaphore_destroy": (None, [c_void_p]),
    "fb_params_slow_stages": (c_uint32, [_P(FbParams)]),
    "fb_pending_slow_stages": (c_uint32, [c_void_p]),
    "fb_pending_wait": (c_int, [c_void_p]),
    "fb_sharded_plan": (c_int, [_P(FbParams), c_uint32, c_int, c_int, c_uint32, c_void_p, c_uint32, _P(c_uint32)]),
    "fb_atmosphere_build_sharded": (c_int, [c_void_p, _P(FbParams), c_uint32, c_void_p, c_int, c_int, c_uint32, c_void_p, _P(c_void_p)]),
    "fb_pending_run_sharded": (c_int, [c_void_p, c_void_p, c_int, c_int, c_uint32, c_void_p]),
    "fb_nccl_version": (c_int, [_P(c_int)]),
    "fb_nccl_unique_id": (c_int, [c_void_p]),
    "fb_nccl_comm_create": (c_int, [c_int, c_int, c_int, c_void_p, _P(c_void_p)]),
    "fb_nccl_comm_destroy": (c_int, [c_void_p]),
}

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C fuzzyblue_b200/csrc`). fuzzyblue_b200 has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in ABI.items():
            fn = getattr(L, name)   # AttributeError = the library does not export what the header declares
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def _check(status: int):
    if status != FB_OK:
        raise FuzzyblueError(status, _lib().fb_last_error().decode())


def _stream(stream) -> c_void_p:
    """Accepts None (legacy default stream), an int handle, or anything with ``.cuda_stream`` (torch)."""
    if stream is None:
        return c_void_p(0)
    if hasattr(stream, "cuda_stream"):
        return c_void_p(stream.cuda_stream)
    return c_void_p(int(stream))


def _ptr(p) -> c_void_p:
    """Device pointer: None, int, or anything with ``.data_ptr()`` (torch tensor)."""
    if p is None:
        return c_void_p(0)
    if hasattr(p, "data_ptr"):
        return c_void_p(p.data_ptr())
    return c_void_p(int(p))


# ---------------------------------------------------------------------------------------------
# Parameters — src/precompute.rs:660-935
# ---------------------------------------------------------------------------------------------
@dataclass
class DensityProfileLayer:   # precompute.rs:660-666
    width: float = 0.0
    exp_term: float = 0.0
    exp_scale: float = 0.0
    linear_term: float = 0.0
    constant_term: float = 0.0


@dataclass
class DensityProfile:        # precompute.rs:674-676
    layers: Tuple[DensityProfileLayer, DensityProfileLayer] = (DensityProfileLayer(), DensityProfileLayer())


@dataclass
class Parameters:
    """``Parameters`` of src/precompute.rs:690-769; ``Parameters()`` equals ``Parameters::default()``
    (:849-935, Earth).  The four Vulkan fields are kept for source compatibility and ignored."""
    usage: int = 0
    dst_stage_mask: int = 0x80          # FRAGMENT_SHADER
    dst_access_mask: int = 0x20         # SHADER_READ
    layout: int = 5                     # SHADER_READ_ONLY_OPTIMAL
    order: int = 4
    transmittance_mu_size: int = 256
    transmittance_r_size: int = 64
    scattering_r_size: int = 32
    scattering_mu_size: int = 128
    scattering_mu_s_size: int = 32
    scattering_nu_size: int = 8
    irradiance_mu_s_size: int = 64
    irradiance_r_size: int = 16
    solar_irradiance: Sequence[float] = (1.474, 1.850, 1.91198)
    sun_angular_radius: float = 0.004675
    bottom_radius: float = 6360.0
    top_radius: float = 6420.0
    rayleigh_density: DensityProfile = field(default_factory=lambda: DensityProfile(
        (DensityProfileLayer(), DensityProfileLayer(0.0, 1.0, -0.125, 0.0, 0.0))))
    rayleigh_scattering: Sequence[float] = (0.005802, 0.013558, 0.033100)
    mie_density: DensityProfile = field(default_factory=lambda: DensityProfile(
        (DensityProfileLayer(), DensityProfileLayer(0.0, 1.0, -0.833333, 0.0, 0.0))))
    mie_scattering: Sequence[float] = (0.003996, 0.003996, 0.003996)
    mie_extinction: Sequence[float] = (0.004440, 0.004440, 0.004440)
    mie_phase_function_g: float = 0.8
    absorbtion_density: DensityProfile = field(default_factory=lambda: DensityProfile(
        (DensityProfileLayer(25.0, 0.0, 0.0, 0.066667, -0.666667),
         DensityProfileLayer(0.0, 0.0, 0.0, -0.066667, 2.666667))))
    absorbtion_extinction: Sequence[float] = (6.5e-4, 1.881e-3, 8.5e-5)
    ground_albedo: Sequence[float] = (0.1, 0.1, 0.1)
    mu_s_min: float = -0.207912

    # precompute.rs:771-793
    def transmittance_extent(self) -> Tuple[int, int]:
        return (self.transmittance_mu_size, self.transmittance_r_size)

    def irradiance_extent(self) -> Tuple[int, int]:
        return (self.irradiance_mu_s_size, self.irradiance_r_size)

    def scattering_extent(self) -> Tuple[int, int, int]:
        return (self.scattering_nu_size * self.scattering_mu_s_size, self.scattering_mu_size, self.scattering_r_size)

    def raw(self) -> FbParams:
        """``ParamsRaw::new``, precompute.rs:964-991."""
        p = FbParams()
        for name, src in (("solar_irradiance", self.solar_irradiance), ("rayleigh_scattering", self.rayleigh_scattering),
                          ("mie_scattering", self.mie_scattering), ("mie_extinction", self.mie_extinction),
                          ("ground_albedo", self.ground_albedo), ("absorption_extinction", self.absorbtion_extinction)):
            setattr(p, name, (c_float * 3)(*src))
        for name in ("sun_angular_radius", "bottom_radius", "top_radius", "mie_phase_function_g", "mu_s_min",
                     "transmittance_mu_size", "transmittance_r_size", "scattering_r_size", "scattering_mu_size",
                     "scattering_mu_s_size", "scattering_nu_size", "irradiance_mu_s_size", "irradiance_r_size"):
            setattr(p, name, getattr(self, name))
        for dst, src in ((p.rayleigh_density, self.rayleigh_density), (p.mie_density, self.mie_density),
                         (p.absorption_density, self.absorbtion_density)):
            for i in range(2):
                l = src.layers[i]
                dst.layers[i].width = l.width
                dst.layers[i].exp_term = l.exp_term
                dst.layers[i].exp_scale = l.exp_scale
                dst.layers[i].linear_term = l.linear_term
                dst.layers[i].constant_term = l.constant_term
        return p

    @staticmethod
    def default_raw() -> Tuple[FbParams, int]:
        p = FbParams()
        _check(_lib().fb_params_default(byref(p)))
        return p, int(_lib().fb_params_default_order())


@dataclass
class DrawParameters:        # src/render.rs:246-252
    inverse_viewproj: Sequence[Sequence[float]]   # [[f32;4];4], each inner array one column
    camera_position: Sequence[float]
    sun_direction: Sequence[float]

    def raw(self) -> FbDrawParams:   # DrawParamsRaw::new, render.rs:262-271
        d = FbDrawParams()
        for c in range(4):
            for r in range(4):
                d.inverse_viewproj[c][r] = float(self.inverse_viewproj[c][r])
        d.camera_position = (c_float * 3)(*[float(v) for v in self.camera_position])
        d.sun_direction = (c_float * 3)(*[float(v) for v in self.sun_direction])
        return d


# ---------------------------------------------------------------------------------------------
# Builder / Atmosphere / PendingAtmosphere — src/precompute.rs
# ---------------------------------------------------------------------------------------------
class Builder:
    """``Builder::new`` (precompute.rs:61-68).  The Vulkan instance/device/cache/queue-family
    arguments collapse into a CUDA device ordinal."""

    def __init__(self, device: int = 0, kernels: int = KERNELS_FAST):
        h = c_void_p()
        _check(_lib().fb_builder_create(device, byref(h)))
        self._h = h
        if kernels != KERNELS_FAST:
            self.set_kernels(kernels)

    @classmethod
    def new(cls, device: int = 0) -> "Builder":
        return cls(device)

    def set_kernels(self, kernels: int):
        _check(_lib().fb_builder_set_kernels(self._h, kernels))

    def set_exportable(self, on: bool = True):
        """Atmospheres built afterwards keep their tables in an fd-exportable allocation (Vulkan interop)."""
        _check(_lib().fb_builder_set_exportable(self._h, 1 if on else 0))

    def trim(self):
        """Return cached device blocks of finished precomputes to the driver."""
        _check(_lib().fb_builder_trim(self._h))

    def device(self) -> int:
        return _lib().fb_builder_device(self._h)

    def sm_count(self) -> int:
        return _lib().fb_builder_sm_count(self._h)

    def measure_peaks(self) -> Tuple[float, float]:
        """(dense FP32 FMA TFLOP/s, SFU Gop/s) measured on this device."""
        a, b = ctypes.c_double(), ctypes.c_double()
        _check(_lib().fb_builder_measure_peaks(self._h, byref(a), byref(b)))
        return a.value, b.value

    def close(self):
        if self._h:
            _lib().fb_builder_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _read(fn, handle, shape, dtype, stream):
    out = np.empty(shape, dtype=dtype)
    _check(fn(handle, out.ctypes.data_as(c_void_p), out.nbytes, _stream(stream)))
    return out


class Atmosphere:
    """A precomputed atmosphere (precompute.rs:1036-1043); owns the three kept tables."""

    def __init__(self, handle: c_void_p, builder: Builder, owned: bool):
        self._h, self._builder, self._owned = handle, builder, owned

    @staticmethod
    def build(builder: Builder, stream, params: Parameters) -> "PendingAtmosphere":
        """``Atmosphere::build(builder, cmd, &params)`` (precompute.rs:1077-1081): enqueue the
        whole precompute on ``stream`` and return immediately."""
        h = c_void_p()
        raw = params.raw()
        _check(_lib().fb_atmosphere_build(builder._h, byref(raw), params.order, _stream(stream), byref(h)))
        return PendingAtmosphere(h, builder, params)

    @staticmethod
    def allocate(builder: Builder, params: Parameters) -> "PendingAtmosphere":
        h = c_void_p()
        raw = params.raw()
        _check(_lib().fb_atmosphere_allocate(builder._h, byref(raw), params.order, byref(h)))
        return PendingAtmosphere(h, builder, params)

    def params_raw(self) -> FbParams:
        p = FbParams()
        _check(_lib().fb_atmosphere_params(self._h, byref(p)))
        return p

    # precompute.rs:2075-2101: device pointers (ints) + extents
    def transmittance(self) -> int:
        p = c_void_p()
        _check(_lib().fb_atmosphere_transmittance(self._h, byref(p), None))
        return p.value

    def transmittance_extent(self) -> Tuple[int, int]:
        e = FbExtent2D()
        _check(_lib().fb_atmosphere_transmittance(self._h, None, byref(e)))
        return (e.width, e.height)

    def scattering(self) -> int:
        p = c_void_p()
        _check(_lib().fb_atmosphere_scattering(self._h, byref(p), None))
        return p.value

    def scattering_extent(self) -> Tuple[int, int, int]:
        e = FbExtent3D()
        _check(_lib().fb_atmosphere_scattering(self._h, None, byref(e)))
        return (e.width, e.height, e.depth)

    def irradiance(self) -> int:
        p = c_void_p()
        _check(_lib().fb_atmosphere_irradiance(self._h, byref(p), None))
        return p.value

    def irradiance_extent(self) -> Tuple[int, int]:
        e = FbExtent2D()
        _check(_lib().fb_atmosphere_irradiance(self._h, None, byref(e)))
        return (e.width, e.height)

    # examples/dump.rs:110-193 — host copies in the linear layout (async on `stream`; sync before use)
    def read_transmittance(self, stream=None) -> np.ndarray:
        w, h = self.transmittance_extent()
        return _read(_lib().fb_atmosphere_read_transmittance, self._h, (h, w, 4), np.float32, stream)

    def read_scattering(self, stream=None) -> np.ndarray:
        w, h, d = self.scattering_extent()
        return _read(_lib().fb_atmosphere_read_scattering, self._h, (d, h, w, 4), np.float16, stream)

    def read_irradiance(self, stream=None) -> np.ndarray:
        w, h = self.irradiance_extent()
        return _read(_lib().fb_atmosphere_read_irradiance, self._h, (h, w, 4), np.float32, stream)

    def export_fd(self):
        """(fd, FbExportLayout) of the block holding the three tables — what a Vulkan caller imports as OPAQUE_FD
        device memory in place of the ``vk::Image`` getters (precompute.rs:2075-2101).  The caller owns ``fd``."""
        fd, lay = c_int(-1), FbExportLayout()
        _check(_lib().fb_atmosphere_export_fd(self._h, byref(fd), byref(lay)))
        return fd.value, lay

    def close(self):
        if self._h and self._owned:
            _lib().fb_atmosphere_destroy(self._h)
        self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_IMAGE_DTYPES = {IMAGE_TRANSMITTANCE: np.float32, IMAGE_IRRADIANCE: np.float32, IMAGE_DELTA_IRRADIANCE: np.float32}


class PendingAtmosphere:
    """An atmosphere being prepared by the GPU (precompute.rs:2103-2120).  Must outlive the work
    enqueued on the stream passed to ``Atmosphere.build``."""

    def __init__(self, handle: c_void_p, builder: Builder, params: Parameters):
        self._h, self._builder, self.params = handle, builder, params

    def acquire_ownership(self, stream, compute_queue_family: int, gfx_queue_family: int):
        """Queue-family ownership transfer (precompute.rs:2147-2201) has no CUDA analogue."""

    def atmosphere(self) -> Atmosphere:                       # :2203-2206
        h = c_void_p()
        _check(_lib().fb_pending_atmosphere(self._h, byref(h)))
        return Atmosphere(h, self._builder, owned=False)

    def wait(self):
        """Block until everything submitted through this object has finished (the fence wait of tests/smoke.rs:147-155)."""
        _check(_lib().fb_pending_wait(self._h))

    def assert_ready(self, check: bool = True, wait: bool = True) -> Atmosphere:   # :2208-2211
        """Consumes the pending, frees the temporaries.  The reference leaves a too-early call undefined; here
        ``wait=True`` (default) first waits for the precompute's own completion events, ``wait=False, check=True``
        raises FuzzyblueError(FB_ERR_NOT_READY) if work is still in flight (the pending stays usable), and
        ``check=False`` trusts the caller as the reference does."""
        if wait and check:
            self.wait()
        h = c_void_p()
        _check(_lib().fb_pending_assert_ready(self._h, 1 if check else 0, byref(h)))
        self._h = c_void_p()
        return Atmosphere(h, self._builder, owned=True)

    def slow_stages(self) -> int:
        """Bit s set: stage s ran the one-thread-per-texel transcription although FAST kernels were asked for."""
        return int(_lib().fb_pending_slow_stages(self._h))

    def run_sharded(self, comm, rank: int, world: int, flags: int = 0, stream=None):
        """The r-slab sharded schedule (fb_sharded_plan) on the images this pending owns, exchanges over NCCL."""
        _check(_lib().fb_pending_run_sharded(self._h, c_void_p(comm.handle if comm is not None else 0), rank, world, flags, _stream(stream)))

    def resubmit(self, stream=None):
        """Replay the recorded command stream (what benches/precompute.rs:138-148 times)."""
        _check(_lib().fb_pending_resubmit(self._h, _stream(stream)))

    def set_readback(self, transmittance=None, scattering=None, irradiance=None):
        """Record the read-back of the finished tables into the command stream: later `resubmit` calls also copy
        them to these host buffers (addresses of pinned memory, or None), overlapped with the last kernels."""
        _check(_lib().fb_pending_set_readback(self._h, c_void_p(transmittance), c_void_p(scattering), c_void_p(irradiance)))

    def launch_count(self) -> int:
        return _lib().fb_pending_launch_count(self._h)

    def run_stage(self, stage: int, order: int = 0, r_begin: int = 0, r_end: int = 0, stream=None):
        _check(_lib().fb_pending_run_stage(self._h, stage, order, r_begin, r_end, _stream(stream)))

    def image(self, image: int) -> Tuple[int, int]:
        p, n = c_void_p(), c_size_t()
        _check(_lib().fb_pending_image(self._h, image, byref(p), byref(n)))
        return p.value, n.value

    def _shape(self, image: int):
        P = self.params
        if image == IMAGE_TRANSMITTANCE:
            return (P.transmittance_r_size, P.transmittance_mu_size, 4)
        if image in (IMAGE_IRRADIANCE, IMAGE_DELTA_IRRADIANCE):
            return (P.irradiance_r_size, P.irradiance_mu_s_size, 4)
        return (P.scattering_r_size, P.scattering_mu_size, P.scattering_nu_size * P.scattering_mu_s_size, 4)

    def upload(self, image: int, host: np.ndarray, stream=None):
        a = np.ascontiguousarray(host, dtype=_IMAGE_DTYPES.get(image, np.float16))
        assert a.shape == self._shape(image), (a.shape, self._shape(image))
        _check(_lib().fb_pending_upload(self._h, image, a.ctypes.data_as(c_void_p), a.nbytes, _stream(stream)))

    def download(self, image: int, stream=None) -> np.ndarray:
        out = np.empty(self._shape(image), dtype=_IMAGE_DTYPES.get(image, np.float16))
        _check(_lib().fb_pending_download(self._h, image, out.ctypes.data_as(c_void_p), out.nbytes, _stream(stream)))
        return out

    def close(self):
        if self._h:
            _lib().fb_pending_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def build_batch(builder: Builder, params: List[Parameters], stream=None) -> List[PendingAtmosphere]:
    """Independent atmospheres (BASELINE.json config 4), overlapped on forked streams.  The C entry point takes ONE
    scattering order for the whole batch, so parameter sets of different orders are refused here."""
    n = len(params)
    if n == 0:
        return []
    if any(p.order != params[0].order for p in params):
        raise ValueError("build_batch: every atmosphere of a batch must use the same scattering order")
    raws = (FbParams * n)(*[p.raw() for p in params])
    outs = (c_void_p * n)()
    _check(_lib().fb_atmosphere_build_batch(builder._h, raws, n, params[0].order, _stream(stream), outs))
    return [PendingAtmosphere(c_void_p(outs[i]), builder, params[i]) for i in range(n)]


def precompute_host(builder: Builder, params: Parameters):
    """Host-buffer end-to-end call: parameters in host memory in, the three tables out."""
    P = params
    T = np.empty((P.transmittance_r_size, P.transmittance_mu_size, 4), np.float32)
    S = np.empty((P.scattering_r_size, P.scattering_mu_size, P.scattering_nu_size * P.scattering_mu_s_size, 4), np.float16)
    E = np.empty((P.irradiance_r_size, P.irradiance_mu_s_size, 4), np.float32)
    raw = P.raw()
    vp = lambda a: a.ctypes.data_as(c_void_p)
    _check(_lib().fb_precompute_host(builder._h, byref(raw), P.order, vp(T), vp(S), vp(E)))
    return T, S, E


# ---------------------------------------------------------------------------------------------
# Renderer — src/render.rs
# ---------------------------------------------------------------------------------------------
class Renderer:
    """``Renderer::new(builder, cache, render_pass, subpass, frames)`` (render.rs:34-40); the
    Vulkan-only arguments are accepted and ignored."""

    def __init__(self, builder: Builder, cache=None, render_pass=None, subpass: int = 0, frames: int = 1):
        h = c_void_p()
        _check(_lib().fb_renderer_create(builder._h, byref(h)))
        self._h, self._builder = h, builder
        self._depth = {}

    def set_depth_buffer(self, frame: int, depth_ptr):
        """render.rs:194-207: remember the depth attachment (device pointer, [h][w] f32) of a frame."""
        self._depth[frame] = depth_ptr

    def draw(self, stream, atmosphere: Atmosphere, frame: int, params: DrawParameters, color_out=None, transmittance_out=None,
             width: int = 0, height: int = 0):
        """``Renderer::draw(cmd, atmosphere, frame, params)`` (render.rs:209-236) + the two
        fragment outputs as explicit device buffers."""
        raw = params.raw()
        _check(_lib().fb_renderer_draw(self._h, atmosphere._h, byref(raw), _ptr(self._depth[frame]), _ptr(color_out),
                                       _ptr(transmittance_out), width, height, _stream(stream)))

    def draw_blend(self, stream, atmosphere: Atmosphere, frame: int, params: DrawParameters, framebuffer, width: int, height: int):
        raw = params.raw()
        _check(_lib().fb_renderer_draw_blend(self._h, atmosphere._h, byref(raw), _ptr(self._depth[frame]), _ptr(framebuffer),
                                             width, height, _stream(stream)))

    def draw_sweep(self, stream, atmosphere: Atmosphere, params: List[DrawParameters], depth, color_out, transmittance_out,
                   width: int, height: int):
        n = len(params)
        raws = (FbDrawParams * n)(*[p.raw() for p in params])
        _check(_lib().fb_renderer_draw_sweep(self._h, atmosphere._h, raws, n, _ptr(depth), _ptr(color_out),
                                             _ptr(transmittance_out), width, height, _stream(stream)))

    def draw_host(self, atmosphere: Atmosphere, params: DrawParameters, depth: np.ndarray):
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        h, w = depth.shape
        color = np.empty((h, w, 4), np.float32)
        transm = np.empty((h, w, 4), np.float32)
        raw = params.raw()
        vp = lambda a: a.ctypes.data_as(c_void_p)
        _check(_lib().fb_renderer_draw_host(self._h, atmosphere._h, byref(raw), vp(depth), vp(color), vp(transm), w, h))
        return color, transm

    def close(self):
        if self._h:
            _lib().fb_renderer_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sky_radiance(atmosphere: Atmosphere, camera, view_ray, sun_direction, n: int, radiance_out, transmittance_out, stream=None):
    """GetSkyRadiance (shaders/render_sky.h:45-109) for n queries; device pointers to [n][3] f32."""
    _check(_lib().fb_sky_radiance(atmosphere._h, _ptr(camera), _ptr(view_ray), _ptr(sun_direction), n, _ptr(radiance_out),
                                  _ptr(transmittance_out), _stream(stream)))


def sun_and_sky_irradiance(atmosphere: Atmosphere, point, normal, sun_direction, n: int, sun_out, sky_out, stream=None):
    """GetSunAndSkyIrradiance (shaders/render_lighting.h:10-28) for n queries."""
    _check(_lib().fb_sun_and_sky_irradiance(atmosphere._h, _ptr(point), _ptr(normal), _ptr(sun_direction), n, _ptr(sun_out),
                                            _ptr(sky_out), _stream(stream)))


def external_memory_read(device: int, fd: int, allocation_bytes: int, offset: int, shape, dtype) -> np.ndarray:
    """Import an exported allocation as a consumer would (here: CUDA is the importer) and copy one table out."""
    out = np.empty(shape, dtype=dtype)
    _check(_lib().fb_external_memory_read_fd(device, fd, allocation_bytes, offset, out.ctypes.data_as(c_void_p), out.nbytes))
    return out
