"""GPU tests of the C-ABI object model: batches, error codes, lifetimes, host-buffer entry points."""
import ctypes
import os

import numpy as np
import pytest

import fuzzyblue_b200 as fb
from fuzzyblue_b200 import api, synthetic

from .conftest import SMOKE_DIMS

pytestmark = pytest.mark.gpu


def sync():
    import torch
    torch.cuda.synchronize()


def test_batch_equals_individual_builds():
    """BASELINE.json config 4 in miniature: distinct atmospheres (Earth-to-Mars radii, random Rayleigh/Mie/ozone) built as
    one batch on forked streams give bit-identical tables to one-at-a-time builds."""
    b = fb.Builder(0)
    params = [fb.Parameters(**SMOKE_DIMS)] + synthetic.random_atmospheres(4, base=fb.Parameters(**SMOKE_DIMS), seed=7)
    batch = fb.build_batch(b, params, None)
    sync()
    for p, pend in zip(params, batch):
        T, S, E = fb.precompute_host(b, p)
        a = pend.assert_ready()
        assert np.array_equal(a.read_scattering(), S) and np.array_equal(a.read_irradiance(), E)
        assert np.array_equal(a.read_transmittance(), T)
        assert np.all(np.isfinite(S.astype(np.float32))) and S.astype(np.float32).max() > 0
    # different physics really gives different tables
    assert not np.array_equal(fb.precompute_host(b, params[1])[1], fb.precompute_host(b, params[2])[1])


def test_error_codes_instead_of_panics():
    b = fb.Builder(0)
    with pytest.raises(fb.FuzzyblueError) as e:
        fb.Atmosphere.build(b, None, fb.Parameters(order=0, **SMOKE_DIMS))
    assert e.value.status == 1
    with pytest.raises(fb.FuzzyblueError) as e:
        fb.Atmosphere.build(b, None, fb.Parameters(scattering_nu_size=1))
    assert e.value.status == 1
    with pytest.raises(fb.FuzzyblueError):
        fb.Builder(1000)
    pend = fb.Atmosphere.allocate(b, fb.Parameters(**SMOKE_DIMS))
    with pytest.raises(fb.FuzzyblueError):
        pend.run_stage(api.STAGE_SCATTERING_DENSITY, order=1)          # scattering_density.comp:22 asserts order >= 2
    with pytest.raises(fb.FuzzyblueError):
        pend.run_stage(api.STAGE_SINGLE_SCATTERING, r_begin=5, r_end=3)
    with pytest.raises(fb.FuzzyblueError):
        api._check(api._lib().fb_pending_upload(pend._h, api.IMAGE_TRANSMITTANCE, ctypes.c_void_p(1), 7, None))
    assert api._lib().fb_last_error()


def test_pending_outlives_and_temporaries_are_freed():
    import torch
    b = fb.Builder(0)
    free0 = torch.cuda.mem_get_info()[0]
    pend = fb.Atmosphere.build(b, None, fb.Parameters())
    sync()
    borrowed = pend.atmosphere().read_irradiance()
    atm = pend.assert_ready()                        # releases the 5 temporaries (4 x 8 MiB + scratch) into the builder's cache
    b.trim()                                         # ... and this returns the cache to the driver
    kept = free0 - torch.cuda.mem_get_info()[0]
    assert np.array_equal(atm.read_irradiance(), borrowed)
    assert kept <= 16 << 20, kept                    # 8 MiB scattering + 256 KiB + 16 KiB, allocator granularity
    atm.close()
    b.trim()
    assert free0 - torch.cuda.mem_get_info()[0] <= 2 << 20
    # a second build reuses the cached blocks of the first: same tables, no fresh allocation
    p1 = fb.Atmosphere.build(b, None, fb.Parameters()); sync(); a1 = p1.assert_ready(); s1 = a1.read_scattering(); a1.close()
    held = torch.cuda.mem_get_info()[0]
    p2 = fb.Atmosphere.build(b, None, fb.Parameters()); sync()
    assert torch.cuda.mem_get_info()[0] == held
    assert np.array_equal(p2.assert_ready().read_scattering(), s1)


def test_non_multiple_of_workgroup_dims_are_fully_written():
    """The reference's truncating dispatch (precompute.rs:1740-1745) leaves edge texels unwritten when a size is not a
    multiple of the workgroup; here every texel is computed."""
    dims = dict(transmittance_mu_size=37, transmittance_r_size=11, irradiance_mu_s_size=13, irradiance_r_size=5,
                scattering_r_size=3, scattering_mu_size=10, scattering_mu_s_size=6, scattering_nu_size=2)
    for kernels in (api.KERNELS_FAST, api.KERNELS_REFERENCE):
        T, S, E = fb.precompute_host(fb.Builder(0, kernels=kernels), fb.Parameters(order=3, **dims))
        assert np.all(T[..., :3] > 0) and np.all(T[..., 3] == 1)
        assert np.all(np.isfinite(S.astype(np.float32))) and (S.astype(np.float32)[..., :3] > 0).mean() > 0.5
        assert np.all(np.isfinite(E)) and E.max() > 0
    from oracle import oracle as O
    # second set: nu_size = 3 is not a power of two (a mu_s tile of the density kernel is then found by division)
    for d in (dims, dict(dims, scattering_nu_size=3, scattering_mu_s_size=5)):
        ref = O.precompute(O.Params(order=3, **d), O.F32)
        Tf, Sf, Ef = fb.precompute_host(fb.Builder(0), fb.Parameters(order=3, **d))
        assert np.max(np.abs(Tf - ref.transmittance) / ref.transmittance) <= 1e-3
        assert np.max(np.abs(Ef - ref.irradiance) / np.maximum(ref.irradiance, 1e-30)) <= 1e-3
        e = np.abs(Sf.astype(np.float64) - ref.scattering) / np.maximum(np.abs(ref.scattering), 2.0 ** -14)
        assert (e > 1e-3).sum() <= 8 and e.max() <= 2e-2, (d, e.max(), int((e > 1e-3).sum()))   # the end-to-end gate of test_parity_gpu.check_compounded


def test_exported_allocation_carries_the_tables():
    """Vulkan interop: the kept block as an fd-exportable allocation.  With CUDA standing in as the
    importer (no Vulkan loader in this image): export, import the descriptor afresh, map it, and the three tables read
    through the import equal the tables read through the Atmosphere."""
    import os
    b = fb.Builder(0)
    b.set_exportable(True)
    p = fb.Parameters(scattering_r_size=8, scattering_mu_size=32, scattering_mu_s_size=8, scattering_nu_size=2)
    pend = fb.Atmosphere.build(b, None, p)
    sync()
    atm = pend.assert_ready()
    fd, lay = atm.export_fd()
    try:
        assert fd >= 0 and lay.allocation_bytes >= lay.irradiance_offset + lay.irradiance_bytes
        T, S, E = atm.read_transmittance(), atm.read_scattering(), atm.read_irradiance()
        assert lay.scattering_bytes == S.nbytes and lay.transmittance_bytes == T.nbytes and lay.irradiance_bytes == E.nbytes
        S2 = api.external_memory_read(0, fd, lay.allocation_bytes, lay.scattering_offset, S.shape, np.float16)
        T2 = api.external_memory_read(0, fd, lay.allocation_bytes, lay.transmittance_offset, T.shape, np.float32)
        E2 = api.external_memory_read(0, fd, lay.allocation_bytes, lay.irradiance_offset, E.shape, np.float32)
        assert np.array_equal(S, S2) and np.array_equal(T, T2) and np.array_equal(E, E2)
        assert float(S.astype(np.float32).max()) > 0 and float(T.max()) > 0
    finally:
        os.close(fd)
    atm.close()
    # the same builder, not exportable: an export request is an error, not a silent copy
    b.set_exportable(False)
    pend = fb.Atmosphere.build(b, None, p)
    sync()
    atm = pend.assert_ready()
    with pytest.raises(fb.FuzzyblueError):
        atm.export_fd()
    atm.close()


def test_assert_ready_reports_not_ready_and_waits():
    """fb_pending_assert_ready(check = 1) queries the completion events of the precompute: FB_ERR_NOT_READY while kernels
    are in flight (the pending stays usable), success after fb_pending_wait -- no device-wide synchronisation."""
    import torch
    b = fb.Builder(0)
    s = torch.cuda.Stream()
    pend = fb.Atmosphere.build(b, s, fb.Parameters())
    for _ in range(20):
        pend.resubmit(s)                       # ~60 ms of queued device work
    with pytest.raises(fb.FuzzyblueError) as ei:
        pend.assert_ready(check=True, wait=False)
    assert ei.value.status == 5                # FB_ERR_NOT_READY
    assert pend.launch_count() > 0             # still a valid object
    pend.wait()
    atm = pend.assert_ready(check=True, wait=False)
    assert np.all(np.isfinite(atm.read_irradiance()))


def test_block_reuse_is_stream_ordered():
    """a released block may still be written by kernels in flight.  Drop a pending early
    (check = 0, work queued on its stream), immediately build another atmosphere of the same dims on ANOTHER stream -- it
    receives the cached blocks -- and its tables must equal an undisturbed build bit for bit."""
    import torch
    b = fb.Builder(0)
    p = fb.Parameters(**SMOKE_DIMS)
    want = fb.precompute_host(b, p)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(4):
        victim = fb.Atmosphere.build(b, s1, fb.Parameters(order=6, **SMOKE_DIMS))
        for _ in range(10):
            victim.resubmit(s1)
        a0 = victim.assert_ready(check=False)  # temporaries go back to the cache with kernels still queued
        a0.close()                             # ... and so does the kept block
        pend = fb.Atmosphere.build(b, s2, p)   # same dims: served from the cache
        pend.wait()
        atm = pend.assert_ready()
        assert np.array_equal(atm.read_scattering().view(np.uint16), want[1].view(np.uint16))
        assert np.array_equal(atm.read_irradiance(), want[2])
        atm.close()


def test_pending_outlives_its_builder():
    """FbPending copied what it needs from the Builder; destroying the builder first is legal."""
    import torch
    b = fb.Builder(0)
    pend = fb.Atmosphere.build(b, None, fb.Parameters(**SMOKE_DIMS))
    api._lib().fb_builder_destroy(b._h)
    b._h = None
    pend.resubmit(None)
    pend.run_stage(api.STAGE_TRANSMITTANCE)
    torch.cuda.synchronize()
    assert np.all(np.isfinite(pend.download(api.IMAGE_SCATTERING).astype(np.float32)))
    pend.close()


@pytest.mark.parametrize("nu,mu_s", [(3, 5), (6, 12), (12, 8), (5, 32)])
def test_non_power_of_two_nu_runs_the_fast_kernels(nu, mu_s):
    """nu_size need not be a power of two for the restructured density kernel (a mu_s tile is T / nu
    columns, found by division); the result matches the oracle and NO stage falls back to the transcription."""
    from oracle import oracle as O
    d = dict(scattering_r_size=4, scattering_mu_size=8, scattering_mu_s_size=mu_s, scattering_nu_size=nu)
    ref = O.precompute(O.Params(order=3, **d), O.F32, keep_history=True)
    b = fb.Builder(0)
    pend = fb.Atmosphere.allocate(b, fb.Parameters(order=3, **d))
    assert api._lib().fb_params_slow_stages(ctypes.byref(fb.Parameters(order=3, **d).raw())) == 0
    for order in (2, 3):
        prev = ref.history["single"] if order == 2 else ref.history[order - 1]
        pend.upload(api.IMAGE_TRANSMITTANCE, ref.transmittance)
        pend.upload(api.IMAGE_DELTA_RAYLEIGH, ref.delta_rayleigh)
        pend.upload(api.IMAGE_DELTA_MIE, ref.delta_mie)
        pend.upload(api.IMAGE_DELTA_IRRADIANCE, prev["delta_irradiance"])
        if order > 2:
            pend.upload(api.IMAGE_DELTA_MULTIPLE_SCATTERING, prev["delta_multiple_scattering"])
        n0 = pend.launch_count()
        pend.run_stage(api.STAGE_SCATTERING_DENSITY, order=order)
        assert pend.launch_count() - n0 == 2               # k_density_prep + k_density_main, not the 1-launch transcription
        got = pend.download(api.IMAGE_SCATTERING_DENSITY).astype(np.float64)
        want = ref.history[order]["scattering_density"]
        e = np.abs(got - want) / np.maximum(np.abs(want), 2.0 ** -14)
        assert e.max() <= 1e-3, (nu, mu_s, order, e.max())
    assert pend.slow_stages() == 0


def test_slow_stage_cliffs_are_reported():
    """Dims the restructured kernels do not cover run the transcription -- and say so (fb_pending_slow_stages)."""
    d = dict(scattering_r_size=2, scattering_mu_size=4, scattering_mu_s_size=2, scattering_nu_size=2, irradiance_mu_s_size=520,
             irradiance_r_size=4)
    p = fb.Parameters(order=2, **d)
    assert api._lib().fb_params_slow_stages(ctypes.byref(p.raw())) == 1 << api.STAGE_SCATTERING_DENSITY
    pend = fb.Atmosphere.build(fb.Builder(0), None, p)
    pend.wait()
    assert pend.slow_stages() == 1 << api.STAGE_SCATTERING_DENSITY
    assert fb.Atmosphere.build(fb.Builder(0), None, fb.Parameters(**SMOKE_DIMS)).slow_stages() == 0


def test_external_semaphore_wrappers_reject_what_is_not_a_semaphore():
    """fb_external_semaphore_*  wrap cudaImportExternalSemaphore / Signal / Wait for the timeline
    semaphore a Vulkan caller exports.  No Vulkan loader exists in this image, so what can be executed is the error
    contract: a descriptor that is not a semaphore is refused with a status (no handle comes back, nothing crashes),
    NULL handles are invalid arguments, destroy(NULL) is a no-op.  Run in a child process so that a driver that reacts
    badly to the bogus descriptor cannot take the other tests' context with it."""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import os, ctypes
        from ctypes import byref, c_void_p
        import fuzzyblue_b200 as fb
        from fuzzyblue_b200 import api
        lib = api._lib()
        b = fb.Builder(0)
        h = c_void_p()
        INVALID = 1                                   # FB_ERR_INVALID_ARGUMENT, include/fuzzyblue.h
        assert lib.fb_external_semaphore_import_fd(0, -1, 1, byref(h)) == INVALID and not h.value
        r, w = os.pipe()
        st = lib.fb_external_semaphore_import_fd(0, r, 1, byref(h))
        assert st != api.FB_OK and not h.value, st
        assert lib.fb_last_error() is not None
        os.close(w)
        try: os.close(r)
        except OSError: pass
        assert lib.fb_external_semaphore_signal(None, 1, None) == INVALID
        assert lib.fb_external_semaphore_wait(None, 1, None) == INVALID
        lib.fb_external_semaphore_destroy(None)
        # the context still works
        T, S, E = fb.precompute_host(b, fb.Parameters(scattering_r_size=2, scattering_mu_size=4, scattering_mu_s_size=4, scattering_nu_size=2, order=2))
        assert float(T.max()) > 0
        print("SEMAPHORE-CONTRACT-OK")
    """)
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "SEMAPHORE-CONTRACT-OK" in out.stdout, (out.stdout[-500:], out.stderr[-1500:])
