"""Property tests of the library's host-only entry points (no GPU is touched): whatever parameter block, rank and world
size a caller hands in, `fb_params_validate`, the extent queries, `fb_params_slow_stages` and `fb_sharded_plan` answer
with a status -- never a crash, never a plan for a block that validation rejects -- and every plan they do return has the
properties the executors rely on (the ranks' slabs tile the r axis, every rank issues the same exchanges in the same
order).  The reference has no error path at all here (every failure is an `.unwrap()` panic)."""
import ctypes
import math
from ctypes import byref, c_uint32

from hypothesis import HealthCheck, assume, given, settings, strategies as st

import fuzzyblue_b200 as fb
from fuzzyblue_b200 import api, sharded

dim = st.one_of(st.integers(-3, 40), st.sampled_from([64, 128, 256, 511, 512, 513, 1024, 4096, 2 ** 20, 2 ** 31 - 1]))
weird_float = st.one_of(st.floats(allow_nan=True, allow_infinity=True, width=32), st.sampled_from([0.0, -0.0, 1e-30, 6360.0, 6420.0]))


def block(d, floats):
    p = fb.Parameters().raw()
    (p.transmittance_mu_size, p.transmittance_r_size, p.scattering_r_size, p.scattering_mu_size, p.scattering_mu_s_size,
     p.scattering_nu_size, p.irradiance_mu_s_size, p.irradiance_r_size) = [ctypes.c_int32(v & 0xffffffff).value if v > 2 ** 31 - 1 else v for v in d]
    if floats is not None:
        p.bottom_radius, p.top_radius, p.mu_s_min, p.mie_phase_function_g, p.sun_angular_radius = floats
    return p


@settings(max_examples=300, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.lists(dim, min_size=8, max_size=8), st.one_of(st.none(), st.lists(weird_float, min_size=5, max_size=5)),
       st.integers(-2, 70), st.integers(-2, 70), st.integers(0, 12), st.integers(0, 7))
def test_host_entry_points_answer_with_a_status(d, floats, rank, world, order, flags):
    lib = api._lib()
    p = block(d, floats)
    ok = lib.fb_params_validate(byref(p))
    assert ok in (api.FB_OK, 1)                                   # FB_OK or FB_ERR_INVALID_ARGUMENT
    ext = api.FbExtent3D()
    assert lib.fb_params_scattering_extent(byref(p), byref(ext)) in (api.FB_OK, 1)
    e2 = api.FbExtent2D()
    assert lib.fb_params_transmittance_extent(byref(p), byref(e2)) in (api.FB_OK, 1)
    assert lib.fb_params_irradiance_extent(byref(p), byref(e2)) in (api.FB_OK, 1)
    lib.fb_params_slow_stages(byref(p))                           # a bit mask; must simply return
    n = c_uint32(0)
    s = lib.fb_sharded_plan(byref(p), order, rank, world, flags, None, 0, byref(n))
    if s == api.FB_OK:
        assert ok == api.FB_OK and 0 <= rank < world and order >= 1 and p.scattering_r_size % world == 0
        assert all(math.isfinite(v) for v in (p.bottom_radius, p.top_radius, p.mu_s_min))
        steps = (sharded.FbShardStep * n.value)()
        assert lib.fb_sharded_plan(byref(p), order, rank, world, flags, steps, n.value, byref(n)) == api.FB_OK
        assert all(0 <= t.op <= 4 for t in steps)
    else:
        assert s == 1


@settings(max_examples=60, deadline=None)
@given(st.sampled_from([1, 2, 4, 8, 16]), st.integers(1, 4), st.sampled_from([2, 3, 8, 32]), st.integers(2, 9), st.booleans())
def test_every_accepted_plan_tiles_the_r_axis_and_agrees_on_its_exchanges(world, levels, nu, order, pipelined):
    assume(world * levels >= 2)                                   # every LUT size is at least 2 (scattering.h:120-131)
    p = fb.Parameters(order=order, scattering_r_size=world * levels, scattering_mu_size=8, scattering_mu_s_size=4, scattering_nu_size=nu,
                      transmittance_mu_size=16, transmittance_r_size=8, irradiance_mu_s_size=8, irradiance_r_size=5)
    flags = sharded.GATHER_RESULT | (sharded.PIPELINE_ALWAYS if pipelined else 0)
    plans = [sharded.plan(p, r, world, flags) for r in range(world)]
    key = lambda s: (s.op, s.stage, s.image, s.order, s.begin, s.end, s.root)
    exch = [[key(s) for s in pl if s.op not in (sharded.SHARD_STAGE, sharded.SHARD_JOIN)] for pl in plans]
    assert all(e == exch[0] for e in exch)
    for stage, stage_order in ((api.STAGE_SINGLE_SCATTERING, None), (api.STAGE_SCATTERING_DENSITY, 2)):
        cover = sorted((s.begin, s.end) for pl in plans for s in pl
                       if s.op == sharded.SHARD_STAGE and s.stage == stage and (stage_order is None or s.order == stage_order))
        merged = [cover[0]]
        for a, b in cover[1:]:
            assert a == merged[-1][1], (stage, cover)                 # contiguous, no overlap, no gap
            merged.append((a, b))
        assert merged[0][0] == 0 and merged[-1][1] == p.scattering_r_size
    rows = sorted((s.begin, s.end) for pl in plans for s in pl
                  if s.op == sharded.SHARD_STAGE and s.stage == api.STAGE_INDIRECT_IRRADIANCE and s.order == 1)
    assert rows[0][0] == 0 and rows[-1][1] == p.irradiance_r_size and all(rows[i][1] == rows[i + 1][0] for i in range(len(rows) - 1))


@settings(max_examples=400, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.binary(min_size=320, max_size=320), st.integers(0, 9), st.integers(0, 8), st.integers(1, 8))
def test_any_320_bytes_are_answered_with_a_status(blob, order, rank, world):
    """The parameter block is plain bytes from the caller (a uniform buffer in the reference): ANY 320 bytes get a status
    from every host-only entry point, and a block that validates has sizes in range and finite geometry."""
    lib = api._lib()
    p = api.FbParams()
    assert ctypes.sizeof(p) == 320
    ctypes.memmove(byref(p), blob, 320)
    ok = lib.fb_params_validate(byref(p))
    assert ok in (api.FB_OK, 1)
    ext, e2, n = api.FbExtent3D(), api.FbExtent2D(), c_uint32(0)
    for st_ in (lib.fb_params_scattering_extent(byref(p), byref(ext)), lib.fb_params_transmittance_extent(byref(p), byref(e2)),
                lib.fb_params_irradiance_extent(byref(p), byref(e2)), lib.fb_sharded_plan(byref(p), order, rank, world, 1, None, 0, byref(n))):
        assert st_ in (api.FB_OK, 1)
    lib.fb_params_slow_stages(byref(p))
    if ok == api.FB_OK:
        sizes = (p.transmittance_mu_size, p.transmittance_r_size, p.scattering_r_size, p.scattering_mu_size, p.scattering_mu_s_size,
                 p.scattering_nu_size, p.irradiance_mu_s_size, p.irradiance_r_size)
        assert all(2 <= v <= 16384 for v in sizes)
        assert math.isfinite(p.bottom_radius) and math.isfinite(p.top_radius) and 0 < p.bottom_radius < p.top_radius
        assert -1.0 <= p.mu_s_min <= 0.0 and abs(p.mie_phase_function_g) < 1.0


def test_a_valid_block_with_one_poisoned_float_is_rejected():
    """Every float of the block, one at a time, as NaN and as +inf: fb_params_validate must refuse (the restructured
    kernels address shared memory from geometry without per-sample clamps)."""
    lib = api._lib()
    base = fb.Parameters().raw()
    assert lib.fb_params_validate(byref(base)) == api.FB_OK
    int_fields = {"transmittance_mu_size", "transmittance_r_size", "scattering_r_size", "scattering_mu_size", "scattering_mu_s_size",
                  "scattering_nu_size", "irradiance_mu_s_size", "irradiance_r_size"}
    raw = bytes(base)
    ints = set()
    for name, _ in api.FbParams._fields_:
        if name in int_fields:
            off = getattr(api.FbParams, name).offset
            ints.update(range(off, off + 4))
    import struct
    refused = 0
    for off in range(0, 320, 4):
        if off in ints:
            continue
        if struct.unpack_from("<f", raw, off)[0] == 0.0 and off not in {getattr(api.FbParams, n).offset for n, _ in api.FbParams._fields_}:
            pass                                                  # padding words are zero in a default block; poisoning them must not matter
        for bad in (float("nan"), float("inf")):
            blob = bytearray(raw)
            struct.pack_into("<f", blob, off, bad)
            p = api.FbParams()
            ctypes.memmove(byref(p), bytes(blob), 320)
            refused += lib.fb_params_validate(byref(p)) != api.FB_OK
    assert refused >= 2 * 40                                       # the block holds more than 40 meaningful floats
