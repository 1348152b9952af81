"""bench.py's pure-Python helpers (no GPU): the render roofline arithmetic and the constants of the bench line."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("fb_bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_render_roofline_fractions():
    b = _bench()
    px = 4 * 3840 * 2160
    geo = px // 4
    r = b.render_roofline(px, geo, 1.0, 70.0, 4600.0)
    flops = geo * 430 + (px - geo) * 290
    sfu = geo * 39 + (px - geo) * 29
    assert abs(r["fp32"]["achieved_tflops"] - flops / 1e-3 / 1e12) < 1e-9
    assert abs(r["sfu"]["achieved_gops"] - sfu / 1e-3 / 1e9) < 1e-6
    assert r["frac"] == max(r["fp32"]["frac"], r["sfu"]["frac"])
    assert r["bound"] in ("fp32", "sfu")
    assert abs(r["geometry_pixel_share"] - 0.25) < 1e-12


def test_render_roofline_without_peaks_is_null():
    b = _bench()
    assert b.render_roofline(10, 0, 1.0, None, None) is None
    assert b.render_roofline(10, 0, 0.0, 70.0, 4600.0) is None


def test_density_tallies_match_survey():
    """5.37e8 density samples per launch at default dims; hoisted-minimal 92 / 63 flop per sample."""
    b = _bench()
    assert b.DENSITY_SAMPLES == 32 * 128 * 256 * 512
    assert b.FLOP_HOISTED[2] == 92 and b.FLOP_HOISTED[3] == 63
    assert b.FLOP_AS_WRITTEN[2] == 223 and b.FLOP_AS_WRITTEN[3] == 184
