"""Pins the CPU oracle against the reference's own golden vectors / KATs.

Fixtures in tests/golden/ are re-encodings of
  imageflow_core/tests/integration/weights.txt, weights_params.txt and graphics/lut.rs:14
made by tests/golden/make_golden.py.  The analytic KATs restate
  imageflow_core/tests/integration/color_conversion.rs:375-402, :701-748, :823-1187.
"""
import gzip
import json
import os

import numpy as np
import pytest

import oracle

G = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1.0e-6  # golden weights are printed with 6 decimals


@pytest.fixture(scope="module")
def golden():
    with gzip.open(os.path.join(G, "weights_golden.json.gz"), "rt") as f:
        return json.load(f)


def _cmp(got, exp):
    assert len(got) == len(exp)
    for (l, r, w), e in zip(got, exp):
        assert len(w) == len(e) == r - l + 1
        assert np.max(np.abs(w.astype(np.float64) - np.array(e))) <= TOL


def test_weights_txt_all_rows(golden):
    rows = golden["weights"]
    assert len(rows) == 660
    assert sorted({r["filter"] for r in rows}) == list(range(1, 31))
    for r in rows:
        _cmp(oracle.weights(r["filter"], r["out"], r["in"]), r["w"])


def test_weights_params_all_rows(golden):
    rows = golden["params"]
    errors = 0
    for r in rows:
        if r.get("error"):
            with pytest.raises(oracle.OracleError) as ei:
                oracle.weights(r["filter"], r["out"], r["in"], r["kernel_scale"], r["lobe_mode"], r["lobe_value"])
            assert ei.value.code == 10  # WeightsError::TotalWeightZero (weights.rs:755-757)
            errors += 1
        else:
            _cmp(oracle.weights(r["filter"], r["out"], r["in"], r["kernel_scale"], r["lobe_mode"], r["lobe_value"]), r["w"])
    assert errors == 6 and len(rows) - errors >= 1673


def test_weights_sum_to_one_and_monotone():
    for f in (2, 6, 14, 4, 24, 27):
        for (i, o) in ((3840, 512), (2160, 512), (1080, 2160), (640, 200), (7, 3)):
            ws = oracle.weights(f, o, i)
            lefts = [l for l, _, _ in ws]
            rights = [r for _, r, _ in ws]
            assert lefts == sorted(lefts) and rights == sorted(rights)
            for l, r, w in ws:
                assert 0 <= l <= r < i
                assert abs(float(np.sum(w.astype(np.float64))) - 1.0) < 1e-5


def test_lut16k_matches_reference_table():
    with gzip.open(os.path.join(G, "lut16k_golden.bin.gz"), "rb") as f:
        ref = np.frombuffer(f.read(), np.uint8)
    assert ref.size == 16384
    assert np.array_equal(oracle.linear_to_srgb_table(), ref)      # color_conversion.rs:375-402


def test_srgb_linear_roundtrip_all_256():
    L = oracle.lib()
    t = oracle.byte_to_float_table(True)
    for v in range(256):                                           # color_conversion.rs:701-748
        assert L.ifo_floatspace_to_srgb(1, float(t[v])) == v
    ts = oracle.byte_to_float_table(False)
    for v in range(256):
        assert L.ifo_floatspace_to_srgb(0, float(ts[v])) == v
    assert t[0] == 0.0 and t[255] == 1.0
    # f64 formula agreement of the forward table
    s = np.arange(256) / 255.0
    lin = np.where(s <= 0.04045, s / 12.92, ((s + 0.055) / 1.055) ** 2.4)
    assert np.max(np.abs(t - lin)) < 5e-7   # f32 evaluation steps (color.rs:85-91) vs f64


def test_uchar_clamp_ff_kats():
    f = oracle.lib().ifo_uchar_clamp_ff                               # color.rs:101-108
    cases = [(-1e9, 0), (-1.5, 0), (-0.7, 0), (-0.2, 0), (0.0, 0), (0.49, 0), (0.5, 1), (1.49, 1), (254.5, 255),
             (255.0, 255), (255.4, 255), (255.5, 255), (300.0, 255), (1e9, 255), (float("nan"), 0), (127.5, 128),
             (float(np.float32(0.49999997)), 0), (40000.0, 255), (-40000.0, 0)]
    for v, e in cases:
        assert f(v) == e, (v, e, f(v))


def _solid(w, h, bgra):
    a = np.zeros((h, w, 4), np.uint8)
    a[:, :] = bgra
    return a


def test_matte_no_double_division_kat():
    # color_conversion.rs:823-986: red a=128, 10x10 -> 5x5 over white matte, expected within +-2
    inp = _solid(10, 10, (0, 0, 255, 128))
    out = np.zeros((5, 5, 4), np.uint8)
    oracle.scale_and_render(inp, out, filter=2, alpha_meaningful=True, compose=oracle.BLEND_WITH_MATTE, matte=(255, 255, 255, 255))
    t = oracle.byte_to_float_table(True)
    af = np.float32(128) / np.float32(255)
    L = oracle.lib()
    er = L.ifo_floatspace_to_srgb(1, float(af * t[255] + (1 - af)))
    eg = L.ifo_floatspace_to_srgb(1, float(1 - af))
    assert np.max(np.abs(out[..., 2].astype(int) - er)) <= 2
    assert np.max(np.abs(out[..., 1].astype(int) - eg)) <= 2
    assert np.max(np.abs(out[..., 0].astype(int) - eg)) <= 2
    assert np.min(out[..., 3]) >= 253


def test_matte_fully_transparent_kat():
    # color_conversion.rs:988-1085: transparent 4x4 -> 2x2 over opaque red => (R=255,G=0,B=0,A=255) +-2
    inp = _solid(4, 4, (0, 0, 0, 0))
    out = np.zeros((2, 2, 4), np.uint8)
    oracle.scale_and_render(inp, out, filter=2, alpha_meaningful=True, compose=oracle.BLEND_WITH_MATTE, matte=(0, 0, 255, 255))
    assert np.max(np.abs(out.astype(int) - np.array([0, 0, 255, 255]))) <= 2


def test_matte_mixed_alpha_kat():
    # color_conversion.rs:1087-1187: 40x10 green bands a=[0,85,170,255] -> 20x5 over blue matte
    inp = np.zeros((10, 40, 4), np.uint8)
    alphas = [0, 85, 170, 255]
    for band, a in enumerate(alphas):
        inp[:, band * 10:(band + 1) * 10] = (0, 255, 0, a)
    out = np.zeros((5, 20, 4), np.uint8)
    oracle.scale_and_render(inp, out, filter=2, alpha_meaningful=True, compose=oracle.BLEND_WITH_MATTE, matte=(255, 0, 0, 255))
    t = oracle.byte_to_float_table(True)
    L = oracle.lib()
    for band, a in enumerate(alphas):
        px = out[2, band * 5 + 2]
        af = np.float32(a) / np.float32(255)
        eg = L.ifo_floatspace_to_srgb(1, float(t[255] * af))
        eb = L.ifo_floatspace_to_srgb(1, float(t[255] * (1 - af)))
        assert abs(int(px[2]) - 0) <= 2 and abs(int(px[1]) - eg) <= 2 and abs(int(px[0]) - eb) <= 2 and px[3] >= 253


def test_color_matrix_sepia_and_invert_kats():
    # color_matrix.rs:5-28 with flow/nodes/color.rs:86-94 (sepia) and :159-167 (invert)
    px = np.array([[[10, 100, 200, 255], [255, 255, 255, 255], [0, 0, 0, 7]]], np.uint8)
    p = px.copy()
    oracle.color_matrix(p, oracle.color_filter_matrix(0))
    b, g, r, a = 10.0, 100.0, 200.0, 255.0
    exp_r = int(np.float32(0.393) * r + np.float32(0.769) * g + np.float32(0.189) * b + 0.5)
    exp_g = int(np.float32(0.349) * r + np.float32(0.686) * g + np.float32(0.168) * b + 0.5)
    exp_b = int(np.float32(0.272) * r + np.float32(0.534) * g + np.float32(0.131) * b + 0.5)
    assert tuple(p[0, 0]) == (exp_b, exp_g, exp_r, 255)
    assert tuple(p[0, 1]) == (239, 255, 255, 255)      # white saturates R,G; B = 0.937*255 = 238.9
    assert tuple(p[0, 2]) == (0, 0, 0, 7)
    q = px.copy()
    oracle.color_matrix(q, oracle.color_filter_matrix(5))
    assert tuple(q[0, 0]) == (245, 155, 55, 255) and tuple(q[0, 2]) == (255, 255, 255, 7)


def test_composite_over_canvas_formula():
    # scaling.rs:254-287 evaluated by hand for one pixel (1x1 -> 1x1, so resample is identity on premult floats)
    t = oracle.byte_to_float_table(True)
    L = oracle.lib()
    for src, dst in [((40, 80, 120, 100), (200, 150, 100, 180)), ((1, 2, 3, 254), (9, 9, 9, 9)), ((50, 60, 70, 0), (5, 6, 7, 0))]:
        inp = _solid(1, 1, src)
        cv = _solid(1, 1, dst)
        oracle.scale_and_render(inp, cv, filter=2, alpha_meaningful=True, compose=oracle.BLEND_WITH_SELF)
        sa = np.float32(src[3]) * np.float32(1 / 255)
        if sa > np.float32(0.994):
            exp = [L.ifo_floatspace_to_srgb(1, float(t[src[c]] * sa)) for c in range(3)] + [255]
        else:
            dc = (np.float32(1) - sa) * (np.float32(1 / 255) * np.float32(dst[3]) + np.float32(0))
            fa = sa + dc
            exp = [L.ifo_floatspace_to_srgb(1, float((t[src[c]] * sa + dc * t[dst[c]]) / fa)) for c in range(3)]
            exp.append(L.ifo_uchar_clamp_ff(float(fa * np.float32(255))))
        assert list(cv[0, 0]) == exp, (src, dst, list(cv[0, 0]), exp)


def test_apply_matte_kats():
    # blend.rs:6-59: a=0 -> matte; a=255 untouched; partial -> linear-light blend
    px = np.array([[[1, 2, 3, 0], [4, 5, 6, 255], [0, 0, 255, 128]]], np.uint8)
    oracle.apply_matte(px, (255, 255, 255, 255))
    assert tuple(px[0, 0]) == (255, 255, 255, 255) and tuple(px[0, 1]) == (4, 5, 6, 255)
    assert px[0, 2, 3] == 255 and px[0, 2, 2] == 255 and abs(int(px[0, 2, 0]) - 186) <= 2


def test_replace_self_opaque_forces_alpha_255_and_identity_on_flat():
    inp = _solid(64, 48, (13, 77, 201, 9))
    out = np.zeros((12, 16, 4), np.uint8)
    oracle.scale_and_render(inp, out, filter=2, alpha_meaningful=False)
    assert np.all(out[..., 3] == 255)                               # scaling.rs:227-232
    assert np.max(np.abs(out[..., :3].astype(int) - np.array([13, 77, 201]))) <= 1


def test_error_paths():
    inp = _solid(4, 4, (1, 2, 3, 4))
    cv = np.zeros((4, 4, 4), np.uint8)
    with pytest.raises(oracle.OracleError) as e:
        oracle.scale_and_render(inp, cv, x=2, y=0, w=3, h=4)        # scaling.rs:24-29
    assert e.value.code == 1
    with pytest.raises(oracle.OracleError) as e:
        oracle.scale_and_render(inp, cv, filter=99)
    assert e.value.code == 13
