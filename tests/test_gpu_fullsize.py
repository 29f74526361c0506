"""GPU suite at BASELINE.json's full sizes.

  * c2 (1920x1080 default scene, depth 8) and c3 (3840x2160 synthetic 10k scene, depth 10): EVERY row of the CUDA
    frame is compared — via per-row CRC32 of the RGBA8 words, of the ray counts and of the primary object ids,
    and per-row ray totals — with the statistics the unmodified reference produced on the CPU
    (tests/golden/fullsize_*.json, written by tests/golden/make_fullsize.py; the 4K frame is ~1.9e11 object
    tests and takes the reference about 15 minutes on 8 cores).
  * c4 (7680x4320): every 16th 4-row band against the reference, plus size-independent properties on the whole
    frame: the union of 8 ranks' cyclic bands equals the single-GPU frame, ray totals add up, the result is
    idempotent, fused and unfused quantise agree.
"""
import zlib

import numpy as np
import pytest

from conftest import load_json

pytestmark = pytest.mark.gpu


def crc_rows(a):
    return [zlib.crc32(np.ascontiguousarray(row).tobytes()) for row in a]


def check_against_fullsize(planes, st, g, rows):
    rgba, rc, ids = planes["rgba8"][0][rows], planes["ray_count"][0][rows], planes["object_id"][0][rows]
    bad_ids = [r for r, a, b in zip(rows, crc_rows(ids), g["object_id_crc"]) if a != b]
    bad_rc = [r for r, a, b in zip(rows, crc_rows(rc), g["ray_count_crc"]) if a != b]
    bad_px = [r for r, a, b in zip(rows, crc_rows(rgba), g["rgba8_crc"]) if a != b]
    assert not bad_ids, "object-id rows differ: %s" % bad_ids[:10]
    assert not bad_rc, "ray-count rows differ: %s" % bad_rc[:10]
    assert not bad_px, "RGBA8 rows differ: %s" % bad_px[:10]
    assert [int(x) for x in rc.astype(np.int64).sum(axis=1)] == g["row_rays"]


def test_c2_full_frame_matches_reference(gpu, renderer_mod, S):
    g = load_json("fullsize_c2.json")
    gpu.set_scene(S.default_scene())
    planes, st = gpu.render([S.default_camera(1920, 16.0 / 9.0).pod()], renderer_mod.default_params(max_depth=8),
                            want=("rgba8", "ray_count", "object_id"))
    assert planes["rgba8"].shape == (1, 1080, 1920)
    check_against_fullsize(planes, st, g, np.array(g["rows"]))
    assert st.total_rays == g["total_rays"] == 2293320 and st.over_range_pixels == 0


def test_c3_full_4k_frame_matches_reference(gpu, renderer_mod, S):
    g = load_json("fullsize_c3.json")
    gpu.set_scene(S.synthetic_scene())
    planes, st = gpu.render([S.default_camera(3840, 16.0 / 9.0).pod()], renderer_mod.default_params(max_depth=10),
                            want=("rgba8", "ray_count", "object_id"))
    assert planes["rgba8"].shape == (1, 2160, 3840)
    check_against_fullsize(planes, st, g, np.array(g["rows"]))
    assert st.total_rays == g["total_rays"]
    assert st.over_range_pixels == sum(g["row_over_range"])
    # the scheduling hint (rtx_params.pixel_order): scan order, then twice in the order of the previous frame's costs —
    # the same frame against the same reference CRCs every time
    for order in (renderer_mod.abi.RTX_ORDER_SCAN, renderer_mod.abi.RTX_ORDER_COST, renderer_mod.abi.RTX_ORDER_COST):
        planes, st = gpu.render([S.default_camera(3840, 16.0 / 9.0).pod()], renderer_mod.default_params(max_depth=10, pixel_order=order),
                                want=("rgba8", "ray_count", "object_id"))
        check_against_fullsize(planes, st, g, np.array(g["rows"]))
        assert st.total_rays == g["total_rays"] and st.over_range_pixels == sum(g["row_over_range"])


def test_c4_8k_bands_and_properties(gpu, renderer_mod, S):
    g = load_json("fullsize_c4.json")
    gpu.set_scene(S.synthetic_scene())
    pod = S.default_camera(7680, 16.0 / 9.0).pod()
    want = ("rgba8", "ray_count", "object_id")
    full, st = gpu.render([pod], renderer_mod.default_params(max_depth=10), want=want)
    assert full["rgba8"].shape == (1, 4320, 7680)
    check_against_fullsize(full, st, g, np.array(g["rows"]))
    # idempotence: a second render is bit-identical (dynamic pixel scheduling must not leak into results)
    again, st2 = gpu.render([pod], renderer_mod.default_params(max_depth=10), want=("rgba8",))
    assert np.array_equal(again["rgba8"], full["rgba8"]) and st2.total_rays == st.total_rays
    # 8-rank cyclic row bands: the union equals the single-GPU frame, ray totals add up
    total = 0
    for r in range(8):
        rows = renderer_mod.global_rows(4320, 4, 8, r)
        part, sr = gpu.render([pod], renderer_mod.default_params(max_depth=10, band_rows=4, n_ranks=8, rank=r), want=want)
        assert part["rgba8"].shape == (1, 540, 7680)
        for k in want:
            assert np.array_equal(part[k][0], full[k][0][rows]), (k, r)
        total += sr.total_rays
    assert total == st.total_rays
    # the separate quantise kernel on the double radiance gives the fused result
    unf, su = gpu.render([pod], renderer_mod.default_params(max_depth=10, fuse_quantise=0, pixel_order=renderer_mod.abi.RTX_ORDER_SCAN), want=("rgba8",))
    assert np.array_equal(unf["rgba8"], full["rgba8"]) and su.launches == 2


def test_c5_full_1080p_orbit_frames_match_reference(gpu, renderer_mod, S):
    """Config C5 at full size: frames 48 and 224 of the orbit see the walls from BEHIND (the pass-through of
    main.cpp:111-113: two depth levels spent, the wall shaded twice), frame 128 is the far side. Every row of the
    three frames against the unmodified reference's per-row CRCs, rendered in one batched call like the orbit itself."""
    g = load_json("fullsize_c5.json")
    cams = S.flythrough_cameras(g["n_frames"], g["width"], 16.0 / 9.0)
    ks = sorted(int(k) for k in g["frames"])
    gpu.set_scene(S.default_scene())
    planes, st = gpu.render([cams[k].pod() for k in ks], renderer_mod.default_params(max_depth=g["depth"]),
                            want=("rgba8", "ray_count", "object_id"))
    assert planes["rgba8"].shape == (len(ks), 1080, 1920)
    total = 0
    for n, k in enumerate(ks):
        gf = g["frames"][str(k)]
        one = {name: a[n:n + 1] for name, a in planes.items()}
        check_against_fullsize(one, st, gf, np.array(gf["rows"]))
        total += gf["total_rays"]
        # back faces really occur in frames 48 and 224: a primary wall hit followed by the same wall again
        if k in (48, 224):
            ids, rc = planes["object_id"][n], planes["ray_count"][n]
            assert ((ids >= 1) & (rc >= 3)).any()
    assert st.total_rays == total and st.over_range_pixels == 0
