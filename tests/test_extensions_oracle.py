"""CPU suite, part 3: the EXTENSIONS (RTX_BOX, sun, Reinhard tone map) — features the reference's README names but whose
code is not in the snapshot. Parity with the reference is therefore unpinned; what CAN be pinned is pinned here:

* a box's four side faces must behave exactly like reference Walls laid over them (same t, same normal) — checked against
  the oracle's Wall and, where oracle/_ref exists, against the unmodified reference's Wall::intersect;
* the z faces (which no reference Wall can represent: its basis is NaN for normals along z, scene.cpp:18) against closed forms;
* the sun and the tone-map specifications in oracle/oracle.c against independent numpy restatements.
"""
import numpy as np
import pytest


def _box_side_walls(S, mat, p, s):
    """The four reference Walls that cover the side faces of Box(p, s): with the reference's basis
    right = normalize(cross(n, z)), up = normalize(cross(right, n)) the rectangle extends from `position` along right/up."""
    (px, py, pz), (sx, sy, sz) = p, s
    return [
        S.Wall(mat, (px, py, pz), (-1, 0, 0), sy, sz),            # -x: right = +y, up = +z
        S.Wall(mat, (px + sx, py + sy, pz), (1, 0, 0), sy, sz),   # +x: right = -y, up = +z
        S.Wall(mat, (px + sx, py, pz), (0, -1, 0), sx, sz),       # -y: right = -x, up = +z
        S.Wall(mat, (px, py + sy, pz), (0, 1, 0), sx, sz),        # +y: right = +x, up = +z
    ]


def _first_hit(oracle, objs, o, d):
    best = (np.inf, None)
    for g in objs:
        dist, nrm, _ = oracle.intersect(g, o, d)
        if dist > 0 and dist < best[0]:
            best = (dist, nrm)
    return best


@pytest.mark.parametrize("which", ["port", "ref"])
def test_box_side_faces_are_reference_walls(request, S, port, which):
    walls_oracle = port if which == "port" else request.getfixturevalue("ref")
    rng = np.random.default_rng(11)
    mat = S.Material((0.3, 0.6, 0.9), 0.4)
    checked = 0
    for _ in range(300):
        p = tuple(rng.uniform(-5, 5, 3))
        s = tuple(rng.uniform(0.2, 4, 3))
        box = S.Box(mat, p, s)
        walls = _box_side_walls(S, mat, p, s)
        # aim at an interior point of a random side face from a random origin at mid height (so no z face is nearer)
        f = int(rng.integers(0, 4))
        uv = rng.uniform(0.05, 0.95, 2)
        face_pt = {0: (p[0], p[1] + uv[0] * s[1], p[2] + uv[1] * s[2]), 1: (p[0] + s[0], p[1] + uv[0] * s[1], p[2] + uv[1] * s[2]),
                   2: (p[0] + uv[0] * s[0], p[1], p[2] + uv[1] * s[2]), 3: (p[0] + uv[0] * s[0], p[1] + s[1], p[2] + uv[1] * s[2])}[f]
        o = (float(rng.uniform(-20, 20)), float(rng.uniform(-20, 20)), p[2] + 0.5 * s[2])
        d = tuple(np.subtract(face_pt, o) * rng.uniform(0.3, 3.0))          # unnormalised, like primary rays
        bt, bn, bhit = port.intersect(box, o, d)
        wt, wn = _first_hit(walls_oracle, walls, o, d)
        # the ray may enter through the z range of another side face first; either way box == nearest of the 4 walls,
        # unless a z face is nearer (then the box's t is smaller than every wall's)
        if bn[2] != 0:
            assert bt < wt
            continue
        assert bhit and bt == wt and tuple(bn) == tuple(wn), (p, s, o, d)
        checked += 1
    assert checked > 250


def test_box_z_faces_closed_form_and_inside_rays(S, port):
    mat = S.Material((1, 1, 1), 0.0)
    box = S.Box(mat, (1.0, -2.0, 0.5), (2.0, 3.0, 1.5))
    # straight down onto the top face (z = 2.0) from z = 10 with an unnormalised direction
    t, n, hit = port.intersect(box, (2.0, 0.0, 10.0), (0.0, 0.0, -4.0))
    assert hit and t == (2.0 - 10.0) / -4.0 and tuple(n) == (0.0, 0.0, 1.0)
    # straight up onto the bottom face: outward normal -z
    t, n, hit = port.intersect(box, (2.0, 0.0, -3.5), (0.0, 0.0, 2.0))
    assert hit and t == (0.5 + 3.5) / 2.0 and n[2] == -1.0
    # from inside: the far face is hit from behind with its OUTWARD normal (like a Wall's back face)
    t, n, hit = port.intersect(box, (2.0, 0.0, 1.0), (1.0, 0.0, 0.0))
    assert hit and t == 1.0 and n[0] == 1.0
    # misses
    assert port.intersect(box, (2.0, 0.0, 10.0), (0.0, 0.0, 1.0))[0] == -1
    assert port.intersect(box, (10.0, 10.0, 10.0), (1.0, 0.0, 0.0))[0] == -1
    # an edge (x = 1, z = 2): the faces -x (0) and +z (5) give the same t; the lower face index wins
    t, n, hit = port.intersect(box, (0.0, 0.0, 3.0), (1.0, 0.0, -1.0))
    assert hit and t == 1.0 and n[0] == -1.0 and n[2] == 0.0
    # nearest-hit order across objects: a box in front of a sphere, ids preserved
    scene = [S.Sphere(mat, (8.0, 0.0, 1.0), 1.0), box]
    dist, _, idx = port.find_closest_hit(scene, (-4.0, 0.0, 1.0), (1.0, 0.0, 0.0))
    assert idx == 1 and dist == 5.0


def test_sun_term_matches_an_independent_restatement(S, port):
    scene = S.default_scene()
    pod = S.default_camera(48, 1.0).pod()
    p = port.default_params()
    p.max_depth = 0                                      # local colour only: closed form per pixel
    base = port.render(scene, pod, params=p)
    p.sun_enabled = 1
    sun = port.render(scene, pod, params=p)
    assert np.array_equal(base["object_id"], sun["object_id"])
    hit = base["object_id"] >= 0
    assert np.array_equal(base["radiance"][~hit], sun["radiance"][~hit])          # the sky is unchanged
    assert (sun["radiance"][hit] >= base["radiance"][hit]).all() and (sun["radiance"][hit] > base["radiance"][hit]).any()
    # independent restatement for the sphere pixels (id 0: centre (1.5,0,0), r .5, colour (0,1,0), kd .9, ks .4, n 50)
    s = np.array([.7, .4, .7]) / np.linalg.norm([.7, .4, .7])
    sun_color = np.array([1.64, 1.27, 0.99])
    cam_pos = np.array([pod.position.x, pod.position.y, pod.position.z])
    tl = np.array([pod.image_top_left.x, pod.image_top_left.y, pod.image_top_left.z])
    dx = np.array([pod.delta_x.x, pod.delta_x.y, pod.delta_x.z])
    dy = np.array([pod.delta_y.x, pod.delta_y.y, pod.delta_y.z])
    worst = 0.0
    for i, j in zip(*np.nonzero(base["object_id"] == 0)):
        d = cam_pos - ((tl + dx * j) + dy * i)
        c, r = np.array([1.5, 0.0, 0.0]), 0.5
        oc = cam_pos - c
        a, b, cc = d @ d, 2 * (d @ oc), oc @ oc - r * r
        proj = (-b - np.sqrt(b * b - 4 * a * cc)) / (2 * a)
        n = (cam_pos + d * proj) - c                      # the normal comes from the true intersection point
        nn = n / np.linalg.norm(n)
        v = -d / np.linalg.norm(d)
        h = (v + s) / np.linalg.norm(v + s)
        ks = max(0.0, s @ nn) * 0.9 + max(0.0, h @ nn) ** 50 * 0.4
        extra = np.array([0.0, 1.0, 0.0]) * sun_color * ks
        worst = max(worst, np.abs((sun["radiance"][i, j] - base["radiance"][i, j]) - extra).max())
    assert worst < 1e-12
    # a black sun changes nothing, bit for bit
    p.sun_color = type(p.sun_color)(0.0, 0.0, 0.0)
    assert np.array_equal(port.render(scene, pod, params=p)["radiance"], base["radiance"])


def _reinhard_numpy(rgb, key, white, saturate=True):
    rgb = np.asarray(rgb, np.float64)
    L = np.maximum(0.2126 * rgb[..., 0] + 0.7152 * rgb[..., 1] + 0.0722 * rgb[..., 2], 0.0)
    L = np.where(np.isnan(L), 0.0, L)
    lavg = np.exp(np.mean(np.log(1e-4 + L), axis=1))
    ls = (key / lavg)[:, None] * L
    ld = ls * (1 + (ls / white ** 2 if white > 0 else 0.0)) / (1 + ls)
    k = np.where(L > 0, ld / np.where(L > 0, L, 1.0), 0.0)
    out = rgb * k[..., None] * 255.0
    q = np.clip(np.trunc(out), 0, 255).astype(np.uint32)
    return (q[..., 0] << 24) | (q[..., 1] << 16) | (q[..., 2] << 8) | 0xFF, lavg


@pytest.mark.parametrize("dtype,white", [(np.float64, 0.0), (np.float32, 0.0), (np.float64, 2.5)])
def test_tonemap_spec_matches_numpy(port, pkg, dtype, white):
    rng = np.random.default_rng(3)
    rgb = (rng.lognormal(-1.0, 1.5, size=(3, 1001, 3))).astype(dtype)           # HDR-ish, ragged pixel count
    rgb[1, :50] = 0.0                                                            # black pixels: L = 0
    p = port.default_params()
    p.tonemap, p.tonemap_white, p.quantise_mode = pkg.abi.RTX_TONEMAP_REINHARD, white, pkg.abi.RTX_QUANT_SATURATE
    got, lavg = port.tonemap(rgb, p)
    exp, lavg_np = _reinhard_numpy(rgb, 0.18, white)
    assert np.allclose(lavg, lavg_np, rtol=1e-8)                                 # 32.32 fixed point vs float mean
    diff = np.abs(((got[..., None] >> np.array([24, 16, 8])) & 255).astype(int) - ((exp[..., None] >> np.array([24, 16, 8])) & 255).astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3
    # exposure invariance: scaling the input scales Lavg and leaves the image within 1 LSB
    got2, lavg2 = port.tonemap((rgb.astype(np.float64) * 8.0).astype(dtype), p)
    assert np.allclose(lavg2[[0, 2]], 8.0 * lavg[[0, 2]], rtol=1e-3)             # (the 1e-4 offset breaks exact scaling)
    d2 = np.abs(((got2[..., None] >> np.array([24, 16, 8])) & 255).astype(int) - ((got[..., None] >> np.array([24, 16, 8])) & 255).astype(int))
    assert np.percentile(d2[[0, 2]], 99) <= 1
