"""Host-side mirror of the reference's scene interface (scene.h) for the ray-tracing hot path.

Same names, constructor argument order, defaults and quirks as the reference classes, so that the
parity tests read like code written against the reference:

    Material(color, metallic=.5, ambient=.1, diffuse=.9, specular=.4, specular_exponent=50)   scene.h:48
    Sphere(mat=DEFAULT_MAT, center=(0,0,0), radius=1.0)                                     scene.h:81
    Wall(mat=DEFAULT_MAT, position=(0,0,0), normal=(0,0,0), length=1.0, width=1.0)          scene.h:70
    Camera(): aspect_ratio, image_width, vfov, position, lookat, vup; init() -> [dx, dy]    scene.h:86-112

Everything here is plain IEEE double arithmetic in the reference's operation order (Python floats);
no rendering happens in this module — `flatten()` turns a scene list into the POD array the C ABI
(include/rtx_b200.h) takes, and that is the only consumer.
"""
import math

from . import abi


def _v(t):
    return (float(t[0]), float(t[1]), float(t[2]))


def _sub(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def _add(a, b):
    return (a[0] + b[0], a[1] + b[1], a[2] + b[2])


def _mul(a, s):
    return (a[0] * s, a[1] * s, a[2] * s)


def _div(a, s):
    return (a[0] / s, a[1] / s, a[2] / s)


def _length(a):  # vec.cpp:3-9
    return math.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])


def _normalize(a):  # vec.cpp:21-24: three divides by length()
    return _div(a, _length(a))


def _cross(u, v):  # vec.cpp:15-19
    return (u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0])


class Material:
    """scene.h:35-49. Note the constructor order: metallic comes BEFORE ambient."""

    def __init__(self, color, metallic=.5, ambient=.1, diffuse=.9, specular=.4, specular_exponent=50):
        self.color = _v(color)
        self.metallic = float(metallic)
        self.ambient = float(ambient)
        self.diffuse = float(diffuse)
        self.specular = float(specular)
        self.specular_exponent = float(specular_exponent)

    def pod(self):
        m = abi.MaterialPOD()
        m.color = abi.Vec3(*self.color)
        m.ambient, m.metallic, m.diffuse = self.ambient, self.metallic, self.diffuse
        m.specular, m.specular_exponent = self.specular, self.specular_exponent
        return m


def default_mat():
    """DEFAULT_MAT (scene.h:3): `Material(RGB(1,1,1), .9, .9, .3, 30)` binds positionally to
    metallic .9, ambient .9, diffuse .3, specular 30, exponent 50 — reproduced, not corrected."""
    return Material((1, 1, 1), .9, .9, .3, 30)


class Sphere:
    """scene.h:75-84."""
    kind = abi.RTX_SPHERE

    def __init__(self, mat=None, center=(0, 0, 0), radius=1.0):
        self.mat = mat if mat is not None else default_mat()
        self.center = _v(center)
        self.radius = float(radius)

    def pod(self):
        o = abi.ObjectPOD()
        o.kind, o.mat = self.kind, self.mat.pod()
        o.p, o.a = abi.Vec3(*self.center), self.radius
        return o


class Wall:
    """scene.h:62-73. `position` is a CORNER of the rectangle (scene.cpp:25-29). The normal is stored
    as given: the library normalises it exactly like the constructor does (scene.h:71)."""
    kind = abi.RTX_WALL

    def __init__(self, mat=None, position=(0, 0, 0), normal=(0, 0, 0), length=1.0, width=1.0):
        self.mat = mat if mat is not None else default_mat()
        self.position = _v(position)
        self.normal = _v(normal)
        self.length = float(length)
        self.width = float(width)

    def pod(self):
        o = abi.ObjectPOD()
        o.kind, o.mat = self.kind, self.mat.pod()
        o.p, o.n = abi.Vec3(*self.position), abi.Vec3(*self.normal)
        o.a, o.b = self.length, self.width
        return o


class Box:
    """EXTENSION — not in the reference snapshot (README.md:21 names a sprint-2 `Box`; no code survives). An axis-aligned
    box: `position` is the minimum corner, `size` the extents. Six Wall-like faces with ONE object id; the exact
    semantics are the header's (include/rtx_b200.h, RTX_BOX) and oracle/oracle.c::box_intersect."""
    kind = abi.RTX_BOX

    def __init__(self, mat=None, position=(0, 0, 0), size=(1, 1, 1)):
        self.mat = mat if mat is not None else default_mat()
        self.position = _v(position)
        self.size = _v(size)

    def pod(self):
        o = abi.ObjectPOD()
        o.kind, o.mat = self.kind, self.mat.pod()
        o.p, o.n = abi.Vec3(*self.position), abi.Vec3(*self.size)
        return o


def flatten(scene):
    """list of Sphere/Wall (/Box) in scene order (== object ids, main.cpp:80) -> ctypes array of rtx_object."""
    arr = (abi.ObjectPOD * max(len(scene), 1))()
    for k, g in enumerate(scene):
        arr[k] = g.pod()
    return arr


class Camera:
    """scene.h:86-112 / scene.cpp:80-106. Field defaults as the reference (scene.h:98-100)."""

    def __init__(self):
        self.position = (0.0, 0.0, -1.0)
        self.lookat = (0.0, 0.0, 0.0)
        self.vup = (0.0, 1.0, 0.0)
        self.aspect_ratio = 1.0
        self.image_width = 640.0
        self.vfov = 90.0
        self.movement_speed = 0.1
        self.image_height = 0.0
        self.focal_length = 0.0
        self.direction = (0.0, 0.0, 0.0)
        self.fov_top_left = (0.0, 0.0, 0.0)
        self.image_top_left = (0.0, 0.0, 0.0)
        # scene.cpp:99-100 shadow these members with locals, so the reference leaves them 0; the returned
        # pair is the only carrier. We keep the returned pair here for convenience.
        self.u = None

    def desc(self):
        d = abi.CameraDesc()
        d.position, d.lookat, d.vup = abi.Vec3(*self.position), abi.Vec3(*self.lookat), abi.Vec3(*self.vup)
        d.vfov, d.aspect_ratio, d.image_width = float(self.vfov), float(self.aspect_ratio), float(self.image_width)
        return d

    def init(self):
        """Camera::init, scene.cpp:80-106, operation for operation (incl. 3.14 and the int() truncation)."""
        position, lookat, vup = _v(self.position), _v(self.lookat), _v(self.vup)
        image_width = float(self.image_width)
        self.image_height = float(int(image_width / self.aspect_ratio))
        self.focal_length = _length(_sub(position, lookat))
        theta = self.vfov * 3.14 / 180.0
        h = math.tan(theta / 2)
        fov_height = 2 * h * self.focal_length
        fov_width = fov_height * (image_width / self.image_height)
        w = _normalize(_sub(position, lookat))
        u = _normalize(_cross(vup, w))
        v = _cross(w, u)
        self.direction = w
        fov_x = _mul(u, fov_width)
        fov_y = _mul(v, -fov_height)
        pixel_delta_x = _div(fov_x, image_width)
        pixel_delta_y = _div(fov_y, self.image_height)
        self.fov_top_left = _sub(_sub(_sub(position, _mul(w, self.focal_length)), _div(fov_x, 2)), _div(fov_y, 2))
        self.image_top_left = _add(self.fov_top_left, _mul(_add(pixel_delta_x, pixel_delta_y), 0.5))
        self.u = [pixel_delta_x, pixel_delta_y]
        return self.u

    # -- moves (scene.cpp:108-165). Like the reference's main loop (main.cpp:154 vs :262-306) they do NOT re-run
    #    init(): image_top_left and the pixel deltas stay where init() put them, only `position` feeds rt_scene.
    def forward_vec(self):
        return _normalize(_v(self.direction))                                        # scene.cpp:108-110

    def right_vec(self):
        return _normalize(_cross(_v(self.direction), _v(self.vup)))                  # scene.cpp:112-114

    def up_vec(self):
        return _normalize(_cross(self.right_vec(), _v(self.direction)))              # scene.cpp:116-119

    def forward(self):
        self.position = _add(_v(self.position), _mul(self.forward_vec(), self.movement_speed))    # scene.cpp:121-123

    def backward(self):
        self.position = _sub(_v(self.position), _mul(self.forward_vec(), self.movement_speed))    # scene.cpp:125-127

    def right(self):
        self.position = _add(_v(self.position), _mul(self.right_vec(), self.movement_speed))      # scene.cpp:129-131

    def left(self):
        self.position = _sub(_v(self.position), _mul(self.right_vec(), self.movement_speed))      # scene.cpp:133-135

    def rotate_left_right(self, angle):
        """scene.cpp:137-145: yaw about z — the planar part of `direction` keeps its length and turns by `angle`;
        then vup = up_vec()."""
        dx, dy, dz = _v(self.direction)
        planar = _length((dx, dy, 0.0))
        yaw = math.atan2(dy, dx) + angle
        self.direction = (math.cos(yaw) * planar, math.sin(yaw) * planar, dz)
        self.vup = self.up_vec()

    def rotate_up_down(self, angle):
        """scene.cpp:147-165: pitch; the result is a unit vector over the old heading. Past the zenith the pitch stays,
        past the nadir it becomes MINUS the old pitch (the reference's own asymmetry, scene.cpp:156). Then vup = up_vec()."""
        dx, dy, dz = _v(self.direction)
        flat = (dx, dy, 0.0)
        pitch = math.atan2(dz, _length(flat))
        target = pitch + angle
        if target > math.pi / 2:
            target = pitch
        if target < -math.pi / 2:
            target = -pitch
        heading = _mul(_normalize(flat), math.cos(target))
        self.direction = (heading[0], heading[1], math.sin(target))
        self.vup = self.up_vec()

    def pod(self):
        """rtx_camera: what rt_scene consumes (main.cpp:132-134). Runs init() if it has not been run."""
        if self.u is None:
            self.init()
        c = abi.CameraPOD()
        c.position = abi.Vec3(*_v(self.position))
        c.image_top_left = abi.Vec3(*self.image_top_left)
        c.delta_x, c.delta_y = abi.Vec3(*self.u[0]), abi.Vec3(*self.u[1])
        c.width, c.height = int(self.image_width), int(self.image_height)
        return c


# ---------------------------------------------------------------------------------------------------
# The reference's data (main.cpp:146-163) and this repo's benchmark configurations.
# ---------------------------------------------------------------------------------------------------

def default_scene():
    """main.cpp:160-163: id0 green sphere, id1 blue wall, id2 green wall."""
    return [
        Sphere(Material((0, 1, 0), 0.5), (1.5, 0, 0), .5),
        Wall(Material((0, 0, 1)), (3.0, 2, 0), (0, -1, 0), 1, 1),
        Wall(Material((0, 1, 0)), (3.0, -3, 0), (0, 1, 0), 2, 2),
    ]


def default_camera(image_width=640, aspect_ratio=1.0):
    """main.cpp:146-154. ASPECT_RATIO = 4/3 is integer division == 1 (main.cpp:25) -> 640x640 by default."""
    cam = Camera()
    cam.aspect_ratio = aspect_ratio
    cam.image_width = image_width
    cam.movement_speed = 0.1
    cam.vfov = 90
    cam.position = (0, 0, 0)
    cam.lookat = (-1, 0, 0)
    cam.vup = (0, 0, -1)
    cam.init()
    return cam


class SplitMix64:
    """PRNG of the synthetic scene."""

    def __init__(self, seed):
        self.state = seed & 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.state = (self.state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def u(self, a=0.0, b=1.0):
        return a + (b - a) * ((self.next() >> 11) * (2.0 ** -53))


def synthetic_scene(n_spheres=10000, n_walls=64, seed=0xB200):
    """Config C3/C4 scene: n_spheres random spheres then n_walls random walls, drawn in the exact order of
    SURVEY.md §8(d)/A.2 (sphere: cx cy cz r R G B metallic; wall: px py pz phi nz length width R G B metallic)."""
    g = SplitMix64(seed)
    scene = []
    for _ in range(n_spheres):
        cx, cy, cz = g.u(4, 64), g.u(-32, 32), g.u(-8, 24)
        r = g.u(.1, .6)
        R, G, B = g.u(.1, 1), g.u(.1, 1), g.u(.1, 1)
        metallic = g.u(0, .8)
        scene.append(Sphere(Material((R, G, B), metallic), (cx, cy, cz), r))
    for _ in range(n_walls):
        px, py, pz = g.u(4, 64), g.u(-32, 32), g.u(-8, 8)
        phi = g.u(0, 6.283185307179586)
        nz = g.u(-.5, .5)
        length, width = g.u(1, 6), g.u(1, 6)
        R, G, B = g.u(.1, 1), g.u(.1, 1), g.u(.1, 1)
        metallic = g.u(0, .8)
        scene.append(Wall(Material((R, G, B), metallic), (px, py, pz), (math.cos(phi), math.sin(phi), nz), length, width))
    return scene


def flythrough_cameras(n_frames=256, image_width=1920, aspect_ratio=16.0 / 9.0):
    """Config C5: radius-6 orbit around the default sphere, Camera::init re-run per frame.
    The reference looks AWAY from `lookat` (main.cpp:133), so lookat = position + unit(position - target)."""
    cams = []
    target = (1.5, 0.0, 0.0)
    for k in range(n_frames):
        th = 2.0 * math.pi * k / n_frames
        pos = (1.5 + 6.0 * math.cos(th), 6.0 * math.sin(th), 0.75 + 0.5 * math.sin(2.0 * th))
        away = _normalize(_sub(pos, target))
        cam = Camera()
        cam.aspect_ratio = aspect_ratio
        cam.image_width = image_width
        cam.vfov = 90
        cam.position = pos
        cam.lookat = _add(pos, away)
        cam.vup = (0, 0, -1)
        cam.init()
        cams.append(cam)
    return cams
