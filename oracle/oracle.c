/*
 * oracle.c — CPU restatement of the reference's ray-tracing hot path in plain C.
 *
 * TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product library never does (there is no CPU fallback).
 *
 * Parity is PINNED: tests/test_oracle.py checks this file bit-for-bit against
 *   (1) the golden vectors the survey captured from the unmodified reference, and
 *   (2) oracle/_ref/libref_oracle.so — the reference's own sources compiled unmodified —
 *       on the default scene, the synthetic 10k scene and random rays,
 * and against the fixtures in tests/golden/ that (2) generated (tests/golden/make_golden.py).
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 * All arithmetic is IEEE double in the reference's operation order; build with -ffp-contract=off and
 * without -ffast-math so that no FMA contraction or reassociation changes a rounding
 * (the reference is built -O3 for baseline x86-64, which has no FMA: CMakeLists.txt:23).
 * Third-party arithmetic: glibc libm pow / sqrt / tan (main.cpp:35,103; vec.cpp:4; scene.cpp:65,70-71,85),
 * the same libm the reference links, so results are bit-identical to it on this image.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "rtx_b200.h"

typedef rtx_vec3 v3;

/* ---- vec3 (vec.h:12-37, vec.cpp:3-57) --------------------------------------------------------- */
static inline v3 mk(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static inline double len2(v3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }            /* vec.cpp:7-9   */
static inline double len(v3 a) { return sqrt(len2(a)); }                                 /* vec.cpp:3-5   */
static inline double dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }        /* vec.cpp:11-14 */
static inline v3 cross(v3 u, v3 v)                                                        /* vec.cpp:15-19 */
{
    return mk(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
}
static inline v3 add(v3 a, v3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }          /* vec.cpp:26-28 */
static inline v3 sub(v3 a, v3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }          /* vec.cpp:32-34 */
static inline v3 neg(v3 a) { return mk(-a.x, -a.y, -a.z); }                               /* vec.cpp:29-31 */
static inline v3 scale(v3 a, double d) { return mk(a.x * d, a.y * d, a.z * d); }          /* vec.cpp:38-40 */
static inline v3 divs(v3 a, double t) { return mk(a.x / t, a.y / t, a.z / t); }           /* vec.cpp:41-43 */
static inline v3 unit(v3 a) { return divs(a, len(a)); }                                   /* vec.cpp:21-24: divides, not a reciprocal */
static inline v3 lerp(v3 a, v3 b, double d)                                               /* vec.cpp:45-49 */
{
    return mk(a.x + d * (b.x - a.x), a.y + d * (b.y - a.y), a.z + d * (b.z - a.z));
}
static inline v3 reflect(v3 v, v3 normal)                                                 /* vec.cpp:51-57 */
{
    v3 nn = unit(normal);
    v3 nv = unit(v);
    double k = 2 * dot(nv, nn);
    return sub(nv, scale(nn, k));
}

/* ---- Collision (scene.h:27-33) ------------------------------------------------------------------ */
typedef struct { double distance; v3 normal; int hit; int index; } collision;
static inline collision miss(void) { collision c = {-1, {0, 0, 0}, 0, -1}; return c; }     /* scene.cpp:34,59 */

/* ---- scene objects with the constructor-time work done once ------------------------------------ */
typedef struct {
    int kind;
    rtx_material mat;
    v3 p;        /* center | corner */
    v3 n;        /* wall normal after Wall's ctor normalised it (scene.h:71) */
    double a, b; /* radius | length,width */
} object;

static void prepare(const rtx_object* in, int n, object* out)
{
    for (int k = 0; k < n; k++) {
        out[k].kind = in[k].kind;
        out[k].mat = in[k].mat;
        out[k].p = in[k].p;
        out[k].a = in[k].a;
        out[k].b = in[k].b;
        out[k].n = in[k].kind == RTX_WALL ? unit(in[k].n) : in[k].kind == RTX_BOX ? in[k].n : mk(0, 0, 0);
    }
}

/* Wall::intersect (scene.cpp:4-35). Distance is the PARAMETER t of the (possibly unnormalised) direction. */
static collision wall_intersect(const object* w, v3 o, v3 d)
{
    double denominator = dot(w->n, d);
    double t = dot(sub(w->p, o), w->n) / denominator;
    if (t > 0) {
        v3 point = add(o, scale(d, t));
        v3 right = unit(cross(w->n, mk(0, 0, 1)));          /* scene.cpp:18 (recomputed per call there) */
        v3 up = unit(cross(right, w->n));                    /* scene.cpp:19 */
        v3 rel = sub(point, w->p);
        double px = dot(rel, right);
        double py = dot(rel, up);
        if (px >= 0 && px <= w->a && py >= 0 && py <= w->b) {
            collision c = {t, w->n, 1, -1};
            return c;
        }
    }
    return miss();
}

/* Sphere::intersect (scene.cpp:40-78). Distance is in WORLD units (projection * |d|), the normal is
 * unnormalised (length r), and "hit" is reported even for negative projections. */
static collision sphere_intersect(const object* s, v3 o, v3 d)
{
    v3 oc = sub(o, s->p);
    double a = len2(d);
    double b = 2 * dot(d, oc);
    double c = len2(oc) - s->a * s->a;
    double det = b * b - 4 * a * c;
    double projection = -1;
    if (det < 0) return miss();
    v3 point = mk(0, 0, 0);
    if (det == 0) {
        point = add(o, scale(d, -b / (2 * a)));
        projection = (-b - sqrt(det)) / a;               /* divides by a, not 2a (scene.cpp:65) */
    } else {
        double p1 = (-b + sqrt(det)) / (2 * a);
        double p2 = (-b - sqrt(det)) / (2 * a);
        projection = p1 < p2 ? p1 : p2;
        point = add(o, scale(d, projection));
    }
    collision r = {projection * len(d), sub(point, s->p), 1, -1};
    return r;
}

/* RTX_BOX — NOT in the reference snapshot (README.md:21 mentions a sprint-2 `Box`; no code survives): parity
 * unpinned, this function IS the specification (include/rtx_b200.h). An axis-aligned box, p = minimum corner,
 * n = (sx, sy, sz) its extents, behaves as six Wall-like faces in the order -x, +x, -y, +y, -z, +z. Each face runs
 * Wall::intersect's arithmetic (scene.cpp:7-29) with its own corner c (the face's minimum corner), outward normal
 * and the positive unit axes spanning it as (right, up) — the reference's own basis derivation (scene.cpp:18-19)
 * is NaN for normals along z, which is why a box cannot be assembled from six reference Walls. The nearest face with
 * t > 0 wins, the lowest face index on equal t; distance is t (parametric, like Wall), the normal is the face's
 * outward normal, never flipped (a ray starting inside hits the far face from behind, like a Wall's back face). */
static collision box_intersect(const object* g, v3 o, v3 d)
{
    static const v3 axis[3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    const double size[3] = {g->n.x, g->n.y, g->n.z};
    collision best = miss();
    double best_t = DBL_MAX;
    for (int f = 0; f < 6; f++) {
        const int ax = f >> 1, hi = f & 1;
        const int r_ax = ax == 0 ? 1 : 0, u_ax = ax == 2 ? 1 : 2;     /* x: (y,z)   y: (x,z)   z: (x,y) */
        const v3 normal = hi ? axis[ax] : neg(axis[ax]);
        const v3 corner = hi ? add(g->p, scale(axis[ax], size[ax])) : g->p;
        double denominator = dot(normal, d);
        double t = dot(sub(corner, o), normal) / denominator;
        if (t > 0) {
            v3 rel = sub(add(o, scale(d, t)), corner);
            double px = dot(rel, axis[r_ax]);
            double py = dot(rel, axis[u_ax]);
            if (px >= 0 && px <= size[r_ax] && py >= 0 && py <= size[u_ax] && t < best_t) {
                best_t = t;
                collision c = {t, normal, 1, -1};
                best = c;
            }
        }
    }
    return best;
}

static inline collision intersect(const object* g, v3 o, v3 d)
{
    return g->kind == RTX_SPHERE ? sphere_intersect(g, o, d) : g->kind == RTX_BOX ? box_intersect(g, o, d) : wall_intersect(g, o, d);
}

/* find_closest_hit (main.cpp:67-84): strict '<' keeps the lowest index on equal distances. */
static collision closest_hit(const object* scene, int n, v3 o, v3 d)
{
    collision best = {DBL_MAX, {0, 0, 0}, 0, -1};
    for (int j = 0; j < n; j++) {
        collision c = intersect(&scene[j], o, d);
        if (c.distance > 0 && c.distance < best.distance) {
            best = c;
            best.index = j;
        }
    }
    return best;
}

/* out_color (main.cpp:28-37): the sign test is on the UNNORMALISED z. */
static v3 sky(v3 v, const rtx_params* p)
{
    if (v.z < 0.0) return p->ground_color;
    v = unit(v);
    return lerp(p->sky_low, p->sky_high, pow(v.z, p->sky_exponent));
}

/* diffuse_shading (main.cpp:42-48) */
static double diffuse_term(v3 pos, v3 normal, v3 light)
{
    v3 l = unit(sub(light, pos));
    double lambert = dot(l, unit(normal));
    return lambert > 0 ? lambert : 0;
}

/* specular (main.cpp:53-62); the exponent is applied by the caller (main.cpp:103) */
static double specular_term(v3 pos, v3 normal, v3 light, v3 view)
{
    view = unit(view);
    normal = unit(normal);
    v3 l = unit(sub(light, pos));
    v3 h = unit(add(view, l));
    double r = dot(h, normal);
    return r > 0 ? r : 0;
}

/* recursive_ray_tracing (main.cpp:89-119). *rays counts find_closest_hit calls (may be NULL). */
static v3 trace(const object* scene, int n, v3 o, v3 d, int remaining, const rtx_params* p, int* rays)
{
    collision col = closest_hit(scene, n, o, d);
    if (rays) (*rays)++;
    if (col.index < 0) return sky(d, p);
    v3 pos = add(o, scale(d, col.distance));
    const rtx_material* m = &scene[col.index].mat;
    double di = diffuse_term(add(o, scale(d, col.distance)), col.normal, p->light_pos);
    double si = pow(specular_term(pos, col.normal, p->light_pos, neg(d)), m->specular_exponent);
    v3 local = scale(m->color, di * m->diffuse + si * m->specular + m->ambient);
    if (p->sun_enabled) {
        /* EXTENSION, not in the reference (rtx_params.sun_enabled, include/rtx_b200.h): SUN_COLOR / SUN_DIRECTION
         * (main.cpp:18-19, defined and never used) as a directional light through the same Blinn-Phong terms;
         * no shadow ray, like the reference's point light. This block IS the specification. */
        v3 s = unit(p->sun_direction);
        v3 nn = unit(col.normal);
        double ls = dot(s, nn);
        double ds = ls > 0 ? ls : 0;
        v3 h = unit(add(unit(neg(d)), s));
        double sps = dot(h, nn);
        double ss = pow(sps > 0 ? sps : 0, m->specular_exponent);
        double ks = ds * m->diffuse + ss * m->specular;
        v3 tint = mk(m->color.x * p->sun_color.x, m->color.y * p->sun_color.y, m->color.z * p->sun_color.z);
        local = add(local, scale(tint, ks));
    }
    if (remaining <= 0) return local;
    v3 start = add(pos, scale(col.normal, p->reflect_offset));   /* normal unnormalised: offset = r*1e-4 on spheres */
    v3 rd = reflect(d, col.normal);
    v3 rt = trace(scene, n, start, rd, remaining - 1, p, rays);
    return lerp(local, rt, m->metallic);
}

/* main.cpp:345: double -> Uint8 is the implicit C conversion. On x86-64 the compiler emits cvttsd2si
 * (32-bit) and keeps the low byte: in-range values truncate toward zero and wrap mod 256; NaN and
 * |v| >= 2^31 produce 0x80000000 whose low byte is 0. Written out so the result does not depend on UB. */
static inline uint32_t to_u8_wrap(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0)) return 0;
    return (uint32_t)(int32_t)v & 0xFFu;
}
static inline uint32_t to_u8_sat(double v)
{
    if (!(v > 0.0)) return 0;
    if (v >= 255.0) return 255;
    return (uint32_t)v;
}
static inline uint32_t pack_rgba(v3 c, int mode)                          /* main.cpp:193,345 */
{
    double r = c.x * 255, g = c.y * 255, b = c.z * 255;
    uint32_t R = mode == RTX_QUANT_SATURATE ? to_u8_sat(r) : to_u8_wrap(r);
    uint32_t G = mode == RTX_QUANT_SATURATE ? to_u8_sat(g) : to_u8_wrap(g);
    uint32_t B = mode == RTX_QUANT_SATURATE ? to_u8_sat(b) : to_u8_wrap(b);
    return (R << 24) | (G << 16) | (B << 8) | 0xFFu;
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ================================================================================================ */

void orc_default_params(rtx_params* p)
{
    memset(p, 0, sizeof *p);
    p->max_depth = 10;                            /* main.cpp:89  */
    p->quantise_mode = RTX_QUANT_WRAP;
    p->fuse_quantise = 1;
    p->light_pos = mk(0, 0, 0);                   /* main.cpp:14  */
    p->ground_color = mk(0.025, 0.05, 0.075);     /* main.cpp:15  */
    p->sky_low = mk(0.36, 0.45, 0.57);            /* main.cpp:16  */
    p->sky_high = mk(0.14, 0.21, 0.49);           /* main.cpp:17  */
    p->reflect_offset = .0001;                    /* main.cpp:111 */
    p->sky_exponent = (double)(float)(1. / 4.);   /* `const float skyGradient = 1. / 4.` main.cpp:34 */
    p->band_rows = 4;
    p->n_ranks = 1;
    p->rank = 0;
    p->frame_offset = 0;
    p->frame_stride = 1;
    p->sun_enabled = 0;                           /* extensions off = the reference */
    p->tonemap = RTX_TONEMAP_NONE;
    p->sun_color = mk(1.64, 1.27, 0.99);          /* main.cpp:18 */
    p->sun_direction = mk(.7, .4, .7);            /* main.cpp:19 */
    p->tonemap_key = 0.18;
    p->tonemap_white = 0.0;
}

int orc_abi_version(void) { return RTX_ABI_VERSION; }

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Camera::init (scene.cpp:80-106). */
void orc_camera_init(const rtx_camera_desc* d, rtx_camera* out)
{
    double image_width = d->image_width;
    double image_height = (int)(image_width / d->aspect_ratio);
    double focal = len(sub(d->position, d->lookat));
    double theta = d->vfov * 3.14 / 180.0;
    double h = tan(theta / 2);
    double fov_height = 2 * h * focal;
    double fov_width = fov_height * (image_width / image_height);
    v3 w = unit(sub(d->position, d->lookat));
    v3 u = unit(cross(d->vup, w));
    v3 v = cross(w, u);
    v3 fov_x = scale(u, fov_width);
    v3 fov_y = scale(v, -fov_height);
    v3 dx = divs(fov_x, image_width);
    v3 dy = divs(fov_y, image_height);
    v3 top_left = sub(sub(sub(d->position, scale(w, focal)), divs(fov_x, 2)), divs(fov_y, 2));
    out->position = d->position;
    out->image_top_left = add(top_left, scale(add(dx, dy), 0.5));
    out->delta_x = dx;
    out->delta_y = dy;
    out->width = (int32_t)image_width;
    out->height = (int32_t)image_height;
}

/* Camera moves (scene.cpp:108-165). The state the reference's methods read and write is (position, direction, vup);
 * init() sets direction = normalize(position - lookat) (scene.cpp:93) and is never re-run by a move (main.cpp:154 vs
 * :262-306), so image_top_left and the pixel deltas stay those of init().
 *   forward_vec = normalize(direction); right_vec = normalize(cross(direction, vup));
 *   up_vec = normalize(cross(right_vec, direction))                                             scene.cpp:108-119
 *   forward/backward/right/left: position +- vec * movement_speed                                scene.cpp:121-135
 *   rotate_left_right(angle): yaw around z through atan2/cos/sin, then vup = up_vec()            scene.cpp:137-145
 *   rotate_up_down(angle): pitch through atan2/sin/cos; beyond +pi/2 the pitch is kept, beyond -pi/2 it becomes
 *   MINUS the old pitch (the reference's own asymmetry, scene.cpp:155-156); then vup = up_vec()  scene.cpp:147-165
 * ops: 'w' 's' 'a' 'd' as the keys of main.cpp:262-306, 'y' = rotate_left_right(args[k]), 'p' = rotate_up_down(args[k]).
 * state_out: position, direction, vup after every step (9 doubles each). */
#define ORC_PI 3.14159265358979323846   /* glibc's M_PI (scene.cpp:155-156); -std=c11 hides the macro */
static v3 cam_right_vec(v3 direction, v3 vup) { return unit(cross(direction, vup)); }
static v3 cam_up_vec(v3 direction, v3 vup) { return unit(cross(cam_right_vec(direction, vup), direction)); }

void orc_camera_walk(const rtx_camera_desc* d, const char* ops, const double* args, int32_t n_ops, double* state_out,
                     rtx_camera* cam_out)
{
    const double movement_speed = 0.1;                       /* main.cpp:149 */
    rtx_camera c;
    orc_camera_init(d, &c);
    v3 position = d->position, vup = d->vup;
    v3 direction = unit(sub(d->position, d->lookat));        /* scene.cpp:93 */
    for (int32_t k = 0; k < n_ops; k++) {
        switch (ops[k]) {
            case 'w': position = add(position, scale(unit(direction), movement_speed)); break;
            case 's': position = sub(position, scale(unit(direction), movement_speed)); break;
            case 'd': position = add(position, scale(cam_right_vec(direction, vup), movement_speed)); break;
            case 'a': position = sub(position, scale(cam_right_vec(direction, vup), movement_speed)); break;
            case 'y': {   /* rotate_left_right: yaw about z, the planar length and z are kept */
                double planar = len(mk(direction.x, direction.y, 0));
                double yaw = atan2(direction.y, direction.x) + args[k];
                direction = mk(cos(yaw) * planar, sin(yaw) * planar, direction.z);
                vup = cam_up_vec(direction, vup);
                break;
            }
            case 'p': {   /* rotate_up_down: unit vector over the old heading; the two clamps are the reference's, sign flip included */
                v3 flat = mk(direction.x, direction.y, 0);
                double pitch = atan2(direction.z, len(flat));
                double target = pitch + args[k];
                if (target > ORC_PI / 2) target = pitch;
                if (target < -ORC_PI / 2) target = -pitch;
                v3 heading = scale(unit(flat), cos(target));
                direction = mk(heading.x, heading.y, sin(target));
                vup = cam_up_vec(direction, vup);
                break;
            }
            default: break;
        }
        const v3 st[3] = {position, direction, vup};
        for (int i = 0; i < 3; i++) {
            state_out[9 * k + 3 * i + 0] = st[i].x;
            state_out[9 * k + 3 * i + 1] = st[i].y;
            state_out[9 * k + 3 * i + 2] = st[i].z;
        }
    }
    if (cam_out) {
        *cam_out = c;
        cam_out->position = position;
    }
}

/* The loop body of rt_scene (main.cpp:129-136) over the given global rows, packed [n_rows][width];
 * every output plane is optional. Returns seconds spent in the loop, -1 on bad arguments. */
double orc_render_rows_params(const rtx_object* objs, int32_t n_objs, const rtx_camera* cam, const rtx_params* p,
                              const int32_t* rows, int32_t n_rows, int32_t n_threads,
                              double* radiance, uint32_t* rgba8, int32_t* object_id, uint8_t* hit_mask,
                              uint8_t* ray_count, uint64_t* total_rays)
{
    if ((!objs && n_objs > 0) || !cam || !rows || n_rows < 0 || !p) return -1;
    object* scene = (object*)malloc(sizeof(object) * (size_t)(n_objs > 0 ? n_objs : 1));
    if (!scene) return -1;
    prepare(objs, n_objs, scene);
    const int W = cam->width;
    uint64_t rays_sum = 0;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#endif
    double t0 = now_s();
    /* work items = (row, 64-pixel column chunk), like oracle/ref_harness.cpp: a few rows still feed every thread */
    enum { CHUNK = 64 };
    const int n_chunks = (W + CHUNK - 1) / CHUNK;
    const long long n_items = (long long)n_rows * n_chunks;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads) reduction(+ : rays_sum)
    for (long long item = 0; item < n_items; item++) {
        const int k = (int)(item / n_chunks);
        const int j0 = (int)(item - (long long)k * n_chunks) * CHUNK;
        const int j1 = j0 + CHUNK < W ? j0 + CHUNK : W;
        const int i = rows[k];
        for (int j = j0; j < j1; j++) {
            v3 center = add(add(cam->image_top_left, scale(cam->delta_x, j)), scale(cam->delta_y, i)); /* main.cpp:132 */
            v3 d = sub(cam->position, center);                                                          /* main.cpp:133 */
            int rays = 0;
            v3 c = trace(scene, n_objs, cam->position, d, p->max_depth, p, &rays);                      /* main.cpp:136 */
            size_t px = (size_t)k * W + j;
            if (radiance) {
                radiance[3 * px + 0] = c.x;
                radiance[3 * px + 1] = c.y;
                radiance[3 * px + 2] = c.z;
            }
            if (rgba8) rgba8[px] = pack_rgba(c, p->quantise_mode);
            if (object_id || hit_mask) {
                collision first = closest_hit(scene, n_objs, cam->position, d);
                if (object_id) object_id[px] = first.index;
                if (hit_mask) hit_mask[px] = first.index >= 0;
            }
            if (ray_count) ray_count[px] = (uint8_t)rays;
            rays_sum += (uint64_t)rays;
        }
    }
    double t1 = now_s();
    free(scene);
    if (total_rays) *total_rays = rays_sum;
    return t1 - t0;
}

/* Same door as ref_render_rows in oracle/ref_harness.cpp: reference literals for every parameter. */
double orc_render_rows(const rtx_object* objs, int32_t n_objs, const rtx_camera* cam, int32_t max_depth,
                       const int32_t* rows, int32_t n_rows, int32_t n_threads,
                       double* radiance, uint32_t* rgba8, int32_t* object_id, uint8_t* hit_mask, uint8_t* ray_count,
                       uint64_t* total_rays)
{
    rtx_params p;
    orc_default_params(&p);
    p.max_depth = max_depth;
    return orc_render_rows_params(objs, n_objs, cam, &p, rows, n_rows, n_threads, radiance, rgba8, object_id, hit_mask,
                                  ray_count, total_rays);
}

/* main.cpp:343-345 over n RGB triples. */
void orc_quantise(const double* rgb, int64_t n, uint32_t* out)
{
    for (int64_t k = 0; k < n; k++) out[k] = pack_rgba(mk(rgb[3 * k], rgb[3 * k + 1], rgb[3 * k + 2]), RTX_QUANT_WRAP);
}
void orc_quantise_mode(const double* rgb, int64_t n, int32_t mode, uint32_t* out)
{
    for (int64_t k = 0; k < n; k++) out[k] = pack_rgba(mk(rgb[3 * k], rgb[3 * k + 1], rgb[3 * k + 2]), mode);
}
/* The float-radiance variant of the product's standalone quantise: float -> double is exact, then main.cpp:345. */
void orc_quantise_f32(const float* rgb, int64_t n, int32_t mode, uint32_t* out)
{
    for (int64_t k = 0; k < n; k++)
        out[k] = pack_rgba(mk((double)rgb[3 * k], (double)rgb[3 * k + 1], (double)rgb[3 * k + 2]), mode);
}

/* EXTENSION, not in the reference (rtx_params.tonemap = RTX_TONEMAP_REINHARD, include/rtx_b200.h; README.md:13 only
 * mentions that tone mapping exists). This function IS the specification the CUDA kernels (csrc/tonemap.cu) follow:
 * Reinhard's global photographic operator per frame, then the 8-bit pack of main.cpp:345 in the chosen mode.
 *   L    = .2126 R + .7152 G + .0722 B, negative or NaN -> 0
 *   sum  = sum over the frame of llrint(log(1e-4 + L) * 2^32)      (32.32 fixed point: order independent)
 *   Lavg = exp(((double)sum / 2^32) / n);  Ls = key / Lavg * L;  Ld = Ls (1 + Ls / white^2) / (1 + Ls)  [white <= 0: no term]
 *   rgb *= Ld / L (0 where L <= 0)
 * Exactly one of rgb64 / rgb32 is non-NULL; log_avg (may be NULL) receives Lavg per frame.
 * FLOAT radiance (rgb32) is processed in FLOAT — luminance, logf, Ls, Ld and the scale are single-precision operations, one
 * rounding each (this file is compiled with -ffp-contract=off); the per-frame constants (Lavg, key / Lavg, 1 / white^2) are
 * computed in double and rounded to float once; the scaled channels go to the 8-bit pack as doubles. */
static double tm_lum(double r, double g, double b)
{
    double l = 0.2126 * r + 0.7152 * g + 0.0722 * b;
    return l > 0.0 ? l : 0.0;
}
static float tm_lum_f(float r, float g, float b)
{
    float l = (0.2126f * r + 0.7152f * g) + 0.0722f * b;
    return l > 0.0f ? l : 0.0f;
}
/* Step 1: ADDS the fixed-point sums of this buffer's frames to sums[n_frames]. */
void orc_tonemap_sums(const double* rgb64, const float* rgb32, int64_t pixels, int32_t n_frames, int64_t* sums)
{
    const double fix = 4294967296.0;
    for (int32_t f = 0; f < n_frames; f++) {
        const int64_t base = (int64_t)f * pixels;
        long long sum = 0;
        if (rgb32) {
            for (int64_t k = base; k < base + pixels; k++)
                sum += llrintf(logf(1e-4f + tm_lum_f(rgb32[3 * k], rgb32[3 * k + 1], rgb32[3 * k + 2])) * 4294967296.0f);
        } else {
            for (int64_t k = base; k < base + pixels; k++)
                sum += llrint(log(1e-4 + tm_lum(rgb64[3 * k], rgb64[3 * k + 1], rgb64[3 * k + 2])) * fix);
        }
        sums[f] += sum;
    }
}
/* Step 2: maps and packs this buffer's pixels with sums taken over pixels_global pixels per frame (>= pixels when the
 * frame's rows are spread over several ranks and the sums were added up over them). */
void orc_tonemap_apply(const double* rgb64, const float* rgb32, int64_t pixels, int32_t n_frames, const int64_t* sums,
                       int64_t pixels_global, const rtx_params* p, uint32_t* out, double* log_avg)
{
    const double fix = 4294967296.0;
    for (int32_t f = 0; f < n_frames; f++) {
        const int64_t base = (int64_t)f * pixels;
        const double mean = ((double)sums[f] / fix) / (double)pixels_global;
        const double lavg = exp(mean);
        const double key_over_avg = p->tonemap_key / lavg;
        const double inv_white2 = p->tonemap_white > 0.0 ? 1.0 / (p->tonemap_white * p->tonemap_white) : 0.0;
        if (log_avg) log_avg[f] = lavg;
        if (rgb32) {
            const float koa = (float)key_over_avg, iw2 = (float)inv_white2;
            for (int64_t k = base; k < base + pixels; k++) {
                float r = rgb32[3 * k], g = rgb32[3 * k + 1], b = rgb32[3 * k + 2];
                float l = tm_lum_f(r, g, b);
                float ls = koa * l;
                float ld = (ls * (1.0f + ls * iw2)) / (1.0f + ls);
                float s = l > 0.0f ? ld / l : 0.0f;
                float R = r * s, G = g * s, B = b * s;
                out[k] = pack_rgba(mk((double)R, (double)G, (double)B), p->quantise_mode);
            }
            continue;
        }
        for (int64_t k = base; k < base + pixels; k++) {
            double r = rgb64[3 * k], g = rgb64[3 * k + 1], b = rgb64[3 * k + 2];
            double l = tm_lum(r, g, b);
            double ls = key_over_avg * l;
            double ld = (ls * (1.0 + ls * inv_white2)) / (1.0 + ls);
            double s = l > 0.0 ? ld / l : 0.0;
            out[k] = pack_rgba(mk(r * s, g * s, b * s), p->quantise_mode);
        }
    }
}
void orc_tonemap(const double* rgb64, const float* rgb32, int64_t pixels, int32_t n_frames, const rtx_params* p, uint32_t* out,
                 double* log_avg)
{
    int64_t* sums = (int64_t*)calloc((size_t)n_frames, sizeof(int64_t));
    orc_tonemap_sums(rgb64, rgb32, pixels, n_frames, sums);
    orc_tonemap_apply(rgb64, rgb32, pixels, n_frames, sums, pixels, p, out, log_avg);
    free(sums);
}

/* ---- function-level doors (same signatures as the ref_* ones) ------------------------------------ */
void orc_intersect(const rtx_object* obj, const v3* origin, const v3* dir, double* distance, v3* normal, int32_t* hit)
{
    object g;
    prepare(obj, 1, &g);
    collision c = intersect(&g, *origin, *dir);
    *distance = c.distance;
    *normal = c.normal;
    *hit = c.hit;
}

void orc_find_closest_hit(const rtx_object* objs, int32_t n, const v3* origin, const v3* dir, double* distance, v3* normal,
                          int32_t* index)
{
    object* scene = (object*)malloc(sizeof(object) * (size_t)(n > 0 ? n : 1));
    prepare(objs, n, scene);
    collision c = closest_hit(scene, n, *origin, *dir);
    *distance = c.distance;
    *normal = c.normal;
    *index = c.index;
    free(scene);
}

void orc_trace_ray(const rtx_object* objs, int32_t n, const v3* origin, const v3* dir, int32_t depth, v3* rgb)
{
    rtx_params p;
    orc_default_params(&p);
    object* scene = (object*)malloc(sizeof(object) * (size_t)(n > 0 ? n : 1));
    prepare(objs, n, scene);
    *rgb = trace(scene, n, *origin, *dir, depth, &p, NULL);
    free(scene);
}

void orc_out_color(const v3* v, v3* rgb)
{
    rtx_params p;
    orc_default_params(&p);
    *rgb = sky(*v, &p);
}
void orc_reflect(const v3* v, const v3* n, v3* out) { *out = reflect(*v, *n); }
double orc_diffuse(const v3* pos, const v3* n, const v3* light) { return diffuse_term(*pos, *n, *light); }
double orc_specular(const v3* pos, const v3* n, const v3* light, const v3* view) { return specular_term(*pos, *n, *light, *view); }
