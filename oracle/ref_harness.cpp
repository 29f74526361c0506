/*
 * ref_harness.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Thin extern "C" door onto the UNMODIFIED reference sources, compiled where they lie under
 * /root/reference by oracle/build_ref.sh into oracle/_ref/libref_oracle.so. Every number this
 * file returns is computed by the reference's own symbols:
 *
 *   out_color, diffuse_shading, specular, find_closest_hit, recursive_ray_tracing, rt_scene
 *                                                            (main.cpp:28-139, external linkage)
 *   Sphere::intersect / Wall::intersect / Camera::init      (scene.cpp:4-106)
 *   vec3::*                                                 (vec.cpp)
 *   main() itself, renamed ref_main by -Dmain=ref_main, run headless on the SDL stub
 *
 * The only arithmetic restated here is (a) the per-pixel loop body of rt_scene (main.cpp:132-136),
 * needed because rt_scene hard-codes the depth default and a transposed frame buffer that only works
 * for square frames (main.cpp:243 vs :136), (b) the chain walk that counts rays (main.cpp:99,111-113),
 * and (c) the byte packing SDL_MapRGB does for the RGBA8888 masks of main.cpp:193.
 * The image is distributed over OpenMP threads ("lines of the image are distributed across hardware
 * threads", README.md:13; here in 64-pixel pieces of a line so that a sample of a few lines still occupies
 * every thread) — the per-pixel function only reads const scene data, so this is safe.
 */
#include <SDL.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "scene.h"        /* /root/reference/scene.h (no include guard: include exactly once) */
#include "rtx_b200.h"

/* ---- the reference's free functions (main.cpp), declared here, defined in main.cpp -------- */
RGB out_color(vec3 v);
double diffuse_shading(vec3 pos, vec3 normal, vec3 light_pos);
double specular(vec3 pos, vec3 normal, vec3 light_pos, vec3 view_dir);
Collision find_closest_hit(const std::vector<std::unique_ptr<SceneGeometry>>& scene, ray r);
RGB recursive_ray_tracing(const std::vector<std::unique_ptr<SceneGeometry>>& scene, ray r, int remaining_iterations);
void rt_scene(std::vector<vec3> u, const std::vector<std::unique_ptr<SceneGeometry>>& scene, const Camera& cam,
              std::vector<std::vector<RGB>>& frame_buffer);
int ref_main(int argc, char* args[]);

/* ---- SDL stub bodies ------------------------------------------------------------------------ */
static int g_frames_before_quit = 1;
static int g_poll_calls = 0;
static SDL_Surface* g_surface = nullptr;
static SDL_PixelFormat g_format;

int SDL_Init(Uint32) { return 0; }
const char* SDL_GetError() { return "stub"; }
SDL_Window* SDL_CreateWindow(const char*, int, int, int, int, Uint32) { return reinterpret_cast<SDL_Window*>(0x1); }
SDL_Surface* SDL_CreateRGBSurface(Uint32, int w, int h, int, Uint32 r, Uint32 g, Uint32 b, Uint32 a)
{
    g_format = SDL_PixelFormat{r, g, b, a};
    SDL_Surface* s = new SDL_Surface;
    s->pixels = std::calloc(static_cast<size_t>(w) * h, 4);
    s->pitch = w * 4;
    s->format = &g_format;
    s->w = w;
    s->h = h;
    g_surface = s;
    return s;
}
SDL_Renderer* SDL_CreateRenderer(SDL_Window*, int, Uint32) { return reinterpret_cast<SDL_Renderer*>(0x1); }
SDL_Texture* SDL_CreateTextureFromSurface(SDL_Renderer*, SDL_Surface*) { return reinterpret_cast<SDL_Texture*>(0x1); }
void SDL_DestroyWindow(SDL_Window*) {}
void SDL_DestroyRenderer(SDL_Renderer*) {}
void SDL_DestroyTexture(SDL_Texture*) {}
void SDL_FreeSurface(SDL_Surface*) {} /* kept alive so ref_run_main can copy it out */
void SDL_Quit() {}
/* The reference drains events in an inner while (main.cpp:253): answer "one event" only when it is time
 * to quit, otherwise "no event" so that the frame is rendered. */
int SDL_PollEvent(SDL_Event* e)
{
    static bool quit_sent = false;
    if (g_poll_calls == 0) quit_sent = false;
    g_poll_calls++;
    /* The iteration that receives SDL_QUIT still renders its frame (main.cpp:250-256, 329), so QUIT is
     * delivered at the start of the last wanted frame. */
    if (!quit_sent && g_poll_calls > g_frames_before_quit) {
        e->type = SDL_QUIT;
        quit_sent = true;
        return 1;
    }
    return 0;
}
/* SDL_MapRGB for a 32-bit surface whose masks are R 0xFF000000, G 0x00FF0000, B 0x0000FF00, A 0x000000FF
 * (main.cpp:193): shift each byte to its mask and set alpha opaque. */
Uint32 SDL_MapRGB(const SDL_PixelFormat*, Uint8 r, Uint8 g, Uint8 b)
{
    return (static_cast<Uint32>(r) << 24) | (static_cast<Uint32>(g) << 16) | (static_cast<Uint32>(b) << 8) | 0xFFu;
}
Uint32 SDL_MapRGBA(const SDL_PixelFormat*, Uint8 r, Uint8 g, Uint8 b, Uint8 a)
{
    return (static_cast<Uint32>(r) << 24) | (static_cast<Uint32>(g) << 16) | (static_cast<Uint32>(b) << 8) | a;
}
int SDL_RenderClear(SDL_Renderer*) { return 0; }
int SDL_RenderCopy(SDL_Renderer*, SDL_Texture*, const SDL_Rect*, const SDL_Rect*) { return 0; }
void SDL_RenderPresent(SDL_Renderer*) {}

/* ---- helpers ---------------------------------------------------------------------------------- */
using Scene = std::vector<std::unique_ptr<SceneGeometry>>;
constexpr int kChunk = 64;   /* pixels per OpenMP work item of ref_render_rows */

static inline vec3 V(const rtx_vec3& v) { return vec3(v.x, v.y, v.z); }
static inline rtx_vec3 P(const vec3& v) { return rtx_vec3{v.x, v.y, v.z}; }

static Scene build_scene(const rtx_object* objs, int n)
{
    Scene scene;
    for (int k = 0; k < n; k++) {
        const rtx_object& o = objs[k];
        /* ctor order (color, metallic, ambient, diffuse, specular, specular_exponent), scene.h:48 */
        Material m(V(o.mat.color), o.mat.metallic, o.mat.ambient, o.mat.diffuse, o.mat.specular, o.mat.specular_exponent);
        if (o.kind == RTX_SPHERE)
            scene.push_back(std::make_unique<Sphere>(m, V(o.p), o.a));
        else if (o.kind == RTX_WALL)
            scene.push_back(std::make_unique<Wall>(m, V(o.p), V(o.n), o.a, o.b));
        else {   /* RTX_BOX and anything else is an extension of this repo: the reference has no such class */
            std::fprintf(stderr, "ref_harness: object kind %d does not exist in the reference\n", o.kind);
            std::abort();
        }
    }
    return scene;
}

/* main.cpp:345 with the conversions the call performs: double -> Uint8 is implicit at the call site. */
static inline Uint32 quantise_like_main(const RGB& val)
{
    return SDL_MapRGB(nullptr, val.x * 255, val.y * 255, val.z * 255);
}

extern "C" {

int ref_abi_version(void) { return RTX_ABI_VERSION; }

int ref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Camera::init (scene.cpp:80-106) through the reference class. */
void ref_camera_init(const rtx_camera_desc* d, rtx_camera* out)
{
    Camera cam;
    cam.aspect_ratio = d->aspect_ratio;
    cam.image_width = d->image_width;
    cam.movement_speed = 0.1;
    cam.vfov = d->vfov;
    cam.position = V(d->position);
    cam.lookat = V(d->lookat);
    cam.vup = V(d->vup);
    std::vector<vec3> u = cam.init();
    out->position = P(cam.position);
    out->image_top_left = P(cam.image_top_left);
    out->delta_x = P(u[0]);
    out->delta_y = P(u[1]);
    out->width = static_cast<int32_t>(cam.image_width);
    out->height = static_cast<int32_t>(cam.image_height);
}

/*
 * Camera moves through the reference class (scene.cpp:108-165): init() once, then one method per step.
 * ops[k]: 'w' forward, 's' backward, 'a' left, 'd' right (the keys of main.cpp:262-306), 'y' rotate_left_right(args[k]),
 * 'p' rotate_up_down(args[k]) (the mouse look main.cpp:319-323 leaves commented out). init() is NOT re-run, exactly
 * like the reference's main loop. state_out receives position, direction, vup (9 doubles) after every step; cam_out
 * what rt_scene would consume after the last step (image_top_left and the deltas are still those of init()).
 */
void ref_camera_walk(const rtx_camera_desc* d, const char* ops, const double* args, int32_t n_ops, double* state_out,
                     rtx_camera* cam_out)
{
    Camera cam;
    cam.aspect_ratio = d->aspect_ratio;
    cam.image_width = d->image_width;
    cam.movement_speed = 0.1;
    cam.vfov = d->vfov;
    cam.position = V(d->position);
    cam.lookat = V(d->lookat);
    cam.vup = V(d->vup);
    std::vector<vec3> u = cam.init();
    for (int32_t k = 0; k < n_ops; k++) {
        switch (ops[k]) {
            case 'w': cam.forward(); break;
            case 's': cam.backward(); break;
            case 'a': cam.left(); break;
            case 'd': cam.right(); break;
            case 'y': cam.rotate_left_right(args[k]); break;
            case 'p': cam.rotate_up_down(args[k]); break;
            default: break;
        }
        const vec3 st[3] = {cam.position, cam.direction, cam.vup};
        for (int i = 0; i < 3; i++) {
            state_out[9 * k + 3 * i + 0] = st[i].x;
            state_out[9 * k + 3 * i + 1] = st[i].y;
            state_out[9 * k + 3 * i + 2] = st[i].z;
        }
    }
    if (cam_out) {
        cam_out->position = P(cam.position);
        cam_out->image_top_left = P(cam.image_top_left);
        cam_out->delta_x = P(u[0]);
        cam_out->delta_y = P(u[1]);
        cam_out->width = static_cast<int32_t>(cam.image_width);
        cam_out->height = static_cast<int32_t>(cam.image_height);
    }
}

/*
 * Renders the given global rows (rows[k], k < n_rows) of one frame; outputs are packed [n_rows][width].
 * radiance comes from recursive_ray_tracing (main.cpp:89); object_id / hit_mask / ray_count (any may be
 * NULL) come from walking the same chain with find_closest_hit + vec3::reflect. Returns the seconds spent
 * in the row loop (std::chrono around the loop, as main.cpp:326-330), or -1 on bad arguments.
 * time_radiance_only != 0 skips the chain walk even if id planes are given (used for timing).
 */
double ref_render_rows(const rtx_object* objs, int32_t n_objs, const rtx_camera* cam, int32_t max_depth,
                       const int32_t* rows, int32_t n_rows, int32_t n_threads,
                       double* radiance, uint32_t* rgba8, int32_t* object_id, uint8_t* hit_mask, uint8_t* ray_count,
                       uint64_t* total_rays)
{
    if (!objs && n_objs > 0) return -1;
    if (!cam || !rows || n_rows < 0) return -1;
    Scene scene = build_scene(objs, n_objs);
    const int W = cam->width;
    const vec3 top_left = V(cam->image_top_left), dx = V(cam->delta_x), dy = V(cam->delta_y), pos = V(cam->position);
    const bool walk = object_id || hit_mask || ray_count || total_rays;
    uint64_t rays = 0;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#endif
    auto t0 = std::chrono::high_resolution_clock::now();
    /* Work items are (row, 64-pixel column chunk) pairs, handed out dynamically: a bounded sample of a few rows of a
     * costly frame (bench.py) still feeds every hardware thread, and a row that crosses a dense part of the scene does
     * not serialise the tail. The per-pixel code below is untouched. */
    const int n_chunks = (W + kChunk - 1) / kChunk;
    const long long n_items = static_cast<long long>(n_rows) * n_chunks;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads) reduction(+ : rays)
    for (long long item = 0; item < n_items; item++) {
        const int k = static_cast<int>(item / n_chunks);
        const int j0 = static_cast<int>(item - static_cast<long long>(k) * n_chunks) * kChunk;
        const int j1 = j0 + kChunk < W ? j0 + kChunk : W;
        const int i = rows[k];
        for (int j = j0; j < j1; j++) {
            /* main.cpp:132-134 */
            auto pixel_center = top_left + dx * j + dy * i;
            auto cam_pixel = pos - pixel_center;
            ray cam_pixel_ray(cam_pixel, pos);
            const size_t px = static_cast<size_t>(k) * W + j;
            if (radiance || rgba8) {
                RGB c = recursive_ray_tracing(scene, cam_pixel_ray, max_depth); /* main.cpp:136 */
                if (radiance) {
                    radiance[3 * px + 0] = c.x;
                    radiance[3 * px + 1] = c.y;
                    radiance[3 * px + 2] = c.z;
                }
                if (rgba8) rgba8[px] = quantise_like_main(c);
            }
            if (walk) {
                ray r = cam_pixel_ray;
                int remaining = max_depth, n = 0, first = -1;
                for (;;) {
                    Collision col = find_closest_hit(scene, r);
                    if (n == 0) first = col.hit_object_index;
                    n++;
                    if (col.hit_object_index < 0 || remaining <= 0) break;
                    vec3 hit = r.get_origin() + r.get_direction() * col.distance; /* main.cpp:99  */
                    point3 start = hit + col.normal * .0001;                      /* main.cpp:111 */
                    vec3 dir = vec3::reflect(r.get_direction(), col.normal);      /* main.cpp:112 */
                    r = ray(dir, start);
                    remaining--;
                }
                if (object_id) object_id[px] = first;
                if (hit_mask) hit_mask[px] = first >= 0;
                if (ray_count) ray_count[px] = static_cast<uint8_t>(n);
                rays += n;
            }
        }
    }
    auto t1 = std::chrono::high_resolution_clock::now();
    if (total_rays) *total_rays = rays;
    return std::chrono::duration<double>(t1 - t0).count();
}

/* The reference's rt_scene itself (main.cpp:124-139), depth default 10, square frames only
 * (the frame buffer is transposed, main.cpp:243). radiance is [H][W][3]. Returns seconds, -1 if W != H. */
double ref_rt_scene(const rtx_object* objs, int32_t n_objs, const rtx_camera_desc* d, double* radiance)
{
    Camera cam;
    cam.aspect_ratio = d->aspect_ratio;
    cam.image_width = d->image_width;
    cam.movement_speed = 0.1;
    cam.vfov = d->vfov;
    cam.position = V(d->position);
    cam.lookat = V(d->lookat);
    cam.vup = V(d->vup);
    auto u = cam.init();
    const int W = static_cast<int>(cam.image_width), H = static_cast<int>(cam.image_height);
    if (W != H) return -1;
    Scene scene = build_scene(objs, n_objs);
    std::vector<std::vector<RGB>> frame_buffer(W, std::vector<RGB>(H, RGB(0, 0, 0))); /* main.cpp:243 */
    auto t0 = std::chrono::high_resolution_clock::now();
    rt_scene(u, scene, cam, frame_buffer);
    auto t1 = std::chrono::high_resolution_clock::now();
    for (int i = 0; i < H; i++)
        for (int j = 0; j < W; j++) {
            const RGB& c = frame_buffer[i][j];
            double* o = radiance + 3 * (static_cast<size_t>(i) * W + j);
            o[0] = c.x;
            o[1] = c.y;
            o[2] = c.z;
        }
    return std::chrono::duration<double>(t1 - t0).count();
}

/* The whole unmodified main() (main.cpp:144-397), headless: renders `frames` frames of the built-in
 * scene, then receives SDL_QUIT. Copies the SDL surface (what the quantise loop main.cpp:338-347 wrote)
 * into surface_out[h][w] and reports its size. The reference prints its own timing log to stdout. */
int ref_run_main(int32_t frames, uint32_t* surface_out, int32_t capacity_pixels, int32_t* w, int32_t* h)
{
    g_frames_before_quit = frames < 1 ? 0 : frames - 1;
    g_poll_calls = 0;
    g_surface = nullptr;
    char arg0[] = "ref";
    char* args[] = {arg0, nullptr};
    int rc = ref_main(1, args);
    if (!g_surface) return -1;
    if (w) *w = g_surface->w;
    if (h) *h = g_surface->h;
    const int n = g_surface->w * g_surface->h;
    if (surface_out && capacity_pixels >= n) std::memcpy(surface_out, g_surface->pixels, static_cast<size_t>(n) * 4);
    std::free(g_surface->pixels);
    delete g_surface;
    g_surface = nullptr;
    return rc;
}

/* main.cpp:343-345 applied to n RGB triples. */
void ref_quantise(const double* rgb, int64_t n, uint32_t* out)
{
    for (int64_t k = 0; k < n; k++) out[k] = quantise_like_main(RGB(rgb[3 * k], rgb[3 * k + 1], rgb[3 * k + 2]));
}

/* ---- function-level known-answer doors ---------------------------------------------------------- */

/* SceneGeometry::intersect of a single object (scene.cpp:4-78). */
void ref_intersect(const rtx_object* obj, const rtx_vec3* origin, const rtx_vec3* dir, double* distance, rtx_vec3* normal,
                   int32_t* hit)
{
    Scene s = build_scene(obj, 1);
    Collision c = s[0]->intersect(ray(V(*dir), V(*origin)));
    *distance = c.distance;
    *normal = P(c.normal);
    *hit = c.hit;
}

/* find_closest_hit (main.cpp:67-84). */
void ref_find_closest_hit(const rtx_object* objs, int32_t n, const rtx_vec3* origin, const rtx_vec3* dir, double* distance,
                          rtx_vec3* normal, int32_t* index)
{
    Scene s = build_scene(objs, n);
    Collision c = find_closest_hit(s, ray(V(*dir), V(*origin)));
    *distance = c.distance;
    *normal = P(c.normal);
    *index = c.hit_object_index;
}

/* recursive_ray_tracing for one ray (main.cpp:89-119). */
void ref_trace_ray(const rtx_object* objs, int32_t n, const rtx_vec3* origin, const rtx_vec3* dir, int32_t depth, rtx_vec3* rgb)
{
    Scene s = build_scene(objs, n);
    *rgb = P(recursive_ray_tracing(s, ray(V(*dir), V(*origin)), depth));
}

void ref_out_color(const rtx_vec3* v, rtx_vec3* rgb) { *rgb = P(out_color(V(*v))); }                    /* main.cpp:28-37 */
void ref_reflect(const rtx_vec3* v, const rtx_vec3* n, rtx_vec3* out) { *out = P(vec3::reflect(V(*v), V(*n))); } /* vec.cpp:51-57 */
double ref_diffuse(const rtx_vec3* pos, const rtx_vec3* n, const rtx_vec3* light) { return diffuse_shading(V(*pos), V(*n), V(*light)); }
double ref_specular(const rtx_vec3* pos, const rtx_vec3* n, const rtx_vec3* light, const rtx_vec3* view)
{
    return specular(V(*pos), V(*n), V(*light), V(*view));
}

} /* extern "C" */
