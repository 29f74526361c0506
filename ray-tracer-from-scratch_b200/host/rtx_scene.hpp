// rtx_scene.hpp — C++ host façade over the C ABI (include/rtx_b200.h).
//
// Mirrors the reference's scene interface so that its main loop body (main.cpp:329-347) becomes two calls:
//
//     rtx::rt_scene(u, scene, cam, frame_buffer);            // was rt_scene(...)            main.cpp:329
//     rtx::update_surface(frame_buffer, pixels, pitch);      // was the quantise loop       main.cpp:338-347
//
// Same class names, constructor argument order and defaults as scene.h / vec.h (Material's metallic-before-
// ambient order, DEFAULT_MAT's positional quirk, Wall normalising its normal, ray(direction, origin), Camera
// fields and init() returning {pixel_delta_x, pixel_delta_y}). The reference's members are private without
// getters (scene.h:64-67,77-78), so here every geometry can DESCRIBE itself as the rtx_object POD the C ABI
// takes: the description is the source of truth. No pixel is computed on the host: there is no CPU fallback,
// a missing GPU is a thrown std::runtime_error.
#pragma once
#include <cmath>
#include <cstring>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <algorithm>
#include <condition_variable>
#include <exception>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "rtx_b200.h"

namespace rtx {

// ---- vec3 (vec.h:12-40): three doubles, the reference's operation order -------------------------------------
class vec3 {
public:
    double x, y, z;
    vec3() : x{0}, y{0}, z{0} {}
    vec3(double xv, double yv, double zv) : x{xv}, y{yv}, z{zv} {}
    double length_squared() const { return x * x + y * y + z * z; }
    double length() const { return std::sqrt(length_squared()); }
    vec3 operator/(double t) const { return {x / t, y / t, z / t}; }
    vec3 normalize() const { return *this / length(); }                       // three divides (vec.cpp:21-24)
    vec3 operator+(const vec3& o) const { return {x + o.x, y + o.y, z + o.z}; }
    vec3 operator-(const vec3& o) const { return {x - o.x, y - o.y, z - o.z}; }
    vec3 operator-() const { return {-x, -y, -z}; }
    vec3 operator*(const vec3& o) const { return {x * o.x, y * o.y, z * o.z}; }
    vec3 operator*(double d) const { return {x * d, y * d, z * d}; }
    static double dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
    static vec3 cross(const vec3& u, const vec3& v) { return {u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x}; }
    static vec3 linear_interp(const vec3& a, const vec3& b, double d) { return {a.x + d * (b.x - a.x), a.y + d * (b.y - a.y), a.z + d * (b.z - a.z)}; }
    static vec3 reflect(const vec3& v, const vec3& normal)
    {
        const vec3 nn = normal.normalize(), nv = v.normalize();
        return nv - nn * (2 * dot(nv, nn));
    }
    rtx_vec3 pod() const { return rtx_vec3{x, y, z}; }
};
using point3 = vec3;
using RGB = vec3;

// ---- ray (scene.h:5-25): constructor order is (direction, origin) -----------------------------------------------
class ray {
    vec3 direction;
    point3 origin;
public:
    ray() {}
    ray(const vec3& dir, const point3& org) : direction{dir}, origin{org} {}
    point3 at(double t) const { return origin + direction * t; }
    vec3 get_direction() const { return direction; }
    vec3 get_origin() const { return origin; }
};

// ---- Material (scene.h:35-49) ------------------------------------------------------------------------------------------
struct Material {
    RGB color;
    double ambient, metallic, diffuse, specular, specular_exponent;
    Material(RGB color, double metallic = .5, double ambient = .1, double diffuse = .9, double specular = .4, double specular_exponent = 50)
        : color{color}, ambient{ambient}, metallic{metallic}, diffuse{diffuse}, specular{specular}, specular_exponent{specular_exponent} {}
    rtx_material pod() const { return rtx_material{color.pod(), ambient, metallic, diffuse, specular, specular_exponent}; }
};
// DEFAULT_MAT (scene.h:3) binds positionally to metallic .9, ambient .9, diffuse .3, specular 30, exponent 50.
inline Material default_mat() { return Material(RGB(1, 1, 1), .9, .9, .3, 30); }

// ---- SceneGeometry / Sphere / Wall (scene.h:51-84) -------------------------------------------------------------
class SceneGeometry {
    Material mat;
public:
    explicit SceneGeometry(Material m) : mat(m) {}
    virtual ~SceneGeometry() {}
    Material get_material() const { return mat; }
    virtual rtx_object describe() const = 0;     // replaces the virtual intersect(): the GPU does the intersecting
};

class Sphere : public SceneGeometry {
    point3 center;
    double radius;
public:
    Sphere(Material m = default_mat(), point3 center = point3(0, 0, 0), double radius = 1.0) : SceneGeometry{m}, center{center}, radius{radius} {}
    rtx_object describe() const override
    {
        rtx_object o{};
        o.kind = RTX_SPHERE;
        o.mat = get_material().pod();
        o.p = center.pod();
        o.a = radius;
        return o;
    }
};

class Wall : public SceneGeometry {
    point3 position;   // a corner of the rectangle
    vec3 normal;       // stored as given; the library normalises it like the reference ctor (scene.h:71)
    double length, width;
public:
    Wall(Material m = default_mat(), point3 position = point3(0, 0, 0), vec3 normal = vec3(0, 0, 0), double length = 1.0, double width = 1.0)
        : SceneGeometry{m}, position{position}, normal{normal}, length{length}, width{width} {}
    rtx_object describe() const override
    {
        rtx_object o{};
        o.kind = RTX_WALL;
        o.mat = get_material().pod();
        o.p = position.pod();
        o.n = normal.pod();
        o.a = length;
        o.b = width;
        return o;
    }
};

// EXTENSION — no reference class (README.md:21 names a sprint-2 `Box`; no code survives): an axis-aligned box given by
// its minimum corner and extents; six Wall-like faces, ONE object id (include/rtx_b200.h, RTX_BOX).
class Box : public SceneGeometry {
    point3 position;   // minimum corner
    vec3 size;         // extents along x, y, z
public:
    Box(Material m = default_mat(), point3 position = point3(0, 0, 0), vec3 size = vec3(1, 1, 1)) : SceneGeometry{m}, position{position}, size{size} {}
    rtx_object describe() const override
    {
        rtx_object o{};
        o.kind = RTX_BOX;
        o.mat = get_material().pod();
        o.p = position.pod();
        o.n = size.pod();
        return o;
    }
};

using Scene = std::vector<std::unique_ptr<SceneGeometry>>;

// ---- Camera (scene.h:86-112, scene.cpp:80-165) ---------------------------------------------------------------------
class Camera {
    vec3 forward_vec() const { return direction.normalize(); }
    vec3 right_vec() const { return vec3::cross(direction, vup).normalize(); }
    vec3 up_vec() const { return vec3::cross(right_vec(), direction).normalize(); }             // scene.cpp:116-119
public:
    vec3 direction, image_top_left;
    double movement_speed = 0.1, aspect_ratio = 1.0, image_width = 640, image_height = 0, focal_length = 0, vfov = 90;
    point3 position = point3(0, 0, -1);
    point3 lookat = point3(0, 0, 0);
    vec3 vup = vec3(0, 1, 0);

    // Camera::init through the library's host code (rtx_camera_init): one implementation of the 3.14 / int()
    // quirks. Returns {pixel_delta_x, pixel_delta_y}; like the reference it is NOT re-run by the move methods.
    std::vector<vec3> init()
    {
        rtx_camera_desc d{position.pod(), lookat.pod(), vup.pod(), vfov, aspect_ratio, image_width};
        rtx_camera c{};
        if (rtx_camera_init(&d, &c) != RTX_OK) throw std::runtime_error("rtx_camera_init failed");
        image_height = c.height;
        focal_length = (position - lookat).length();
        direction = (position - lookat).normalize();
        image_top_left = vec3(c.image_top_left.x, c.image_top_left.y, c.image_top_left.z);
        return {vec3(c.delta_x.x, c.delta_x.y, c.delta_x.z), vec3(c.delta_y.x, c.delta_y.y, c.delta_y.z)};
    }
    void forward() { position = position + forward_vec() * movement_speed; }     // scene.cpp:121-123
    void backward() { position = position - forward_vec() * movement_speed; }
    void right() { position = position + right_vec() * movement_speed; }
    void left() { position = position - right_vec() * movement_speed; }
    // The mouse look the reference implements but leaves commented out in its loop (scene.cpp:137-165,
    // main.cpp:319-323). They change direction and vup only; init() ignores `direction`, so like the reference a
    // rotation shows in the image only through later right()/left() moves unless the caller re-aims lookat itself.
    void rotate_left_right(double angle)
    {
        // yaw about z (scene.cpp:137-145): the planar part of `direction` keeps its length and turns by `angle`, z stays
        const double planar = vec3(direction.x, direction.y, 0).length();
        const double yaw = std::atan2(direction.y, direction.x) + angle;
        direction = vec3(std::cos(yaw) * planar, std::sin(yaw) * planar, direction.z);
        vup = up_vec();
    }
    void rotate_up_down(double angle)
    {
        // pitch (scene.cpp:147-165): the result is a UNIT vector over the old heading. Past the zenith the pitch stays
        // what it was; past the nadir the reference takes MINUS the old pitch (scene.cpp:156) — kept as is.
        const vec3 flat(direction.x, direction.y, 0);
        const double pitch = std::atan2(direction.z, flat.length());
        double target = pitch + angle;
        if (target > M_PI / 2) target = pitch;
        if (target < -M_PI / 2) target = -pitch;
        const vec3 heading = flat.normalize() * std::cos(target);
        direction = vec3(heading.x, heading.y, std::sin(target));
        vup = up_vec();
    }

    rtx_camera pod(const std::vector<vec3>& u) const
    {
        rtx_camera c{};
        c.position = position.pod();
        c.image_top_left = image_top_left.pod();
        c.delta_x = u.at(0).pod();
        c.delta_y = u.at(1).pod();
        c.width = static_cast<int32_t>(image_width);
        c.height = static_cast<int32_t>(image_height);
        return c;
    }
};

// The synthetic scene of configs C3 / C4: splitmix64 seeded 0xB200, U() = (next() >> 11) * 2^-53,
// 10 000 spheres then 64 walls drawn in the survey's order. Same objects as scene.py::synthetic_scene.
inline Scene synthetic_scene(int n_spheres = 10000, int n_walls = 64, uint64_t seed = 0xB200)
{
    uint64_t state = seed;
    auto next = [&state]() {
        state += 0x9E3779B97F4A7C15ull;
        uint64_t z = state;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    auto U = [&next](double a, double b) { return a + (b - a) * (static_cast<double>(next() >> 11) * (1.0 / 9007199254740992.0)); };
    Scene scene;
    for (int k = 0; k < n_spheres; k++) {
        const double cx = U(4, 64), cy = U(-32, 32), cz = U(-8, 24), r = U(.1, .6);
        const double R = U(.1, 1), G = U(.1, 1), B = U(.1, 1), metallic = U(0, .8);
        scene.push_back(std::make_unique<Sphere>(Material(RGB(R, G, B), metallic), point3(cx, cy, cz), r));
    }
    for (int k = 0; k < n_walls; k++) {
        const double px = U(4, 64), py = U(-32, 32), pz = U(-8, 8), phi = U(0, 6.283185307179586), nz = U(-.5, .5);
        const double length = U(1, 6), width = U(1, 6);
        const double R = U(.1, 1), G = U(.1, 1), B = U(.1, 1), metallic = U(0, .8);
        scene.push_back(std::make_unique<Wall>(Material(RGB(R, G, B), metallic), point3(px, py, pz), vec3(std::cos(phi), std::sin(phi), nz), length, width));
    }
    return scene;
}

// ---- the renderer: one rtx_ctx -------------------------------------------------------------------------------------------
class Renderer {
    friend class ShardedRenderer;
    rtx_ctx* ctx = nullptr;
    std::vector<rtx_object> uploaded;      // what the device holds: compared by content before every frame
    bool have_uploaded = false;
    void check(int rc, const char* what) const
    {
        if (rc != RTX_OK) throw std::runtime_error(std::string(what) + ": " + rtx_status_string(rc) + ": " + rtx_last_error(ctx));
    }
public:
    rtx_params params;
    rtx_stats stats{};
    explicit Renderer(int device = 0)
    {
        rtx_default_params(&params);
        const int rc = rtx_create(&ctx, device);
        if (rc != RTX_OK) throw std::runtime_error(std::string("rtx_create: ") + rtx_status_string(rc) + ": " + rtx_last_error(nullptr));
    }
    ~Renderer() { rtx_destroy(ctx); }
    Renderer(const Renderer&) = delete;
    Renderer& operator=(const Renderer&) = delete;

    rtx_ctx* raw() const { return ctx; }
    void enable_peer_access(int peer_device) { check(rtx_enable_peer_access(ctx, peer_device), "rtx_enable_peer_access"); }
    // rtx_render with caller-built structures (the sharded host uses it); stats land in this->stats.
    void render_raw(const rtx_camera* cams, int n, const rtx_params& p, const rtx_outputs& out)
    {
        check(rtx_render(ctx, cams, n, &p, &out, &stats), "rtx_render");
    }

    void set_scene(const Scene& scene)
    {
        std::vector<rtx_object> objs;
        objs.reserve(scene.size());
        for (const auto& g : scene) objs.push_back(g->describe());
        upload(std::move(objs));
    }

    // The reference reads the scene vector afresh on every frame (main.cpp:329), so an element replaced or edited in
    // place must show in the next frame: the scene is re-described per call (O(N), nothing next to a frame) and
    // uploaded again whenever its CONTENT differs from what the device holds — not merely its address or size.
    void sync_scene(const Scene& scene) { sync_described(describe_scene(scene)); }

    static std::vector<rtx_object> describe_scene(const Scene& scene)
    {
        std::vector<rtx_object> objs;
        objs.reserve(scene.size());
        for (const auto& g : scene) objs.push_back(g->describe());
        return objs;
    }
    // The same with a scene that has already been described (the multi-GPU host describes it once for all GPUs).
    void sync_described(const std::vector<rtx_object>& objs)
    {
        if (have_uploaded && objs.size() == uploaded.size() &&
            (objs.empty() || std::memcmp(objs.data(), uploaded.data(), objs.size() * sizeof(rtx_object)) == 0))
            return;
        upload(std::vector<rtx_object>(objs));
    }

    // rt_scene (main.cpp:124-139): fills frame_buffer[i][j] (row i, column j) with the radiance of every pixel.
    // The scene is uploaded on first use and whenever its content has changed (sync_scene).
    void rt_scene(const std::vector<vec3>& u, const Scene& scene, const Camera& cam, std::vector<std::vector<RGB>>& frame_buffer)
    {
        sync_scene(scene);
        const rtx_camera c = cam.pod(u);
        const size_t W = c.width, H = c.height;
        if (frame_buffer.size() < H) throw std::out_of_range("frame_buffer has fewer rows than image_height");   // .at(i), main.cpp:136
        radiance.resize(W * H * 3);
        rtx_outputs out{};
        out.memory = RTX_MEM_HOST;
        out.radiance_f64 = radiance.data();
        check(rtx_render(ctx, &c, 1, &params, &out, &stats), "rtx_render");
        for (size_t i = 0; i < H; i++) {
            std::vector<RGB>& row = frame_buffer[i];
            if (row.size() < W) throw std::out_of_range("frame_buffer row shorter than image_width");
            const double* src = &radiance[i * W * 3];
            for (size_t j = 0; j < W; j++) row[j] = RGB(src[3 * j], src[3 * j + 1], src[3 * j + 2]);
        }
    }

    // One call for both stages when only the 8-bit surface is wanted (what main.cpp:329-347 produces):
    // pixels[i * pitch/4 + j] = RGBA8888 word. pitch in bytes, as SDL_Surface::pitch.
    void render_surface(const std::vector<vec3>& u, const Scene& scene, const Camera& cam, uint32_t* pixels, int pitch)
    {
        sync_scene(scene);
        const rtx_camera c = cam.pod(u);
        const size_t W = c.width, H = c.height;
        rtx_outputs out{};
        out.memory = RTX_MEM_HOST;
        if (static_cast<size_t>(pitch) == W * 4) {
            out.rgba8 = pixels;
            check(rtx_render(ctx, &c, 1, &params, &out, &stats), "rtx_render");
        } else {
            surface.resize(W * H);
            out.rgba8 = surface.data();
            check(rtx_render(ctx, &c, 1, &params, &out, &stats), "rtx_render");
            for (size_t i = 0; i < H; i++)
                for (size_t j = 0; j < W; j++) pixels[i * (pitch / 4) + j] = surface[i * W + j];
        }
    }

    // The quantise loop (main.cpp:338-347) on a frame buffer the caller already holds (with params.tonemap set: the
    // tone-map extension followed by the same pack).
    void update_surface(const std::vector<std::vector<RGB>>& frame_buffer, size_t H, size_t W, uint32_t* pixels, int pitch)
    {
        radiance.resize(W * H * 3);
        for (size_t i = 0; i < H; i++)
            for (size_t j = 0; j < W; j++) {
                const RGB& v = frame_buffer.at(i).at(j);
                double* dst = &radiance[(i * W + j) * 3];
                dst[0] = v.x; dst[1] = v.y; dst[2] = v.z;
            }
        surface.resize(W * H);
        if (params.tonemap == RTX_TONEMAP_REINHARD)   // extension: global operator before the pack (off by default)
            check(rtx_tonemap(ctx, nullptr, radiance.data(), static_cast<int64_t>(W * H), 1, &params, surface.data(), RTX_MEM_HOST, nullptr, &stats),
                  "rtx_tonemap");
        else
            check(rtx_quantise(ctx, nullptr, radiance.data(), static_cast<int64_t>(W * H), params.quantise_mode, surface.data(), RTX_MEM_HOST, &stats),
                  "rtx_quantise");
        for (size_t i = 0; i < H; i++)
            for (size_t j = 0; j < W; j++) pixels[i * (pitch / 4) + j] = surface[i * W + j];
    }

private:
    void upload(std::vector<rtx_object>&& objs)
    {
        have_uploaded = false;
        check(rtx_set_scene(ctx, objs.data(), static_cast<int32_t>(objs.size())), "rtx_set_scene");
        uploaded = std::move(objs);
        have_uploaded = true;
    }
    std::vector<double> radiance;
    std::vector<uint32_t> surface;
};

// ---- several GPUs, one process -----------------------------------------------------------------------------------
// The reference's frame loop (main.cpp:250-375) is one thread calling rt_scene once per frame. Here the same call fans
// out: one rtx_ctx and one persistent host thread per GPU, the frame's rows dealt to the GPUs in cyclic bands (band b -> GPU
// b mod N; rtx_params.band_rows / n_ranks / rank), and every trace kernel stores its finished pixels AT THEIR GLOBAL
// POSITION in one shared surface — no gather step, no reassembly:
//   * surface(): a pinned, mapped host surface, the stand-in for SDL's surface->pixels (main.cpp:193,344); each GPU
//     writes its rows over its own PCIe link (zero copy);
//   * render_to_device(): the frame in GPU 0's memory instead, peers storing over NVLink (rtx_enable_peer_access).
// A device may be listed more than once ({0, 0}: two contexts on one GPU) — the same code path, useful for testing.
class ShardedRenderer {
    std::vector<std::unique_ptr<Renderer>> gpus;
    std::vector<int> devices;
    uint32_t* host_surface = nullptr;      // pinned + mapped, W * H words
    uint32_t* host_alias = nullptr;        // what the kernels dereference
    uint32_t* device_frame = nullptr;      // on devices[0]
    size_t surface_words = 0, device_words = 0;
    int band_rows;

    // One persistent host thread per GPU (the contexts are created here, driven there): a frame is one job handed to all of
    // them; the caller's thread waits. Exceptions travel back to the caller.
    std::vector<std::thread> workers;
    std::mutex mtx;
    std::condition_variable cv_job, cv_done;
    std::function<void(int, Renderer&)> job;
    std::vector<std::exception_ptr> errors;
    int generation = 0, pending = 0;
    bool stopping = false;

    void worker_loop(int g)
    {
        int seen = 0;
        for (;;) {
            std::function<void(int, Renderer&)> mine;
            {
                std::unique_lock<std::mutex> lk(mtx);
                cv_job.wait(lk, [&] { return stopping || generation != seen; });
                if (stopping) return;
                seen = generation;
                mine = job;
            }
            try { mine(g, *gpus[g]); } catch (...) { errors[g] = std::current_exception(); }
            std::lock_guard<std::mutex> lk(mtx);
            if (--pending == 0) cv_done.notify_all();
        }
    }
    template <class F>
    void on_every_gpu(F&& body)
    {
        {
            std::unique_lock<std::mutex> lk(mtx);
            job = std::forward<F>(body);
            std::fill(errors.begin(), errors.end(), std::exception_ptr());
            pending = static_cast<int>(gpus.size());
            generation++;
            cv_job.notify_all();
            cv_done.wait(lk, [&] { return pending == 0; });
        }
        for (auto& e : errors)
            if (e) std::rethrow_exception(e);
    }
    void render_into(const std::vector<vec3>& u, const Scene& scene, const Camera& cam, uint32_t* frame_alias)
    {
        const rtx_camera c = cam.pod(u);
        const std::vector<rtx_object> described = Renderer::describe_scene(scene);     // once per frame, for all GPUs
        on_every_gpu([&](int g, Renderer& r) {
            r.sync_described(described);
            rtx_params p = params;
            p.band_rows = band_rows;
            p.n_ranks = static_cast<int32_t>(gpus.size());
            p.rank = g;
            rtx_outputs out{};
            out.memory = RTX_MEM_DEVICE;
            out.frame_mode = RTX_FRAME_STORE;
            out.frame_rgba8 = frame_alias;
            r.render_raw(&c, 1, p, out);      // returns when this GPU's kernel is complete: its pixels are in place
        });
    }
public:
    rtx_params params;                     // applied to every GPU (band fields are set per GPU)
    explicit ShardedRenderer(const std::vector<int>& device_list, int band_rows = 4) : devices(device_list), band_rows(band_rows)
    {
        if (devices.empty()) throw std::invalid_argument("ShardedRenderer: no devices");
        rtx_default_params(&params);
        for (int d : devices) gpus.push_back(std::make_unique<Renderer>(d));
        for (size_t g = 1; g < gpus.size(); g++) gpus[g]->enable_peer_access(devices[0]);
        errors.resize(gpus.size());
        for (size_t g = 0; g < gpus.size(); g++) workers.emplace_back([this, g] { worker_loop(static_cast<int>(g)); });
    }
    ~ShardedRenderer()
    {
        {
            std::lock_guard<std::mutex> lk(mtx);
            stopping = true;
        }
        cv_job.notify_all();
        for (auto& t : workers) t.join();
        if (host_surface) rtx_host_free(gpus[0]->raw(), host_surface);
        if (device_frame) rtx_buffer_free(gpus[0]->raw(), device_frame);
    }
    size_t size() const { return gpus.size(); }
    Renderer& gpu(size_t g) { return *gpus.at(g); }

    // The shared host surface for a W x H frame (allocated on first use / size change), RGBA8888 words, pitch = W * 4.
    uint32_t* surface(size_t W, size_t H)
    {
        if (surface_words != W * H) {
            if (host_surface) rtx_host_free(gpus[0]->raw(), host_surface);
            host_surface = host_alias = nullptr;
            void* p = nullptr;
            gpus[0]->check(rtx_host_alloc(gpus[0]->raw(), W * H * 4, &p), "rtx_host_alloc");
            host_surface = static_cast<uint32_t*>(p);
            void* d = nullptr;
            gpus[0]->check(rtx_host_device_pointer(gpus[0]->raw(), p, &d), "rtx_host_device_pointer");
            host_alias = static_cast<uint32_t*>(d);
            surface_words = W * H;
        }
        return host_surface;
    }

    // rt_scene + quantise (main.cpp:329-347) over all GPUs into the shared host surface; returns it.
    const uint32_t* render_surface(const std::vector<vec3>& u, const Scene& scene, const Camera& cam)
    {
        const rtx_camera c = cam.pod(u);
        surface(static_cast<size_t>(c.width), static_cast<size_t>(c.height));
        render_into(u, scene, cam, host_alias);
        return host_surface;
    }

    // The same with the assembled frame left in the first GPU's memory (device pointer, W * H words).
    const uint32_t* render_to_device(const std::vector<vec3>& u, const Scene& scene, const Camera& cam)
    {
        const rtx_camera c = cam.pod(u);
        const size_t words = static_cast<size_t>(c.width) * c.height;
        if (device_words != words) {
            if (device_frame) rtx_buffer_free(gpus[0]->raw(), device_frame);
            device_frame = nullptr;
            void* p = nullptr;
            gpus[0]->check(rtx_buffer_alloc(gpus[0]->raw(), words * 4, &p), "rtx_buffer_alloc");
            device_frame = static_cast<uint32_t*>(p);
            device_words = words;
        }
        render_into(u, scene, cam, device_frame);
        return device_frame;
    }

    // Copies a frame returned by render_to_device into host memory (blocking).
    void read_device_frame(const uint32_t* device_frame_ptr, uint32_t* host, size_t words)
    {
        gpus[0]->check(rtx_buffer_read(gpus[0]->raw(), device_frame_ptr, host, words * 4), "rtx_buffer_read");
    }

    // Sum over the GPUs of the last frame's statistics; raytracing_ms is the slowest GPU's.
    rtx_stats stats() const
    {
        rtx_stats s{};
        for (const auto& g : gpus) {
            s.total_rays += g->stats.total_rays;
            s.sphere_tests += g->stats.sphere_tests;
            s.wall_tests += g->stats.wall_tests;
            s.over_range_pixels += g->stats.over_range_pixels;
            s.launches += g->stats.launches;
            s.raytracing_ms = std::max(s.raytracing_ms, g->stats.raytracing_ms);
            s.total_ms = std::max(s.total_ms, g->stats.total_ms);
            s.max_luminance = std::max(s.max_luminance, g->stats.max_luminance);
        }
        return s;
    }
};

// Free-function form with the reference's exact signature shape (main.cpp:124-125); uses one process-wide renderer.
inline Renderer& default_renderer()
{
    static Renderer r(0);
    return r;
}
inline void rt_scene(std::vector<vec3> u, const Scene& scene, const Camera& cam, std::vector<std::vector<RGB>>& frame_buffer)
{
    default_renderer().rt_scene(u, scene, cam, frame_buffer);
}

}  // namespace rtx
