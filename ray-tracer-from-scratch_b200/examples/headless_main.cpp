// headless_main.cpp — the reference's main loop (main.cpp:144-397) without SDL, on the GPU path.
//
// Same camera and scene set-up (main.cpp:146-163), same per-frame sequence — key events move the camera
// (main.cpp:262-306; init() is NOT re-run after a move, exactly like the reference), then
//     rt_scene(u, scene, cam, frame_buffer)            main.cpp:329
//     quantise loop into the RGBA8888 surface          main.cpp:338-347
// — and the same timing log at exit (main.cpp:384-392) with the reference's stage names, plus the device-side
// CUDA-event times. The window/renderer/texture calls are replaced by a binary PPM (or raw RGBA8888) file.
//
//   rtx_headless [--width 640] [--aspect 1] [--depth 10] [--frames 3] [--keys wwad] [--out frame.ppm] [--raw frame.rgba]
//                [--png frame.png] [--sun 1] [--tonemap 1] [--box 1] [--devices 0,1,...] [--band-rows 4] [--scene default|synthetic]
//                [--accel 1] [--to-device 1] [--pixel-order 0|1|2]   (RTX_ORDER_AUTO / _SCAN / _COST)
// --devices: more than one entry renders every frame on several GPUs from THIS process (rtx::ShardedRenderer: one context
// and one host thread per GPU, cyclic row bands, every kernel storing its pixels straight into one pinned host surface);
// a device may be repeated (0,0 = two contexts on one GPU). --scene synthetic = the 10 064-object scene of configs C3/C4.
// --to-device 1 (with --devices): the frame is assembled in the FIRST GPU's memory instead (peers store over NVLink) and
// read back once at the end.
// --keys: one key event per frame — w s a d as in the reference (main.cpp:262-306); j l i k = the mouse look the reference
// leaves commented out (rotate_left_right(+-0.05), rotate_up_down(+-0.05), main.cpp:319-323); anything else = no event.
// --sun / --tonemap / --box switch on this repo's EXTENSIONS (default-off; include/rtx_b200.h): the sun of main.cpp:18-19
// as a directional light, Reinhard's operator (saturating pack) in the surface update, a red box added to the scene.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <numeric>
#include <string>

#include "rtx_scene.hpp"

using namespace rtx;

// Minimal PNG writer (8-bit RGB, "stored" deflate blocks: no zlib needed).
static void put32(std::vector<unsigned char>& v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back(static_cast<unsigned char>(x >> s)); }
static uint32_t crc32_of(const unsigned char* p, size_t n)
{
    static uint32_t table[256];
    if (!table[1]) for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}
static void png_chunk(FILE* f, const char tag[4], const std::vector<unsigned char>& data)
{
    std::vector<unsigned char> buf(tag, tag + 4);
    buf.insert(buf.end(), data.begin(), data.end());
    std::vector<unsigned char> len, crc;
    put32(len, static_cast<uint32_t>(data.size()));
    put32(crc, crc32_of(buf.data(), buf.size()));
    std::fwrite(len.data(), 1, 4, f);
    std::fwrite(buf.data(), 1, buf.size(), f);
    std::fwrite(crc.data(), 1, 4, f);
}
static void write_png(const std::string& path, const std::vector<uint32_t>& surface, int W, int H)
{
    std::vector<unsigned char> raw;
    raw.reserve(static_cast<size_t>(H) * (W * 3 + 1));
    for (int i = 0; i < H; i++) {
        raw.push_back(0);   // filter type 0
        for (int j = 0; j < W; j++) {
            const uint32_t p = surface[static_cast<size_t>(i) * W + j];
            raw.push_back(static_cast<unsigned char>(p >> 24)); raw.push_back(static_cast<unsigned char>(p >> 16)); raw.push_back(static_cast<unsigned char>(p >> 8));
        }
    }
    std::vector<unsigned char> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;
    for (size_t pos = 0; pos < raw.size() || pos == 0; pos += 65535) {
        const size_t n = std::min<size_t>(65535, raw.size() - pos);
        z.push_back(pos + n >= raw.size() ? 1 : 0);
        z.push_back(static_cast<unsigned char>(n)); z.push_back(static_cast<unsigned char>(n >> 8));
        z.push_back(static_cast<unsigned char>(~n)); z.push_back(static_cast<unsigned char>((~n) >> 8));
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
        if (raw.empty()) break;
    }
    for (unsigned char c : raw) { a = (a + c) % 65521u; b = (b + a) % 65521u; }
    put32(z, (b << 16) | a);
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot open " + path);
    const unsigned char sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    std::fwrite(sig, 1, 8, f);
    std::vector<unsigned char> ihdr;
    put32(ihdr, static_cast<uint32_t>(W)); put32(ihdr, static_cast<uint32_t>(H));
    const unsigned char tail[5] = {8, 2, 0, 0, 0};
    ihdr.insert(ihdr.end(), tail, tail + 5);
    png_chunk(f, "IHDR", ihdr);
    png_chunk(f, "IDAT", z);
    png_chunk(f, "IEND", {});
    std::fclose(f);
}

static void write_outputs(const std::vector<uint32_t>& surface, int W, int H, const std::string& out_ppm, const std::string& out_raw,
                          const std::string& out_png)
{
    if (!out_ppm.empty()) {
        FILE* f = std::fopen(out_ppm.c_str(), "wb");
        if (!f) throw std::runtime_error("cannot open " + out_ppm);
        std::fprintf(f, "P6\n%d %d\n255\n", W, H);
        for (uint32_t p : surface) {
            const unsigned char rgb[3] = {static_cast<unsigned char>(p >> 24), static_cast<unsigned char>(p >> 16), static_cast<unsigned char>(p >> 8)};
            std::fwrite(rgb, 1, 3, f);
        }
        std::fclose(f);
    }
    if (!out_raw.empty()) {
        FILE* f = std::fopen(out_raw.c_str(), "wb");
        if (!f) throw std::runtime_error("cannot open " + out_raw);
        std::fwrite(surface.data(), 4, surface.size(), f);
        std::fclose(f);
    }
    if (!out_png.empty()) write_png(out_png, surface, W, H);
}

// One key event (main.cpp:262-306); init() is NOT re-run after a move, exactly like the reference.
static void apply_key(Camera& cam, char key)
{
    switch (key) {
        case 'w': cam.forward(); break;
        case 's': cam.backward(); break;
        case 'a': cam.left(); break;
        case 'd': cam.right(); break;
        // the mouse look of main.cpp:319-323 (commented out there) at full deflection, x_input / y_input = -+1
        case 'j': cam.rotate_left_right(0.05); break;
        case 'l': cam.rotate_left_right(-0.05); break;
        case 'i': cam.rotate_up_down(0.05); break;
        case 'k': cam.rotate_up_down(-0.05); break;
        default: break;
    }
}

int main(int argc, char* argv[])
{
    int width = 640, depth = 10, frames = 3;
    double aspect = 1.0;   // ASPECT_RATIO = 4/3 is integer division = 1 in the reference (main.cpp:25)
    std::string keys, out_ppm = "frame.ppm", out_raw, out_png;
    bool ext_sun = false, ext_tonemap = false, ext_box = false, accel = false, to_device = false;
    int pixel_order = RTX_ORDER_AUTO;
    std::vector<int> devices = {0};
    int band_rows = 4;
    std::string scene_name = "default";
    for (int k = 1; k + 1 < argc; k += 2) {
        const std::string a = argv[k];
        if (a == "--width") width = std::atoi(argv[k + 1]);
        else if (a == "--aspect") aspect = std::atof(argv[k + 1]);
        else if (a == "--depth") depth = std::atoi(argv[k + 1]);
        else if (a == "--frames") frames = std::atoi(argv[k + 1]);
        else if (a == "--keys") keys = argv[k + 1];
        else if (a == "--out") out_ppm = argv[k + 1];
        else if (a == "--raw") out_raw = argv[k + 1];
        else if (a == "--png") out_png = argv[k + 1];
        else if (a == "--sun") ext_sun = std::atoi(argv[k + 1]) != 0;
        else if (a == "--tonemap") ext_tonemap = std::atoi(argv[k + 1]) != 0;
        else if (a == "--box") ext_box = std::atoi(argv[k + 1]) != 0;
        else if (a == "--accel") accel = std::atoi(argv[k + 1]) != 0;
        else if (a == "--pixel-order") pixel_order = std::atoi(argv[k + 1]);
        else if (a == "--to-device") to_device = std::atoi(argv[k + 1]) != 0;
        else if (a == "--band-rows") band_rows = std::atoi(argv[k + 1]);
        else if (a == "--scene") scene_name = argv[k + 1];
        else if (a == "--devices") {
            devices.clear();
            const std::string list = argv[k + 1];
            for (size_t pos = 0; pos <= list.size();) {
                const size_t comma = std::min(list.find(',', pos), list.size());
                devices.push_back(std::atoi(list.substr(pos, comma - pos).c_str()));
                pos = comma + 1;
            }
        }
        else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    try {
        Camera cam;                                   // main.cpp:146-154
        cam.aspect_ratio = aspect;
        cam.image_width = width;
        cam.movement_speed = 0.1;
        cam.vfov = 90;
        cam.position = point3(0, 0, 0);
        cam.lookat = point3(-1, 0, 0);
        cam.vup = vec3(0, 0, -1);
        auto u = cam.init();

        Scene scene;                                  // main.cpp:156-163
        if (scene_name == "synthetic") {
            scene = synthetic_scene();                // configs C3 / C4
        } else {
            scene.push_back(std::make_unique<Sphere>(Material(RGB(0, 1, 0), 0.5), point3(1.5, 0, 0), .5));
            scene.push_back(std::make_unique<Wall>(Material(RGB(0, 0, 1)), point3(3.0, 2, 0), vec3(0, -1, 0), 1, 1));
            scene.push_back(std::make_unique<Wall>(Material(RGB(0, 1, 0)), point3(3.0, -3, 0), vec3(0, 1, 0), 2, 2));
        }

        const int H = static_cast<int>(cam.image_height), W = width;
        std::vector<uint32_t> surface(static_cast<size_t>(W) * H);
        const int pitch = W * 4;
        std::vector<std::vector<RGB>> frame_buffer(H, std::vector<RGB>(W, RGB(0, 0, 0)));   // [row][column]

        if (ext_box) scene.push_back(std::make_unique<Box>(Material(RGB(0.9, 0.2, 0.2), 0.6), point3(2.5, -1.0, -0.8), vec3(1.0, 1.2, 0.9)));

        if (devices.size() > 1) {
            // several GPUs: the two calls of the single-GPU loop below become ONE (trace + quantise fused, pixels stored by the
            // kernels into the shared surface); the log keeps the reference's stage names
            ShardedRenderer sharded(devices, band_rows);
            sharded.params.max_depth = depth;
            if (ext_sun) sharded.params.sun_enabled = 1;
            if (accel) sharded.params.accel = RTX_ACCEL_GRID;
            sharded.params.pixel_order = pixel_order;
            std::vector<int64_t> rt_times;
            const uint32_t* pixels = nullptr;
            for (int frame = 0; frame < frames; frame++) {
                if (frame < static_cast<int>(keys.size())) apply_key(cam, keys[frame]);
                auto t0 = std::chrono::high_resolution_clock::now();
                pixels = to_device ? sharded.render_to_device(u, scene, cam) : sharded.render_surface(u, scene, cam);
                auto t1 = std::chrono::high_resolution_clock::now();
                rt_times.push_back(std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count());
            }
            if (to_device) sharded.read_device_frame(pixels, surface.data(), surface.size());
            else std::copy(pixels, pixels + surface.size(), surface.begin());
            write_outputs(surface, W, H, out_ppm, out_raw, out_png);
            const rtx_stats st = sharded.stats();
            const int64_t mean_us = rt_times.empty() ? 0 : std::accumulate(rt_times.begin(), rt_times.end(), int64_t{0}) / static_cast<int64_t>(rt_times.size());
            std::cout << "Number of frames: " << frames << " : " << mean_us / 1000 << " ms average frame time\n";
            std::cout << "   " << mean_us << " microseconds for average raytracing\n";
            std::cout << "   0 milliseconds for surface average update\n";
            std::cout << "   fastest frame: " << *std::min_element(rt_times.begin(), rt_times.end()) << " microseconds (the first frame uploads the scene and allocates)\n";
            std::cout << "   device (CUDA events): " << st.raytracing_ms << " ms raytracing on the slowest of " << devices.size() << " GPUs, "
                      << st.total_rays << " rays, " << st.over_range_pixels << " over-range pixels in the last frame\n";
            return 0;
        }
        Renderer renderer(devices[0]);
        renderer.params.max_depth = depth;
        if (accel) renderer.params.accel = RTX_ACCEL_GRID;
        renderer.params.pixel_order = pixel_order;
        if (ext_sun) renderer.params.sun_enabled = 1;
        if (ext_tonemap) {
            renderer.params.tonemap = RTX_TONEMAP_REINHARD;
            renderer.params.quantise_mode = RTX_QUANT_SATURATE;
        }
        std::vector<int64_t> total_times, rt_times, surface_update_times;
        std::vector<double> device_rt_ms, device_surface_ms;
        for (int frame = 0; frame < frames; frame++) {
            if (frame < static_cast<int>(keys.size())) apply_key(cam, keys[frame]);   // one key event per frame (main.cpp:262-306)
            auto rt_start = std::chrono::high_resolution_clock::now();
            renderer.rt_scene(u, scene, cam, frame_buffer);
            auto rt_end = std::chrono::high_resolution_clock::now();
            device_rt_ms.push_back(renderer.stats.raytracing_ms);
            renderer.update_surface(frame_buffer, H, W, surface.data(), pitch);
            auto surface_end = std::chrono::high_resolution_clock::now();
            device_surface_ms.push_back(renderer.stats.surface_update_ms);
            rt_times.push_back(std::chrono::duration_cast<std::chrono::microseconds>(rt_end - rt_start).count());
            surface_update_times.push_back(std::chrono::duration_cast<std::chrono::milliseconds>(surface_end - rt_end).count());
            total_times.push_back(std::chrono::duration_cast<std::chrono::milliseconds>(surface_end - rt_start).count());
        }

        write_outputs(surface, W, H, out_ppm, out_raw, out_png);

        auto mean = [](const std::vector<int64_t>& v) { return v.empty() ? 0 : std::accumulate(v.begin(), v.end(), int64_t{0}) / static_cast<int64_t>(v.size()); };
        auto meand = [](const std::vector<double>& v) { return v.empty() ? 0.0 : std::accumulate(v.begin(), v.end(), 0.0) / v.size(); };
        std::cout << "Number of frames: " << frames << " : " << mean(total_times) << " ms average frame time\n";
        std::cout << "   " << mean(rt_times) << " microseconds for average raytracing\n";
        std::cout << "   " << mean(surface_update_times) << " milliseconds for surface average update\n";
        std::cout << "   device (CUDA events): " << meand(device_rt_ms) << " ms raytracing, " << meand(device_surface_ms) << " ms surface update, "
                  << renderer.stats.over_range_pixels << " over-range pixels in the last frame\n";
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "rtx_headless: %s\n", e.what());
        return 1;
    }
}
