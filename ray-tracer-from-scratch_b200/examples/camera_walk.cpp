// camera_walk.cpp — the façade's Camera (host/rtx_scene.hpp) driven like the reference's main loop drives its own
// (main.cpp:146-154 set-up, :262-306 keys, scene.cpp:108-165 methods): init() once, then one method per step; prints
// position, direction and vup after every step as hex doubles. Host code only — no GPU is touched, which is what lets
// the CPU test suite compare it with the reference's arithmetic (tests/golden/camera_walks.json).
//
//   rtx_camera_walk px py pz  lx ly lz  ux uy uz  vfov aspect width  [op[:arg]]...
//   op: w s a d (moves), y:<angle> rotate_left_right, p:<angle> rotate_up_down; numbers in any strtod format (hex ok)
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "rtx_scene.hpp"

int main(int argc, char** argv)
{
    using namespace rtx;
    if (argc < 13) {
        std::fprintf(stderr, "usage: rtx_camera_walk px py pz lx ly lz ux uy uz vfov aspect width [op[:arg]]...\n");
        return 2;
    }
    double v[12];
    for (int k = 0; k < 12; k++) v[k] = std::strtod(argv[1 + k], nullptr);
    Camera cam;
    cam.position = point3(v[0], v[1], v[2]);
    cam.lookat = point3(v[3], v[4], v[5]);
    cam.vup = vec3(v[6], v[7], v[8]);
    cam.vfov = v[9];
    cam.aspect_ratio = v[10];
    cam.image_width = v[11];
    cam.movement_speed = 0.1;
    try {
        cam.init();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "rtx_camera_walk: %s\n", e.what());
        return 1;
    }
    for (int k = 13; k < argc; k++) {
        const char op = argv[k][0];
        const double arg = (std::strlen(argv[k]) > 2 && argv[k][1] == ':') ? std::strtod(argv[k] + 2, nullptr) : 0.0;
        switch (op) {
            case 'w': cam.forward(); break;
            case 's': cam.backward(); break;
            case 'a': cam.left(); break;
            case 'd': cam.right(); break;
            case 'y': cam.rotate_left_right(arg); break;
            case 'p': cam.rotate_up_down(arg); break;
            default: std::fprintf(stderr, "unknown op %s\n", argv[k]); return 2;
        }
        std::printf("%a %a %a  %a %a %a  %a %a %a\n", cam.position.x, cam.position.y, cam.position.z, cam.direction.x, cam.direction.y,
                    cam.direction.z, cam.vup.x, cam.vup.y, cam.vup.z);
    }
    return 0;
}
