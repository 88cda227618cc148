// curvis_image.cpp — minimal C++ driver over curvis.hpp: renders one frame with the
// reference's default scene (settings/defaults/*.toml) on synthetic backgrounds and writes a
// binary PPM.  Usage: curvis_image <ellis|interstellar> <W> <H> <max_iter> <max_radius> <delta> <out.ppm>
// With no GPU it exits 3 after printing the library's error (there is no CPU fallback).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "curvis.hpp"

static std::vector<uint8_t> decodable_background(uint32_t w, uint32_t h, bool negative) {
    std::vector<uint8_t> v((size_t)w * h * 4);
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < w; ++x) {
            uint8_t* t = &v[((size_t)y * w + x) * 4];
            uint8_t r = (uint8_t)(x & 255), g = (uint8_t)(y & 255);
            t[0] = negative ? (uint8_t)~r : r;
            t[1] = negative ? (uint8_t)~g : g;
            t[2] = (uint8_t)((((x >> 8) & 15) << 4) | ((y >> 8) & 15));
            t[3] = 255;
        }
    return v;
}

template <class M>
static int run(M metric, uint32_t W, uint32_t H, uint32_t max_iter, double R, double delta, const char* out) {
    const double pi = 3.14159265358979323846;
    curvis::Camera cam({0.0, 5.0, pi / 2.0, 0.0}, {-1.0, 0.0, 0.0}, {0.0, 0.0, 1.0}, 15.0, 43.0, W, H);
    curvis::RelativisticSystem<M> sys(metric, curvis::SphericalImage(decodable_background(1024, 512, false), 1024, 512),
                                      curvis::SphericalImage(decodable_background(1024, 512, true), 1024, 512), cam);
    curvis_stats st;
    curvis::ImageRgb8 img = sys.render_image(max_iter, R, delta, &st);
    std::printf("steps=%llu positive=%llu negative=%llu not_escaped=%llu kernel_ms=%.3f\n",
                (unsigned long long)st.total_steps, (unsigned long long)st.n_positive, (unsigned long long)st.n_negative,
                (unsigned long long)st.n_not_escaped, st.kernel_ms);
    if (FILE* f = std::fopen(out, "wb")) {
        std::fprintf(f, "P6\n%u %u\n255\n", img.width, img.height);
        std::fwrite(img.data.data(), 1, img.data.size(), f);
        std::fclose(f);
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc != 8) { std::fprintf(stderr, "usage: %s <ellis|interstellar> W H max_iter max_radius delta out.ppm\n", argv[0]); return 2; }
    try {
        const uint32_t W = (uint32_t)std::atoi(argv[2]), H = (uint32_t)std::atoi(argv[3]), it = (uint32_t)std::atoi(argv[4]);
        const double R = std::atof(argv[5]), d = std::atof(argv[6]);
        if (!std::strcmp(argv[1], "ellis")) return run(curvis::EllisMetric(1.0), W, H, it, R, d, argv[7]);
        return run(curvis::InterstellarMetric(0.1, 1e-4, 1.0), W, H, it, R, d, argv[7]);
    } catch (const curvis::Error& e) {
        std::fprintf(stderr, "curvis error %d: %s\n", e.code, e.what());
        return 3;
    }
}
