// curvis.hpp — C++ host mirror of the reference's scene types over the C ABI
// (include/curvis_gpu.h).  Same names, argument meaning and error behaviour as the Rust crate:
//
//   curvis::EllisMetric / InterstellarMetric / FlatSphericalMetric   src/metrics.rs:399-505
//   curvis::Camera                                                   src/cameras.rs:30-172
//   curvis::SphericalImage                                           src/images.rs:51-105
//   curvis::RelativisticSystem<M>::render_image(max_iterations, max_radius, delta)
//                                                                    src/systems.rs:68-73, :307-330
//
// Where the reference panics, these throw curvis::Error carrying the curvis_status.  Header
// only; link against libcurvis_b200.so.  The per-ray arithmetic is NOT here — it runs in the
// sm_100a kernels behind curvis_render_image.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "../../include/curvis_gpu.h"

namespace curvis {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

inline void check(int code, const curvis_ctx* ctx = nullptr) {
    if (code != CURVIS_OK) throw Error(code, curvis_last_error(ctx));
}

using Vector3 = std::array<double, 3>;
using Vector4 = std::array<double, 4>;
using Matrix3 = std::array<double, 9>;  // row-major

// Orientation::new(forward, up)  (src/algebra.rs:16-38)
class Orientation {
  public:
    Orientation(const Vector3& forward, const Vector3& up) : forward_(forward) {
        check(curvis_orientation(forward.data(), up.data(), rot_.data(), inv_.data(), up_.data()));
    }
    const Vector3& forward() const { return forward_; }
    const Vector3& up() const { return up_; }
    const Matrix3& rotation_matrix() const { return rot_; }
    const Matrix3& inverse_rotation_matrix() const { return inv_; }

  private:
    Vector3 forward_, up_{};
    Matrix3 rot_{}, inv_{};
};

struct EllisMetric {  // EllisMetric::new(rho), src/metrics.rs:404-414
    explicit EllisMetric(double rho) : c{CURVIS_METRIC_ELLIS, 0, rho, 0.0, 0.0} { check(curvis_metric_validate(&c)); }
    curvis_metric c;
};
struct InterstellarMetric {  // InterstellarMetric::new(m, a, rho), src/metrics.rs:441-459
    InterstellarMetric(double m, double a, double rho) : c{CURVIS_METRIC_INTERSTELLAR, 0, rho, m, a} { check(curvis_metric_validate(&c)); }
    curvis_metric c;
};
struct FlatSphericalMetric {  // FlatSphericalMetric::new(), src/metrics.rs:496-498
    FlatSphericalMetric() : c{CURVIS_METRIC_FLAT, 0, 0.0, 0.0, 0.0} {}
    curvis_metric c;
};

// Camera::new(position, forward_world, up_world, focal_length, sensor_diagonal, w, h)  (src/cameras.rs:79-122)
class Camera {
  public:
    Camera(const Vector4& position, const Vector3& forward_world, const Vector3& up_world, double focal_length,
           double sensor_diagonal, uint32_t resolution_width, uint32_t resolution_height)
        : forward_(forward_world), up_(up_world), diagonal_(sensor_diagonal) {
        check(curvis_camera_init(&c_, position.data(), forward_world.data(), up_world.data(), focal_length,
                                 sensor_diagonal, resolution_width, resolution_height));
    }
    Vector4 position() const { return {c_.position[0], c_.position[1], c_.position[2], c_.position[3]}; }
    void update_position(const Vector4& p) { for (int i = 0; i < 4; ++i) c_.position[i] = p[i]; }   // :135-140
    void update_orientation(const Vector3& forward_world, const Vector3& up_world) {                 // :143-146
        check(curvis_orientation(forward_world.data(), up_world.data(), c_.cam_to_world, nullptr, nullptr));
        forward_ = forward_world; up_ = up_world;
    }
    uint32_t resolution_width() const { return c_.resolution_width; }
    uint32_t resolution_height() const { return c_.resolution_height; }
    const curvis_camera& c() const { return c_; }

  private:
    curvis_camera c_{};
    Vector3 forward_, up_;
    double diagonal_;
};

// SphericalImage::new(img, forward, up)  (src/images.rs:71-89); img = RGBA8 texels, row-major.
class SphericalImage {
  public:
    SphericalImage(std::vector<uint8_t> rgba8, uint32_t width, uint32_t height,
                   const Vector3& forward = {1.0, 0.0, 0.0}, const Vector3& up = {0.0, 0.0, 1.0})
        : rgba8_(std::move(rgba8)), width_(width), height_(height), orientation_(forward, up) {
        if (rgba8_.size() != (size_t)width * height * 4) throw Error(CURVIS_ERR_INVALID_ARGUMENT, "rgba8 size != width*height*4");
    }
    const Orientation& orientation() const { return orientation_; }
    void set_forward_up(const Vector3& forward, const Vector3& up) { orientation_ = Orientation(forward, up); }  // :102-104
    uint32_t width() const { return width_; }
    uint32_t height() const { return height_; }
    const uint8_t* data() const { return rgba8_.data(); }

  private:
    std::vector<uint8_t> rgba8_;
    uint32_t width_, height_;
    Orientation orientation_;
};

// RGB8 frame, row-major: the payload of the DynamicImage::ImageRgb8 render_image returns.
struct ImageRgb8 {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> data;
    const uint8_t* pixel(uint32_t x, uint32_t y) const { return &data[((size_t)y * width + x) * 3]; }
};

// RelativisticSystem<M>  (src/systems.rs:68-73, :283-330)
template <class M>
class RelativisticSystem {
  public:
    RelativisticSystem(M metric, SphericalImage background_positive, SphericalImage background_negative, Camera camera,
                       const std::vector<int>& devices = {})
        : metric(std::move(metric)), background_positive(std::move(background_positive)),
          background_negative(std::move(background_negative)), camera(std::move(camera)) {
        check(curvis_ctx_create(devices.empty() ? nullptr : devices.data(), (int)devices.size(), &ctx_));
        try {
            upload(+1, this->background_positive);
            upload(-1, this->background_negative);
        } catch (...) {
            curvis_ctx_destroy(ctx_);
            throw;
        }
    }
    ~RelativisticSystem() { curvis_ctx_destroy(ctx_); }
    RelativisticSystem(const RelativisticSystem&) = delete;
    RelativisticSystem& operator=(const RelativisticSystem&) = delete;

    // render_image(&self, max_iterations: u32, max_radius: f64, delta: f64) -> DynamicImage   (:307-330)
    ImageRgb8 render_image(uint32_t max_iterations, double max_radius, double delta, curvis_stats* stats = nullptr) const {
        ImageRgb8 img;
        img.width = camera.resolution_width();
        img.height = camera.resolution_height();
        img.data.resize((size_t)img.width * img.height * 3);
        curvis_sim sim{max_iterations, CURVIS_FRAME_LOCAL, max_radius, delta, CURVIS_PRECISION_F64, CURVIS_SAMPLING_NEAREST,
                       CURVIS_INTEGRATOR_EULER, CURVIS_COORDINATES_SPHERICAL, 0.0, 0.0};
        check(curvis_render_image(ctx_, &metric.c, &camera.c(), &sim, img.data.data(), stats), ctx_);
        return img;
    }

    M metric;
    SphericalImage background_positive, background_negative;
    Camera camera;

  private:
    void upload(int side, const SphericalImage& im) {
        check(curvis_set_background(ctx_, side, im.data(), im.width(), im.height(),
                                    im.orientation().inverse_rotation_matrix().data()), ctx_);
    }
    curvis_ctx* ctx_ = nullptr;
};

}  // namespace curvis
