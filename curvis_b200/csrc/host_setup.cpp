// host_setup.cpp — CPU-side scene setup of the C ABI: the arithmetic the reference does once
// per camera/orientation on the host before any ray is fired.  Compiled with
// -ffp-contract=off so the doubles handed to the kernel equal the ones nalgebra 0.33.0
// (Cargo.lock:623-624) produces for the reference:
//   Orientation::new                        src/algebra.rs:16-38
//   rotation_matrix_from_forward_up_pairs   src/algebra.rs:64-74   (Rotation3::face_towards)
//   Camera::new                             src/cameras.rs:79-122
//   {Ellis,Interstellar,Flat}Metric::new    src/metrics.rs:404-414, :441-459, :496-498
#include <cmath>
#include <cstring>
#include "../../include/curvis_gpu.h"
#include "host_error.h"
#include "launch_host.h"

namespace {

struct Vec3 { double x, y, z; };

inline double norm(const Vec3& v) { return std::sqrt((v.x * v.x + v.y * v.y) + v.z * v.z); }
inline Vec3 normalized(const Vec3& v) { const double n = norm(v); return {v.x / n, v.y / n, v.z / n}; }
inline Vec3 cross(const Vec3& a, const Vec3& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

struct Mat3 {
    double m[3][3];
    Mat3 transposed() const {
        Mat3 t;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) t.m[j][i] = m[i][j];
        return t;
    }
    Mat3 operator*(const Mat3& b) const {  // k accumulated 0,1,2 left to right
        Mat3 o;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) o.m[i][j] = (m[i][0] * b.m[0][j] + m[i][1] * b.m[1][j]) + m[i][2] * b.m[2][j];
        return o;
    }
    Vec3 operator*(const Vec3& v) const {
        return {(m[0][0] * v.x + m[0][1] * v.y) + m[0][2] * v.z,
                (m[1][0] * v.x + m[1][1] * v.y) + m[1][2] * v.z,
                (m[2][0] * v.x + m[2][1] * v.y) + m[2][2] * v.z};
    }
};

// Rotation3::face_towards(dir, up): columns (up x z' normalised, z' x x' normalised, z' = dir normalised)
Mat3 face_towards(const Vec3& dir, const Vec3& up) {
    const Vec3 zc = normalized(dir);
    const Vec3 xc = normalized(cross(up, zc));
    const Vec3 yc = normalized(cross(zc, xc));
    return Mat3{{{xc.x, yc.x, zc.x}, {xc.y, yc.y, zc.y}, {xc.z, yc.z, zc.z}}};
}

}  // namespace

namespace curvis {

// r(l) of the three metrics, reference operation order (metrics.rs:417-418, :465-472, :502).
double host_shape_r(const curvis_metric& g, double l) {
    const double pi = 3.14159265358979323846264338327950288;
    switch (g.kind) {
    case CURVIS_METRIC_ELLIS: return std::sqrt(g.rho * g.rho + l * l);
    case CURVIS_METRIC_INTERSTELLAR:
        if (std::fabs(l) > g.a) {
            const double x = 2.0 * (std::fabs(l) - g.a) / (pi * g.m);
            return g.rho + g.m * (x * std::atan(x) - std::log(1.0 + x * x) / 2.0);
        }
        return g.rho;
    default: return l;
    }
}

double host_sin(double x) { return std::sin(x); }

// CURVIS_COORDINATES_CARTESIAN: the spherical basis at the camera position (the reference's convention for (theta, phi),
// src/algebra.rs:118-126).
void host_camera_basis(double theta, double phi, double n[3], double e_theta[3], double e_phi[3]) {
    const double st = std::sin(theta), ct = std::cos(theta), sp = std::sin(phi), cp = std::cos(phi);
    n[0] = st * cp; n[1] = st * sp; n[2] = ct;
    e_theta[0] = ct * cp; e_theta[1] = ct * sp; e_theta[2] = -st;
    e_phi[0] = -sp; e_phi[1] = cp; e_phi[2] = 0.0;
}

}  // namespace curvis

extern "C" int curvis_orientation(const double forward[3], const double up[3],
                                  double rot[9], double inv_rot[9], double up_orthogonal[3]) {
    if (!forward || !up) return curvis::set_thread_error(CURVIS_ERR_INVALID_ARGUMENT, "curvis_orientation: null forward/up");
    const Vec3 f{forward[0], forward[1], forward[2]}, u{up[0], up[1], up[2]};
    if (norm(cross(f, u)) == 0.0)  // algebra.rs:19-21
        return curvis::set_thread_error(CURVIS_ERR_PARALLEL_VECTORS, "Forward and up vectors must not be parallel");
    const Mat3 previous = face_towards({1.0, 0.0, 0.0}, {0.0, 0.0, 1.0});
    const Mat3 next = face_towards(f, u);
    const Mat3 r = next * previous.transposed();
    const Mat3 rinv = r.transposed();
    if (rot) std::memcpy(rot, r.m, sizeof r.m);
    if (inv_rot) std::memcpy(inv_rot, rinv.m, sizeof rinv.m);
    if (up_orthogonal) {
        const Vec3 o = r * Vec3{0.0, 0.0, 1.0};
        up_orthogonal[0] = o.x; up_orthogonal[1] = o.y; up_orthogonal[2] = o.z;
    }
    return CURVIS_OK;
}

extern "C" int curvis_camera_init(curvis_camera* cam, const double position[4],
                                  const double forward[3], const double up[3],
                                  double focal_length, double sensor_diagonal,
                                  uint32_t resolution_width, uint32_t resolution_height) {
    if (!cam || !position) return curvis::set_thread_error(CURVIS_ERR_INVALID_ARGUMENT, "curvis_camera_init: null argument");
    if (!(focal_length > 0.0)) return curvis::set_thread_error(CURVIS_ERR_INVALID_ARGUMENT, "focal_length must be greater than 0");
    if (!(sensor_diagonal > 0.0)) return curvis::set_thread_error(CURVIS_ERR_INVALID_ARGUMENT, "sensor_diagonal must be greater than 0");
    if (resolution_width == 0 || resolution_height == 0)
        return curvis::set_thread_error(CURVIS_ERR_INVALID_ARGUMENT, "resolution_width and resolution_height must be greater than 0");
    const int rc = curvis_orientation(forward, up, cam->cam_to_world, nullptr, nullptr);
    if (rc != CURVIS_OK) return rc;
    const double aspect = (double)resolution_width / (double)resolution_height;   // cameras.rs:107
    const double aspect_squared = aspect * aspect;                                // :108
    cam->sensor_height = std::sqrt((sensor_diagonal * sensor_diagonal) / (aspect_squared + 1.0));  // :109
    cam->sensor_width = aspect * cam->sensor_height;                              // :110
    std::memcpy(cam->position, position, 4 * sizeof(double));
    cam->focal_length = focal_length;
    cam->resolution_width = resolution_width;
    cam->resolution_height = resolution_height;
    return CURVIS_OK;
}

extern "C" int curvis_metric_validate(const curvis_metric* metric) {
    if (!metric) return curvis::set_thread_error(CURVIS_ERR_INVALID_ARGUMENT, "curvis_metric_validate: null metric");
    switch (metric->kind) {
    case CURVIS_METRIC_ELLIS:
        if (!(metric->rho > 0.0)) return curvis::set_thread_error(CURVIS_ERR_INVALID_METRIC, "The rho parameter for Ellis Metrics must be positive.");
        return CURVIS_OK;
    case CURVIS_METRIC_INTERSTELLAR:
        if (!(metric->m > 0.0)) return curvis::set_thread_error(CURVIS_ERR_INVALID_METRIC, "The mass parameter for Interstellar Metrics must be positive.");
        if (!(metric->a > 0.0)) return curvis::set_thread_error(CURVIS_ERR_INVALID_METRIC, "The angular momentum parameter for Interstellar Metrics must be positive.");
        if (!(metric->rho > 0.0)) return curvis::set_thread_error(CURVIS_ERR_INVALID_METRIC, "The rho parameter for Interstellar Metrics must be positive.");
        return CURVIS_OK;
    case CURVIS_METRIC_FLAT:
        return CURVIS_OK;
    default:
        return curvis::set_thread_error(CURVIS_ERR_INVALID_ARGUMENT, "unknown metric kind");
    }
}
