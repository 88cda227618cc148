// efficient.h — host side of the table-based renderer, RelativisticSystem::render_image_efficient
// (reference src/systems.rs:333-527): the adaptive sampler of the escape-angle function
// (src/sampling.rs:46-245) and the per-segment interpolation coefficients (interp 1.0.3).
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>
#include "../../include/curvis_gpu.h"

namespace curvis {

// Device parameters of the per-pixel pass (efficient_kernel.cu).
struct Background;
struct CameraBlock;

struct EscapeTable {
    std::vector<double> alphas, escapes, signs;       // the sampler's output (systems.rs:458-486)
    std::vector<double> m_e, c_e, m_s, c_s;           // interp_slice segments: value = m[i]*x + c[i]
    uint64_t evaluations = 0, steps = 0;
    uint32_t passes = 0;              // device launches (speculative look-ahead merges refinement passes)
    uint32_t refinement_passes = 0;   // passes of the sampler's loop (sampling.rs:90-112)
};

// Integrates n photons leaving the camera position with the given tangent-space directions
// (3 doubles each) on the device and returns their final records.
using BatchIntegrate = std::function<int(const double* dirs, size_t n, curvis_ray_record* out)>;

// doubly_sample_function(-0.1 pi, 1.1 pi, ...) over compute_escape_angle (systems.rs:437-486).
// Returns a curvis_status; `err` explains a failure.
int build_escape_table(const curvis_metric& metric, double l_camera, uint32_t alphas_num, uint32_t max_iterations_sampling,
                       double threshold_1, double threshold_2, const BatchIntegrate& integrate, EscapeTable& table, std::string& err);

// vector3_from_theta_phi (algebra.rs:118-126) and rotation_from_two_vectors (algebra.rs:92-101;
// returns false where the reference panics).  Row-major 3x3.
void host_vector3_from_theta_phi(double theta, double phi, double out[3]);
bool host_rotation_from_two_vectors(const double v1[3], const double v2[3], double m[9]);

}  // namespace curvis
