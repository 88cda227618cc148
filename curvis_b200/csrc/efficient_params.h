// efficient_params.h — kernel argument of the per-pixel pass of render_image_efficient.
#pragma once
#include <cuda_runtime.h>
#include "frame_params.h"

namespace curvis {

struct EfficientParams {
    CameraBlock cam;
    uint32_t width, height, row_begin, row_end;
    double cam_pos_bg[3];      // vector3_from_theta_phi(theta_cam, phi_cam)              systems.rs:392-396
    double rot_bg[9];          // rotation_from_two_vectors(x, cam_pos_bg), row-major     systems.rs:409
    const double* alphas;      // the sampler's table (device)                             systems.rs:458-486
    const double *m_e, *c_e, *m_s, *c_s;   // interp_slice segments
    uint32_t n_points, n_segments;
    Background bg[2];
    uint8_t* out_rgb8;
    double* dbg;               // optional (alpha, angle, space) per pixel
    DeviceCounters* counters;
};

cudaError_t launch_efficient_pixels(const EfficientParams& p, int sm_count, cudaStream_t stream);

}  // namespace curvis
