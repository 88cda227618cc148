// launch.h — host-callable launchers of the render kernels (one per precision TU).
#pragma once
#include <cuda_runtime.h>
#include "frame_params.h"

namespace curvis {

// fp64 parity kernel (render_f64.cu, compiled with -fmad=false).
cudaError_t launch_render_f64(const FrameParams& p, int metric_kind, int sm_count, cudaStream_t stream);

// FMA-only micro-kernels used as the measured compute-roofline denominator (peak_kernels.cu).
cudaError_t measure_fma_peak(int sm_count, cudaStream_t stream, double* fp64_tflops, double* fp32_tflops);

}  // namespace curvis
