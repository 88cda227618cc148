// launch_host.h — host-only declarations shared by host_setup.cpp (g++) and curvis_abi.cu.
#pragma once
#include "../../include/curvis_gpu.h"
namespace curvis {
double host_shape_r(const curvis_metric& metric, double l);
double host_sin(double x);
void host_camera_basis(double theta, double phi, double n[3], double e_theta[3], double e_phi[3]);
}
