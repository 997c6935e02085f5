// omc_altmin.cuh -- alternating minimisation kernels (K7 / K8); see omc_altmin_impl below.
#pragma once
#include "omc_device.cuh"
namespace omc {}
