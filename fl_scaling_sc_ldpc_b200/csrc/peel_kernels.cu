// peel_kernels.cu -- placeholder translation unit, filled in by the peeling decoder (K3).
#include "common.cuh"
