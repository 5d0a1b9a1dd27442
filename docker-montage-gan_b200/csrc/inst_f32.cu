#include "launchers.cuh"
MGR_INSTANTIATE(f32, float)
