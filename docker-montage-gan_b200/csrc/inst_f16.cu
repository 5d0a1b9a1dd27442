#include "launchers.cuh"
MGR_INSTANTIATE(f16, __half)
