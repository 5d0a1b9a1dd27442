#include "launchers.cuh"
MGR_INSTANTIATE(bf16, __nv_bfloat16)
