#include "inst_lean.cuh"
ENTERP_DEFINE_LEAN(float, false)
