#include "inst_rows.cuh"
ENTERP_DEFINE_ROWS_PART(float, true, true, false)
