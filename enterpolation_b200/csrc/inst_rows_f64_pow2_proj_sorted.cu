#include "inst_rows.cuh"
ENTERP_DEFINE_ROWS_PART(double, true, true, false)
