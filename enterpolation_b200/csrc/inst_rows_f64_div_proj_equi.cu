#include "inst_rows.cuh"
ENTERP_DEFINE_ROWS_PART(double, false, true, true)
