// fp64 instantiations of the sfft column passes (a translation unit per precision, so that they compile side by side)
#include "sfft_cols.cuh"
SFFT_COLS_ENTRY(double, double2, f64)
