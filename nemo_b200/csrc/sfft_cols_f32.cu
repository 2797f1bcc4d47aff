// fp32 instantiations of the sfft column passes
#include "sfft_cols.cuh"
SFFT_COLS_ENTRY(float, float2, f32)
