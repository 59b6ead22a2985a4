// tc_row_kernel instantiations: TILES = 2, rows per tile = 1 (see ddd1d_tc_inst.inc)
#define DDD1D_TC_TILES 2
#define DDD1D_TC_RPT 1
#define DDD1D_TC_NAME t2
#define DDD1D_TC_POOL 1
#include "ddd1d_tc_inst.inc"
