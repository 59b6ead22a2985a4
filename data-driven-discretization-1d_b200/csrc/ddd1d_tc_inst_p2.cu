// tc_row_kernel instantiations: TILES = 1, rows per tile = 2 (see ddd1d_tc_inst.inc)
#define DDD1D_TC_TILES 1
#define DDD1D_TC_RPT 2
#define DDD1D_TC_NAME p2
#include "ddd1d_tc_inst.inc"
