// Host-side view of the tensor-core engine: one entry per compiled (TILES, RPT, NL, PREC) instantiation
// of tc::tc_row_kernel.  The instantiations live in their own translation units (ddd1d_tc_inst_*.cu) so
// that they compile in parallel.
#pragma once
#include <cuda_runtime.h>

namespace ddd1d {
struct Work;
struct Tableau;
namespace tc {
struct TcParams;

struct TcEntry {
  const void* kernel;      // for cudaFuncSetAttribute / occupancy queries
  void (*launch)(const TcParams&, const Work&, const Tableau&, int grid, cudaStream_t);
  int tiles, rpt, nl, prec;
  int teams, slots_per_team;   // row teams per CTA; rows a team keeps in flight (3: TMEM block pool)
  int num_points;          // N of a row
  int threads, smem_bytes, tmem_cols;
  int slots_per_cta, rows_per_slot, sc_stride;   // scratch: grid * slots_per_cta * sc_stride floats
  int blob_hidden_bytes, blob_last_bytes;
};

// false: this combination is not compiled
bool lookup(int tiles, int rpt, int nl, int prec, TcEntry* out);
bool lookup_t1(int nl, int prec, TcEntry* out);
bool lookup_t2(int nl, int prec, TcEntry* out);
bool lookup_t2_pool(int nl, int prec, TcEntry* out);
bool lookup_t4(int nl, int prec, TcEntry* out);
bool lookup_p2(int nl, int prec, TcEntry* out);
bool lookup_p4(int nl, int prec, TcEntry* out);

}  // namespace tc
}  // namespace ddd1d
