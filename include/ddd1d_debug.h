/* Test-only entry points of libddd1d (not part of the drop-in surface). */
#ifndef DDD1D_DEBUG_H_
#define DDD1D_DEBUG_H_
#ifdef __cplusplus
extern "C" {
#endif
/* One 128-position tile of a 32 -> nout (16 | 32), 5-tap conv through the tcgen05 path
 * (same shared-memory descriptors, 3xTF32 split and TMEM read-back as the row kernel).
 *   x      device float [132][32]   inputs at positions -2..129
 *   w_cat  device float [5*8][2*nout][4]   filters packed as B planes: rows 0..nout-1 the TF32-rounded
 *          weights, rows nout..2*nout-1 their TF32-rounded remainders
 *   out    device float [128][nout] */
int ddd1d_debug_tc_probe(int device, const float* x, const float* w_cat, float* out, int nout, void* stream);
/* MMA issue-rate microbenchmark: `blocks` CTAs each issue reps x 20 (tap, ci-block) steps of the
 * tensor engine's MMA pattern (mode: see csrc/ddd1d_tc.cuh) and report the clocks until completion. */
int ddd1d_debug_tc_rate(int device, int variant, int reps, int blocks, long long* cycles_host);
/* Tensor-pipe / CUDA-core overlap experiment: warp 0 streams reps x 10 fp16 MMA steps (hidden-layer shape)
 * while warps 4..7 run `iters` rounds of a CUDA-core workload.  mode bit 0 = run the MMAs; mode >> 1 =
 * workload (0 none, 1 FFMA, 2 STS.128 + LDS.128, 3 packed fp16 splits, 4 tcgen05.ld, 5 SHFL, 6 LDG,
 * 7 LDS.128, 8 STS.128).
 * cycles_host[2*b] = clocks of the MMA stream, cycles_host[2*b + 1] = clocks of the CUDA-core stream. */
int ddd1d_debug_tc_overlap(int device, int mode, int reps, int iters, int blocks, long long* cycles_host);
/* TMEM-resident A operand probe (tcgen05.cp 128x256b / 4x256b, tcgen05.shift.down, tcgen05.mma with A in TMEM,
 * .ashift): raw TMEM dumps after every step, out_host uint32 [10][128][16] (scripts/tc_shift_probe.py). */
int ddd1d_debug_tc_shift_probe(int device, unsigned int* out_host);
/* Write-after-read stress of the TMEM A operand: `chain` MMAs read A, a tcgen05.cp overwrites it, one more MMA reads
 * the new rows; out_host uint32 [2][128][16] + 1 (mismatch count over `rounds` repetitions). */
int ddd1d_debug_tc_war_probe(int device, int chain, int rounds, unsigned int* out_host);
/* Rate of the TMEM-resident-A tile-layer pattern (6 x cp.128x256b, 30 MMAs of N = nb with .ashift, 24 x cp.4x256b):
 * cycles_host[4 * block + issuer] = clocks for `reps` tile-layers.  flags bit 0: patches, bit 1: .ashift. */
int ddd1d_debug_tc_ta_rate(int device, int nb, int reps, int issuers, int flags, int blocks, long long* cycles_host);
/* Collector-buffer reuse: clocks for `reps` pairs of fp16 MMAs of width n sharing an operand (mode 0 plain pair with
 * the same B, 1 tcgen05.mma.ws with B kept in collector b0, 2 same A with .collector::a::fill / ::lastuse, 3 same A
 * without qualifiers). */
int ddd1d_debug_tc_reuse_rate(int device, int n, int mode, int reps, long long* cycles_host);
#ifdef __cplusplus
}
#endif
#endif
