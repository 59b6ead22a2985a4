"""Summarise an .ncu-rep (ncu --set full --import-source on) into the small JSON kept under profiles/:
selected raw metrics of the first kernel plus the source lines with the most warp-stall samples.
  python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01/name.json"""
import csv, io, json, subprocess, sys

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
           'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
           'sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
           'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
           # the tensor core with its operand path (the unit the small-N MMAs of this engine keep busy: a 128 x 16
           # fp16 A tile costs 32 clocks of shared-memory wavefronts whatever N is)
           'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed',
           'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
           'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed',
           'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed',
           'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed',
           'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
           'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio']


def ncu_csv(rep, *args):
  out = subprocess.run(['ncu', '-i', rep, '--csv'] + list(args), capture_output=True, text=True).stdout
  return list(csv.reader(io.StringIO(out)))


def main(rep, dest):
  rows = ncu_csv(rep, '--page', 'raw')
  hdr, units, vals = rows[0], rows[1], rows[2]
  col = {h: i for i, h in enumerate(hdr)}
  summary = {'Kernel Name': vals[col['Kernel Name']], 'Block Size': vals[col['Block Size']],
             'Grid Size': vals[col['Grid Size']]}
  for m in METRICS:
    if m in col:
      summary[m] = {'unit': units[col[m]], 'value': vals[col[m]]}
  src = ncu_csv(rep, '--page', 'source', '--print-source', 'cuda,sass')
  lines, fn = [], None
  for r in src:
    if r and r[0] == 'File Path':
      fn = r[1].split('/')[-1]
    elif len(r) > 8 and r[0].isdigit():
      try:
        lines.append((fn, int(r[0]), r[1].strip(), int(r[6]), int(r[7])))
      except ValueError:
        pass
  tot_s = sum(l[3] for l in lines) or 1
  tot_i = sum(l[4] for l in lines) or 1
  lines.sort(key=lambda l: -l[3])
  summary['top_stall_lines'] = [{'file': l[0], 'line': l[1], 'source': l[2][:120], 'pct_samples': round(100.0 * l[3] / tot_s, 2),
                                 'pct_instructions': round(100.0 * l[4] / tot_i, 2)} for l in lines[:25]]
  with open(dest, 'w') as f:
    json.dump(summary, f, indent=1)
  print('wrote', dest)


if __name__ == '__main__':
  main(sys.argv[1], sys.argv[2])
