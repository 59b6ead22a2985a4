"""profiles/traffic.json from the ncu summaries under profiles/r02 (DRAM bytes per launch of the `ncu --set full`
captures; bench.py scales them to its launch and reports them as roofline.traffic)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
CAPTURES = {     # key: (summary, batch, N, RK steps in the captured launch)
    'c2/tensor': ('tc_c2_ncu_full.json', 4096, 256, 20), 'c3/tensor': ('tc_c3_ncu_full.json', 4096, 128, 20),
    'c2s/tensor': ('tc_c2s_ncu_full.json', 16384, 64, 20), 'c5/ffma': ('weno_c5_ncu_full.json', 8192, 2048, 5),
    'c1b/ffma': ('warp_c1b_ncu_full.json', 65536, 64, 20)}
table = {}
for key, (name, batch, n, steps) in CAPTURES.items():
  with open(os.path.join(ROOT, 'profiles', 'r02', name)) as f:
    d = json.load(f)
  total = sum(float(d[m]['value']) * UNIT[d[m]['unit']] for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
  table[key] = {'batch': batch, 'num_points': n, 'rk_steps_in_capture': steps, 'dram_bytes_per_launch_in_capture': total,
                'dram_bytes_per_rk_step': total / steps, 'algorithmic_bytes_per_rk_step': 8.0 * batch * n,
                'source': 'profiles/r02/%s (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)' % name}
with open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w') as f:
  json.dump(table, f, indent=1)
print(json.dumps(table, indent=1))
