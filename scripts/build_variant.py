"""Build a variant of libddd1d.so with extra nvcc flags into data-driven-discretization-1d_b200/variants/ for
same-box A/B runs (scripts/gpu_ab2.sh).  Usage: python scripts/build_variant.py <name> [-DFLAG=..] ..."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
obj_dir = os.path.join(g.OBJ_DIR, 'variant_' + name)
os.makedirs(obj_dir, exist_ok=True)
os.makedirs(os.path.join(g.PKG_DIR, 'variants'), exist_ok=True)


def compile_unit(unit):
  obj = os.path.join(obj_dir, unit[:-3] + '.o')
  cmd = ['nvcc'] + g.NVCC_FLAGS + extra + ['-I', os.path.join(ROOT, 'include'), '-c', '-o', obj,
                                          os.path.join(g.CSRC, unit)]
  subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
  return obj


with ThreadPoolExecutor(max_workers=len(g.PRODUCT_UNITS)) as pool:
  objs = list(pool.map(compile_unit, g.PRODUCT_UNITS))
target = os.path.join(g.PKG_DIR, 'variants', 'libddd1d_%s.so' % name)
subprocess.run(['nvcc', '--shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', target] + objs, check=True)
print(target)
