"""Build a kernel variant of libtrimem_b200.so for timing experiments.
   python scripts/build_variant.py NAME -DTM_GRAD_MINB=3 ...   ->  trimem_b200/lib/variants/NAME.so
Run with TRIMEM_B200_LIB=trimem_b200/lib/variants/NAME.so python scripts/kernel_bench.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trimem_b200 import build as b
name = sys.argv[1]
defs = [a[2:] for a in sys.argv[2:] if a.startswith("-D")]
out = os.path.join(ROOT, "trimem_b200", "lib", "variants", name + ".so")
log = b.build_cuda(force=True, ptxas_info=True, defines=defs, out=out)
for line in log.splitlines():
    if "k_grad" in line or "k_props" in line or "spill" in line or "registers" in line:
        print(line[:160])
print("built", out)
