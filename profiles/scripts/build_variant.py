"""Build a kernel variant next to the product library: gymwipe_b200/lib/variants/lib_<name>.so

    python profiles/scripts/build_variant.py <name> [-DFLAG ...]

Select it at run time with GYMWIPE_B200_LIB=<path> (A/B experiments only; the product is lib/libgymwipe_b200.so)."""
import os, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from gymwipe_b200 import _native as N

name, flags = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(N.LIB_DIR, "variants")
os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, "lib_%s.so" % name)
cmd = ["nvcc"] + N.NVCC_FLAGS + flags + ["-o", out, os.path.join(N.CSRC, "gw_kernels.cu")]
subprocess.check_call(cmd)
print(out)
