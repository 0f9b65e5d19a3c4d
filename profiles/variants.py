"""Build A/B variants of the 5x6 step kernels:  python profiles/variants.py NAME [-DFLAG ...] [--maxrregcount N]
Recompiles ppad_step_inst.cu for geometry 0 with the extra flags and links it with the regular objects of the other
translation units (build/obj, from __graft_entry__.build()) into build/variants/NAME.so.  Use with PPAD_B200_LIB=..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
name, flags = sys.argv[1], sys.argv[2:]
g.build()
out_dir = os.path.join(ROOT, "build", "variants")
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, name + "_g0.o")
subprocess.check_call(["nvcc"] + g.NVCC_FLAGS + ["-DPPAD_GEOM=0"] + flags + ["-c", "-o", obj, os.path.join(g.CSRC, "ppad_step_inst.cu")])
objs = [os.path.join(g.OBJ_DIR, o) for _, _, o in g.UNITS if o != "ppad_step_inst_g0.o"] + [obj]
subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", os.path.join(out_dir, name + ".so")] + objs)
print("built", os.path.join(out_dir, name + ".so"))
