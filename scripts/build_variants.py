"""Experiment builds of the accumulate kernel: recompiles the hot BLS12-377 translation unit with extra -D flags and links it
with the library's other objects into variants/lib_<name>.so (git-ignored, travels to the GPU box).  A variant is selected at
run time with B200MSM_LIB=<path> (2022-entries_b200/__init__.py).  Also records registers / spills of k_accumulate per variant.

  python scripts/build_variants.py name1="-DFLAG ..." name2="..."        (run after `make -C 2022-entries_b200/csrc`)
"""
import os, re, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "2022-entries_b200", "csrc")
OUT = os.path.join(ROOT, "variants")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
BASE = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xptxas", "-v", "-Xcompiler", "-fPIC"]


def build(item):
    name, flags = item
    obj = os.path.join(OUT, f"hot377_{name}.o")
    r = subprocess.run([NVCC] + BASE + flags.split() + ["-c", "-o", obj, os.path.join(CSRC, "inst_hot_377.cu")], capture_output=True, text=True)
    if r.returncode:
        return name, "COMPILE FAILED\n" + r.stderr[-2000:]
    others = [os.path.join(CSRC, "build", o) for o in ("msm.o", "inst_hot_381.o", "inst_hot_377x.o", "inst_cold_377.o", "inst_cold_381.o", "inst_cold_377x.o")]
    so = os.path.join(OUT, f"lib_{name}.so")
    subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so, obj] + others)
    m = re.search(r"k_accumulateINS_9Bls12_377ELb0.*?\n.*?\n\s*(\d+ bytes stack frame.*?)\n.*?Used (\d+) registers", r.stderr, re.S)
    return name, (f"regs {m.group(2)}, {m.group(1)}" if m else "?") + f"   [{flags}]"


def build_full(item):
    """every translation unit with the flags (layout experiments that touch cold and hot code alike)"""
    name, flags = item
    objs = []
    for tu in ("msm", "inst_hot_377", "inst_hot_381", "inst_hot_377x", "inst_cold_377", "inst_cold_381", "inst_cold_377x"):
        obj = os.path.join(OUT, f"full_{name}_{tu}.o")
        r = subprocess.run([NVCC] + BASE + flags.split() + ["-c", "-o", obj, os.path.join(CSRC, tu + ".cu")], capture_output=True, text=True)
        if r.returncode:
            return name, "COMPILE FAILED " + tu + "\n" + r.stderr[-1500:]
        objs.append(obj)
    subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", os.path.join(OUT, f"lib_{name}.so")] + objs)
    return name, f"full build   [{flags}]"


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--full" in sys.argv:
        for a in sys.argv[1:]:
            if a != "--full":
                print("%s: %s" % build_full(a.split("=", 1)), flush=True)
        return
    items = [a.split("=", 1) for a in sys.argv[1:]]
    with ThreadPoolExecutor(4) as ex:
        for name, info in ex.map(build, items):
            print(f"{name}: {info}", flush=True)


if __name__ == "__main__":
    main()
