#!/usr/bin/env python3
"""Builds oracle/_ref/libcombined_msm.so: the native half of the reference's fastest prize-1a entry
(P1A/combined-top-solutions/combined-msm/MSM.cu, one translation unit, no dependency besides the CUDA runtime),
compiled for sm_100 FROM THE SOURCES WHERE THEY LIE -- nothing is copied into this repository.

Why not plain `nvcc MSM.cu`: the entry's planning kernels spin on shared memory through an inline-PTX helper that issues a
plain `ld.shared.u32` (yrrid-ff-ec/asm.cu:13-18, used e.g. in Partition1024.cu:132-137).  A weak load may be moved by the
assembler, and ptxas 12.9 does move it out of the spin loop (SASS of partition1024Kernel: the back edge of the loop returns
to the vote, behind the LDS), so on this toolchain MSMRun never returns -- at any size, on any GPU (measured on the B200:
profiles/r01_reference_entry.txt).  This script runs nvcc's own build steps (`nvcc -dryrun`) and, between the compiler and
the assembler, rewrites the 157 scalar `ld.shared.u32` of the generated PTX to `ld.volatile.shared.u32` (the vector loads
of the hot kernel are left alone).  Sources unmodified; the patch is to a build intermediate and is stated wherever a number
from this library is quoted.  `--plain` builds without the patch.

Test infrastructure only: nothing in the product, tests/ or bench.py loads this library (scripts/retime_reference_entry.py does).
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
P1A = os.environ.get("REF_ROOT", "/root/reference") + "/open-division/prize1-msm/prize1a-msm-gpu"
ENTRIES = {"combined": (P1A + "/combined-top-solutions/combined-msm/MSM.cu", "libcombined_msm"),      # the fastest entry (default)
           "yrrid": (P1A + "/yrrid/yrrid-msm/MSM.cu", "libyrrid_msm")}                                 # its parent: same driver, yrrid's own field code
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def patch_ptx(path):
    """--volatile=planning: in every function of the PTX except the two arithmetic kernels (computeBucketSums, reduceBuckets) and
    the table precompute, make ALL shared-memory loads and stores volatile, so the assembler keeps the program order the
    entry's hand-written inter-warp protocols (byte stores polled by another warp, Partition1024.cu:124-147, 223-240) rely on."""
    out, keep, n = [], False, 0
    for line in open(path):
        m = re.match(r"\s*(?:\.visible\s+|\.weak\s+)?\.(entry|func)\s+(?:\([^)]*\)\s*)?(\w+)", line)
        if m:
            keep = any(k in m.group(2) for k in ("computeBucketSums", "reduceBuckets", "precomputePoints"))
        if not keep:
            new = re.sub(r"\b(ld|st)\.shared\.", r"\1.volatile.shared.", line)
            n += new != line
            line = new
        out.append(line)
    open(path, "w").writelines(out)
    print(f"patched {n} shared loads / stores outside the arithmetic kernels")


def main():
    if "--patch-ptx" in sys.argv:
        return patch_ptx(sys.argv[sys.argv.index("--patch-ptx") + 1])
    plain = "--plain" in sys.argv
    wide = "--volatile=planning" in sys.argv     # every shared load AND store of the planning kernels volatile (not only the u32 loads)
    diag = "--diag" in sys.argv                  # build oracle/ref_entry_diag.cu (includes the entry's MSM.cu) instead: adds MSMDiagOneBatch
    entry = "yrrid" if "--entry=yrrid" in sys.argv else "combined"
    REF, stem = ENTRIES[entry]
    out_dir = os.path.join(HERE, "_ref")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, stem + ("_plain" if plain else "") + ("_v2" if wide else "") + ("_diag" if diag else "") + ".so")
    src = [os.path.join(HERE, "ref_entry_diag.cu"), "-I", os.path.dirname(REF)] if diag else [REF]
    cmd = [NVCC, "-O3", "-gencode", "arch=compute_100,code=sm_100", "-w", "-shared", "-Xcompiler", "-fPIC", "-o", out] + src
    dry = subprocess.run(cmd + ["-dryrun"], capture_output=True, text=True, check=True).stderr
    steps = [l[3:] for l in dry.splitlines() if l.startswith("#$ ")]
    script, patched = ["set -e"], 0
    for s in steps:
        m = re.match(r"([A-Za-z_]+)=", s)
        if m:                                        # nvcc's environment lines: keep the ones the later steps use
            if m.group(1) in ("PATH", "LD_LIBRARY_PATH", "CICC_PATH", "NVVMIR_LIBRARY_DIR", "TOP"):
                script.append("export " + s)
            continue
        if s.startswith("rm "):
            s = s.replace("rm ", "rm -f ", 1)        # nvcc's clean-up of files an earlier step may not have left behind
        script.append(s)
        if "cicc" in s.split()[0] and not plain:
            m = re.search(r'-o "([^"]+\.ptx)"', s)
            if not m:
                raise SystemExit("could not find the PTX output of the compiler step")
            if wide:
                script.append(f"python3 {os.path.join(HERE, 'build_ref_entry.py')} --patch-ptx \"{m.group(1)}\"")
            else:
                script.append(f"sed -i 's/ld\\.shared\\.u32/ld.volatile.shared.u32/g' \"{m.group(1)}\"")
                script.append(f"echo patched $(grep -c 'ld.volatile.shared.u32' \"{m.group(1)}\") scalar shared loads")
            patched += 1
    if not plain and patched != 1:
        raise SystemExit(f"expected one compiler step, found {patched}")
    steps_file = os.path.join(out_dir, f"build_steps_{entry}{'_v2' if wide else ''}{'_diag' if diag else ''}.sh")
    with open(steps_file, "w") as f:
        f.write("\n".join(script) + "\n")
    rc = subprocess.run(["bash", steps_file]).returncode
    if rc != 0:
        raise SystemExit(f"a build step failed (exit {rc}); the steps are in {steps_file}")
    print("built", out)


if __name__ == "__main__":
    main()
