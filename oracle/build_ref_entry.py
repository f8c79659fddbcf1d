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

Second, decisive, incompatibility (found on the B200 with oracle/ref_entry_diag.cu: partition1024Kernel never returns, with
or without the volatile patch): the entry addresses shared memory through inline PTX with RAW window offsets
(`st.shared.b8 [highBits*60 + 4096 + ...]`, Partition1024.cu:208-233; asm.cu / Support.cu helpers) and assumes its
`extern __shared__` array starts at offset 0 of the shared window.  That held for sm_80 builds in 2022; on sm_90 / sm_100 the
first 1 KB of the window is reserved and the compiler places the dynamic array at 0x400 (SASS of this very kernel:
`MOV R17, 0x400 ... ATOMS.ADD [R22]` for `atomicAdd(&counts[..])`, while the hand-written stores go to `R16 * 0x3c + 0x1000`).
The counter a warp bumps through the C array and the counter it resets through the raw address are then 1 KB apart, and the
`while(!__all_sync(processed))` loop of Partition1024.cu:223 never ends.  `--smem-base=1024` (default for the patched build)
adds that base to the shared-window operand of every shared-memory instruction INSIDE INLINE-ASM BLOCKS of the generated
PTX (`// begin inline asm` ... `// end inline asm`), which is exactly the set of raw accesses; compiler-generated accesses
are untouched.  Again a patch to a build intermediate, stated wherever a number from this library is quoted.

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


def patch_ptx(path, wide, smem_base):
    """The build-intermediate patches (see the module docstring).  wide: all shared loads AND stores of the planning
    kernels volatile instead of only the scalar u32 loads."""
    out, keep, inline, n_vol, n_base = [], False, False, 0, 0
    for line in open(path):
        m = re.match(r"\s*(?:\.visible\s+|\.weak\s+)?\.(entry|func)\s+(?:\([^)]*\)\s*)?(\w+)", line)
        if m:
            keep = any(k in m.group(2) for k in ("computeBucketSums", "reduceBuckets", "precomputePoints"))
        if "begin inline asm" in line:
            inline = True
        elif "end inline asm" in line:
            inline = False
        elif inline and smem_base and re.search(r"\b(ld|st|atom|red|cp\.async\.\w+)\.(volatile\.)?shared\b", line):
            new = re.sub(r"\[(%r\d+)\]", lambda mm: f"[{mm.group(1)}+{smem_base}]", line, count=1)      # the shared-window operand
            n_base += new != line
            line = new
        if wide and not keep:
            new = re.sub(r"\b(ld|st)\.shared\.", r"\1.volatile.shared.", line)
        else:
            new = line.replace("ld.shared.u32", "ld.volatile.shared.u32")
        n_vol += new != line
        out.append(new)
    open(path, "w").writelines(out)
    print(f"patched: {n_vol} shared accesses made volatile, {n_base} raw shared-window operands rebased by {smem_base}")


def main():
    if "--patch-ptx" in sys.argv:
        return patch_ptx(sys.argv[sys.argv.index("--patch-ptx") + 1], "--volatile=planning" in sys.argv,
                         next((int(a.split("=")[1]) for a in sys.argv if a.startswith("--smem-base=")), 1024))
    plain = "--plain" in sys.argv
    wide = "--volatile=planning" in sys.argv     # every shared load AND store of the planning kernels volatile (not only the u32 loads)
    smem_base = next((int(a.split("=")[1]) for a in sys.argv if a.startswith("--smem-base=")), 1024)
    diag = "--diag" in sys.argv                  # build oracle/ref_entry_diag.cu (includes the entry's MSM.cu) instead: adds MSMDiagOneBatch
    entry = "yrrid" if "--entry=yrrid" in sys.argv else "combined"
    REF, stem = ENTRIES[entry]
    out_dir = os.path.join(HERE, "_ref")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, stem + ("_plain" if plain else "") + ("_v2" if wide else "") + ("_diag" if diag else "") + ".so")
    src = [os.path.join(HERE, "ref_entry_diag.cu"), "-I", os.path.dirname(REF)] if diag else [REF]
    cmd = [NVCC, "-O3", "-gencode", "arch=compute_100,code=sm_100", "-w", "-shared", "-Xcompiler", "-fPIC", "-o", out] + (["-lineinfo"] if diag else []) + src
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
            script.append(f"python3 {os.path.join(HERE, 'build_ref_entry.py')} --patch-ptx \"{m.group(1)}\" {'--volatile=planning' if wide else ''} --smem-base={smem_base}")
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
