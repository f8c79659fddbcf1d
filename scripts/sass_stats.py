"""Static SASS statistics of k_accumulate<Bls12_377, false> in an object / library: instruction histogram of the accumulate loop
(the span of the last backward branch) and of the multiplier subroutine it calls, per mixed add.
  python scripts/sass_stats.py <file.o|.so> [kernel-name-fragment] [--dump]"""
import collections, re, subprocess, sys
path = sys.argv[1]
txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
pat = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "k_accumulateINS_9Bls12_377ELb0"
m = re.search(r"Function : (\S*" + pat + r"\S*)(.*?)(?=\n\s*Function :|\Z)", txt, re.S)
body = m.group(2)
ins = [(int(a, 16), re.sub(r"^@!?U?P\w+\s+", "", t.strip())) for a, t in re.findall(r"/\*([0-9a-f]{4,})\*/\s+([^;]+);", body)]
back = [(a, int(re.search(r"0x([0-9a-f]+)", t).group(1), 16)) for a, t in ins if t.startswith("BRA") and re.search(r"0x([0-9a-f]+)", t) and int(re.search(r"0x([0-9a-f]+)", t).group(1), 16) < a]
calls = sorted({int(re.search(r"0x([0-9a-f]+)", t).group(1), 16) for a, t in ins if t.startswith("CALL")})
loop = max(back, key=lambda b: b[0] - b[1]) if back else None
def hist(lo, hi):
    c = collections.Counter(t.split()[0] for a, t in ins if lo <= a <= hi)
    return sum(c.values()), c
print(f"{path}: {len(ins)} instructions in the function")
if loop:
    n, c = hist(loop[1], loop[0])
    ncall = sum(v for k, v in c.items() if k.startswith("CALL"))
    print(f" loop body {n} instr, {ncall} calls:", ", ".join(f"{k} {v}" for k, v in c.most_common(14)))
    tot = collections.Counter(c)
    rets = sorted(a for a, t in ins if t.startswith("RET"))
    for k, ca in enumerate(calls):
        end = min([r for r in rets if r > ca] or [ins[-1][0]])
        n2, c2 = hist(ca, end)
        times = sum(1 for a, t in ins if loop[1] <= a <= loop[0] and t.startswith("CALL") and f"0x{ca:x}" in t)
        print(f" subroutine @0x{ca:x}: {n2} instr, called {times}x per iteration:", ", ".join(f"{k} {v}" for k, v in c2.most_common(10)))
        for kk, v in c2.items():
            tot[kk] += v * times
    n3 = sum(tot.values())
    wide = tot["IMAD.WIDE.U32"] + tot.get("IMAD.WIDE", 0)
    other_imad = sum(v for k, v in tot.items() if k.startswith("IMAD") and not k.startswith("IMAD.WIDE"))
    print(f" per mixed add: {n3} warp-instructions, IMAD.WIDE {wide}, other IMAD.* {other_imad}, MOV {tot['MOV']}, LDL/STL {sum(v for k, v in tot.items() if k[:3] in ('LDL', 'STL'))}")
if "--dump" in sys.argv:
    print(body)
