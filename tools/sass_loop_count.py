#!/usr/bin/env python
"""Static issue-cost estimate of a kernel's outer loop from its SASS: instructions between the head of the outermost
backward branch and the branch itself, minus inner loops (rarely-run refinement), packed FP32 instructions counted twice
(a packed f32x2 instruction occupies the FP32 pipe for two issue cycles: tools/pipe_probe.py)."""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
ins, on = [], False
for line in txt.splitlines():
    if "Function :" in line:
        on = pat in line
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if on and m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
def target(t):
    m = re.search(r"BRA\s+(?:\w+,\s*)?(0x[0-9a-f]+)", t)
    return int(m.group(1), 16) if m else None
if not ins:
    sys.exit(f"no function matching {pat!r}")
back = [(a, target(t)) for a, t in ins if "BRA" in t and target(t) is not None and target(t) < a]
outer = max(back, key=lambda x: x[0] - x[1])
inner = [b for b in back if b != outer and outer[1] <= b[1] and b[0] <= outer[0]]
# a forward branch that jumps over an inner loop marks the start of the skipped (rare) region
skip = []
for a, t in ins:
    tg = target(t)
    if "BRA" in t and tg and tg > a and outer[1] <= a <= outer[0]:
        for ia, it in inner:
            if it <= a <= ia and tg > ia:
                skip.append((a, tg))
hist = collections.Counter()
for a, t in ins:
    if not (outer[1] <= a <= outer[0]):
        continue
    if any(lo < a < hi for lo, hi in skip):
        continue
    op = re.sub(r"^@!?U?P\w+\s+", "", t).split()[0].split(".")[0]
    hist[op] += 1
packed = sum(c for o, c in hist.items() if o in ("FFMA2", "FMUL2", "FADD2"))
total = sum(hist.values())
print(f"{pat}: loop {outer[1]:#x}..{outer[0]:#x}, skipped {[(hex(a), hex(b)) for a, b in skip]}")
print(f"  instructions {total}, packed {packed}, issue cycles (packed x2) {total + packed}")
print("  " + ", ".join(f"{o} {c}" for o, c in hist.most_common(24)))
