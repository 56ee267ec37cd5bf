"""Debug helper: how far apart (in issued FP64 instructions) are dependent DFMAs in each barrier-separated phase of
a kernel's SASS?  ptxas orders the half-band stages; a distance of 1-2 means the FP64 pipe waits on its own latency.
Usage: cuobjdump -sass -fun <mangled> obj.o | python tools/sass_chains.py"""
import re, sys
ins = re.compile(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)\s*(.*?);")
phase, phases = [], []
for line in sys.stdin:
    m = ins.search(line)
    if not m:
        continue
    op, args = m.group(1), m.group(2)
    if op.startswith("BAR"):
        phases.append(phase); phase = []
    else:
        phase.append((op, [a.strip().replace(".reuse", "") for a in args.split(",")]))
phases.append(phase)
for i, ph in enumerate(phases):
    last = {}
    n = 0
    dists = []
    lds = sum(1 for op, _ in ph if op.startswith("LDS"))
    for op, a in ph:
        if op in ("DFMA", "DMUL", "DADD"):
            n += 1
            srcs = a[1:]
            d = min([n - last[s] for s in srcs if s in last] or [99])
            dists.append(d)
            last[a[0]] = n
    if n >= 20:
        dd = sorted(dists)
        short = sum(1 for d in dists if d <= 2) / n
        print(f"phase {i:2d}: {len(ph):5d} instr, {n:4d} FP64, {lds:3d} LDS, dep distance median {dd[n // 2]:2d}, <=2: {short:.0%}, <=4: {sum(1 for d in dists if d <= 4) / n:.0%}")
