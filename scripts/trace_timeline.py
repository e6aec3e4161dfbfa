"""Chronological view of an attn_trace.py log: python scripts/trace_timeline.py gpurun_out/trace_68.log [block]"""
import re, sys
f = sys.argv[1]
b0 = int(sys.argv[2]) if len(sys.argv) > 2 else 16
span = int(sys.argv[3]) if len(sys.argv) > 3 else 3400
ev = []
for line in open(f):
    m = re.match(r'(blk\s+(\d+))?\s*(WG0|WG1|MMA0|MMA1):\s+(.*)', line)
    if not m:
        continue
    if m.group(2):
        blk = int(m.group(2))
    for kv in m.group(4).split():
        k, v = kv.split('=')
        v = int(v)
        if v > 0:
            ev.append((v, m.group(3), blk, k))
ev.sort()
t0 = [v for v, w, b, k in ev if w == 'WG0' and b == b0 and k == 'enter'][0]
for v, w, b, k in ev:
    if t0 - 200 <= v <= t0 + span:
        print(f"{v - t0:6d} {'' if w in ('WG0', 'MMA0') else ' ' * 34}{w:5s} g={b} {k}")
