"""Per-kernel counts of the Blackwell-specific SASS mnemonics in the built libraries (no GPU needed).
   python scripts/sass_summary.py > profiles/r2_sass_summary.txt
UTCHMMA = tcgen05.mma (".2CTA" = cta_group::2), LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG / UTMASTG = TMA load / store,
UTCBAR = tcgen05.commit, HMMA = legacy mma.sync, MUFU.EX2 = exp2 on the XU pipe, FFMA2 = packed fp32x2 FMA."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "MUFU.EX2", "FFMA2", "F2FP.BF16", "F2FP.F16",
       "SYNCS", "ELECT"]
for lib in ("libtaste_b200.so", "libtaste_b200_f16.so"):
    path = os.path.join(ROOT, "taste_spokenlm_b200", lib)
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    print(f"== {lib} ({os.path.getsize(path)} bytes), cuobjdump -sass, sm_100a ==")
    cur, counts, total = None, collections.OrderedDict(), collections.Counter()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            cur = cur.replace("taste::", "")
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        total[cur] += 1
        for p in PAT:
            if op.startswith(p):
                counts[cur][p] += 1
                break
    w = max(len(k) for k in counts)
    print("kernel".ljust(w), "instrs", " ".join(p.rjust(12) for p in PAT))
    for k, c in counts.items():
        print(k.ljust(w), str(total[k]).rjust(6), " ".join(str(c.get(p, 0)).rjust(12) for p in PAT))
    print()
