"""Top stall sites per kernel from an `ncu --page source --csv` export (SASS view):
   python profiles/ncu_source_hot.py gpurun_out/r02/greedy_full_source.csv [top_n] [kernel_index,...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
only = set(int(x) for x in sys.argv[3].split(",")) if len(sys.argv) > 3 else None
kern, hdr, body, blocks = None, None, [], []
for r in rows:
    if r and r[0] == "Kernel Name":
        if kern is not None:
            blocks.append((kern, hdr, body))
        kern, hdr, body = r[1], None, []
    elif r and r[0] == "Address":
        hdr = r
    elif kern is not None and hdr is not None and len(r) >= 5:
        body.append(r)
if kern is not None:
    blocks.append((kern, hdr, body))
for bi, (k, h, b) in enumerate(blocks):
    if only is not None and bi not in only:
        continue
    si = h.index("# Samples")
    stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_")]
    tot = sum(int(r[si] or 0) for r in b)
    print(f"== [{bi}] {k[:140]}  ({len(b)} SASS lines, {tot} samples)")
    agg = {}
    for r in b:
        for i in stall_cols:
            if i < len(r) and r[i]:
                agg[h[i]] = agg.get(h[i], 0) + int(r[i])
    print("   stall reasons:", ", ".join(f"{n[6:]} {100 * v / max(tot, 1):.0f}%" for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
    for r in sorted(b, key=lambda r: -int(r[si] or 0))[:top]:
        why = sorted(((int(r[i]), h[i][6:]) for i in stall_cols if i < len(r) and r[i]), reverse=True)[:2]
        print(f"   {100 * int(r[si] or 0) / max(tot, 1):5.1f}%  {r[1].strip()[:90]:90s} {' '.join(f'{n}:{v}' for v, n in why)}")
