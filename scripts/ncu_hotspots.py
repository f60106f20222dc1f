"""Per-instruction stall hot spots of one kernel from an ncu report that was captured with --import-source on:
    ncu -i rep --page source --csv > src.csv ;  python scripts/ncu_hotspots.py src.csv [top]"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1]))]
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tot = sum(int(r[ix['# Samples']]) for r in data)
print("total samples", tot, " total warp instructions", sum(int(r[ix['Instructions Executed']]) for r in data))
keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(int(r[ix[k]]) for r in data) for k in keys}
print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:n_top]:
    stalls = {k: int(r[ix[k]]) for k in keys}
    big = [(k[6:], v) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:3] if v]
    print(f"{int(r[ix['# Samples']]):6d} {100*int(r[ix['# Samples']])/tot:5.1f}%  exec={r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:58]:58s} {big}")
