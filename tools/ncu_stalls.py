"""Stall-reason totals and the hottest SASS instructions of each kernel in an `ncu --page source --csv --print-source sass` export."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25; which = int(sys.argv[3]) if len(sys.argv) > 3 else None
rows = list(csv.reader(open(path)))
kern = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = dict(name=r[1], hdr=None, rows=[]); kern.append(cur); continue
    if cur is None: continue
    if r and r[0] == "Address": cur["hdr"] = r; continue
    if cur["hdr"] and len(r) == len(cur["hdr"]): cur["rows"].append(r)
for ki, k in enumerate(kern):
    if which is not None and ki != which: continue
    h = {n: i for i, n in enumerate(k["hdr"])}
    stalls = [n for n in k["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
    tot = {s: 0 for s in stalls}; samp = 0
    for r in k["rows"]:
        for s in stalls:
            try: tot[s] += int(r[h[s]])
            except ValueError: pass
        try: samp += int(r[h["# Samples"]])
        except ValueError: pass
    print(f"== kernel {ki}: {k['name'][:60]}  samples={samp}")
    print("   " + "  ".join(f"{s[6:]}={v} ({100*v/max(samp,1):.0f}%)" for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
    rs = sorted(k["rows"], key=lambda r: -int(r[h["# Samples"]] or 0))[:top]
    for r in rs:
        st = sorted(((int(r[h[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
        print(f"   {r[h['# Samples']]:>6s}  {r[h['Address']][-5:]}  {r[h['Source']][:90]:90s} {st}")
