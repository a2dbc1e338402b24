"""Split an `ncu --page source --csv` SASS listing at barriers and summarise each segment:
samples, warp instructions executed, FFMA share, top stall reasons."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
segs = []
cur = dict(n=0, samples=0, inst=0, ffma=0, lds=0, first=None, st=collections.Counter())
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    op = src.split()[1] if src.startswith("@") else src.split()[0]
    ie = int(float(r[ix["Instructions Executed"]] or 0)); sm = int(float(r[ix["# Samples"]] or 0))
    if cur["first"] is None: cur["first"] = r[ix["Address"]]
    cur["n"] += 1; cur["samples"] += sm; cur["inst"] += ie
    if op.startswith("FFMA"): cur["ffma"] += ie
    if op.startswith("LDS"): cur["lds"] += ie
    for s in stalls:
        v = r[ix[s]]
        if v and v != "0": cur["st"][s] += int(float(v))
    if op.startswith("BAR"):
        segs.append(cur)
        cur = dict(n=0, samples=0, inst=0, ffma=0, lds=0, first=None, st=collections.Counter())
segs.append(cur)
tot_s = sum(s["samples"] for s in segs); tot_i = sum(s["inst"] for s in segs)
print("total samples %d, warp inst %d, ffma %d (%.1f%%)" % (tot_s, tot_i, sum(s["ffma"] for s in segs), 100.0 * sum(s["ffma"] for s in segs) / max(1, tot_i)))
for k, s in enumerate(segs):
    if s["samples"] < tot_s * 0.003: continue
    top = ", ".join("%s %.0f%%" % (a.replace("stall_", ""), 100.0 * b / max(1, s["samples"])) for a, b in s["st"].most_common(6))
    print("seg %2d @%s: %5d sass, samples %5.1f%%, inst %5.1f%% (ffma %4.1f%%, lds %4.1f%%) | %s" % (
        k, s["first"][-5:], s["n"], 100.0 * s["samples"] / tot_s, 100.0 * s["inst"] / tot_i, 100.0 * s["ffma"] / max(1, s["inst"]), 100.0 * s["lds"] / max(1, s["inst"]), top))
