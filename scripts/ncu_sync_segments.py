"""Split an `ncu --page source --csv --print-source sass` dump of a warp-specialised kernel at its synchronisation
instructions (setmaxnreg, bar.sync, mbarrier try_wait / arrive) and print samples, instruction mix and stall reasons
per segment.  usage: ncu_sync_segments.py dump.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = 0
segs = []; cur = {'start': 0, 'samples': 0, 'inst': 0, 'ffma2': 0, 'loc': 0, 'st': {}}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for i, r in enumerate(data):
    src = r[ix['Source']]
    if any(k in src for k in ('USETMAXREG', 'BAR.SYNC', 'SYNCS.PHASECHK', 'SYNCS.ARRIVE')):
        segs.append(cur); cur = {'start': i, 'samples': 0, 'inst': 0, 'ffma2': 0, 'loc': 0, 'st': {}, 'mark': src.strip()[:50]}
    s = int(r[ix['# Samples']] or 0); n = int(r[ix['Instructions Executed']] or 0)
    cur['samples'] += s; cur['inst'] += n
    if 'FFMA2' in src: cur['ffma2'] += n
    if 'LDL' in src or 'STL' in src: cur['loc'] += n
    for c in stall_cols:
        v = int(r[ix[c]] or 0)
        if v: cur['st'][c] = cur['st'].get(c, 0) + v
    tot += s
segs.append(cur)
print("total samples %d" % tot)
for s in segs:
    if s['samples'] < tot * 0.004: continue
    top = sorted(s['st'].items(), key=lambda kv: -kv[1])[:5]
    print("seg@%5d %-50s samples %5.1f%% inst %10d ffma2 %3.0f%% local %8d | %s" % (
        s['start'], s.get('mark', ''), 100 * s['samples'] / tot, s['inst'], 100 * s['ffma2'] / max(1, s['inst']), s['loc'],
        ', '.join('%s %d%%' % (k[6:], 100 * v / max(1, s['samples'])) for k, v in top)))
