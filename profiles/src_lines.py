"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per CUDA source line
(first kernel instance): share of executed warp instructions and of stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
hdrs = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
first_fn = None
lines = {}
for hi_n, hi in enumerate(hdrs):
    fn = rows[hi - 1][1] if rows[hi - 1] and rows[hi - 1][0] == 'Function Name' else ''
    if first_fn is None: first_fn = fn
    h = rows[hi]
    ie, isamp = h.index('Instructions Executed'), h.index('# Samples')
    end = hdrs[hi_n + 1] - 2 if hi_n + 1 < len(hdrs) else len(rows)
    for r in rows[hi + 1:end]:
        if len(r) <= max(ie, isamp) or r[2] != '-':   # keep the per-line rows (Address == '-')
            continue
        key = (rows[hi - 2][1].split('/')[-1] if rows[hi - 2][0] == 'File Path' else '', r[0])
        try: v, s = float(r[ie] or 0), float(r[isamp] or 0)
        except ValueError: continue
        a = lines.setdefault(key, [0, 0, r[1]])
        a[0] += v; a[1] += s
# both kernel instances are summed; shares are what matters
tot = sum(a[0] for a in lines.values()); ts = sum(a[1] for a in lines.values())
print("total warp-inst %.3g, samples %d" % (tot, ts))
for (f, ln), a in sorted(lines.items(), key=lambda kv: (kv[0][0], int(kv[0][1]))):
    if a[0] / tot * 100 >= thresh or a[1] / ts * 100 >= thresh:
        print("%-22s %4s inst %5.1f%% samp %5.1f%%  %s" % (f[:22], ln, a[0] / tot * 100, a[1] / ts * 100, a[2].strip()[:100]))
