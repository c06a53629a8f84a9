"""Hot CUDA source lines of each kernel in an ncu report: `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > cs.csv;
python scripts/ncu_lines.py cs.csv [top]` prints, per kernel, the lines with the most executed warp instructions and stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = fn = hdr = None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name":
        fn = r[1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) > 8 and r[0].isdigit():
        num = lambda name: int(r[hdr.index(name)]) if r[hdr.index(name)].isdigit() else 0
        ie = num("Instructions Executed"); smp = num("# Samples"); te = num("Thread Instructions Executed")
        a = agg.setdefault((fn, cur, int(r[0])), [0, 0, 0, r[1]]); a[0] += ie; a[1] += smp; a[2] += te
for kern in sorted(set(k[0] for k in agg)):
    tot = sum(v[0] for k, v in agg.items() if k[0] == kern); ts = sum(v[1] for k, v in agg.items() if k[0] == kern)
    print("=====", kern[:70], "inst", tot, "samples", ts)
    items = sorted([(v[0], v[1], v[2], k[1], k[2], v[3]) for k, v in agg.items() if k[0] == kern], reverse=True)[:top]
    for ie, smp, te, f, l, src in items:
        print("%5.1f%% inst %5.1f%% smp %4.1f thr  %s:%d  %s" % (100 * ie / tot, 100 * smp / max(ts, 1), te / max(ie, 1), f.replace('pm_', ''), l, src.strip()[:100]))
