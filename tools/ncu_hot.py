"""Hot instructions of one kernel from `ncu -i rep --page source --csv --kernel-name regex:NAME` output.
usage: python tools/ncu_hot.py src.csv [min_percent]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for si, s in enumerate(start):
    hdr = rows[s]
    ix = {h: i for i, h in enumerate(hdr)}
    end = start[si + 1] - 1 if si + 1 < len(start) else len(rows)
    body = [r for r in rows[s + 1:end] if len(r) == len(hdr)]
    tot = sum(int(r[ix['# Samples']]) for r in body)
    print(rows[s - 1][1] if s else "", 'total samples', tot, 'instructions', len(body))
    for n, r in enumerate(body):
        smp = int(r[ix['# Samples']])
        if smp > tot * thr / 100:
            print("%5d %-72s %6d %5.1f%%  long_sb %s bar %s wait %s short_sb %s  | smem wavefronts %s ideal %s" % (
                n, r[ix['Source']].strip()[:72], smp, 100 * smp / tot, r[ix['stall_long_sb']], r[ix['stall_barrier']], r[ix['stall_wait']],
                r[ix['stall_short_sb']], r[ix['L1 Wavefronts Shared']], r[ix['L1 Wavefronts Shared Ideal']]))
    break
