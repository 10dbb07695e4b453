"""Summarise an ncu report: key raw metrics per kernel and a per-phase (barrier-delimited) breakdown of the SASS
(instructions, FP64 instructions, shared-memory wavefronts, stall samples).  Usage:
  ncu -i rep --page raw --csv > raw.csv; ncu -i rep --page source --csv > src.csv; python tools/ncu_phases.py raw.csv src.csv n_eles"""
import csv, collections, re, sys
raw, src, ne = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows = list(csv.reader(open(raw)))
hdr = rows[0]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "smsp__cycles_active.avg"]
keys += [k for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")][:70])
    for k in keys:
        if k in hdr:
            print("   %-88s %s" % (k, r[hdr.index(k)]))
rows = list(csv.reader(open(src)))
kern = None; data = collections.OrderedDict(); h = None
for r in rows:
    if r and r[0] == "Kernel Name": kern = r[1]; data.setdefault(kern, []); continue
    if r and r[0] == "Address": h = r; continue
    if kern and len(r) > 10: data[kern].append(r)
ix = h.index("Instructions Executed"); isamp = h.index("# Samples"); iw = h.index("L1 Wavefronts Shared"); iwi = h.index("L1 Wavefronts Shared Ideal")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
for k, v in data.items():
    half = len(v) // 2
    if [r[1] for r in v[:half]] == [r[1] for r in v[half:]]: v = v[:half]  # listing repeated
    print(k[:60], len(v))
    seg = []; new = lambda i: dict(n=0, ins=0, fp=0, s=0, w=0, wi=0, lds=0, ldg=0, start=i, st=collections.Counter()); cur = new(0)
    for i, r in enumerate(v):
        cur['n'] += 1; cur['ins'] += int(r[ix]); cur['s'] += int(r[isamp]); cur['w'] += int(r[iw]); cur['wi'] += int(r[iwi])
        for c in stall_cols: cur['st'][h[c]] += int(r[c])
        if re.search(r'\b(DFMA|DMUL|DADD)', r[1]): cur['fp'] += int(r[ix])
        if re.search(r'\b(LDS|STS)', r[1]): cur['lds'] += int(r[ix])
        if re.search(r'\b(LDG|STG|LDGSTS)', r[1]): cur['ldg'] += int(r[ix])
        if 'BAR.SYNC' in r[1] or ('EXIT' in r[1] and int(r[ix]) > 0):
            seg.append(cur); cur = new(i + 1)
    seg.append(cur)
    T = sum(s['ins'] for s in seg); S = sum(s['s'] for s in seg)
    for s in seg:
        if s['ins'] == 0: continue
        top = ", ".join("%s %.0f%%" % (a.replace("stall_", ""), 100 * b / max(1, s['s'])) for a, b in s['st'].most_common(4))
        print('  seg@%5d n=%5d instr/el %6.0f fp64 %5.0f lds/sts %5.0f ldg/stg %4.0f smemwave %5.0f (ideal %5.0f) samples %4.1f%% | %s' % (
            s['start'], s['n'], s['ins'] / ne, s['fp'] / ne, s['lds'] / ne, s['ldg'] / ne, s['w'] / ne, s['wi'] / ne, 100 * s['s'] / S, top))
    print('  total instr/el %.0f' % (T / ne))
