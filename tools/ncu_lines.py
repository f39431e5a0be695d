#!/usr/bin/env python
"""Aggregates `ncu --page source --csv --print-source cuda,sass` output per CUDA source line:
   warp instructions executed, thread instructions, stall samples.  usage: ncu_lines.py file.csv [top]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur_file = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])   # (file, line) -> [inst, thread inst, samples, text]
line_no = None; line_txt = ""
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; iI = r.index("Instructions Executed"); iT = r.index("Thread Instructions Executed"); iS = r.index("# Samples"); continue
    if r[0] != "":
        line_no = int(r[0]); line_txt = r[1]; agg[(cur_file, line_no)][3] = line_txt; continue
    if r[2] in ("...", "-") or hdr is None: continue
    try:
        a = agg[(cur_file, line_no)]
        a[0] += int(r[iI]); a[1] += int(r[iT]); a[2] += int(r[iS])
    except ValueError:
        pass
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[2] for a in agg.values())
print("total warp inst %d, samples %d" % (tot_i, tot_s))
rows = sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]
for (f, l), a in rows:
    print("%5.2f%% inst %5.2f%% smp  thr/inst %4.1f  %s:%d  %s" % (100.0 * a[0] / tot_i, 100.0 * a[2] / max(tot_s, 1), a[1] / max(a[0], 1), f, l, a[3].strip()[:110]))

# phase buckets of ms_kernels.cu (line ranges of the current tree; adjust when the file moves)
if len(sys.argv) > 3 and sys.argv[3] == "phases":
    B = [(35, 146, "helpers (st_v4, rec, ring_slot, dense_base)"), (147, 207, "block scan / bitonic"), (208, 234, "journal_raw"),
         (235, 291, "wheel"), (292, 362, "emit_one"), (626, 746, "node programs"), (747, 878, "commit"), (1040, 1130, "small scans"),
         (1220, 1318, "prologue + ticket fetch"), (1319, 1362, "injector"), (1363, 1390, "endpoint setup"), (1391, 1473, "PA1 load"),
         (1474, 1493, "PA2 seen test"), (1494, 1592, "PB ordering"), (1593, 1625, "PC first-sight insert"), (1626, 1726, "seq families"),
         (1727, 1774, "winners + counts"), (1775, 1850, "scan + gset"), (1851, 1925, "PD claims"), (1926, 1964, "PE1 recv records"),
         (1965, 1975, "own_map"), (1976, 2094, "PE2 emissions"), (2095, 2143, "epilogue")]
    ph = collections.defaultdict(lambda: [0, 0])
    for (f, l), a in agg.items():
        name = "other:" + f
        if f == "ms_kernels.cu":
            for lo, hi, nm in B:
                if lo <= l <= hi: name = nm; break
        ph[name][0] += a[0]; ph[name][1] += a[2]
    print("==== phases")
    for nm, a in sorted(ph.items(), key=lambda kv: -kv[1][0]):
        print("%5.1f%% inst %5.1f%% smp  %s" % (100.0 * a[0] / tot_i, 100.0 * a[1] / max(tot_s, 1), nm))
