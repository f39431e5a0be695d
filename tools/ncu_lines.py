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

# phase buckets of ms_kernels.cu: line ranges found from marker comments in the current source
if len(sys.argv) > 3 and sys.argv[3] == "phases":
    import os
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "maelstrom_b200", "csrc", "ms_kernels.cu")).read().split("\n")
    marks = [("helpers (st_v4, rec, ring_slot, dense_base)", "namespace msd {"), ("block scan / bitonic", "__device__ uint64_t block_excl_scan("),
             ("emit ctx / journal_raw", "struct EmitCtx {"), ("wheel", "timing wheel (pooled chains)"), ("emit_one", "void emit_one("),
             ("gen clients / barrier / snapshot / release", "closed-loop clients (MS_KIND_GEN_CLIENT)"), ("node programs", "// ------------------------------------------------------------------ node programs"),
             ("commit", "// ------------------------------------------------------------------ round commit"), ("small scans + services", "emit_count_meta(uint32_t workload"),
             ("prologue + ticket fetch", "template <int CLS, int WL, bool FIX = false>"), ("injector", "injector slice"), ("endpoint setup", "endpoint CTA"),
             ("PA1 load", "// PA1:"), ("PA2 seen test", "// PA2:"), ("PB ordering", "// PB:"), ("PC first-sight insert", "// PC:"),
             ("seq families / services", "Raft / txn-list-append node: the step is sequential"), ("winners + counts", "resolve winners and publish packed counts"),
             ("scan + gset", "const uint64_t tot = block_excl_scan(aux"), ("PD claims", "// PD:"), ("PE1 recv records", "// PE1:"), ("PE2 emissions", "// PE2:"),
             ("epilogue", "ticket epilogue"), ("journal kernels", "k_journal_expand (K3)")]
    B = []
    for name, needle in marks:
        ln = next((i + 1 for i, l in enumerate(src) if needle in l), None)
        if ln: B.append([ln, name])
    B.sort()
    ph = collections.defaultdict(lambda: [0, 0])
    for (f, l), a in agg.items():
        name = "other:" + f
        if f == "ms_kernels.cu":
            name = "other:ms_kernels.cu"
            for lo, nm in B:
                if l >= lo: name = nm
        ph[name][0] += a[0]; ph[name][1] += a[2]
    print("==== phases")
    for nm, a in sorted(ph.items(), key=lambda kv: -kv[1][0]):
        print("%5.1f%% inst %5.1f%% smp  %s" % (100.0 * a[0] / tot_i, 100.0 * a[1] / max(tot_s, 1), nm))
