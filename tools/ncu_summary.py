"""Summarise an .ncu-rep (raw + source pages) into text. Usage: ncu_summary.py file.ncu-rep"""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__block_size', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__t_sectors_srcunit_tex_op_write.sum', 'sm__inst_executed_pipe_lsu.sum']
for vals in rows[2:]:
    print("== kernel:", vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
    for i, h in enumerate(hdr):
        if h in want:
            print(f"  {h:72s} {units[i]:14s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]; data = [r for r in rows[2:] if len(r) == len(h2)]
iS, iE, iSm = h2.index('Source'), h2.index('Instructions Executed'), h2.index('# Samples')
tot = sum(int(r[iE]) for r in data); ts = sum(int(r[iSm]) for r in data)
print(f"SASS instructions: {len(data)}, executed warp-inst {tot}, samples {ts}")
stall = [(h, i) for i, h in enumerate(h2) if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[i] or 0) for r in data) for h, i in stall}
print("stall reasons:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
byop = collections.Counter(); samp = collections.Counter()
for r in data:
    t = r[iS].split()
    op = t[1] if t[0].startswith('@') else t[0]
    byop[op] += int(r[iE]); samp[op] += int(r[iSm])
print("top opcodes by executed:", [(o, f"{c/tot*100:.1f}%") for o, c in byop.most_common(14)])
print("top opcodes by stall samples:", [(o, f"{c/ts*100:.1f}%") for o, c in samp.most_common(10)])
