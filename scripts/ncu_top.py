"""Summarise an .ncu-rep: key metrics + top stall instructions.  usage: python scripts/ncu_top.py rep [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct", "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct",
        "launch__registers_per_thread", "launch__grid_size", "sm__throughput.avg.pct", "l1tex__throughput.avg.pct",
        "launch__shared_mem_per_block_dynamic", "sm__inst_executed.sum", "smsp__inst_executed.sum"]
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(w) or h == w for w in want):
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
print(rows[0][:2])
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
f = lambda x: float(x) if x.replace('.', '', 1).isdigit() else 0.0
tot = sum(f(r[ix['# Samples']]) for r in data)
sc = [h for h in hdr if h.startswith('stall_') and 'Not' not in h]
agg = sorted(((h, sum(f(r[ix[h]]) for r in data)) for h in sc), key=lambda t: -t[1])[:8]
print("total samples", tot, agg)
for r in sorted(data, key=lambda r: -f(r[ix['# Samples']]))[:topn]:
    st = sorted(((h, f(r[ix[h]])) for h in sc), key=lambda t: -t[1])[:2]
    print(r[ix['Address']][-5:], r[ix['# Samples']], r[ix['Instructions Executed']], r[ix['Source']][:80], st)
