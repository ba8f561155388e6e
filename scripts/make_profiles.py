"""Copies the judged evidence from gpurun_out/ (scratch) into profiles/ (tracked): bench lines, the ncu
launch list of the default bench command, and text summaries of the full ncu captures.
usage: python scripts/make_profiles.py r01"""
import csv, io, json, os, subprocess, sys, collections

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")
os.makedirs(P, exist_ok=True)


def ncu_csv(rep, page):
    return list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"],
                                                      capture_output=True, text=True).stdout)))


def summarise(rep, out, want_extra=()):
    rows = ncu_csv(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_active.avg",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum"] + list(want_extra)
    lines = [f"# ncu --set full summary of {os.path.basename(rep)} (one launch; cold-cache, serialised replay)", ""]
    for h, u, v in zip(hdr, units, vals):
        if any(h == w or h.endswith("." + w) or h.endswith(w) for w in want):
            lines.append(f"{h} [{u}] = {v}")
    src = ncu_csv(rep, "source")
    shdr = src[1]; ix = {h: i for i, h in enumerate(shdr)}; data = src[2:]
    f = lambda x: float(x) if x.replace('.', '', 1).isdigit() else 0.0
    sc = [h for h in shdr if h.startswith("stall_") and "Not" not in h]
    tot = sum(f(r[ix["# Samples"]]) for r in data)
    agg = sorted(((h, sum(f(r[ix[h]]) for r in data)) for h in sc), key=lambda t: -t[1])[:8]
    lines += ["", f"warp-state samples: {int(tot)}; by reason: " + ", ".join(f"{h[6:]}={int(v)}" for h, v in agg), "",
              "top instructions by samples (address tail, samples, executions, SASS, top stall):"]
    for r in sorted(data, key=lambda r: -f(r[ix["# Samples"]]))[:25]:
        st = sorted(((h, f(r[ix[h]])) for h in sc), key=lambda t: -t[1])[0]
        lines.append(f"  {r[ix['Address']][-5:]} {r[ix['# Samples']]:>7} {r[ix['Instructions Executed']]:>10} "
                     f"{r[ix['Source']][:70]:70s} {st[0][6:]}={int(st[1])}")
    mn = collections.Counter()
    for r in data:
        op = r[ix["Source"]].replace("@P0", "").replace("@!P0", "").split()
        for tok in op[:2]:
            for key in ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR", "SYNCS"):
                if tok.startswith(key):
                    mn[tok] += 1
    lines += ["", "Blackwell-specific SASS in this kernel: " + ", ".join(f"{k} x{v}" for k, v in sorted(mn.items()))]
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


# bench lines
rec = os.path.join(G, "record.log")
if os.path.exists(rec):
    out = []
    for line in open(rec):
        if line.startswith("=== python bench.py") or line.startswith("{") or "passed" in line or line.startswith("smoke ok"):
            out.append(line.rstrip())
    open(os.path.join(P, f"{tag}_bench_lines_1gpu.log"), "w").write("\n".join(out) + "\n")
    print("wrote bench lines")
for name, src in (("bench_2gpu.log", "r2d_n2.log"), ("bench_8gpu.log", "r2d_n8.log")):
    p = os.path.join(G, src if os.path.exists(os.path.join(G, src)) else name)
    if os.path.exists(p):
        keep = [l for l in open(p) if l.startswith("{") or l.startswith("=== python") or "passed" in l]
        open(os.path.join(P, f"{tag}_{name}"), "w").writelines(keep)
lc = os.path.join(G, "launches_bench_c2.csv")
if os.path.exists(lc):
    rows = list(csv.reader(open(lc)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
    keep = ["ID,Kernel Name,Grid Size,Block Size,gpu__time_duration.sum,unit"]
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr): continue
        name = r[ix["Kernel Name"]].split("(")[0]
        keep.append(",".join([r[ix["ID"]], name, r[ix["Grid Size"]].replace(",", " "), r[ix["Block Size"]].replace(",", " "),
                              r[ix["Metric Value"]], r[ix["Metric Unit"]]]))
        v = float(r[ix["Metric Value"]]); u = r[ix["Metric Unit"]]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    keep.append("")
    keep.append("# per-kernel totals (us); `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` under ncu "
                "--metrics gpu__time_duration.sum --clock-control none")
    for k, (c, t) in agg.items():
        keep.append(f"# {k}: n={c} total={t:.1f}us avg={t / c:.1f}us")
    open(os.path.join(P, f"{tag}_launches_bench_c2.csv"), "w").write("\n".join(keep) + "\n")
    print("wrote launch list")
def traffic_bytes(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    tot = 0.0
    for h, u, v in zip(hdr, units, vals):
        if h in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[u]
            tot += float(v) * mult
    return tot


traffic = {}
for wl, rep in (("c2", "prof_gemm_c2.ncu-rep"), ("c3a", "prof_gemm_c3a.ncu-rep"), ("c3a_scan", "prof_scan_c3a.ncu-rep"),
                ("c5", "prof_gemm_c5.ncu-rep")):
    p = os.path.join(G, rep)
    if os.path.exists(p):
        traffic[wl] = {"bytes_per_launch": traffic_bytes(p), "source": f"profiles/{tag}_ncu_*: dram__bytes_read.sum + dram__bytes_write.sum, one launch"}
if traffic:
    json.dump(traffic, open(os.path.join(P, f"{tag}_traffic.json"), "w"), indent=1)
    print("wrote traffic", traffic)

for rep, out in (("prof_gemm_c2.ncu-rep", f"{tag}_ncu_gemm_topk_c2.txt"), ("prof_scan_c3a.ncu-rep", f"{tag}_ncu_scan_topk_c3a.txt"),
                 ("prof_gemm_c3a.ncu-rep", f"{tag}_ncu_gemm_topk_c3a.txt"),
                 ("prof_gemm_c5.ncu-rep", f"{tag}_ncu_gemm_topk_c5.txt"),
                 ("prof_rerank_c2.ncu-rep", f"{tag}_ncu_rerank_c2.txt")):
    p = os.path.join(G, rep)
    if os.path.exists(p):
        summarise(p, os.path.join(P, out))

# in-kernel cycle / clock diagnostics (scripts/gpu_record.sh)
for name in ("clock_diag_c2.log", "clock_diag_c2_pair.log", "clock_diag_c5_4Mrows.log", "clock_diag_c5.log"):
    src = os.path.join(G, name)
    if os.path.exists(src):
        keep = [l for l in open(src) if l.startswith("{") or (("[gemm_topk]" in l or "[gemm_topk_sk]" in l) and "clk" + "" in l + "clk")]
        keep = [l for l in keep if l.startswith("{") or "[gemm_topk_sk]" in l or ("ns" in l and "cycles" in l)]
        # the seed pass prints too (a few ten thousand cycles); keep the main pass lines only
        keep = [l for l in keep if l.startswith("{") or "[gemm_topk_sk]" in l or int(l.split("cycles")[1].split()[0]) > 500000]
        if name == "clock_diag_c5.log":
            keep = keep[-11:]  # one launch: ten roles + the cycle line
        open(os.path.join(P, f"{tag}_{name}"), "w").writelines(
            ["# gemm_debug bits: 32 = print cycles/ns of CTA 0; +4 no epilogue; +7 no loads after the ring fill, no epilogue; "
             "+23 also no MMA issue (barrier handshakes only); +19 no loads, no MMA issue, epilogue ON (epilogue alone).  "
             "Results of debug runs are invalid by design.\n"] + keep)
        print("wrote", name)
