#!/usr/bin/env python
"""Prints a table of every bench.py JSON line in a log (headline, sustained, secondary entries)."""
import json
import sys


def row(r, tag=""):
    if "error" in r or r.get("name") == "ingest":
        print(tag, json.dumps(r)[:600])
        return
    rf = r.get("roofline") or {}
    e = r.get("e2e") or {}
    c = r.get("cpu_baseline") or {}
    x = r.get("exchange") or {}
    print(f"{tag}{r.get('name', 'HEAD'):9s} N={r.get('n_gpus')} ms/step {r['ms_per_step']:8.3f} value {r['value']:11.1f} full {r.get('qps_full_index', 0):10.1f} | "
          f"{rf.get('kernel', '')[:10]} {rf.get('bound')} {rf.get('achieved', 0):8.1f} {rf.get('unit')} frac {rf.get('frac', 0):.3f} "
          f"kern_ms {rf.get('kernel_ms', 0):.3f} | e2e ms {e.get('ms_per_step', 0):.3f} (serial {e.get('serial_ms_per_step', 0):.3f}) | merge_ms {x.get('merge_ms', 0):.4f} | "
          f"cpu {c.get('value', 0):.2f} ({c.get('cores')}) fb {r.get('fallback_queries')} launches {r.get('gpu_launches')} host_enq_ms {r.get('host_enqueue_ms_per_step', 0):.3f}"
          + (f" shard_kern_ms {x.get('kernel_ms_per_shard')}" if x.get('kernel_ms_per_shard') else ""))


for path in sys.argv[1:]:
    for l in open(path):
        if not l.startswith("{"):
            if l.startswith("==="):
                print(l.rstrip()[:200])
            continue
        j = json.loads(l)
        row(j)
        if j.get("sustained"):
            s = j["sustained"]
            print(f"   sustained: ms/step {s['ms_per_step']:.3f} kern_ms {s['kernel_ms']:.3f} {s['achieved']:.1f} {s['unit']} "
                  f"frac_burst {s['frac_of_burst_peak']:.3f} sm_mhz {s['sm_mhz']} {s['reasons']}")
        if j.get("clocks"):
            print("   clocks:", j["clocks"])
        for r in j.get("secondary", []):
            row(r, "   ")
