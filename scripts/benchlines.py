"""Prints the bench JSON lines of a log in one row each (workload, step, value, e2e, dominant kernel, roofline)."""
import json, sys
for l in open(sys.argv[1]):
    if not l.startswith('{'):
        continue
    j = json.loads(l)
    if j.get("impl") == "reference":
        print("reference:", round(j["value"], 2), j["unit"], j["cpu_baseline"]["sample"])
        continue
    r = j["roofline"]
    lat = j["e2e"].get("call_latency_ms") or {}
    print(f'{j["config"]["workload"][:58]:58s} step {j["ms_per_step"]:.3f} ms value {j["value"]:.1f} | e2e {j["e2e"]["value"]:.1f} '
          f'({j["e2e"]["ms_per_step"]:.3f} ms, p50 {lat.get("median", 0):.3f} p99 {lat.get("p99", 0):.3f}) | {r["kernel"]} '
          f'{r["kernel_ms"]:.3f} ms {r["bound"]} {r["achieved"]:.0f} {r["unit"]} frac {r["frac"]:.3f} | fb {j.get("fallback_queries")} '
          f'clk {j["clocks"]["sm_mhz"]} {j["clocks"]["reasons"]}')
