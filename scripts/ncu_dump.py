"""Dump the SASS page of an .ncu-rep in address order: samples, executions, top stalls.  usage: ncu_dump.py rep [lo hi]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 62
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
f = lambda x: float(x) if x.replace('.', '', 1).isdigit() else 0.0
sc = [h for h in hdr if h.startswith('stall_') and 'Not' not in h]
base = int(data[0][ix['Address']], 16)
for r in data:
    a = int(r[ix['Address']], 16) - base
    if a < lo or a > hi: continue
    st = sorted(((h[6:], int(f(r[ix[h]]))) for h in sc), key=lambda t: -t[1])[:2]
    st = " ".join(f"{h}={v}" for h, v in st if v)
    print(f"{a:05x} {int(f(r[ix['# Samples']])):6d} {int(f(r[ix['Instructions Executed']])):9d}  {r[ix['Source']][:72]:72s} {st}")
