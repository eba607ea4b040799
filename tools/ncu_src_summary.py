"""Summarises the source page of an .ncu-rep: per-stall-reason clocks over the instructions that executed `count`
times (e.g. the unrolled loop body), and the instructions with most samples.
usage: python tools/ncu_src_summary.py report.ncu-rep <instructions-executed value> [layers per body]"""
import csv, subprocess, sys, io
rep, count = sys.argv[1], int(sys.argv[2])
per = int(sys.argv[3]) if len(sys.argv) > 3 else 1
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg, idx = {}, []
for i, r in enumerate(data):
    if int(r[ix["Instructions Executed"]]) == count:
        idx.append(i)
        for s in stalls:
            agg[s] = agg.get(s, 0) + int(r[ix[s]])
sel = agg["stall_selected"]
print(len(idx), "instructions, rows", idx[0], "-", idx[-1])
for k, v in sorted(agg.items(), key=lambda x: -x[1]):
    if v:
        print("%-24s %6d  ~%4.0f clk per layer" % (k, v, v / sel * len(idx) / per))
print("total ~%.0f clk per layer" % (sum(agg.values()) / sel * len(idx) / per))
top = sorted(idx, key=lambda i: -int(data[i][ix["# Samples"]]))[:int(sys.argv[4]) if len(sys.argv) > 4 else 30]
for i in sorted(top):
    r = data[i]
    st = sorted(((s.replace("stall_", ""), int(r[ix[s]])) for s in stalls if int(r[ix[s]]) > 0), key=lambda x: -x[1])[:3]
    print(i, r[ix["Source"]].strip()[:64].ljust(64), r[ix["# Samples"]].rjust(5), st)
