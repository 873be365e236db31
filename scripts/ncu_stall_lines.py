"""Per-source-line stall samples and instruction shares of one kernel from an
`ncu --set full --import-source on` report (the view that showed where k_dst3 waits):

  python scripts/ncu_stall_lines.py gpurun_out/r01g_full.ncu-rep k_dst3 [top_n]

Prints, for the first matching launch, the source lines ranked by warp-stall samples with
their share of executed instructions."""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
funcs, cur = [], None
for r in csv.reader(raw.splitlines()):
    if r and r[0] == "Function Name":
        cur = {"name": r[1], "lines": []}
        funcs.append(cur)
    elif r and r[0] == "Line No" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and "hdr" in cur and r and r[0] != "File Path":
        cur["lines"].append(r)
if not funcs:
    sys.exit("no launch of %s in %s" % (kern, rep))
f = funcs[0]
h = f["hdr"]
i_s, i_i = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
rows = [(int(r[i_s]), int(r[i_i]), r[0], r[1]) for r in f["lines"] if r[0] and r[i_s].isdigit() and r[i_i].isdigit()]
ts, ti = sum(x[0] for x in rows) or 1, sum(x[1] for x in rows) or 1
print("%s: %d stall samples, %d warp instructions" % (f["name"], ts, ti))
for s, i, line, src in sorted(rows, reverse=True)[:top]:
    print("%5.1f%% stalls %5.1f%% instr  L%-5s %s" % (100.0 * s / ts, 100.0 * i / ti, line, src[:110]))
