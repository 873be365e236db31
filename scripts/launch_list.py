"""Condense an `ncu --metrics gpu__time_duration.sum --csv` log into one row per kernel
(python scripts/launch_list.py gpurun_out/x_launches_raw.csv profiles/rNN_launches.csv "note")."""
import csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
acc, order = {}, []
for r in rows[1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"^void ", "", r[kn])
    name = re.sub(r"\(.*$", "", name)
    t = float(r[mv].replace(",", ""))
    t *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(r[mu], 1.0)
    if name not in acc:
        acc[name] = [0, 0.0]
        order.append(name)
    acc[name][0] += 1
    acc[name][1] += t
tot = sum(v[1] for v in acc.values())
with open(sys.argv[2], "w") as f:
    if len(sys.argv) > 3:
        f.write("# %s\n" % sys.argv[3])
    f.write("kernel,launches,total_us,mean_us,share\n")
    for n in sorted(order, key=lambda k: -acc[k][1]):
        c, t = acc[n]
        f.write('"%s",%d,%.1f,%.2f,%.4f\n' % (n, c, t, t / c, t / tot))
