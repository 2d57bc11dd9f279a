#!/usr/bin/env python
"""Summarise an ncu report (and optionally a launch-list CSV) into a small markdown file for profiles/.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [gpurun_out/launches.csv] > profiles/r01_xxx.md
"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
           "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def stalls(rep, kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel, "--launch-count", "1"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return None
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    n = 0
    seen = set()
    for d in rows[2:]:
        if len(d) != len(hdr) or not d[col["# Samples"]].isdigit() or d[col["Address"]] in seen:
            continue
        seen.add(d[col["Address"]])
        n += int(d[col["# Samples"]])
        for s in st:
            tot[s] += int(d[col[s]])
    return n, tot


def main():
    rep = sys.argv[1]
    hdr, units, data = raw(rep)
    col = {h: i for i, h in enumerate(hdr)}
    print("# ncu summary of `%s`\n" % rep.split("/")[-1])
    print("Captured with `ncu --set full --clock-control none --import-source on` (cold-cache, serialised replays: compare shares, not absolutes).\n")
    print("| kernel | grid | time us | dram rd MB | dram wr MB | dram % | sm % | warps active % | regs | warp-instr |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    names = []
    for d in data:
        name = d[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("gs::", "")
        names.append(name)
        v = {m: d[col[m]] if m in col else "" for m in METRICS}
        def f(x, s=1.0):
            try:
                return "%.1f" % (float(x) * s)
            except ValueError:
                return x
        tu = units[col["gpu__time_duration.sum"]]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(tu, 1.0)
        def mb(m):
            u = units[col[m]]
            return f(v[m], {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0))
        print("| %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |" % (
            name, d[col["Grid Size"]].replace(" ", ""), f(v["gpu__time_duration.sum"], scale), mb("dram__bytes_read.sum"), mb("dram__bytes_write.sum"),
            f(v["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]), f(v["sm__throughput.avg.pct_of_peak_sustained_elapsed"]),
            f(v["sm__warps_active.avg.pct_of_peak_sustained_active"]), v["launch__registers_per_thread"].split(".")[0], v["smsp__inst_executed.sum"].split(".")[0]))
    print()
    for k in sorted(set(names)):
        r = stalls(rep, k.split("<")[0])
        if not r or not r[0]:
            continue
        n, tot = r
        top = ", ".join("%s %.0f%%" % (s.replace("stall_", ""), 100.0 * c / n) for s, c in tot.most_common(6))
        print("- `%s` warp-stall samples (first launch): %s" % (k, top))
    if len(sys.argv) > 2:
        rows = [l for l in open(sys.argv[2]) if not l.startswith("==")]
        r = list(csv.DictReader(rows))
        agg = collections.OrderedDict()
        for x in r:
            n = x["Kernel Name"].split("(")[0].replace("void ", "").replace("gs::", "")
            a = agg.setdefault(n, [0, 0.0])
            a[0] += 1
            a[1] += float(x["Metric Value"]) / 1000.0
        tot = sum(a[1] for a in agg.values())
        print("\n## launch list `%s` (`--metrics gpu__time_duration.sum --clock-control none`)\n" % sys.argv[2].split("/")[-1])
        print("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|")
        for n, (c, t) in agg.items():
            print("| %s | %d | %.1f | %.1f | %.1f%% |" % (n, c, t, t / c, 100 * t / tot))


if __name__ == "__main__":
    main()
