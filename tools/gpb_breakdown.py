import csv, collections, sys
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]; ik = hdr.index("Kernel Name"); im = hdr.index("Metric Name"); iv = hdr.index("Metric Value"); iid = hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        per.setdefault((r[iid], r[ik].split("(")[0][:40]), {})[r[im]] = float(r[iv].replace(",", ""))
    items = list(per.items())
    half = len(items) // 2
    print(f)
    tot = 0
    for (i, k), m in items[half:]:
        t = m["gpu__time_duration.sum"] / 1e3
        tot += t
        print(f"  {k:36s} {t:8.1f} us  rd {m['dram__bytes_read.sum']/1e6:8.1f} MB wr {m['dram__bytes_write.sum']/1e6:8.1f} MB  fp64 {m['sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active']:5.1f}%  issue {m['smsp__issue_active.avg.pct_of_peak_sustained_active']:5.1f}%")
    print("  total", round(tot, 1), "us")
