"""Per-kernel table (mean over launches) from an `ncu --page raw --csv` export: duration, DRAM bytes, achieved DRAM GB/s and
its fraction of the measured HBM copy peak, SM / tensor-pipe utilisation, occupancy -- the per-kernel evidence north_star
asks for on the motion / modulation / alignment kernels.  usage: ncu_table.py raw.csv [hbm_peak_gbs]"""
import collections, csv, re, sys


def num(row, col, name, units, scale=None):
    if name not in col or row[col[name]] in ("", "n/a"):
        return None
    v = float(row[col[name]].replace(",", ""))
    u = units[col[name]]
    if scale == "bytes":
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    if scale == "rate":
        v *= {"byte/s": 1.0, "Kbyte/s": 1e3, "Mbyte/s": 1e6, "Gbyte/s": 1e9, "Tbyte/s": 1e12}.get(u, 1.0)
    if scale == "us":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u.replace("second", "s").replace("usecond", "us").replace("nsecond", "ns").replace("msecond", "ms"), 1.0)
    return v


def main(path, peak=6550.7):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("spm::", "").replace("<unnamed>::", "")
        a = agg.setdefault(name, collections.defaultdict(float))
        a["n"] += 1
        for key, metric, scale in (("us", "gpu__time_duration.sum", "us"), ("rd", "dram__bytes_read.sum", "bytes"),
                                   ("wr", "dram__bytes_write.sum", "bytes"),
                                   ("rate", "dram__bytes.sum.per_second", "rate"),
                                   ("sm", "sm__throughput.avg.pct_of_peak_sustained_elapsed", None),
                                   ("tensor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", None),
                                   ("warps", "sm__warps_active.avg.pct_of_peak_sustained_active", None),
                                   ("l2", "lts__throughput.avg.pct_of_peak_sustained_elapsed", None),
                                   ("grid", "launch__grid_size", None), ("regs", "launch__registers_per_thread", None)):
            v = num(r, col, metric, units, scale)
            if v is not None:
                a[key] += v
        for h in hdr:   # warp-state stall reasons (WarpStateStats): keep the sums, report the two largest
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[col[h]] not in ("", "n/a"):
                a["stall:" + h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] += float(r[col[h]].replace(",", ""))
    print("%-46s %4s %9s %10s %9s %7s %6s %7s %7s %6s %8s %5s" % ("kernel", "n", "us", "dram MB", "GB/s", "of HBM", "SM %", "tens %", "warps %", "L2 %", "grid", "regs") + "  top stalls (warps per issue)")
    for name, a in agg.items():
        n = a["n"]
        us, byt = a["us"] / n, (a["rd"] + a["wr"]) / n
        if byt == 0 and a["rate"] > 0:   # light section sets export the DRAM rate, not the byte counts
            byt = a["rate"] / n * us * 1e-6
        gbs = byt / us / 1e3 if us > 0 else 0.0
        stalls = sorted(((v / n, k[6:]) for k, v in a.items() if k.startswith("stall:") and k[6:] not in ("selected", "not_selected")), reverse=True)[:2]
        print("%-46s %4d %9.1f %10.2f %9.0f %7.2f %6.1f %7.1f %7.1f %6.1f %8d %5d  %s" % (
            name[:46], n, us, byt / 1e6, gbs, gbs / peak, a["sm"] / n, a["tensor"] / n, a["warps"] / n, a["l2"] / n,
            int(a["grid"] / n), int(a["regs"] / n), ", ".join("%s %.1f" % (k, v) for v, k in stalls)))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 6550.7)
