"""Summarise an `ncu --page raw --csv` dump: one block per profiled launch with the metrics the roofline needs."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct']


def _to_bytes(val, unit):
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(val.replace(",", "")) * mult.get(unit, 1.0)


def _to_us(val, unit):
    mult = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    return float(val.replace(",", "")) * mult.get(unit.replace("second", "s").replace("usecond", "us"), 1.0)


def write_json(path, out, source):
    """--json: per kernel name (template arguments kept, parameter list cut) the mean duration and DRAM bytes per
    launch -- the file bench.py reads `roofline.traffic` from."""
    import collections, json, re
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[2:]:
        full = re.sub(r"\(.*", "", r[col["Kernel Name"]])
        base = re.sub(r"<.*", "", full).replace("void ", "").strip()
        dur = _to_us(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]])
        byt = _to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + \
            _to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        for key in {full.replace("void ", "").strip(), base}:
            a = agg.setdefault(key, [0, 0.0, 0.0])
            a[0] += 1; a[1] += dur; a[2] += byt
    d = {"source": source, "kernels": {k: {"launches": v[0], "duration_us": v[1] / v[0], "dram_bytes_per_launch": v[2] / v[0],
                                           "dram_gbs": v[2] / v[1] / 1e3} for k, v in agg.items()}}
    json.dump(d, open(out, "w"), indent=1)
    print("wrote", out, "(%d kernel names)" % len(agg))


def main(path, extra=()):
    if len(extra) >= 2 and extra[0] == "--json":
        return write_json(path, extra[1], extra[2] if len(extra) > 2 else "ncu --set full")
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('---', r[col['Kernel Name']][:90], 'grid', r[col.get('launch__grid_size', 0)])
        for w in list(WANT) + [h for h in hdr if any(e in h for e in extra)]:
            if w in col and r[col[w]] not in ('', 'n/a'):
                print('   %-78s %s %s' % (w, r[col[w]], units[col[w]]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2:])
