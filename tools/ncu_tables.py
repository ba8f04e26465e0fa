#!/usr/bin/env python
"""Turn ncu CSV exports into the markdown tables of profiles/rNN_summary.md.

    python tools/ncu_tables.py launches profiles/r02_launches.csv          # share of the step per kernel
    python tools/ncu_tables.py raw profiles/r02_c2_raw.csv [more.csv ...]  # per-kernel --set full metrics

The CSVs come from `ncu --csv --log-file ...` (launch list) and `ncu -i X.ncu-rep --page raw --csv` (raw page)."""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    m = re.match(r"(?:void )?([A-Za-z0-9_]+)(<[^>]*>)?", name)
    t = ""
    if m and m.group(2):
        k = re.search(r"(\d+)", m.group(2))
        t = "<%s>" % k.group(1) if k else ""
    return (m.group(1) if m else name) + t


def launches(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[h]
    iK, iM, iV, iID = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("ID")
    per = OrderedDict()
    seen = set()
    for r in rows[h + 1:]:
        if len(r) <= iV or r[iM] != "gpu__time_duration.sum":
            continue
        if r[iID] in seen:
            continue
        seen.add(r[iID])
        k = short(r[iK])
        per.setdefault(k, []).append(float(r[iV].replace(",", "")))
    tot = sum(sum(v) for v in per.values())
    print("| kernel | launches | us / launch (mean) | share of the captured time |")
    print("|---|---|---|---|")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        print("| `%s` | %d | %.1f | %.1f %% |" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    print("\n%d launches, %.1f ms in total (serialised, cold cache under ncu)" % (len(seen), tot / 1e6))


WANT = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/inst"),
        ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %")]


def raw(paths):
    print("| kernel | " + " | ".join(w[1] for w in WANT) + " |")
    print("|---|" + "---|" * len(WANT))
    for path in paths:
        rows = list(csv.reader(open(path)))
        H, U = rows[0], rows[1]
        for r in rows[2:]:
            if len(r) < len(H):
                continue
            cells = []
            for key, _ in WANT:
                if key in H:
                    i = H.index(key)
                    v = r[i]
                    try:
                        f = float(v.replace(",", ""))
                        v = ("%.3g" % f) if abs(f) < 1e6 else ("%.3e" % f)
                    except ValueError:
                        pass
                    cells.append("%s %s" % (v, U[i]) if U[i] not in ("", "%") else v)
                else:
                    cells.append("-")
            print("| `%s` | %s |" % (short(r[H.index("Kernel Name")]), " | ".join(cells)))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.argv[2:])
