"""Per-kernel summary of an ncu --csv launch list (metrics gpu__time_duration.sum [+ dram__bytes_read.sum, dram__bytes_write.sum]).
    python tools/ncu_summary.py launches.csv [--from-last KERNEL_SUBSTRING [--grid X]] [--meta key=value ...] > summary.json
--from-last: only the launches from the last launch whose name contains the substring (and whose grid x equals --grid) on -- e.g. the
last verify prefill of a run that first warms up."""
import collections
import csv
import json
import sys


def main():
    args = sys.argv[1:]
    path = args[0]
    start_sub, grid_x, meta = None, None, {}
    i = 1
    while i < len(args):
        if args[i] == "--from-last": start_sub = args[i + 1]; i += 2
        elif args[i] == "--grid": grid_x = args[i + 1]; i += 2
        elif args[i] == "--meta":
            i += 1
            while i < len(args) and "=" in args[i] and not args[i].startswith("--"):
                k, v = args[i].split("=", 1); meta[k] = int(v) if v.isdigit() else v; i += 1
        else: i += 1
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    col = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Grid Size", "Metric Name", "Metric Value")}
    launches = collections.OrderedDict()
    for r in data:
        if len(r) <= col["Metric Value"]: continue
        L = launches.setdefault(r[col["ID"]], {"name": r[col["Kernel Name"]], "grid": r[col["Grid Size"]]})
        L[r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", ""))
    seq = list(launches.values())
    if start_sub:
        idx = [i for i, L in enumerate(seq) if start_sub in L["name"] and (grid_x is None or L["grid"].strip("()").split(",")[0].strip() == grid_x)]
        seq = seq[idx[-1]:] if idx else seq
    per = collections.OrderedDict()
    tot = {"ns": 0.0, "rd": 0.0, "wr": 0.0}
    for L in seq:
        name = L["name"].split("(")[0].strip()
        k = per.setdefault(name, {"launches": 0, "ms": 0.0, "read_gb": 0.0, "write_gb": 0.0})
        k["launches"] += 1
        k["ms"] += L.get("gpu__time_duration.sum", 0.0) / 1e6
        k["read_gb"] += L.get("dram__bytes_read.sum", 0.0) / 1e9
        k["write_gb"] += L.get("dram__bytes_write.sum", 0.0) / 1e9
        tot["ns"] += L.get("gpu__time_duration.sum", 0.0); tot["rd"] += L.get("dram__bytes_read.sum", 0.0); tot["wr"] += L.get("dram__bytes_write.sum", 0.0)
    out = dict(meta)
    out.update({"kernels": len(seq), "gpu_time_ms_serialised": tot["ns"] / 1e6, "dram__bytes_read": tot["rd"], "dram__bytes_write": tot["wr"],
                "per_kernel": {k: {kk: (round(vv, 6) if isinstance(vv, float) else vv) for kk, vv in v.items()} for k, v in per.items()}})
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
