#!/usr/bin/env python
"""Merge `ncu -i X.ncu-rep --page raw --csv` tables into profiles/ncu_traffic.json (DRAM bytes per launch, duration,
DRAM / issue utilisation, registers per kernel) and keep a trimmed copy of the table under profiles/.

    python tools/ncu_to_traffic.py gpurun_out/r2_c2_step_raw.csv "C2 ..."  [more csv/label pairs]
"""
from __future__ import annotations

import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEEP = ["Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def num(v):
    try:
        return float(str(v).replace(",", ""))
    except ValueError:
        return None


def main():
    out_path = ROOT / "profiles" / "ncu_traffic.json"
    data = json.loads(out_path.read_text()) if out_path.exists() else {"kernels": {}}
    sources = [data.get("source", "")]
    args = sys.argv[1:]
    for path, label in zip(args[0::2], args[1::2]):
        rows = [r for r in csv.reader(open(path)) if r]
        hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        hdr, units = rows[hi], rows[hi + 1]
        col = {h: i for i, h in enumerate(hdr)}
        trimmed = [[h for h in KEEP if h in col], [units[col[h]] for h in KEEP if h in col]]
        for r in rows[hi + 2:]:
            if len(r) < len(hdr):
                continue
            name = r[col["Kernel Name"]].replace("void ", "").split("(")[0].replace("lifu::", "")
            def get(h, table=SCALE):
                if h not in col:
                    return None
                v = num(r[col[h]])
                return None if v is None else v * table.get(units[col[h]], 1.0)
            rec = {"dram_read_MB": get("dram__bytes_read.sum"), "dram_write_MB": get("dram__bytes_write.sum"),
                   "ncu_us": get("gpu__time_duration.sum"),
                   "dram_pct": get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", {}),
                   "issue_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active", {}),
                   "registers": get("launch__registers_per_thread", {}), "capture": label}
            key = name if name not in data["kernels"] or data["kernels"][name].get("capture") == label else f"{name} [{label}]"
            data["kernels"][key] = rec
            trimmed.append([r[col[h]] for h in KEEP if h in col])
        dst = ROOT / "profiles" / (Path(path).stem + ".csv")
        with open(dst, "w", newline="") as f:
            csv.writer(f).writerows(trimmed)
        sources.append(f"{dst.relative_to(ROOT)} ({label})")
    data["source"] = "; ".join(s for s in sources if s)
    out_path.write_text(json.dumps(data, indent=1))
    print("kernels:", len(data["kernels"]))


if __name__ == "__main__":
    main()
