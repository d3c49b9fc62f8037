#!/usr/bin/env python
"""Extract the DRAM traffic (and a few headline counters) of one kernel from an ncu --set full capture into the small JSON
that bench.py reads for `roofline.traffic`:   python tools/ncu_traffic.py X.ncu-rep KERNEL_REGEX OUT.json"""
import csv
import json
import re
import subprocess
import sys

KEYS = {"dram_bytes_read": "dram__bytes_read.sum", "dram_bytes_write": "dram__bytes_write.sum",
        "duration_us": "gpu__time_duration.sum", "lts_hit_rate_pct": "lts__t_sector_hit_rate.pct",
        "lts_sectors_tex_read": "lts__t_sectors_srcunit_tex_op_read.sum",
        "dram_throughput_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "registers": "launch__registers_per_thread", "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def main():
    rep, pat, out = sys.argv[1:4]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    sel = [r for r in rows[2:] if re.search(pat, r[ki])]
    if not sel:
        raise SystemExit(f"no kernel matching {pat}")
    r = sel[-1]
    d = {"kernel": r[ki], "capture": rep.split("/")[-1], "launches_in_capture": len(sel)}
    for k, m in KEYS.items():
        if m in hdr:
            i = hdr.index(m)
            v = float(r[i].replace(",", ""))
            d[k] = v * UNIT.get(units[i], 1.0)
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps(d))


if __name__ == "__main__":
    main()
