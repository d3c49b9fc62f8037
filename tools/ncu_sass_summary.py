#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass` for one kernel launch:
instructions executed by opcode, the hottest SASS lines by stall samples, and stall-reason totals."""
import csv
import subprocess
import sys
from collections import Counter


def main():
    rep = sys.argv[1]
    kernel_idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass",
                          "--launch-skip", str(kernel_idx), "--launch-count", "1"],
                         capture_output=True, text=True).stdout.splitlines()
    rows = list(csv.reader(out))
    name = rows[0][1]
    hdr = rows[1]
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    ix = {h: i for i, h in enumerate(hdr)}
    def f(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0
    tot_inst = sum(f(r, "Instructions Executed") for r in body)
    tot_samp = sum(f(r, "# Samples") for r in body)
    print(f"kernel: {name}\nSASS lines {len(body)}, warp instructions executed {tot_inst:.0f}, samples {tot_samp:.0f}")
    ops = Counter()
    for r in body:
        op = r[ix["Source"]].strip().split()
        op = [o for o in op if not o.startswith("@")]
        ops[op[0].split(".")[0] if op else "?"] += f(r, "Instructions Executed")
    print("\nby opcode (warp instr, share):")
    for op, n in ops.most_common(18):
        print(f"  {op:10s} {n:14.0f} {100 * n / tot_inst:6.2f}%")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print("\nstall reasons (all samples):")
    st = {s: sum(f(r, s) for r in body) for s in stalls}
    for s, n in sorted(st.items(), key=lambda x: -x[1])[:8]:
        print(f"  {s:28s} {n:10.0f} {100 * n / max(tot_samp, 1):6.2f}%")
    print(f"\ntop {top} SASS lines by samples:")
    for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:top]:
        reasons = sorted(((f(r, s), s) for s in stalls), reverse=True)[:2]
        rs = " ".join(f"{s[6:]}={n:.0f}" for n, s in reasons if n > 0)
        print(f"  {f(r, '# Samples'):8.0f} {100 * f(r, '# Samples') / max(tot_samp, 1):5.2f}%  exec {f(r, 'Instructions Executed'):11.0f}  "
              f"{r[ix['Source']].strip()[:70]:70s} {rs}")


if __name__ == "__main__":
    main()
