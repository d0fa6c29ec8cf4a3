#!/usr/bin/env python
"""Per-instruction L1 traffic of one kernel from an `ncu --set full --import-source on` report.

  python tools/ncu_l1_requests.py gpurun_out/prof.ncu-rep knn_warp_kernel 131072

Prints, for every memory instruction of the first matching launch, how often it ran per unit (the third argument:
rows, tiles, ... the launch processed), its L1 tag requests (global) + shared-memory wavefronts per unit and per
execution.  A warp-wide load that needs many more requests than 128-byte lines it has to touch is uncoalesced: this
listing is how the k-NN scan (16 requests per load for 2 KB per warp) and the nine-load feature loop were found
(profiles/r01_kernels_ncu_full.md).
"""
import csv
import subprocess
import sys

MEM = ("LDG", "STG", "LDS", "STS", "LDSM", "ATOM", "RED", "UTMALDG", "UTMASTG", "LDTM", "STTM", "SHFL")


def main():
    rep, kernel, units = sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}", "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(rows[0][1] if len(rows[0]) > 1 else kernel)
    total_req = total_inst = 0
    for k, r in enumerate(rows[2:]):
        if len(r) < len(hdr):
            break  # next launch
        n = int(r[col["Instructions Executed"]] or 0)
        total_inst += n
        src = r[col["Source"]].strip()
        if n == 0 or not any(m in src for m in MEM):
            continue
        req = int(r[col["L1 Tag Requests Global"]] or 0) + int(r[col["L1 Wavefronts Shared"]] or 0)
        total_req += req
        print(f"{k:5d} {src[:60]:60s} exec/unit {n / units:8.2f}  req/unit {req / units:8.2f}  req/exec {req / n:6.2f}")
    print(f"total: {total_req / units:.1f} requests + wavefronts per unit, {total_inst / units:.1f} warp instructions per unit")


if __name__ == "__main__":
    main()
