#!/usr/bin/env python
"""Turn the ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/rNN_launch_shares.md [steps]
  python tools/summarize_ncu.py full     gpurun_out/prof.ncu-rep  profiles/rNN_edge_mlp_metrics.md
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
]


def launches(src, dst, steps=None):
    """A pass of the hot call starts at `prep_kernel` and ends at `fsq_quantize_kernel`: only launches inside whole
    passes are counted (model set-up before the first pass and a pass cut off by `-c` are dropped), and the number
    of passes is counted from the list itself unless `steps` is given."""
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
        rows.append((name, v))
    starts = [i for i, (n, _) in enumerate(rows) if n.startswith("prep_kernel")]
    ends = [i for i, (n, _) in enumerate(rows) if n.startswith("fsq_quantize_kernel")]
    dropped = 0
    if starts and ends and ends[-1] > starts[0]:
        dropped = starts[0] + len(rows) - 1 - ends[-1]
        rows = rows[starts[0]:ends[-1] + 1]
        passes = sum(1 for n, _ in rows if n.startswith("fsq_quantize_kernel"))
    else:
        passes = steps or 1
    steps = steps or passes
    agg = collections.OrderedDict()
    tot = 0.0
    for name, v in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    with open(dst, "w") as fh:
        fh.write(f"# ncu launch list ({src}): `--metrics gpu__time_duration.sum --clock-control none`; {steps} whole passes of the "
                 f"hot call ({dropped} launches outside them dropped)\n\n")
        fh.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's live event timing.\n\n")
        fh.write("| kernel | share of step | ms / step | launches / step | avg us |\n|---|---|---|---|---|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            fh.write(f"| `{k}` | {t / tot * 100:.2f} % | {t / steps / 1e3:.3f} | {c / steps:.1f} | {t / c:.1f} |\n")
        fh.write(f"\ntotal {tot / steps / 1e3:.3f} ms / step over {sum(c for c, _ in agg.values())} launches\n")


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as fh:
        fh.write(f"# ncu --set full --clock-control none ({src})\n\n")
        for r in rows[2:]:
            fh.write(f"## {r[idx['Kernel Name']]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in idx:
                    fh.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
            st = [(float(r[i].replace(',', '')), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for i, h in enumerate(hdr)
                  if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued")]
            tot = sum(v for v, _ in st) or 1.0
            fh.write("\nwarp stall samples: " + ", ".join(f"{n} {v / tot * 100:.0f}%" for v, n in sorted(st, reverse=True)[:8]) + "\n\n")


def traffic(src, dst, residues):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the two edge-MLP kernels -> the JSON bench.py reads
    (`roofline.traffic`); `residues` = the batch the capture ran on."""
    import json

    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        if "edge_msg_t_kernel" in name:
            key = "msg"
        elif "edge_mlp_tc_kernel" in name:
            key = "msg" if ", 0>" in name or "(int)0>" in name else "upd"
        else:
            continue
        tot = sum(float(r[idx[k]].replace(",", "")) * scale[units[idx[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        out.setdefault(key, []).append(tot)
    doc = {k: sum(v) / len(v) for k, v in out.items()}
    doc["residues"] = int(residues)
    doc["source"] = f"ncu --set full --clock-control none, {src} (dram__bytes_read.sum + dram__bytes_write.sum per launch, scaled to this batch)"
    with open(dst, "w") as fh:
        json.dump(doc, fh, indent=1)
    print(doc)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else None)
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 131072)
    else:
        full(sys.argv[2], sys.argv[3])
