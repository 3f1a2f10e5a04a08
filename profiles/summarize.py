#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full --import-source on) into the text summary committed under
profiles/: per kernel the roofline-relevant metrics, the top stall reasons and the hottest CUDA
source lines.   usage: python profiles/summarize.py gpurun_out/x.ncu-rep [profiles/ncu_metrics.json [workload]] > profiles/x.summary.md"""
import csv
import io
import subprocess
import sys

sys.path.insert(0, __file__.rsplit("/", 1)[0])
import src_hot  # noqa: E402

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "CTAs/SM limit (registers)"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM limit (shared memory)"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instruction"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
]


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def traffic_json(rows, idx, units, rep, out, workload="c2"):
    """profiles/ncu_metrics.json: per workload and kernel the per-launch DRAM bytes and the issue-slot / FP64-pipe
    utilisation of the committed capture (read by bench.py -> roofline.traffic / issue_slot_frac / fp64_pipe_frac).
    The file is merged: other workloads' entries are kept."""
    import json
    import os
    import re
    res = {}
    if os.path.exists(out):
        res = json.load(open(out))
    w = res.setdefault(workload, {})
    for r in rows[2:]:
        m = re.search(r"(\w+_kernel)", r[idx["Kernel Name"]])
        if not m:
            continue
        rd = float(r[idx["dram__bytes_read.sum"]]) * UNIT[units[idx["dram__bytes_read.sum"]]]
        wr = float(r[idx["dram__bytes_write.sum"]]) * UNIT[units[idx["dram__bytes_write.sum"]]]
        w[m.group(1)] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes": rd + wr,
                         "issue_slot_frac": float(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]]) / 100,
                         "fp64_pipe_frac": float(r[idx["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]]) / 100,
                         "warp_instructions": float(r[idx["smsp__inst_executed.sum"]]),
                         "duration_ms_under_ncu": float(r[idx["gpu__time_duration.sum"]]) *
                         {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[idx["gpu__time_duration.sum"]], 1.0),
                         "source": rep.rsplit("/", 1)[-1] + " (ncu --set full --clock-control none, one launch per kernel)"}
    json.dump(res, open(out, "w"), indent=1)


def main(rep, traffic_out=None, workload="c2"):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    if traffic_out:
        traffic_json(rows, idx, units, rep, traffic_out, workload)
    print(f"# ncu summary of `{rep.rsplit('/', 1)[-1]}` (ncu --set full --clock-control none)\n")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        print(f"## {name[:90]}\n")
        print("| metric | value |\n|---|---|")
        for key, label in WANT:
            if key in idx:
                print(f"| {label} (`{key}`) | {r[idx[key]]} {units[idx[key]]} |")
        st = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                try:
                    st.append((float(r[idx[h]]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        print("\nTop warp stall reasons (warps stalled per issue-active cycle): " +
              ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:6]) + "\n")
    for pat in ("cover_kernel", "plan_gen_kernel", "path_kernel"):
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv",
                              "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
        if "Instructions Executed" not in src:
            continue
        tmp = f"/tmp/_src_{pat}.csv"
        open(tmp, "w").write(src)
        print(f"## hottest source lines of {pat} (share of warp instructions / of stall samples)\n\n```")
        src_hot.main(tmp, 25)
        print("```\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else "c2")
