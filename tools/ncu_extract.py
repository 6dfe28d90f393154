"""Turn an .ncu-rep into a small tracked JSON (profiles/*.json) that bench.py reads instead of literals.

    python tools/ncu_extract.py gpurun_out/prof.ncu-rep profiles/r2_movegen_warp.json [--kernel regex] [--calls-per-block 8]

Per captured launch: kernel name, grid, duration and the raw metrics that the rooflines quote (issue
utilisation, lanes per warp instruction, warp / thread instruction counts, pipe utilisation, DRAM bytes,
tensor-pipe activity).  The JSON also records the git commit of the tree it was extracted in.
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_active",
    "smsp__thread_inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sass__thread_inst_executed_true_per_opcode",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second", "lts__t_bytes.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_active.avg",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_elapsed.avg.per_second",
]
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0, "usecond": 1e-3, "msecond": 1.0, "second": 1e3,
              "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--kernel", default=".*")
    ap.add_argument("--units-per-block", type=float, default=0.0, help="e.g. 8 movegen calls per block: adds per-unit figures")
    ap.add_argument("--units", type=float, default=0.0, help="units (e.g. movegen calls) every captured launch worked on, when the grid does not tell")
    ap.add_argument("--note", default="")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    pat = re.compile(args.kernel)
    launches = []
    for r in data:
        name = r[col["Kernel Name"]]
        if not pat.search(name):
            continue
        m = {}
        for h, i in col.items():
            base = h.split(".TriageCompute.")[-1] if ".TriageCompute." in h else h
            if base in WANT and r[i] not in ("", "no data", "n/a"):
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                u = units[i]
                if u in UNIT_SCALE and (base.startswith("dram__bytes") or base.startswith("lts__t_bytes") or base.startswith("gpu__time")) \
                        and "per_second" not in base:
                    v *= UNIT_SCALE[u]
                    u = "byte" if "byte" in u else "ms"
                m[base] = {"value": v, "unit": u}
        rec = {"kernel": re.sub(r"\(.*", "", name), "grid": r[col["Grid Size"]], "block": r[col["Block Size"]], "metrics": m}
        if args.units or (args.units_per_block and "launch__grid_size" in m):
            n = args.units or m["launch__grid_size"]["value"] * args.units_per_block
            rec["units"] = n
            if "sm__inst_executed.sum" in m:
                rec["warp_instructions_per_unit"] = m["sm__inst_executed.sum"]["value"] / n
            if "smsp__thread_inst_executed.sum" in m:
                rec["thread_instructions_per_unit"] = m["smsp__thread_inst_executed.sum"]["value"] / n
            elif "sass__thread_inst_executed_true_per_opcode" in m:
                rec["thread_instructions_per_unit"] = m["sass__thread_inst_executed_true_per_opcode"]["value"] / n
            if "dram__bytes_read.sum" in m and "dram__bytes_write.sum" in m:
                rec["dram_bytes_per_unit"] = (m["dram__bytes_read.sum"]["value"] + m["dram__bytes_write.sum"]["value"]) / n
        launches.append(rec)
    try:
        commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True,
                                cwd=os.path.dirname(os.path.abspath(__file__))).stdout.strip()
    except Exception:
        commit = ""
    out = {"report": os.path.basename(args.rep), "extracted_at_commit": commit, "note": args.note, "launches": launches}
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(f"{len(launches)} launches -> {args.out}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
