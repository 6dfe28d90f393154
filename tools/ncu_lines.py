"""Per-source-line view of one kernel of an .ncu-rep captured with --import-source on / -lineinfo.

    python tools/ncu_lines.py gpurun_out/x.ncu-rep movegen_solo_kernel [--top 40] [--cubin movegen_warp]

`ncu --page source --csv` lists SASS instructions with their counters in program order; `nvdisasm -g`
of the same cubin (extracted from the in-tree .so, so run this on the tree the report was taken from)
lists the same instructions with their source lines.  Joined by position, this prints for every source
line: static SASS instructions, warp instructions executed, average active threads, and the stall
samples (total / no-instruction), i.e. where the instruction count and the instruction-cache misses
of a kernel come from.
"""
import argparse
import collections
import csv
import glob
import io
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(so, cubin_hint, kernel, want_len=-1):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
    out = []
    for cub in sorted(glob.glob(os.path.join(tmp, "*.cubin"))):
        base = os.path.basename(cub)
        if cubin_hint and not base.startswith(cubin_hint + "."):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
        cur_fn, line = None, None
        fns = collections.OrderedDict()
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                cur_fn = m.group(1)
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if cur_fn and kernel in cur_fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
                fns.setdefault(cur_fn, []).append((line, ln.split("*/", 1)[1].strip().rstrip(";")))
        if fns:
            # template instantiations: the one whose length matches the report, else the first
            out = next((v for v in fns.values() if len(v) == want_len), next(iter(fns.values())))
            break
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--cubin", default="")
    ap.add_argument("--so", default=os.path.join(ROOT, "tetris_reinforcement_learning_b200", "libtrl_b200.so"))
    ap.add_argument("--ranges", default="", help="comma separated a-b source line ranges to total, e.g. 123-153,154-175")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.rep, "--page", "source", "--csv", "-k", "regex:" + args.kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # the page lists every captured launch: "Kernel Name" row, header row, one row per SASS instruction
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    pick = next((i for i in starts if re.search(args.kernel, rows[i][1])), starts[0] if starts else -1)
    end = next((i for i in starts if i > pick), len(rows))
    hdr_i = next(i for i in range(pick, end) if rows[i] and rows[i][0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    inst = [r for r in rows[hdr_i + 1:end] if len(r) == len(hdr)]
    sass = sass_lines(args.so, args.cubin, args.kernel, len(inst))
    if len(sass) != len(inst):
        print(f"warning: {len(sass)} SASS instructions in the .so vs {len(inst)} in the report (different build?)")
    n = min(len(sass), len(inst))
    agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])   # static, executed, thread-executed, samples, no_inst samples
    tot = [0, 0, 0, 0, 0]
    for k in range(n):
        line = sass[k][0]
        r = inst[k]
        v = [1, int(r[col["Instructions Executed"]]), int(r[col["Thread Instructions Executed"]]),
             int(r[col["# Samples"]]), int(r[col["stall_no_inst"]]) if "stall_no_inst" in col else 0]
        for j in range(5):
            agg[line][j] += v[j]
            tot[j] += v[j]
    print(f"kernel {args.kernel}: {tot[0]} SASS instructions ({tot[0] * 16 / 1024:.1f} KB), {tot[1]:.3e} warp instructions, "
          f"{tot[2] / max(tot[1], 1):.1f} threads/instruction, {tot[3]} samples, {100.0 * tot[4] / max(tot[3], 1):.1f} % no-instruction")
    print(f"{'line':>22} {'static':>6} {'exec %':>7} {'thr':>5} {'samp %':>7} {'noinst %':>8}")
    for line, v in sorted(agg.items(), key=lambda kv: -kv[1][3])[:args.top]:
        name = f"{line[0]}:{line[1]}" if line else "?"
        print(f"{name:>22} {v[0]:6d} {100.0 * v[1] / tot[1]:7.2f} {v[2] / max(v[1], 1):5.1f} {100.0 * v[3] / tot[3]:7.2f} {100.0 * v[4] / max(v[3], 1):8.1f}")
    if args.ranges:
        print("ranges (movegen_warp.cu lines):")
        for rg in args.ranges.split(","):
            a, b = (int(x) for x in rg.split("-"))
            s = [0, 0, 0, 0, 0]
            for line, v in agg.items():
                if line and a <= line[1] <= b and line[0].startswith(args.cubin or ""):
                    for j in range(5):
                        s[j] += v[j]
            print(f"  {rg:>9}: static {s[0]:5d}  executed {100.0 * s[1] / tot[1]:6.2f} %  threads {s[2] / max(s[1], 1):5.1f}  "
                  f"samples {100.0 * s[3] / tot[3]:6.2f} %  no-inst {100.0 * s[4] / max(s[3], 1):5.1f} %")


if __name__ == "__main__":
    main()
