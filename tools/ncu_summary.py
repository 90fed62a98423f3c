#!/usr/bin/env python
"""Summaries of an `ncu --set full --import-source on` report for profiles/:
  python tools/ncu_summary.py REPORT.ncu-rep OUT_PREFIX
writes OUT_PREFIX.csv (selected raw metrics, metric,unit,value) and OUT_PREFIX_hot_lines.txt (instruction and
stall-sample shares per source line, from `--page source --print-source cuda,sass`... aggregated by file:line)."""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

KEEP = re.compile(r"^(Kernel Name|Block Size|Grid Size|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum$|"
                  r"dram__throughput\.avg\.pct_of_peak_sustained_elapsed|launch__registers_per_thread$|"
                  r"launch__occupancy_limit_(registers|shared_mem|warps)|launch__shared_mem_per_block_dynamic|"
                  r"sm__inst_executed\.avg\.per_cycle_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__inst_executed\.sum$|smsp__thread_inst_executed_per_inst_executed\.ratio|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|"
                  r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|lts__t_sector_hit_rate\.pct|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|smsp__cycles_active\.avg$)")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units, vals = rows[0], rows[1], rows[2]
    with open(out + ".csv", "w") as f:
        f.write("metric,unit,value\n")
        for n, u, v in zip(names, units, vals):
            if KEEP.match(n):
                f.write(f'{n},{u},"{v}"\n' if "," in v else f"{n},{u},{v}\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    if not src.strip():
        return
    # the page is a sequence of per-file sections: "File Path",<path> / "Function Name",.. / header / rows; a row with a
    # line number is a CUDA source line (its counters already aggregate the SASS rows that follow it)
    inst, samp, text, by_file = defaultdict(float), defaultdict(float), {}, defaultdict(lambda: [0.0, 0.0])
    path, hdr = "", None
    for r in csv.reader(io.StringIO(src)):
        if len(r) == 2 and r[0] == "File Path":
            path, hdr = r[1].rsplit("/", 1)[-1], None
            continue
        if len(r) >= 2 and r[0] == "Line No":
            hdr = {c: i for i, c in enumerate(r) if c not in ("Source",)}
            continue
        if hdr is None or len(r) < 8 or not r[0].strip().isdigit():
            continue
        try:
            i_, p_ = float(r[hdr["Instructions Executed"]] or 0), float(r[hdr["Warp Stall Sampling (All Samples)"]] or 0)
        except ValueError:
            continue
        key = (path, int(r[0]))
        inst[key] += i_
        samp[key] += p_
        text.setdefault(key, r[1].strip())
        by_file[path][0] += i_
        by_file[path][1] += p_
    ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
    with open(out + "_hot_lines.txt", "w") as f:
        f.write(f"total inst {ti:.0f} samples {ts:.0f}\n")
        for fn, (a, b) in sorted(by_file.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{fn} {100 * a / ti:.1f}% inst {100 * b / ts:.1f}% samp\n")
        for k in sorted(inst, key=lambda k: -samp[k])[:60]:
            f.write(f"{k[0]}:{k[1]:4d} inst {100 * inst[k] / ti:5.1f}%  samp {100 * samp[k] / ts:5.1f}%  {text[k][:140]}\n")


if __name__ == "__main__":
    main()
