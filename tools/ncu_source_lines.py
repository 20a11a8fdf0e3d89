#!/usr/bin/env python
"""Top source lines of a kernel in an .ncu-rep captured with --import-source on (-lineinfo build): warp-stall samples and
executed instructions per CUDA source line, from `ncu --page source --csv --print-source cuda,sass`.
    python tools/ncu_source_lines.py report.ncu-rep <kernel name regex> [top N]"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          f"regex:{kern}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    lines = []
    fname = ""
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1]
        if len(r) > 8 and r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0].strip().isdigit():
            d = dict(zip(hdr, r))
            try:
                lines.append((int(d["# Samples"]), int(d["Instructions Executed"]), fname.split("/")[-1], int(r[0]), r[1].strip()))
            except ValueError:
                pass
    tot_s = sum(l[0] for l in lines) or 1
    tot_i = sum(l[1] for l in lines) or 1
    print(f"# {kern}: {tot_s} stall samples, {tot_i} warp instructions over {len(lines)} source lines")
    print("# share of samples | share of instructions | file:line | source")
    for s, i, f, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"{100.0 * s / tot_s:5.1f}% {100.0 * i / tot_i:5.1f}%  {f}:{ln:<4d} {src[:110]}")


if __name__ == "__main__":
    main()
