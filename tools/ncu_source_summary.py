#!/usr/bin/env python
"""Per-source-line summary of one kernel from an ncu report captured with --import-source on (needs -lineinfo).

    python tools/ncu_source_summary.py gpurun_out/prof.ncu-rep [top] [sort: samp|inst] [kernel name (first launch of it)]
Prints, per CUDA source line: share of executed warp instructions, share of stall samples, average active threads.
"""
import csv
import io
import subprocess
import sys


def main():
    rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
    key = sys.argv[3] if len(sys.argv) > 3 else "samp"
    sel = ["--kernel-name", sys.argv[4], "--launch-count", "1"] if len(sys.argv) > 4 else []
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + sel, capture_output=True, text=True).stdout
    hdr, lines, fname = None, [], ""
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0]:   # a source line (SASS rows have an empty line number)
            i_inst, i_thr, i_samp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
            try:
                inst, thr, samp = float(r[i_inst]), float(r[i_thr]), float(r[i_samp])
            except ValueError:
                continue
            if inst > 0 or samp > 0:
                lines.append((inst, thr, samp, fname, r[0], r[1].strip()))
    total = sum(x[0] for x in lines) or 1.0
    stot = sum(x[2] for x in lines) or 1.0
    print(f"total warp instructions {total:.3e}, samples {stot:.0f}")
    for inst, thr, samp, f, ln, src in sorted(lines, key=lambda x: -(x[2] if key == "samp" else x[0]))[:top]:
        print(f"{100 * inst / total:5.1f}% inst  {100 * samp / stot:5.1f}% samp  thr {thr / inst if inst else 0:5.1f}  {f}:{ln}  {src[:120]}")


if __name__ == "__main__":
    main()
