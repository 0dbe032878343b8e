#!/usr/bin/env python
"""Per-source-line instruction and stall-sample counts of an .ncu-rep captured with --import-source on.

    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top_n] [env_steps]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    units = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         stdout=subprocess.PIPE, text=True).stdout
    rows, fname = [], None
    hdr = None
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] == "":
            continue
        d = dict(zip(hdr[:2] + ["Address", "Sass"] + hdr[4:], r))
        try:
            inst = int(d["Instructions Executed"]); thr = int(d["Thread Instructions Executed"]); smp = int(d["# Samples"])
        except (KeyError, ValueError):
            continue
        rows.append((fname, int(r[0]), r[1].strip(), inst, thr, smp))
    tot_i = sum(x[3] for x in rows) or 1
    tot_s = sum(x[5] for x in rows) or 1
    print("total warp-instructions %d, samples %d%s" % (tot_i, tot_s, (", per unit %.2f" % (tot_i / units)) if units else ""))
    byfile = {}
    for f, ln, src, i, t, s in rows:
        b = byfile.setdefault(f, [0, 0]); b[0] += i; b[1] += s
    for f, (i, s) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print("  %-22s inst %5.1f%%  samples %5.1f%%" % (f, 100.0 * i / tot_i, 100.0 * s / tot_s))
    print("top lines by instructions:")
    for f, ln, src, i, t, s in sorted(rows, key=lambda x: -x[3])[:top]:
        print("  %5.2f%% i  %5.2f%% s  thr/inst %4.1f  %s:%d  %s" % (100.0 * i / tot_i, 100.0 * s / tot_s, t / max(i, 1), f, ln, src[:90]))


if __name__ == "__main__":
    main()
