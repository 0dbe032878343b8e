#!/usr/bin/env python
"""Executed-instruction count per SASS line of an .ncu-rep (captured with --import-source on), divided by `units`.

    python tools/ncu_sass.py gpurun_out/x.ncu-rep units [min_per_unit]
"""
import csv
import subprocess
import sys


def main():
    rep, units = sys.argv[1], float(sys.argv[2])
    floor = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
    on = False
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "Address":
            on = True
            continue
        if on and len(r) > 6:
            try:
                i, t = int(r[5]), int(r[6])
            except ValueError:
                continue
            if i / units >= floor:
                print("%7.2f %5.1f  %s" % (i / units, t / max(i, 1), r[1][:110]))


if __name__ == "__main__":
    main()
