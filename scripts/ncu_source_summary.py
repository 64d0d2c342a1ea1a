#!/usr/bin/env python
"""Export the per-source-line hot spots (stall samples) of every kernel in an .ncu-rep as small text files.
usage: ncu_source_summary.py <report.ncu-rep> <outdir>"""
import csv, io, subprocess, sys, collections, os
rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
# the source page prints one CSV table per kernel launch, separated by header lines
with open(os.path.join(out, "prof_source.csv"), "w") as f:
    f.write(txt)
subprocess.run(["gzip", "-f", os.path.join(out, "prof_source.csv")])
