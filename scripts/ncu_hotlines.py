#!/usr/bin/env python
"""Top source lines of a kernel by warp-stall samples, from an ncu report captured with --import-source on.

    python scripts/ncu_hotlines.py gpurun_out/prof_<kernel>.ncu-rep [N]
"""
import csv
import subprocess
import sys


def main(path, top=30):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    fname, hdr, rows = "?", None, []
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[2] == "-" and r[0].isdigit():
            d = dict(zip(hdr, r))
            stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit()}
            rows.append((int(d["# Samples"]), fname, int(r[0]), r[1].strip(), stalls, int(d["Instructions Executed"])))
    tot = sum(x[0] for x in rows)
    print("total samples %d" % tot)
    for n, f, line, src, stalls, inst in sorted(rows, key=lambda x: -x[0])[:top]:
        s = ", ".join("%s %d" % kv for kv in sorted(stalls.items(), key=lambda kv: -kv[1])[:3] if kv[1])
        print("%5.1f%% %-22s:%-4d inst %-8d [%s]  %s" % (100.0 * n / max(tot, 1), f[:22], line, inst, s, src[:90]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
