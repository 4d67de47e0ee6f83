"""Top source lines by warp-stall samples from an .ncu-rep captured with --import-source on."""
import csv
import subprocess
import sys


def main(path, top=25):
    txt = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    files = [i for i, r in enumerate(rows) if r and r[0] == "File Path"]
    for s_i, s in enumerate(starts):
        hdr = rows[s]
        si = hdr.index("# Samples")
        end = starts[s_i + 1] if s_i + 1 < len(starts) else len(rows)
        fname = ""
        for f in files:
            if f < s:
                fname = rows[f][1] if len(rows[f]) > 1 else ""
        lines = []
        for r in rows[s + 1:end]:
            if r and r[0].isdigit():
                try:
                    lines.append((int(r[si]), int(r[0]), r[1][:110]))
                except Exception:
                    pass
        tot = sum(l[0] for l in lines)
        if tot == 0:
            continue
        print(f"== {fname}  total samples {tot}")
        for smp, ln, src in sorted(lines, reverse=True)[:top]:
            print(f"  {smp:7d} {100.0 * smp / tot:5.1f}%  L{ln}: {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
