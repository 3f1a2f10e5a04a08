"""Aggregate an `ncu --page source --print-source sass,cuda --csv` dump per CUDA source line.
usage: python profiles/src_hot.py dump.csv [top_n]"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr, cur_file, out = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            ie = hdr.index("Instructions Executed")
            ns = hdr.index("# Samples")
            continue
        if hdr is None or r[0] == "":
            continue
        try:
            out.append((int(r[ie]), int(r[ns]), cur_file, r[0], r[1].strip()[:95]))
        except ValueError:
            pass
    tot = sum(o[0] for o in out) or 1
    tots = sum(o[1] for o in out) or 1
    print(f"total warp instructions {tot}, samples {tots}")
    for o in sorted(out, key=lambda x: -x[0])[:top]:
        print(f"{o[0] / tot * 100:5.1f}% inst {o[1] / tots * 100:5.1f}% smp  {o[2]}:{o[3]:>4s}  {o[4]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
