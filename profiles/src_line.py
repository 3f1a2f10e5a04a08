"""SASS instructions (with executed counts) attributed to one CUDA source line of an
`ncu --page source --print-source sass,cuda --csv` dump.   usage: python profiles/src_line.py dump.csv file.cu LINE [max]"""
import csv
import sys


def main(path, fname, line, top=60):
    rows = list(csv.reader(open(path)))
    cur_file, hdr, on, n = None, None, False, 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            ie = hdr.index("Instructions Executed")
            continue
        if hdr is None or r[0] == "Function Name":
            continue
        if r[0] != "":
            on = cur_file == fname and r[0] == str(line)
            if on:
                print("LINE", r[0], r[1][:100], "inst", r[ie])
            continue
        if on and n < top and len(r) > ie and r[3] != "...":
            print(f"   {r[ie]:>10s}  {r[3]}")
            n += 1


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 60)
