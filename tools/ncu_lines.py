"""Per-CUDA-line instruction counts and stall samples of one kernel from an .ncu-rep (source page, cuda,sass view),
grouped into named line ranges.

    python tools/ncu_lines.py rep.ncu-rep file.cu "name:lo-hi,name:lo-hi,..." [kernel index]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, fname, groups = sys.argv[1], sys.argv[2], sys.argv[3]
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    # the report has one block per (kernel launch, file); a block starts with a "File Path" row
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if row and row[0] == "File Path":
            cur = {"file": row[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(row)
    sel = [b for b in blocks if b["file"].endswith(fname)]
    b = sel[which]
    header = next(r for r in b["rows"] if r and r[0] == "Line No")
    i_inst = header.index("Instructions Executed")
    i_smp = header.index("# Samples")
    per_line = {}
    for r in b["rows"]:
        if len(r) <= i_inst or not r[0].isdigit() or r[2] != "-":      # cuda rows have "-" in the Address column
            continue
        per_line[int(r[0])] = (int(r[i_inst] or 0), int(r[i_smp] or 0))
    tot_i = sum(v[0] for v in per_line.values()) or 1
    tot_s = sum(v[1] for v in per_line.values()) or 1
    print(f"# {rep} [{fname}, launch {which}]: {tot_i} warp instructions, {tot_s} stall samples attributed to this file")
    print(f"{'phase':28s} {'instr %':>8s} {'stall %':>8s}")
    for g in groups.split(","):
        name, rng = g.split(":")
        lo, hi = (int(x) for x in rng.split("-"))
        ii = sum(v[0] for k, v in per_line.items() if lo <= k <= hi)
        ss = sum(v[1] for k, v in per_line.items() if lo <= k <= hi)
        print(f"{name:28s} {100 * ii / tot_i:8.1f} {100 * ss / tot_s:8.1f}")
    top = sorted(per_line.items(), key=lambda kv: -kv[1][0])[:12]
    print("top lines by instructions:", ", ".join(f"{k}:{100 * v[0] / tot_i:.1f}%" for k, v in top))
    top = sorted(per_line.items(), key=lambda kv: -kv[1][1])[:12]
    print("top lines by stall samples:", ", ".join(f"{k}:{100 * v[1] / tot_s:.1f}%" for k, v in top))


if __name__ == "__main__":
    main()
