"""Hot spots of a .ncu-rep source page: aggregates warp-stall samples per SASS opcode class and lists the
top instructions with their dominant stall reason.  usage: python tools/ncu_hot.py rep.ncu-rep [topN]"""
import csv
import subprocess
import sys
from collections import Counter


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    data = []
    for r in rows:
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(r)
    ci = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = 0
    by_reason = Counter()
    by_op = Counter()
    items = []
    for r in data:
        n = int(r[ci["# Samples"]] or 0)
        tot += n
        s = r[ci["Source"]].split()
        op = s[1] if s and s[0].startswith("@") and len(s) > 1 else (s[0] if s else "?")
        by_op[op.split(".")[0]] += n
        st = {h: int(r[ci[h]] or 0) for h in stall_cols}
        for h, v in st.items():
            by_reason[h] += v
        items.append((n, r[ci["Address"]], r[ci["Source"]][:90], max(st, key=st.get) if n else "", int(r[ci["Instructions Executed"]] or 0), op.split(".")[0]))
    print("total samples", tot)
    print("by stall reason:", [(k, round(100.0 * v / tot, 1)) for k, v in by_reason.most_common(10)])
    print("by opcode:", [(k, round(100.0 * v / tot, 1)) for k, v in by_op.most_common(14)])
    ninst = sum(i[4] for i in items)
    ops = Counter()
    for i in items:
        ops[i[5]] += i[4]
    print("executed warp-instructions by opcode (%):", [(k, round(100.0 * v / ninst, 1)) for k, v in ops.most_common(18)])
    items.sort(reverse=True)
    for n, addr, src, why, ne, op in items[:top]:
        print("%6d %5.2f%% %-14s %s" % (n, 100.0 * n / tot, why, src))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
