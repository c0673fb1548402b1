"""profiles/roofline_traffic.json from an ncu CSV of one bench step:
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \\
      -k regex:k_stage_march --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu
usage: python tools/ncu_traffic.py gpurun_out/traffic.csv cells_per_launch [launches_per_step]
Averages the LAST launches_per_step launches (the timed step; the ones before are warm-up)."""
import csv
import json
import sys
from collections import defaultdict


def to_bytes(v, u):
    k = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * k[u]


def main(path, cells, per_step=8):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    ci = {h: i for i, h in enumerate(hdr)}
    per = defaultdict(dict)
    for r in rows[start:]:
        if len(r) != len(hdr):
            continue
        per[int(r[ci["ID"]])][r[ci["Metric Name"]]] = (r[ci["Metric Value"]].replace(",", ""), r[ci["Metric Unit"]])
    ids = sorted(per)[-per_step:]
    rd = [to_bytes(*per[i]["dram__bytes_read.sum"]) for i in ids]
    wr = [to_bytes(*per[i]["dram__bytes_write.sum"]) for i in ids]
    tot = [a + b for a, b in zip(rd, wr)]
    out = {"kernel": "k_stage_pipe (k_stage_march for the shapes it does not cover)", "cells_per_launch": cells, "launches_averaged": len(ids),
           "dram_bytes_per_launch": sum(tot) / len(tot), "dram_read_bytes_per_launch": sum(rd) / len(rd),
           "dram_write_bytes_per_launch": sum(wr) / len(wr), "dram_bytes_per_cell": sum(tot) / len(tot) / cells,
           "per_launch_bytes": tot,
           "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, bench.py --steps 1 --warmup 1 (timed step's launches)"}
    json.dump(out, open("profiles/roofline_traffic.json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 8)
