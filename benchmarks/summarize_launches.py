"""profiles/traffic.json from the ncu launch list of bench.py (gpu__time_duration + dram bytes per decode-regime kernel launch:
cb200::stream_kernel, or cb200::decode_kernel where AUTO falls back to it).

    python benchmarks/summarize_launches.py profiles/r01_bench_launches.csv profiles/traffic.json
"""
import collections
import csv
import json
import sys


def main() -> None:
    src, dst = sys.argv[1], sys.argv[2]
    per = collections.defaultdict(dict)
    with open(src) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if "decode_kernel" not in r["Kernel Name"] and "stream_kernel" not in r["Kernel Name"]:
            continue
        per[(int(r["ID"]), r["Grid Size"])][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    groups = collections.defaultdict(list)
    for (_, grid), m in per.items():
        groups[grid].append(m)
    shapes, tot_bytes, tot_n = [], 0.0, 0
    for grid, ms in groups.items():
        n = len(ms)
        rd = sum(m.get("dram__bytes_read.sum", 0.0) for m in ms) / n
        wr = sum(m.get("dram__bytes_write.sum", 0.0) for m in ms) / n
        ns = sum(m.get("gpu__time_duration.sum", 0.0) for m in ms) / n
        shapes.append({"grid": grid, "launches": n, "avg_ns": ns, "avg_dram_read_bytes": rd, "avg_dram_write_bytes": wr})
        tot_bytes += (rd + wr) * n
        tot_n += n
    out = {
        "decode_kernel_dram_bytes_per_launch": tot_bytes / max(tot_n, 1),
        "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
        "-k regex:stream_kernel|decode_kernel on bench.py --steps 2 --warmup 1 (" + src + "), averaged over all decode-regime launches",
        "per_shape": shapes,
    }
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out)[:300])


if __name__ == "__main__":
    main()
