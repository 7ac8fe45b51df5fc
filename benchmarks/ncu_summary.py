"""Turn an .ncu-rep (read here, on the CPU box) into the text summary committed under profiles/.

    python benchmarks/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_decode_ncu_summary.txt [--stalls]

Writes every raw metric of the first kernel in the report (name, value, unit) and, with --stalls, the instructions with
the most warp-stall samples from the source page (needs a -lineinfo build and --import-source on).
"""
import csv
import io
import subprocess
import sys


def main() -> None:
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, first = rows[0], rows[1], rows[2]
    name = first[hdr.index("Kernel Name")]
    lines = [
        f"# ncu --set full --clock-control none --import-source on (one launch), kernel: {name}",
        "# command: see benchmarks/profile_target.py; numbers under a profiler are NOT bench values",
    ]
    start = hdr.index("Kernel Name") + 1
    for i in sorted(range(start, len(hdr)), key=lambda j: hdr[j]):
        if first[i] == "" or hdr[i] in ("Context", "Stream", "Block Size", "Grid Size", "Device", "CC", "Section Name"):
            continue
        lines.append(f"{hdr[i]:100s} {first[i]:>14s} {units[i]}")
    if "--stalls" in sys.argv:
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        h = srows[1]
        isrc, isamp = h.index("Source"), h.index("# Samples")
        stall_cols = [j for j, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        data = []
        for r in srows[2:]:
            if len(r) < len(h):
                continue
            try:
                n = int(r[isamp] or 0)
            except ValueError:
                continue
            top = sorted(((int(r[j] or 0), h[j]) for j in stall_cols), reverse=True)[:2]
            data.append((n, r[isrc].strip(), top))
        total = sum(d[0] for d in data) or 1
        lines.append("")
        lines.append(f"# warp-state samples by SASS instruction (top 25 of {total} samples)")
        for n, text, top in sorted(data, reverse=True)[:25]:
            lines.append(f"{n:7d} {100.0 * n / total:5.1f}%  {text[:90]:90s} {top[0][1]}={top[0][0]} {top[1][1]}={top[1][0]}")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(f"wrote {out}: {len(lines)} lines, kernel {name}")


if __name__ == "__main__":
    main()
