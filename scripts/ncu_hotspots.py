"""Per-source-line hot spots of an `ncu --set full --import-source on` report (read on the CPU box):
   python scripts/ncu_hotspots.py <report.ncu-rep> [kernel-substring] [top N]
Aggregates `ncu --page source --print-source sass,cuda` per CUDA source line: share of warp-stall samples and of executed warp instructions."""
import csv, io, subprocess, sys


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    fn, hdr, lines = None, None, {}
    seen = set()

    def flush():
        if fn is None or not lines or (want and want not in fn) or fn in seen:
            return
        seen.add(fn)
        ts = sum(v[0] for v in lines.values()) or 1
        ti = sum(v[1] for v in lines.values()) or 1
        print(f"== {fn}  ({ts} samples, {ti} warp instructions)")
        col = 1 if "--by-inst" in sys.argv else 0
        for ln, v in sorted(lines.items(), key=lambda kv: -kv[1][col])[:top]:
            print(f"  {ln:>5s} {100 * v[0] / ts:5.1f}% samples {100 * v[1] / ti:5.1f}% inst   {v[2][:150]}")
        print()

    for r in rows:
        if not r:
            continue
        if r[0] == "Function Name":
            flush()
            fn, hdr, lines = r[1], None, {}
        elif r[0] == "Line No":
            hdr = r
        elif hdr and fn and r[0] not in ("", "File Path") and len(r) > 8:
            try:
                i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
                lines[r[0]] = [int(r[i_s]), int(r[i_i]), r[1].strip()]
            except Exception:
                pass
    flush()


if __name__ == "__main__":
    main()
