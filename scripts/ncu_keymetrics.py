"""Key metrics of an `ncu --set full` report (read on the CPU box): python scripts/ncu_keymetrics.py <report.ncu-rep> [out.json]
Prints one block per captured launch; with a second argument also writes the DRAM bytes per kernel (bench.py's roofline.traffic)."""
import csv, io, json, subprocess, sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    seen, traffic = set(), {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d["Kernel Name"]
        key = (name, d.get("launch__grid_size", ""))          # the same kernel at another grid size is another pass (set-up of the camera / the volume pass)
        if key in seen:
            continue
        seen.add(key)
        print("----")
        print(f"  {'Kernel Name':<80s} {name}")
        for w in WANT:
            if w in d:
                print(f"  {w:<80s} {d[w]:>16s} {u.get(w, '')}")
        def num(k):
            v = float(d[k].replace(",", ""))
            return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u.get(k, "byte"), 1.0)
        traffic[name] = {"dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"), "kernel": name}
    if len(sys.argv) > 2:
        key = lambda n: ("k_tile<GBUFFER>" if "k_tile<3" in n or "k_tile<1" in n else "k_tile<DEPTH>" if "k_tile<0" in n else
                         "k_tile<MOMENTS>" if "k_tile<4" in n else "k_visibility" if "k_visibility" in n else
                         "k_mom_visibility" if "k_mom_visibility" in n else ("k_mom_filter<X>" if "k_mom_filter<1" in n else "k_mom_filter<Y>") if "k_mom_filter" in n else n)
        with open(sys.argv[2], "w") as f:
            json.dump({"source": f"ncu --set full capture {rep} (per launch)", "kernels": {key(n): v for n, v in traffic.items()}}, f, indent=1)


if __name__ == "__main__":
    main()
