"""Extract the headline metrics of every kernel in an .ncu-rep (ncu --set full) into a markdown table.
usage: python profiles/extract_ncu.py gpurun_out/x.ncu-rep > profiles/x.md"""
import csv
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg": "hmma_cyc",
    "sm__cycles_elapsed.max": "cycles",
    "sm__cycles_elapsed.max.per_second": "sm_clk",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ%",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue%",
    "lts__t_sector_hit_rate.pct": "l2hit%",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram%",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm%",
}


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = {h: i for i, h in enumerate(hdr)}
    names = [w for w in WANT if w in cols]
    print("| kernel | " + " | ".join(WANT[n] for n in names) + " |")
    print("|---|" + "---|" * len(names))
    for r in rows[2:]:
        k = r[cols["Kernel Name"]].split("(")[0]
        cells = []
        for n in names:
            v, u = r[cols[n]], units[cols[n]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:,.4g}"
            except ValueError:
                pass
            cells.append(f"{v} {u}".strip())
        print(f"| {k} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
