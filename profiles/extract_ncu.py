"""Extract the headline metrics of every kernel in one or more .ncu-rep files (ncu --set full) into a markdown table and
into profiles/ncu_kernels.json -- the file bench.py reads `roofline.traffic` from.

usage: python profiles/extract_ncu.py --config 2 --tag r2 gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...] > profiles/r2_ncu_full_kernels.md

For every kernel name only the LAST captured launch is kept (captures are taken after warm-up launches).  Metrics:
time, DRAM bytes read / written, L2 -> SM bytes (crossbar to L1TEX, includes TMA), tensor-pipe active % (the counter the
north star asks for), tensor instructions, registers, occupancy, issue %, L2 hit %, DRAM %, SM %.
"""
import argparse
import csv
import json
import os
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time_us", "time"),
    ("dram__bytes_read.sum", "dram_bytes_read", "dram_rd"),
    ("dram__bytes_write.sum", "dram_bytes_write", "dram_wr"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm_bytes", "L2->SM"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct", "tensor%"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "tensor_hmma_cycles", "hmma_cyc"),
    ("sm__inst_executed_pipe_tensor_realtime.sum", "tensor_inst", "tensor_inst"),
    ("sm__cycles_elapsed.max", "cycles", "cycles"),
    ("launch__registers_per_thread", "registers", "regs"),
    ("launch__grid_size", "grid", "grid"),
    ("launch__block_size", "block", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct", "occ%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct", "issue%"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct", "l2hit%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct", "sm%"),
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def find_col(cols, name):
    """ncu prefixes some metrics with their section (e.g. 'TPC.TriageCompute.sm__pipe_tensor...')."""
    if name in cols:
        return cols[name]
    for h, i in cols.items():
        if h.endswith("." + name):
            return i
    return None


def read(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = {h: i for i, h in enumerate(hdr)}
    res = {}
    for r in rows[2:]:
        name = r[cols["Kernel Name"]].split("(")[0].replace("void ", "").strip()
        rec = {"kernel": name}
        for metric, key, _ in WANT:
            i = find_col(cols, metric)
            if i is None or r[i] in ("", "n/a"):
                continue
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            rec[key] = v * UNIT.get(units[i], 1.0)
        res[name] = rec            # keep the last launch of each kernel
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+")
    ap.add_argument("--config", type=int, default=2, help="SURVEY.md §8 config the capture ran (bench.py --config)")
    ap.add_argument("--tag", default="r2")
    ap.add_argument("--json", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_kernels.json"))
    a = ap.parse_args()
    kernels = {}
    for rep in a.reports:
        for name, rec in read(rep).items():
            rec.update(capture=f"{a.tag}: {os.path.basename(rep)}", survey_config=a.config)
            kernels[name] = rec
    keys = [(k, short) for _, k, short in WANT]
    print("| kernel | " + " | ".join(s for _, s in keys) + " |")
    print("|---|" + "---|" * len(keys))
    for name, rec in kernels.items():
        cells = []
        for k, _ in keys:
            v = rec.get(k)
            if v is None:
                cells.append("-")
            elif k.endswith("bytes") or k.startswith("dram_bytes") or k == "l2_to_sm_bytes":
                cells.append(f"{v / 1e6:,.1f} MB")
            elif k == "time_us":
                cells.append(f"{v:,.1f} us")
            else:
                cells.append(f"{v:,.4g}")
        print(f"| {name} | " + " | ".join(cells) + " |")
    old = {}
    if os.path.isfile(a.json):
        try:
            old = {(r["kernel"], r.get("survey_config")): r for r in json.load(open(a.json)).get("kernels", [])}
        except Exception:
            old = {}
    for name, rec in kernels.items():
        old[(name, a.config)] = rec
    json.dump({"source": "profiles/extract_ncu.py (ncu --set full --clock-control none); one record per kernel and config, last capture wins",
               "kernels": list(old.values())}, open(a.json, "w"), indent=1)
    print(f"\n{len(kernels)} kernel(s) -> {a.json}", file=sys.stderr)


if __name__ == "__main__":
    main()
