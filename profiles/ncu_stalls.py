"""Stall-reason summary and source-level hot regions of one kernel in an .ncu-rep (ncu --set full --import-source on).
usage: python profiles/ncu_stalls.py gpurun_out/prof.ncu-rep [kernel-name-substring] > profiles/x.txt"""
import csv
import subprocess
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main(path, pat=""):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    ki = hdr.index("Kernel Name")
    row = next(r for r in rows[2:] if pat in r[ki])
    print(f"# {row[ki][:100]}")
    keep = ("gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
            "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "lts__t_sector_hit_rate.pct")
    for h, v in zip(hdr, row):
        if h in keep:
            print(f"{h:90s} {v}")
    print("# warps stalled per issue-active cycle, by reason")
    for h, v in zip(hdr, row):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            print(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:24s} {num(v):6.2f}")

    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + pat] if pat else []),
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = next(r for r in rows if r and r[0] == "Address")
    idx = {h: i for i, h in enumerate(hdr)}
    seen, data = set(), []
    for r in rows:
        if len(r) > idx["stall_wait"] and r[0].startswith("0x") and r[0] not in seen:
            seen.add(r[0])
            data.append(r)
    tot = sum(num(r[idx["# Samples"]]) for r in data)
    ex = sum(num(r[idx["Instructions Executed"]]) for r in data)
    print(f"# source page: {len(data)} SASS instructions, {tot:.0f} stall samples, {ex / 1e6:.1f} M warp instructions executed")
    print("# regions of 64 instructions with > 1.2 % of the samples: offset, samples %, warp instructions executed, top opcodes")
    base = int(data[0][0], 16)
    for i in range(0, len(data), 64):
        ch = data[i:i + 64]
        s = sum(num(r[idx["# Samples"]]) for r in ch)
        if s / tot <= 0.012:
            continue
        e = sum(num(r[idx["Instructions Executed"]]) for r in ch)
        ops = {}
        for r in ch:
            tok = [o for o in r[idx["Source"]].split() if not o.startswith("@")]
            op = tok[0].split(".")[0] if tok else ""
            ops[op] = ops.get(op, 0) + 1
        top = " ".join(f"{k}x{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
        print(f"  +{int(ch[0][0], 16) - base:05x}  {100 * s / tot:5.1f} %  {e / 1e6:7.2f} M  {top}")
    print("# instructions with the most samples")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in sorted(data, key=lambda r: -num(r[idx["# Samples"]]))[:16]:
        s = num(r[idx["# Samples"]])
        st = sorted(((num(r[idx[h]]), h) for h in stalls), reverse=True)[:2]
        print(f"  +{int(r[0], 16) - base:05x}  {100 * s / tot:5.2f} %  {r[idx['Source']].strip()[:70]:70s} " + " ".join(f"{h[6:]}={v:.0f}" for v, h in st if v > 0))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
