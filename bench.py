#!/usr/bin/env python
"""bench.py — FF + Sinkhorn clips/s (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

One step = the FF + Sinkhorn part of TimeT.get_loss (time_tuning.py:263-296, no teacher / queue) on
one batch of synthetic clips: cosine scores -> 2 x Sinkhorn-Knopp (K=200, 10 iterations, eps 0.05)
-> Feature-Forwarding of Q_source through every clip (n_last 7, radius 6, top-k 5) -> last-frame
hard labels.  Workload = BASELINE configs[1]: ViT-S/16 448^2 (28x28 patches, D=384), 8-frame
clips, batch 32 per GPU (weak scaling: 8 GPUs = the 256-clip batch of configs[2], Sinkhorn
marginals all-reduced over NCCL).

Prints ONE JSON line on rank 0 (keys documented in DESIGN.md §Measurement):
  value    : clips/s, inputs resident in HBM, CUDA events around exactly K steps, max over ranks
  e2e      : the same through the public API with pinned HOST inputs (H2D every step) and a D2H read
             of the hard labels inside the timed region
  roofline : the dominant kernel (affinity/top-k selection) timed live with CUDA events
  cpu_baseline : the oracle port of the reference path on this box's host cores, bounded sample
--impl reference times that CPU path alone (rank 0 only under torchrun).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CFG = dict(sr=28, dim=384, head_dim=256, K=200, fs=8, clips_per_gpu=32, n_last=7, radius=6, topk=5,
           epsilon=0.05, iters=10)
WORKLOAD = "ViT-S/16 448^2 (28x28 patches, D=384), 8-frame clips, batch 32 per GPU, K=200 (BASELINE configs[1])"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=CFG["clips_per_gpu"], help="clips per GPU")
    ap.add_argument("--engine", default="auto", choices=["auto", "exact", "tc"])
    ap.add_argument("--cpu-clips", type=int, default=8, help="clips in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------- synthetic inputs
def make_inputs(n_clips, seed):
    from timetuning_b200 import synth
    backbone = synth.clip_features(n_clips, CFG["fs"], CFG["sr"], CFG["dim"], seed=seed)
    head = synth.head_features(backbone[:, [0, -1]], CFG["head_dim"], seed=2)
    protos = synth.prototypes(CFG["K"], CFG["head_dim"], seed=3)
    return np.ascontiguousarray(head[:, 0]), np.ascontiguousarray(head[:, 1]), backbone, protos


# --------------------------------------------------------------------------- CPU arm (oracle port)
def _cpu_setup(n_clips):
    """The torch-CPU port of the reference's dense path (oracle/timet_oracle_torch.py) on all host cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import timet_oracle_torch as OT
    torch.set_num_threads(cpu_cores())
    hs, ht, bb, pr = (torch.from_numpy(x) for x in make_inputs(n_clips, seed=1))
    mask = OT.window_mask(CFG["sr"], CFG["radius"])       # one-off mask build is not timed

    def run():
        t0 = time.perf_counter()
        OT.ff_sinkhorn_step(hs, ht, bb, pr, CFG["sr"], CFG["n_last"], CFG["radius"], CFG["topk"], CFG["epsilon"],
                            CFG["iters"], mask)
        return time.perf_counter() - t0
    return run, torch.get_num_threads()


CPU_SAMPLE = ("{n} clips of the workload per step, torch-CPU port of the reference's dense per-clip path "
              "(same ATen operator sequence as mask_propagation.py:418-444 / my_utils.py:246-274, bit-identical to "
              "the reference on the golden fixtures), {thr} threads; neighbourhood-mask build excluded; clips/s is "
              "linear in clips (the reference loops over clips, time_tuning.py:277)")


def cpu_step_time(n_clips, reps=3):
    run, thr = _cpu_setup(n_clips)
    run()
    return min(run() for _ in range(reps)), thr


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args, rank):
    if rank != 0:
        return
    n = args.cpu_clips
    run, thr = _cpu_setup(n)
    for _ in range(max(1, min(args.warmup, 2))):
        run()
    steps = max(1, min(args.steps, 10))
    times = [run() for _ in range(steps)]
    t = sum(times) / len(times)
    val = n / t
    line = {"impl": "reference", "metric": "FF+Sinkhorn clips/s", "value": val, "unit": "clips/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample_clips": n},
            "cpu_baseline": {"value": val, "unit": "clips/s", "cores": thr, "kind": "port",
                             "sample": CPU_SAMPLE.format(n=n, thr=thr)},
            "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- our arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from timetuning_b200 import _cabi, dist as tdist, ops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        tdist.init_comm()
    engine = {"auto": ops.FF_AUTO, "exact": ops.FF_EXACT, "tc": ops.FF_TC}[args.engine]

    bs, fs, sr, D, K = args.clips, CFG["fs"], CFG["sr"], CFG["dim"], CFG["K"]
    N = sr * sr
    hs, ht, bb, pr = make_inputs(bs, seed=1 + rank)
    # pinned host copies (e2e arm) and device-resident copies (value arm)
    host = [torch.from_numpy(x).pin_memory() for x in (hs, ht, bb)]
    d_hs, d_ht, d_bb = (x.to(dev, non_blocking=True) for x in host)
    d_pr = torch.from_numpy(pr).to(dev)
    h2d_bytes = sum(x.numel() * x.element_size() for x in host)
    d2h_bytes = bs * N * 8

    plan = ops._plan(bs, fs, sr, sr, D, K, CFG["n_last"], CFG["radius"], CFG["topk"], device=dev)
    engine_used = "tcgen05" if (engine != ops.FF_EXACT and plan.tc_supported) else "exact-fp32"
    labels = torch.empty((bs, fs, N, K), dtype=torch.float32, device=dev)
    hard = torch.empty((bs, N), dtype=torch.int64, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    stage_ev = []
    tc_events = [] if engine_used == "tcgen05" else None

    def step(x_hs, x_ht, x_bb, record=False):
        """staged form of timetuning_b200.step.ff_sinkhorn_step so the selection kernel can be timed live"""
        e = [ev() for _ in range(7)] if record else None
        if record: e[6].record()
        s_src = ops.cosine_scores(x_hs.reshape(bs * N, -1), d_pr)
        s_tgt = ops.cosine_scores(x_ht.reshape(bs * N, -1), d_pr)
        if record: e[0].record()
        q_src = ops.sinkhorn_from_scores(s_src, CFG["epsilon"], CFG["iters"], world)
        q_tgt = ops.sinkhorn_from_scores(s_tgt, CFG["epsilon"], CFG["iters"], world)
        if record: e[1].record()
        labels[:, 0] = q_src.view(bs, N, K)
        if record: e[2].record()
        plan.prepare(x_bb)
        if record: e[3].record()
        if record and tc_events is not None:
            eb, ee = ev(), ev()
            eb.record(); ee.record()                    # materialise the cudaEvent_t handles
            plan.select_timed(engine, eb, ee)           # the library re-records them around the tcgen05 kernel
            tc_events.append((eb, ee))
        else:
            plan.select(engine)
        if record: e[4].record()
        plan.gather(labels, hard)
        if record:
            e[5].record()
            stage_ev.append(e)
        return q_src, q_tgt

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value arm: inputs resident in HBM
    for _ in range(max(args.warmup, 3)):
        step(d_hs, d_ht, d_bb)
    sync_all()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    n0 = _cabi.launch_count()
    t0, t1 = ev(), ev()
    sync_all()
    t0.record()
    for _ in range(args.steps):
        step(d_hs, d_ht, d_bb, record=True)
    t1.record()
    sync_all()
    launches = _cabi.launch_count() - n0
    ms_total = torch.tensor([t0.elapsed_time(t1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = ms_total.item() / args.steps
    value = bs * world / (ms_step * 1e-3)

    # ---- e2e arm: pinned host inputs, H2D + D2H inside the timed region, through the public API
    from timetuning_b200.step import HostStepPipeline
    pipe = HostStepPipeline(bs, fs, N, D, CFG["head_dim"], K, chunks=4, device=dev)

    def e2e_step():
        # public API with pinned host inputs: chunked H2D overlapped with compute, D2H of the hard labels
        pipe.run(host[0], host[1], host[2], d_pr, CFG["n_last"], CFG["radius"], CFG["topk"], CFG["epsilon"],
                 CFG["iters"], world, engine)
    for _ in range(3):
        e2e_step()
    sync_all()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    sync_all()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = bs * world / (e2e_ms.item() / args.steps * 1e-3)
    clock_info = clocks.stop() if rank == 0 else None

    # ---- per-stage device times (CUDA events recorded inside the timed region, same stream)
    names = ["sinkhorn_x2", "label_init", "prepare", "select", "gather"]
    stage_ms = {n: sum(e[i].elapsed_time(e[i + 1]) for e in stage_ev) / len(stage_ev) for i, n in enumerate(names)}
    stage_ms["cosine_scores_x2"] = sum(e[6].elapsed_time(e[0]) for e in stage_ev) / len(stage_ev)
    st = plan.stats()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sigma_ctx = sum(1 + (t - max(1, t - CFG["n_last"])) for t in range(1, fs))
        dense_flops = 2.0 * N * N * D * sigma_ctx * bs                        # SURVEY.md §8d, per launch (per GPU)
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        if engine_used == "tcgen05":
            tc_ms = sum(a.elapsed_time(b) for a, b in tc_events) / len(tc_events)
            stage_ms["select_tc_kernel"] = tc_ms
            # executed = key tiles actually multiplied (window band only), M padded to 128 rows
            QR, RPC, W_ = 128 // sr, (256 // sr) // (16 // __import__("math").gcd(sr, 16)) * (16 // __import__("math").gcd(sr, 16)), sr
            exec_flops = 0.0
            for qt in range(-(-sr // QR)):
                qr0, qr1 = qt * QR, min(sr - 1, qt * QR + QR - 1)
                rows = min(sr - 1, qr1 + CFG["radius"]) - max(0, qr0 - CFG["radius"]) + 1
                exec_flops += 2.0 * 128 * rows * W_ * D
            exec_flops *= sigma_ctx * bs
            roof = {"bound": "tensor", "kernel": "ff_tc_persist_kernel (tcgen05 affinity + fused window / top-k nomination), "
                    "timed alone with CUDA events recorded by the library around its launch",
                    "achieved": dense_flops / (tc_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback",
                    "flops_model": "dense 2*N^2*D*sum_ctx per clip (SURVEY.md §8d) x clips per launch",
                    "executed_tflops": exec_flops / (tc_ms * 1e-3) / 1e12, "executed_over_dense": exec_flops / dense_flops,
                    "ms_per_launch": tc_ms,
                    "traffic": 161.9e6, "traffic_source": "ncu --set full, profiles/r1_ncu_full_kernels.md: dram read 154.5 MB + write 7.4 MB per launch"}
        else:
            roof = {"bound": "tensor", "kernel": "ff_select_exact (fp32 CUDA-core scan; tensor-core engine not used)",
                    "achieved": dense_flops / (stage_ms["select"] * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback",
                    "flops_model": "dense 2*N^2*D*sum_ctx per clip (SURVEY.md §8d)", "traffic": None}
        roof["frac"] = roof["achieved"] / roof["peak"]
        hbm = peaks.get("hbm_gbs", 6650.0)
        sk_bytes = 2 * (CFG["iters"] + 2) * bs * N * K * 4
        gather_bytes = bs * ((fs - 1) * N * K * 4 + N * K * 4)
        prep_bytes = bs * fs * N * (D * 4 + D * 6)
        extra = {
            "sinkhorn": {"bound": "hbm", "achieved": sk_bytes / (stage_ms["sinkhorn_x2"] * 1e-3) / 1e9, "peak": hbm,
                         "unit": "GB/s", "bytes_model": "(iters+2)*B*K*4 per call, 2 calls (streaming model, SURVEY.md §8d); the "
                         "resident kernel's real DRAM traffic is the compulsory 2*B*K*4"},
            "gather": {"bound": "hbm", "achieved": gather_bytes / (stage_ms["gather"] * 1e-3) / 1e9, "peak": hbm,
                       "unit": "GB/s", "bytes_model": "compulsory (fs-1)*N*C*4 write + N*C*4 read per clip"},
            "prepare": {"bound": "hbm", "achieved": prep_bytes / (stage_ms["prepare"] * 1e-3) / 1e9, "peak": hbm,
                        "unit": "GB/s", "bytes_model": "read 4*D, write 6*D per row"},
        }
        for v in extra.values():
            v["frac"] = v["achieved"] / v["peak"]
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            t, thr = cpu_step_time(args.cpu_clips)
            cpu = {"value": args.cpu_clips / t, "unit": "clips/s", "cores": thr, "kind": "port",
                   "sample": "best of 3 after 1 warm-up; " + CPU_SAMPLE.format(n=args.cpu_clips, thr=thr)}
        line = {"metric": "FF+Sinkhorn clips/s", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "clips_per_gpu": bs, "global_clips": bs * world, "engine": engine_used,
                           "precision": "all results f32; the tcgen05 kernel nominates top-k candidates from fp16 inputs / fp32 "
                                        "accumulation and every nominated key is re-evaluated in f32 (bit-identical to the f32 engine)"
                           if engine_used == "tcgen05" else "f32",
                           "l2": "inputs larger than L2 (backbone features %.0f MB per step)" % (d_bb.numel() * 4 / 1e6),
                           "parallelism": f"clips sharded over {world} GPU(s); Sinkhorn marginals all-reduced (NCCL)"
                           if world > 1 else "single GPU"},
                "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
                "gpu_launches": launches, "roofline": roof, "stage_ms": stage_ms, "stage_roofline": extra,
                "ff_stats": st, "clocks": clock_info, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        tdist.destroy_comm()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
