#!/usr/bin/env python
"""bench.py — FF + Sinkhorn clips/s (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1..5] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

One step = the FF + Sinkhorn part of TimeT.get_loss (time_tuning.py:263-296, no teacher / queue) on one batch of
synthetic clips: cosine scores -> 2 x Sinkhorn-Knopp -> Feature-Forwarding of Q_source through every clip ->
last-frame hard labels.  Default workload = BASELINE configs[1] (--config 2 in SURVEY.md §8 numbering): ViT-S/16 448^2
(28x28 patches, D=384), 8-frame clips, batch 32 per GPU, K=200; weak scaling (8 GPUs = the 256-clip batch of configs[2]).
Other SURVEY §8 configs: --config 1 (configs[0], 224^2 batch 2), 3 (configs[2]: 256 clips GLOBAL, --scaling strong ->
256/N per GPU), 4 (configs[3]: DAVIS-style eval, 80-frame 60x60 video, radius 12, top-k 7, C=11, one video per GPU;
step = propagate + bilinear up-sampling + argmax, no Sinkhorn), 5 (configs[4]: ViT-B/8 56x56, D=768, 16 frames, K=300,
8 clips per GPU).

Prints ONE JSON line on rank 0 (keys documented in DESIGN.md §5):
  value        clips/s, inputs resident in HBM, CUDA events around exactly K steps, max over ranks
  e2e          the same through the public API with pinned HOST inputs (H2D every step) and a D2H read of the result
  roofline     the dominant kernel (tcgen05 affinity + top-k nomination) timed live with CUDA events: EXECUTED FLOPs
               (tiles issued) / time / measured peak; the dense-equivalent figure is kept beside it
  dist_parity  (N > 1) multi-GPU results checked before timing: distributed Sinkhorn vs the reference's own 2-rank
               fixture (N = 2), vs an fp64 evaluation and vs the single-GPU kernel on the global batch; sharded FF ==
               single process.  A failure exits non-zero.
  cpu_baseline the reference's own CPU path (kind "reference", from the shipped baseline/_ref copy) or its torch-CPU
               port (kind "port") on this box's host cores, bounded sample
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

# SURVEY.md §8 config table (1-based; BASELINE.json configs[i-1])
CONFIGS = {
    1: dict(sr=14, dim=384, head_dim=256, K=200, fs=4, clips=2, n_last=7, radius=6, topk=5, kind="train",
            name="ViT-S/16 224^2 (14x14 patches, D=384), 4-frame clips, batch 2, K=200 (BASELINE configs[0])"),
    2: dict(sr=28, dim=384, head_dim=256, K=200, fs=8, clips=32, n_last=7, radius=6, topk=5, kind="train",
            name="ViT-S/16 448^2 (28x28 patches, D=384), 8-frame clips, batch 32 per GPU, K=200 (BASELINE configs[1])"),
    3: dict(sr=28, dim=384, head_dim=256, K=200, fs=8, clips=256, n_last=7, radius=6, topk=5, kind="train",
            name="ViT-S/16 448^2 (28x28 patches, D=384), 8-frame clips, batch 256 GLOBAL sharded over the GPUs, K=200 "
                 "(BASELINE configs[2])"),
    4: dict(sr=60, dim=384, head_dim=256, K=11, fs=80, clips=1, n_last=7, radius=12, topk=7, kind="eval", out_res=480,
            name="DAVIS-style propagation eval, 480p ViT-S/8 (60x60 patches, D=384), 80-frame video, radius 12, "
                 "top-k 7, n_last 7, C=11, one video per GPU (BASELINE configs[3])"),
    5: dict(sr=56, dim=768, head_dim=256, K=300, fs=16, clips=8, n_last=7, radius=6, topk=5, kind="train",
            name="ViT-B/8 448^2 (56x56 patches, D=768), 16-frame clips, batch 8 per GPU, K=300 (BASELINE configs[4])"),
}
EPSILON, ITERS = 0.05, 10


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="SURVEY.md §8 config number")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="weak: clips per GPU fixed (default); strong: the config's clip count is GLOBAL (default for --config 3)")
    ap.add_argument("--clips", type=int, default=None, help="override: clips per GPU (weak) / global clips (strong)")
    ap.add_argument("--engine", default="auto", choices=["auto", "exact", "tc"])
    ap.add_argument("--cpu-clips", type=int, default=None, help="clips in the bounded CPU-baseline sample")
    ap.add_argument("--cpu-impl", default="auto", choices=["auto", "reference", "port"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--overlap", action="store_true", help="run the cosine-scores + Sinkhorn chain on a second stream under the "
                    "Feature-Forwarding kernels (measured: no gain, see profiles/r2_experiments.md); default is one stream")
    return ap.parse_args()


def resolve(args, world):
    cfg = dict(CONFIGS[args.config])
    scaling = args.scaling or ("strong" if args.config == 3 else "weak")
    n = args.clips if args.clips is not None else cfg["clips"]
    if scaling == "strong":
        if n % world:
            raise SystemExit(f"--scaling strong: {n} clips do not shard evenly over {world} GPUs")
        cfg["clips_per_gpu"], cfg["global_clips"] = n // world, n
    else:
        cfg["clips_per_gpu"], cfg["global_clips"] = n, n * world
    cfg["scaling"] = scaling
    return cfg


# --------------------------------------------------------------------------- synthetic inputs
def make_inputs(cfg, n_clips, seed):
    from timetuning_b200 import synth
    backbone = synth.clip_features(n_clips, cfg["fs"], cfg["sr"], cfg["dim"], seed=seed)
    if cfg["kind"] == "eval":
        ann = np.stack([synth.blob_label_map(cfg["out_res"], cfg["K"] - 1, seed=seed + 50 + b) for b in range(n_clips)])
        return dict(backbone=backbone, annotation=ann)
    head = synth.head_features(backbone[:, [0, -1]], cfg["head_dim"], seed=2)
    protos = synth.prototypes(cfg["K"], cfg["head_dim"], seed=3)
    return dict(head_src=np.ascontiguousarray(head[:, 0]), head_tgt=np.ascontiguousarray(head[:, 1]), backbone=backbone,
                prototypes=protos)


# --------------------------------------------------------------------------- CPU arm
def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _cpu_setup(cfg, n_clips, want):
    """Returns (run() -> seconds, threads, kind, sample description).  kind "reference": the reference's OWN code
    (time_tuning.TimeT.get_scores x2 + the per-clip make_seg_maps loop, or mask_propagation.propagate_labels + the eval
    tail) imported from the shipped baseline/_ref copy; kind "port": oracle/timet_oracle_torch.py, the same ATen
    operator sequence restated (bit-identical to the reference on the golden fixtures)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    torch.set_num_threads(cpu_cores())
    thr = torch.get_num_threads()
    x = {k: torch.from_numpy(v) for k, v in make_inputs(cfg, n_clips, seed=1).items()}
    sr = cfg["sr"]
    kind = "port"
    if want in ("auto", "reference"):
        import ref_loader
        if ref_loader.available():
            kind = "reference"
        elif want == "reference":
            raise SystemExit("--cpu-impl reference: no copy of the reference (run oracle/make_ref.py in the build container)")
    if kind == "reference":
        import ref_loader
        import ref_step
        mu, mp, tt, _ = ref_loader.load()
        if cfg["kind"] == "eval":
            fe = ref_loader.fake_feature_extractor(sr)
            first = [mp.to_one_hot(x["annotation"][b], cfg["K"]).unsqueeze(0) for b in range(n_clips)]
            R = cfg["out_res"]

            def body():
                for b in range(n_clips):                                                  # mask_propagation.py:819-824
                    maps = torch.stack(mp.propagate_labels(cfg["n_last"], cfg["radius"], cfg["topk"], fe, x["backbone"][b],
                                                           first[b], True), dim=0)
                    up = torch.nn.functional.interpolate(maps, size=(R, R), mode="bilinear", align_corners=False)
                    up.max(dim=1)
        else:
            head = torch.stack([x["head_src"], x["head_tgt"]], 1)
            model = ref_step.build_timet(head, x["backbone"], x["prototypes"], sr)

            def body():
                ref_step.ff_sinkhorn_step(model, x["head_src"], x["head_tgt"], x["backbone"], cfg["n_last"], cfg["radius"],
                                          cfg["topk"], EPSILON, ITERS)
        sample = ("{n} clip(s) of the workload per step through the reference's OWN code (baseline/_ref copy of "
                  "time_tuning.py / mask_propagation.py / my_utils.py, unmodified) on the CPU, {thr} torch threads; the "
                  "neighbourhood mask (a 4-deep Python loop, mask_propagation.py:377-391) is built once in the untimed "
                  "warm-up and cached by the reference itself (:473-476); clips/s is linear in clips (the reference loops "
                  "over clips, time_tuning.py:277)")
    else:
        import timet_oracle_torch as OT
        mask = OT.window_mask(sr, cfg["radius"])
        if cfg["kind"] == "eval":
            import timet_oracle as O
            first = [torch.from_numpy(O.to_one_hot(x["annotation"][b].numpy(), cfg["K"])).unsqueeze(0) for b in range(n_clips)]
            R = cfg["out_res"]

            def body():
                for b in range(n_clips):
                    maps = torch.stack(OT.propagate_labels(cfg["n_last"], cfg["radius"], cfg["topk"], sr, x["backbone"][b],
                                                           first[b], mask))
                    up = torch.nn.functional.interpolate(maps, size=(R, R), mode="bilinear", align_corners=False)
                    up.max(dim=1)
        else:
            def body():
                OT.ff_sinkhorn_step(x["head_src"], x["head_tgt"], x["backbone"], x["prototypes"], sr, cfg["n_last"],
                                    cfg["radius"], cfg["topk"], EPSILON, ITERS, mask)
        sample = ("{n} clip(s) of the workload per step, torch-CPU PORT of the reference's dense per-clip path (same ATen "
                  "operator sequence as mask_propagation.py:418-444 / my_utils.py:246-274, bit-identical to the reference "
                  "on the golden fixtures), {thr} threads; neighbourhood-mask build excluded; clips/s is linear in clips")

    def run():
        t0 = time.perf_counter()
        body()
        return time.perf_counter() - t0
    return run, thr, kind, sample.format(n=n_clips, thr=thr)


def default_cpu_clips(cfg):
    # about 10-30 s of CPU work for warm-up + the timed repetitions
    return {1: 2, 2: 8, 3: 8, 4: 1, 5: 1}[cfg_number(cfg)]


def cfg_number(cfg):
    return next(k for k, v in CONFIGS.items() if v["name"] == cfg["name"])


def cpu_sample_cfg(cfg):
    """The bounded CPU sample of the eval / ViT-B workloads uses fewer frames (stated in `sample`)."""
    c = dict(cfg)
    if cfg_number(cfg) == 4:
        c["fs"] = 6        # first 6 of the 80 frames (the reference's 60x60 r12 mask build alone takes > 10 s, untimed)
    if cfg_number(cfg) == 5:
        c["fs"] = 4
    return c


def run_reference_arm(args, rank):
    if rank != 0:
        return
    cfg = resolve(args, 1)
    ccfg = cpu_sample_cfg(cfg)
    n = args.cpu_clips or default_cpu_clips(cfg)
    run, thr, kind, sample = _cpu_setup(ccfg, n, args.cpu_impl)
    warm = max(1, min(args.warmup, 2))
    for _ in range(warm):
        run()
    steps = max(1, min(args.steps, 10))
    times = [run() for _ in range(steps)]
    t = sum(times) / len(times)
    # frames scale the work linearly too (one label_propagation per target frame); normalise a shortened sample
    scale = (ccfg["fs"] - 1) / (cfg["fs"] - 1)
    val = n / t * scale
    if ccfg["fs"] != cfg["fs"]:
        sample += f"; sample clips have {ccfg['fs']} of the {cfg['fs']} frames, clips/s scaled by {scale:.4f} (work is at least linear in target frames: the first frames have fewer contexts, so this over-states the CPU path)"
    line = {"impl": "reference", "metric": "FF+Sinkhorn clips/s", "value": val, "unit": "clips/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"], "sample_clips": n},
            "cpu_baseline": {"value": val, "unit": "clips/s", "cores": thr, "kind": kind, "sample": sample},
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
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- multi-GPU parity (before timing)
def _sinkhorn_fp64(scores, eps, iters):
    """fp64 evaluation of my_utils.sinkhorn on the GLOBAL batch (scaling-vector form) on the GPU.
    A checker for dist_parity, written out here: bench.py's GPU arm does not import oracle/."""
    import torch
    E = torch.exp(scores.double() / eps)
    B, K = E.shape
    b = torch.ones(B, dtype=torch.float64, device=E.device)
    a = torch.ones(K, dtype=torch.float64, device=E.device)
    for _ in range(iters):
        a = (1.0 / K) / (E * b[:, None]).sum(0)
        b = (1.0 / B) / (E * a[None]).sum(1)
    Q = E * a[None]
    return Q / Q.sum(1, keepdim=True)


def dist_parity(rank, world, dev):
    """The checks of tests/dist_check.py, run by every bench at N > 1 so that the driver's scaling runs carry them."""
    import torch
    import torch.distributed as dist
    import timetuning_b200 as tb
    from timetuning_b200 import dist as tdist, synth
    res = {"cases": [], "max_abs_err": 0.0, "outside_tol": 0}
    cases = []
    if world == 2:      # produced by the reference itself under 2 gloo ranks (oracle/make_golden.py)
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", "sinkhorn_ws2_b512_k200.npz")))
        cases.append(("reference-2-rank-fixture", g["scores"], float(g["epsilon"]), int(g["iters"]), g["q"]))
    cases.append(("cfg3-slice (resident kernel)", synth.cosine_scores(4 * 784 * world, 200, seed=123), 0.05, 10, None))
    # 64 clips per rank (BASELINE configs[2] at 4 GPUs): rows beyond shared memory -> the hybrid one-launch kernel
    cases.append(("cfg3 64 clips per rank (hybrid kernel)", synth.cosine_scores(64 * 784 * world, 200, seed=124), 0.05, 10, None))
    for name, sc, eps, iters, want in cases:
        full_scores = torch.from_numpy(sc).to(dev)
        rows = tdist.shard_range(sc.shape[0], rank, world)
        q = tb.sinkhorn_from_scores(full_scores[rows.start:rows.stop].contiguous(), eps, iters, world_size=world)
        gathered = [torch.empty_like(q) for _ in range(world)]
        dist.all_gather(gathered, q)
        full = torch.cat(gathered).double()
        ref = torch.from_numpy(want).to(dev).double() if want is not None else _sinkhorn_fp64(full_scores, eps, iters)
        err = (full - ref).abs()
        bad = int((err > 1e-4 + 1e-5 * ref.abs()).sum().item())
        single = tb.sinkhorn_from_scores(full_scores, eps, iters, world_size=1).double()      # same kernels, one GPU, global batch
        err1 = float((full - single).abs().max().item())
        res["cases"].append({"case": name, "rows": int(sc.shape[0]), "max_abs_err": float(err.max().item()), "outside_tol": bad,
                             "max_abs_vs_single_gpu_global_batch": err1})
        res["max_abs_err"] = max(res["max_abs_err"], float(err.max().item()))
        res["outside_tol"] += bad + int(err1 > 1e-5)
    # clips sharded by rank == one process doing all clips (bit for bit)
    bs_local, fs, sr, D, C = 2, 4, 14, 64, 8
    feats = torch.from_numpy(synth.clip_features(bs_local * world, fs, sr, D, seed=7)).to(dev)
    first = torch.from_numpy(np.stack([synth.soft_labels(sr * sr, C, seed=30 + b) for b in range(bs_local * world)])).to(dev)
    clips = tdist.shard_range(bs_local * world, rank, world)
    lab, _ = tb.propagate_labels_batched(feats[clips.start:clips.stop].contiguous(), first[clips.start:clips.stop].contiguous(), 7, 6, 5)
    gl = [torch.empty_like(lab) for _ in range(world)]
    dist.all_gather(gl, lab)
    ref, _ = tb.propagate_labels_batched(feats, first, 7, 6, 5)
    res["ff_sharded_equal"] = bool(torch.equal(torch.cat(gl), ref))
    ok = torch.tensor([1 if (res["outside_tol"] == 0 and res["ff_sharded_equal"]) else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    res["ok"] = bool(ok.item())
    return res


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
    from timetuning_b200.step import HostStepPipeline, StepRunner

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    parity = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        tdist.init_comm()
        parity = dist_parity(rank, world, dev)
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"error": "multi-GPU parity check failed", "dist_parity": parity}), flush=True)
            tdist.destroy_comm()
            dist.destroy_process_group()
            raise SystemExit(3)
    cfg = resolve(args, world)
    engine = {"auto": ops.FF_AUTO, "exact": ops.FF_EXACT, "tc": ops.FF_TC}[args.engine]
    bs, fs, sr, D, K = cfg["clips_per_gpu"], cfg["fs"], cfg["sr"], cfg["dim"], cfg["K"]
    N = sr * sr
    is_eval = cfg["kind"] == "eval"
    x = make_inputs(cfg, bs, seed=1 + rank)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    runner = StepRunner(bs, fs, sr, D, cfg["head_dim"], K, cfg["n_last"], cfg["radius"], cfg["topk"], EPSILON, ITERS,
                        world_size=world, engine=engine, overlap=args.overlap, device=dev) if not is_eval else None
    host = {k: torch.from_numpy(v).pin_memory() for k, v in x.items() if k != "prototypes"}
    d_in = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    d_pr = torch.from_numpy(x["prototypes"]).to(dev) if not is_eval else None
    h2d_bytes = sum(v.numel() * v.element_size() for k, v in host.items())

    if is_eval:
        R = cfg["out_res"]
        first = [ops.one_hot_first_seg(d_in["annotation"][b], K) for b in range(bs)]
        plan = ops._plan(1, fs, sr, sr, D, K, cfg["n_last"], cfg["radius"], cfg["topk"], device=dev)

        def dev_step(record=None):
            outs = []
            for b in range(bs):
                outs.append(ops.propagate_labels_eval(cfg["n_last"], cfg["radius"], cfg["topk"], d_in["backbone"][b], first[b], R,
                                                      engine=engine, events=record))
            return outs
        d2h_bytes = bs * (fs - 1) * R * R * 8
    else:
        plan = runner.plan

        def dev_step(record=None):
            return runner.run(d_in["head_src"], d_in["head_tgt"], d_in["backbone"], d_pr, events=record)
        d2h_bytes = bs * N * 8
    engine_used = "tcgen05" if (engine != ops.FF_EXACT and plan.tc_supported) else "exact-fp32"

    # ---- value arm: inputs resident in HBM
    for _ in range(max(args.warmup, 3)):
        dev_step()
    sync_all()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    n0 = _cabi.launch_count()
    t0, t1 = ev(), ev()
    records = []
    sync_all()
    t0.record()
    for _ in range(args.steps):
        rec = {}
        dev_step(rec)
        records.append(rec)
    t1.record()
    sync_all()
    launches = _cabi.launch_count() - n0
    ms_total = torch.tensor([t0.elapsed_time(t1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = ms_total.item() / args.steps
    value = bs * world / (ms_step * 1e-3)

    # ---- e2e arm: pinned host inputs, H2D + D2H inside the timed region, through the public API
    e2e = None
    if not args.no_e2e:
        if is_eval:
            d_feats = torch.empty_like(d_in["backbone"])
            d_ann = torch.empty_like(d_in["annotation"])
            out_host = torch.empty((bs, fs - 1, cfg["out_res"], cfg["out_res"]), dtype=torch.int64).pin_memory()

            def e2e_step():
                d_feats.copy_(host["backbone"], non_blocking=True)
                d_ann.copy_(host["annotation"], non_blocking=True)
                for b in range(bs):
                    pred = ops.propagate_labels_eval(cfg["n_last"], cfg["radius"], cfg["topk"], d_feats[b],
                                                     ops.one_hot_first_seg(d_ann[b], K), cfg["out_res"], engine=engine)
                    out_host[b].copy_(pred, non_blocking=True)
        else:
            pipe = HostStepPipeline(bs, fs, N, D, cfg["head_dim"], K, chunks=4, device=dev)

            def e2e_step():
                pipe.run(host["head_src"], host["head_tgt"], host["backbone"], d_pr, cfg["n_last"], cfg["radius"], cfg["topk"],
                         EPSILON, ITERS, world, engine)
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
        mine = e2e_ms.item() / args.steps
        if world > 1:
            dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        e2e_step_ms = e2e_ms.item() / args.steps
        per_rank = torch.tensor([h2d_bytes / (mine * 1e-3) / 1e9], device=dev)
        per_rank_all = [torch.empty_like(per_rank) for _ in range(world)] if world > 1 else [per_rank]
        if world > 1:
            dist.all_gather(per_rank_all, per_rank)
        e2e = {"value": bs * world / (e2e_step_ms * 1e-3), "unit": "clips/s", "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": d2h_bytes, "h2d_gbs_per_rank": [round(float(t.item()), 2) for t in per_rank_all],
               "bound": "PCIe host->device copy of the fp32 backbone features (compute is overlapped chunk-wise)"}
    clock_info = clocks.stop() if rank == 0 else None

    # ---- per-stage device times (CUDA events recorded inside the timed region, each on the stream its stage runs on)
    def mean_ms(a, b):
        v = [r[a].elapsed_time(r[b]) for r in records if a in r and b in r]
        return sum(v) / len(v) if v else None
    stage_ms = {}
    for name, (a, b) in {"cosine_scores_x2": ("scores0", "scores1"), "sinkhorn_x2": ("scores1", "sinkhorn1"),
                         "prepare": ("prep0", "prep1"), "select": ("prep1", "select1"), "select_tc_kernel": ("tc0", "tc1"),
                         "gather": ("gather0", "gather1"), "eval_tail": ("gather1", "tail1")}.items():
        v = mean_ms(a, b)
        if v is not None:
            stage_ms[name] = v
    st = plan.stats()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sigma_ctx = sum(1 + (t - max(1, t - cfg["n_last"])) for t in range(1, fs))
        clips_per_launch = 1 if is_eval else bs
        dense_flops = 2.0 * N * N * D * sigma_ctx * clips_per_launch                    # SURVEY.md §8d, per launch
        # peak: burst figure when the SM clock stayed at its maximum during the timed region (short step, no power cap),
        # else the sustained one (B200_PROFILING.md)
        at_max = bool(clock_info and clock_info.get("sm_mhz") and clock_info.get("sm_max_mhz")
                      and clock_info["sm_mhz"] >= 0.97 * clock_info["sm_max_mhz"] and "sw_power_cap" not in clock_info["reasons"])
        if peaks:
            peak = peaks["bf16_tflops"] if at_max else peaks["bf16_tflops_sustained"]
            peak_src = ("MEASURED_PEAKS.json bf16_tflops (burst: SM clock at max, no power cap during the timed region)" if at_max
                        else "MEASURED_PEAKS.json bf16_tflops_sustained (clock below max / power cap during the timed region)")
        else:
            peak, peak_src = 1590.0, "fallback 1.59 PFLOP/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"
        traffic, traffic_src = ncu_traffic("ff_tc_persist_kernel", args.config)
        if engine_used == "tcgen05" and "select_tc_kernel" in stage_ms:
            tc_ms = stage_ms["select_tc_kernel"]
            exec_flops = plan.tc_executed_flops
            roof = {"bound": "tensor", "kernel": "ff_tc_persist_kernel (tcgen05 affinity + fused window / top-k nomination), "
                    "timed alone with CUDA events recorded by the library around its launch",
                    "achieved": exec_flops / (tc_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s", "peak_source": peak_src + " (of measured)",
                    "flops_model": "EXECUTED: key tiles issued inside the window band, M padded to 128 (timet_ff_tc_executed_flops) "
                                   "= what the tensor pipe runs; SURVEY.md §8d %TC",
                    "executed_gflop_per_launch": exec_flops / 1e9, "dense_gflop_per_launch": dense_flops / 1e9,
                    "executed_over_dense": exec_flops / dense_flops,
                    "dense_equiv_tflops": dense_flops / (tc_ms * 1e-3) / 1e12, "ms_per_launch": tc_ms,
                    "traffic": traffic, "traffic_source": traffic_src}
        else:
            sel_ms = stage_ms.get("select", float("nan"))
            roof = {"bound": "tensor", "kernel": "ff_select_exact (fp32 CUDA-core scan; tensor-core engine not used)",
                    "achieved": dense_flops / (sel_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s", "peak_source": peak_src,
                    "flops_model": "dense 2*N^2*D*sum_ctx per clip (SURVEY.md §8d)", "traffic": None}
        roof["frac"] = roof["achieved"] / roof["peak"]
        hbm = peaks.get("hbm_gbs", 6650.0)
        extra = {}
        if "sinkhorn_x2" in stage_ms:
            B = bs * N
            sk_bytes = 2 * (ITERS + 2) * B * K * 4
            extra["sinkhorn"] = {"bound": "hbm", "achieved": sk_bytes / (stage_ms["sinkhorn_x2"] * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                 "bytes_model": "(iters+2)*B*K*4 per call, 2 calls (streaming model, SURVEY.md §8d); the resident "
                                                "kernel's real DRAM traffic is the compulsory 2*B*K*4",
                                 "compulsory_gbs": 2 * 2 * B * K * 4 / (stage_ms["sinkhorn_x2"] * 1e-3) / 1e9,
                                 "note": "second stream under the Feature-Forwarding kernels" if args.overlap else "same stream"}
        if "gather" in stage_ms:
            gather_bytes = clips_per_launch * ((fs - 1) * N * K * 4 + N * K * 4)
            extra["gather"] = {"bound": "hbm", "achieved": gather_bytes / (stage_ms["gather"] * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                               "bytes_model": "compulsory (fs-1)*N*C*4 write + N*C*4 read per clip"}
        if "prepare" in stage_ms:
            prep_bytes = clips_per_launch * fs * N * (D * 4 + D * 2)
            extra["prepare"] = {"bound": "hbm", "achieved": prep_bytes / (stage_ms["prepare"] * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                "bytes_model": "read 4*D, write 2*D (fp16 operand) + 4 (inverse norm) per row"}
        for v in extra.values():
            v["frac"] = v["achieved"] / v["peak"]
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ccfg = cpu_sample_cfg(cfg)
            n = args.cpu_clips or default_cpu_clips(cfg)
            run, thr, kind, sample = _cpu_setup(ccfg, n, args.cpu_impl)
            run()
            t = min(run() for _ in range(3))
            scale = (ccfg["fs"] - 1) / (cfg["fs"] - 1)
            if ccfg["fs"] != cfg["fs"]:
                sample += (f"; sample clips have {ccfg['fs']} of the {cfg['fs']} frames, clips/s scaled by {scale:.4f} (the first frames "
                           "have fewer contexts, so this over-states the CPU path)")
            cpu = {"value": n / t * scale, "unit": "clips/s", "cores": thr, "kind": kind, "sample": "best of 3 after 1 warm-up; " + sample}
        if world > 1:
            mode = runner.sinkhorn_mode if not is_eval else "none"
            launch = {"resident": "one resident launch per call", "hybrid": "one hybrid launch per call (rows partly re-read from L2 / HBM)",
                      "streaming": "one launch per pass"}.get(mode, mode)
            if not is_eval and runner.sinkhorn_pair_mode == "dual" and not args.overlap:
                launch = "ONE launch for both calls of the step, side by side on half of the SMs each, one exchange channel each"
            if mode == "streaming" or not ops._comm.get("p2p"):
                sk_path = f"nccl (one ncclAllReduce of K floats per Sinkhorn pass on the library's own communicator; {launch})"
            else:
                sk_path = f"p2p (in-kernel NVLink peer stores + flags; {launch})"
            par = f"clips sharded over {world} GPUs, no data-path collective for Feature-Forwarding; Sinkhorn K-vector marginals: {sk_path}"
        else:
            sk_path, par = "single", "single GPU"
        line = {"metric": "FF+Sinkhorn clips/s", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": cfg["scaling"],
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": cfg["name"], "survey_config": args.config, "clips_per_gpu": bs, "global_clips": bs * world,
                           "engine": engine_used,
                           "precision": "all results f32; the tcgen05 kernel nominates top-k candidates from fp16 inputs / fp32 "
                                        "accumulation and every nominated key is re-evaluated in f32 (bit-identical to the f32 engine)"
                           if engine_used == "tcgen05" else "f32",
                           "l2": "inputs larger than L2 (backbone features %.0f MB per step)" % (d_in["backbone"].numel() * 4 / 1e6),
                           "streams": "cosine scores + Sinkhorn on a second stream, joined before the gather" if (args.overlap and not is_eval) else "one stream",
                           "parallelism": par},
                "e2e": e2e, "gpu_launches": launches, "roofline": roof, "stage_ms": stage_ms, "stage_roofline": extra,
                "ff_stats": st, "clocks": clock_info, "cpu_baseline": cpu, "sinkhorn_path": sk_path}
        if parity is not None:
            line["dist_parity"] = parity
        print(json.dumps(line), flush=True)
    if world > 1:
        tdist.destroy_comm()
        dist.destroy_process_group()


def ncu_traffic(kernel, config):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full summary (profiles/ncu_kernels.json, written by
    profiles/extract_ncu.py); only valid for the workload it was captured on."""
    path = os.path.join(ROOT, "profiles", "ncu_kernels.json")
    try:
        data = json.load(open(path))
    except Exception:
        return None, "no committed ncu summary"
    for row in data.get("kernels", []):
        if kernel in row["kernel"] and row.get("survey_config") == config:
            return row["dram_bytes_read"] + row["dram_bytes_write"], (
                f"ncu --set full, profiles/ncu_kernels.json ({row.get('capture', '?')}): dram read "
                f"{row['dram_bytes_read'] / 1e6:.1f} MB + write {row['dram_bytes_write'] / 1e6:.1f} MB per launch")
    return None, f"no ncu capture of {kernel} for config {config}"


if __name__ == "__main__":
    main()
