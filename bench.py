#!/usr/bin/env python
"""Benchmark of the dual-encoder contrastive pre-training step (BASELINE.json metric: clips/sec, TSF-B 16f x 224^2).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...   # reference algorithm on the host CPU cores (oracle port)

One "step" = the body of the reference's train() loop (main_pretrain.py:486-530) on one synthetic batch:
zero_grad -> model(frames, tokens, norm_embed=True) -> CLIPLoss -> backward -> AdamW.step -> logit_scale clamp.
`value`  : clips/s with the batch already resident in HBM when the timed region starts (CUDA events, max over ranks).
`e2e`    : the same loop through the public API with HOST inputs: per step an H2D copy of the pinned batch and a D2H
           read of the loss are inside the timed region.
`roofline`: tensor-pipe fraction of the dominant kernel family (the tcgen05 GEMM), measured with CUDA events around
           every GEMM launch during extra instrumented steps (same workload), against MEASURED_PEAKS.json.
`cpu_baseline`: the oracle port (oracle/dual_encoder.py, reference algorithm in fp32 PyTorch) timed on the host cores.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "clips/sec dual-encoder pretrain TSF-B 16f x 224^2"
FALLBACK_PEAK_TFLOPS = 1400.0   # B200_PROFILING.md: sustained cuBLAS bf16 under the 1 kW cap ("of fallback")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager"])
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU (BASELINE config: 64)")
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--model", default="base", choices=["base", "large", "large336"],
                    help="base = CLIP_OPENAI_TIMESFORMER_BASE (the BASELINE metric); large / large336 = the TSF-L/14 factories "
                         "(BASELINE config 5 is large336 at --frames 32), ours arm only")
    ap.add_argument("--use-checkpoint", action="store_true",
                    help="model(..., use_checkpoint=True): recompute each SpaceTimeBlock in backward (main_pretrain.py --use-checkpoint)")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------- FLOP model
def flops_per_clip_train(T=16, n=196, D=768, depth=12, p=16, E=256, L=77, W=512, layers=12):
    N = 1 + T * n
    f_blk = 32 * N * D * D + 4 * D * ((T * n) * (T + n + 2) + 2 * N)
    f_vis = depth * f_blk + 2 * (T * n) * D * 3 * p * p + 2 * D * E
    f_txt = layers * (24 * L * W * W + 4 * L * L * W) + 2 * W * E
    return 3.0 * (f_vis + f_txt)


# --------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------- model / data
def randomise_zero_init(model, seed=0):
    """Zero-init trap: time attention / temporal embedding start at exactly 0 in the reference; randomise so
    that the time path does real work (same arithmetic cost either way, but keeps values finite and non-degenerate)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "timeattn" in n or "temporal_embed" in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
            elif "alpha_" in n:
                p.fill_(0.5)


MODELS = {   # factory, image size, kwargs of flops_per_clip_train
    "base": ("CLIP_OPENAI_TIMESFORMER_BASE", 224, dict(n=196, D=768, depth=12, p=16, W=512)),
    "large": ("CLIP_OPENAI_TIMESFORMER_LARGE", 224, dict(n=256, D=1024, depth=24, p=14, W=768)),
    "large336": ("CLIP_OPENAI_TIMESFORMER_LARGE_336PX", 336, dict(n=576, D=1024, depth=24, p=14, W=768)),
}


def make_batch(batch, frames, seed, size=224):
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 3, frames, size, size, generator=g)
    text = torch.zeros(batch, 77, dtype=torch.int64)
    for b in range(batch):
        ln = int(torch.randint(4, 21, (1,), generator=g))
        text[b, 0] = 49406
        text[b, 1:1 + ln] = torch.randint(1, 49406, (ln,), generator=g)
        text[b, 1 + ln] = 49407
    return x, text


def param_groups(model):
    """main_pretrain.py:199-213."""
    wd, nwd = [], []
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (nwd if (p.ndim < 2 or 'bias' in n or 'ln' in n or 'bn' in n) else wd).append(p)
    return [{"params": wd, "weight_decay": 0.01}, {"params": nwd, "weight_decay": 0.0}]


# --------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))
    from lavila_b200 import _lib, ops
    from lavila_b200.models import models as M
    from lavila_b200.models.loss import CLIPLoss

    torch.manual_seed(0)
    factory, img_size, flop_kw = MODELS[args.model]
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):     # the factories print like the reference's; stdout carries ONE JSON line
        model = getattr(M, factory)(num_frames=args.frames, project_embed_dim=256)
    randomise_zero_init(model)
    model.to(dev)
    crit = CLIPLoss(use_vissl=True, cache_labels=True, rank=rank, world_size=world)
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], bucket_cap_mb=200)
    opt = torch.optim.AdamW(param_groups(model), lr=3e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)

    B = args.batch
    frames_h, text_h = make_batch(B, args.frames, 1234 + rank, img_size)
    frames_h, text_h = frames_h.pin_memory(), text_h.pin_memory()
    frames_d, text_d = frames_h.to(dev), text_h.to(dev)

    def step(fr, tx):
        opt.zero_grad(set_to_none=True)
        out = net(fr, tx, use_checkpoint=args.use_checkpoint, norm_embed=True)
        ld = crit(out)
        ld["loss"].backward()
        opt.step()
        model.logit_scale.data.clamp_(0, 4.6052)
        return ld["loss"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(k, host_inputs):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        e0.record()
        last = None
        for _ in range(k):
            if host_inputs:
                fr, tx = frames_h.to(dev, non_blocking=True), text_h.to(dev, non_blocking=True)
                last = float(step(fr, tx).item())           # D2H read of the step's loss, every step
            else:
                last = step(frames_d, text_d)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), _lib.launch_count() - n0, last

    for _ in range(max(3, args.warmup)):
        loss = step(frames_d, text_d)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(loss)), "non-finite loss in warm-up"

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, launches, loss = timed(args.steps, host_inputs=False)
    clocks = sampler.stop() if sampler else None
    ms_step = ms / args.steps
    value = B * world / (ms_step / 1e3)

    e2e = None
    if not args.no_e2e:
        timed(1, host_inputs=True)
        ms_e, _, _ = timed(args.steps, host_inputs=True)
        e2e = {"value": B * world / (ms_e / args.steps / 1e3), "unit": "clips/s",
               "h2d_bytes_per_step": int(frames_h.numel() * 4 + text_h.numel() * 8) * world,
               "d2h_bytes_per_step": 4 * world}

    roof = None
    if not args.no_roofline:
        # every rank runs the instrumented steps (DDP's gradient all-reduce is collective); rank 0 reports its own GEMMs
        roof = gemm_roofline(lambda: step(frames_d, text_d), ops, torch)
    barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.model == "base":
        cpu = cpu_baseline(args.frames, budget_s=25.0)

    if rank == 0:
        fl = flops_per_clip_train(T=args.frames, **flop_kw)
        line = {
            "metric": METRIC if args.model == "base" and args.frames == 16 else
            "clips/sec dual-encoder pretrain %s %df x %d^2" % (factory, args.frames, img_size),
            "value": round(value, 3), "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": round(ms_step, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "%s dual-encoder pretrain step (fwd + CLIPLoss + bwd + AdamW), "
                                   "%d frames x %d^2, batch %d per GPU" % (factory, args.frames, img_size, B),
                       "global_batch": B * world, "per_gpu_batch": B, "parallelism": "dp%d" % world,
                       "use_checkpoint": bool(args.use_checkpoint),
                       "l2": "inputs (%.0f MB/step) and activations exceed the 126 MB L2; no explicit flush" % (frames_h.numel() * 4 / 1e6),
                       "weights": "random init (no network for checkpoints)",
                       "model_tflop_per_clip": round(fl / 1e12, 4),
                       "model_tflops_achieved": round(value * fl / 1e12 / world, 1)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "loss": float(loss), "max_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def gemm_roofline(step_fn, ops, torch):
    """Instrument every tcgen05 GEMM launch of two extra steps with CUDA events on the launching stream."""
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", FALLBACK_PEAK_TFLOPS))
    which = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback (B200_PROFILING.md)"
    rec = []
    orig = ops.gemm

    def timed_gemm(A, B, M, N, K, out, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(A, B, M, N, K, out, **kw)
        e1.record()
        rec.append((e0, e1, 2.0 * M * N * K))
        return r

    ops.gemm = timed_gemm
    try:
        step_fn()
        rec.clear()
        step_fn()
        torch.cuda.synchronize()
    finally:
        ops.gemm = orig
    tot_ms = sum(a.elapsed_time(b) for a, b, _ in rec)
    tot_fl = sum(f for _, _, f in rec)
    ach = tot_fl / (tot_ms / 1e3) / 1e12
    traffic, traffic_note = None, None
    try:   # DRAM bytes per launch from the committed ncu --set full capture of the same kernel (profiles/)
        t = json.load(open(os.path.join(ROOT, "profiles", "gemm_traffic_r01.json")))
        traffic, traffic_note = t["mean_dram_bytes_per_launch"], t["source"]
    except Exception:
        pass
    return {"bound": "tensor", "kernel": "lv::gemm2::gemm2_bf16_kernel (tcgen05 cta_group::2, all %d GEMM launches of one step)" % len(rec),
            "achieved": round(ach, 1), "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 4), "traffic": traffic,
            "traffic_source": traffic_note, "flops_per_launch": tot_fl / max(1, len(rec)),
            "peak_source": which, "flops_per_step": tot_fl, "gemm_ms_per_step": round(tot_ms, 3)}


# --------------------------------------------------------------------------------------------- CPU arms (oracle port)
def oracle_step_fn(frames):
    import torch
    from oracle import dual_encoder as O
    cfg = O.tsf_base_config(num_frames=frames)
    params = {k: v.clone().requires_grad_(True) for k, v in O.init_params(cfg, seed=0).items()}
    opt = torch.optim.AdamW([{"params": list(params.values())}], lr=3e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)

    def step(batch, seed):
        x, text = O.synthetic_batch(cfg, batch, seed=seed, frames=frames)
        opt.zero_grad(set_to_none=True)
        out = O.clip_forward(x, text, params, cfg, norm_embed=True)
        loss = O.clip_loss(out["image_embed"], out["text_embed"], out["logit_scale"])["loss"]
        loss.backward()
        opt.step()
        return float(loss)

    return step


def host_threads():
    """Threads the CPU arms may use: scheduler affinity, capped by the cgroup CPU quota and by 32 (the oracle's fp32
    GEMMs stop scaling past that, and oversubscribing a shared host is catastrophic: 250 s/clip was observed)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(float(q) / float(per))))
    except Exception:
        pass
    return max(1, min(n, 32))


def cpu_baseline(frames, budget_s=25.0):
    import torch
    cores = host_threads()
    torch.set_num_threads(cores)
    step = oracle_step_fn(frames)
    t0 = time.time()
    step(1, 1)                                  # calibrate on one clip (also warms the thread pool)
    t1 = time.time() - t0
    batch = int(max(1, min(4, budget_s // max(t1, 1e-3) - 1)))
    t0 = time.time()
    step(batch, 2)
    dt = time.time() - t0
    return {"value": round(batch / dt, 4), "unit": "clips/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "1 fp32 training step (fwd+CLIPLoss+bwd+AdamW) of the oracle port at batch %d, %d frames, after a "
                      "1-clip calibration step (%.1f s)" % (batch, frames, t1)}


def run_reference(args):
    """Reference arm: the reference's algorithm (oracle port, fp32 PyTorch on the host cores) on the same config/metric;
    each step is a bounded sample (small batch) so the whole run ends within minutes.  Rank 0 only."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch
    cores = host_threads()
    torch.set_num_threads(cores)
    step = oracle_step_fn(args.frames)
    t0 = time.time()
    step(1, 0)
    t1 = time.time() - t0
    total = args.steps + args.warmup
    batch = int(max(1, min(4, 150.0 / max(total * t1, 1e-3))))
    for i in range(args.warmup):
        step(batch, 10 + i)
    t0 = time.time()
    for i in range(args.steps):
        step(batch, 100 + i)
    dt = (time.time() - t0) / max(1, args.steps)
    v = batch / dt
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": "clips/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "CLIP_OPENAI_TIMESFORMER_BASE dual-encoder pretrain step (fwd + CLIPLoss + bwd + AdamW), "
                                   "%d frames x 224^2; bounded sample: batch %d per step on the host CPU" % (args.frames, batch),
                       "global_batch": batch, "parallelism": "cpu"},
            "cpu_baseline": {"value": round(v, 4), "unit": "clips/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "%d steps of batch %d" % (args.steps, batch)},
            "e2e": {"value": round(v, 4), "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_eager(args):
    """Not part of the driver contract: the oracle port run as PyTorch-eager on the GPU under bf16 autocast -- the
    'reference PyTorch-eager on 1 B200' baseline of BASELINE.md (B1), since /root/reference cannot travel."""
    import torch
    from oracle import dual_encoder as O
    dev = torch.device("cuda", 0)
    cfg = O.tsf_base_config(num_frames=args.frames)
    params = {k: v.to(dev).requires_grad_(True) for k, v in O.init_params(cfg, seed=0).items()}
    opt = torch.optim.AdamW([{"params": list(params.values())}], lr=3e-5, weight_decay=0.01)
    B = args.batch
    x, text = make_batch(B, args.frames, 1234)
    x, text = x.to(dev), text.to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = O.clip_forward(x, text, params, cfg, norm_embed=True)
            loss = O.clip_loss(out["image_embed"], out["text_embed"], out["logit_scale"])["loss"]
        loss.backward()
        opt.step()
        return loss

    for _ in range(max(2, args.warmup)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"impl": "eager-port", "metric": METRIC, "value": round(B / (ms / 1e3), 3), "unit": "clips/s",
                      "ms_per_step": round(ms, 2), "batch": B, "dtype": "bf16 autocast", "loss": float(loss),
                      "max_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "eager":
        run_eager(a)
    else:
        run_ours(a)
