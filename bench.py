#!/usr/bin/env python
"""Benchmark of the dual-encoder contrastive pre-training step (BASELINE.json metric: clips/sec, TSF-B 16f x 224^2).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...   # the UNMODIFIED reference (baseline/_ref) on the host CPU cores
    python bench.py --impl eager [--batch B]                  # the UNMODIFIED reference as PyTorch-eager on the B200

One "step" = the body of the reference's train() loop (main_pretrain.py:486-530) on one synthetic batch:
zero_grad -> model(frames, tokens, norm_embed=True) -> CLIPLoss -> backward -> AdamW.step -> logit_scale clamp.
`value`  : clips/s with the batch already resident in HBM when the timed region starts (CUDA events, max over ranks).
`e2e`    : the same loop through the public API with HOST inputs: per step an H2D copy of the pinned batch and a D2H
           read of the loss are inside the timed region.
`roofline`: tensor-pipe fraction of the dominant kernel family (the tcgen05 GEMM), measured with CUDA events around
           every GEMM launch during extra instrumented steps (same workload), against MEASURED_PEAKS.json.
`block_roofline`: CUDA events around every SpaceTimeBlock forward and backward node (the unit the 0.5x target is defined on).
`eager_baseline`: the unmodified reference (baseline/_ref: its own factory, CLIP, CLIPLoss, train() body) run as PyTorch-eager
           under bf16 autocast on the same GPU, in the same process, before our arm (N = 1 only) -- the >= 6x target's denominator.
`cpu_baseline`: the same reference modules in fp32 on the host cores (kind "reference"); the oracle port only if
           baseline/_ref is absent (kind "port").
`loss_check` (N > 1): the fused NVLink gather + loss kernel against the NCCL all_gather route and fp32 torch on the same
           embeddings, before the timed region; a mismatch above 1e-5 fails the run.
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
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--no-narrator", action="store_true", help="skip the narrator leg (BASELINE config 4) of the default N = 1 run")
    ap.add_argument("--amp", default="bf16", choices=["bf16", "fp16"], help="--impl eager: autocast dtype (fp16 = the "
                    "reference's literal mode, torch.cuda.amp.autocast + GradScaler, main_pretrain.py:223,490)")
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
    eager = None
    if rank == 0 and world == 1 and not args.no_eager_baseline and args.model == "base":
        eager = eager_baseline(args, dev)          # measured first: the reference's eager step needs most of the HBM

    from lavila_b200 import _lib, engine, ops
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

    copy_stream = torch.cuda.Stream(device=dev)

    def h2d_async():
        with torch.cuda.stream(copy_stream):
            fr = frames_h.to(dev, non_blocking=True)
            tx = text_h.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return fr, tx, ev

    def timed(k, host_inputs):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        e0.record()
        last = None
        if host_inputs:
            # Every step's inputs come from pinned HOST memory and its loss is read back; both transfers are inside the
            # timed region.  The copy of step i+1's batch is enqueued on a copy stream before step i's loss is read, so it
            # overlaps compute the way a pin_memory DataLoader with a CUDA prefetcher does; step 0's copy is exposed.
            nxt = h2d_async()
            for i in range(k):
                fr, tx, ev = nxt
                torch.cuda.current_stream().wait_event(ev)
                loss_t = step(fr, tx)
                fr.record_stream(torch.cuda.current_stream())
                tx.record_stream(torch.cuda.current_stream())
                if i + 1 < k:
                    nxt = h2d_async()
                last = float(loss_t.item())                 # D2H read of the step's loss, every step
        else:
            for _ in range(k):
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

    loss_check = None
    if world > 1:
        loss_check = multi_gpu_loss_check(model, crit, frames_d, text_d, rank, world, dev)

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

    roof = block_roof = None
    if not args.no_roofline:
        # every rank runs the instrumented steps (DDP's gradient all-reduce is collective); rank 0 reports its own kernels
        roof = gemm_roofline(lambda: step(frames_d, text_d), ops, torch)
        block_roof = block_roofline(lambda: step(frames_d, text_d), engine, torch, B, args.frames, flop_kw)
    barrier()

    ddp_exposed = None
    if world > 1:
        # step time with and without DDP's gradient all-reduce (no_sync): names the limiter of the 1 -> N curve
        def step_nosync(fr, tx):
            with net.no_sync():
                return step(fr, tx)
        k = max(2, min(4, args.steps))
        ms_sync, _, _ = timed(k, host_inputs=False)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            step_nosync(frames_d, text_d)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step(frames_d, text_d)        # one synchronised step so every rank's gradients / weights agree again
        ddp_exposed = {"ms_per_step_sync": round(ms_sync / k, 3), "ms_per_step_no_sync": round(float(t.item()) / k, 3),
                       "ddp_allreduce_exposed_ms": round((ms_sync - float(t.item())) / k, 3), "steps": k}

    narr = inp = None
    if rank == 0 and world == 1 and not args.no_narrator and args.model == "base":
        # free the dual-encoder first: the narrator (TSF-L/14 + GPT-2 XL, 1.9 B parameters) is measured on an empty device
        peak_mem_gb = round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)
        del opt, net, model, crit, frames_d, text_d
        import gc
        gc.collect()
        engine.SHADOW.clear()
        torch.cuda.empty_cache()
        try:
            narr = narrator_leg(dev)
        except Exception as e:          # the headline metric must not be lost to the secondary leg
            narr = {"error": repr(e)[:300]}
        try:
            inp = input_pipeline_leg(dev)
        except Exception as e:
            inp = {"error": repr(e)[:300]}
    else:
        peak_mem_gb = round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)

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
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "block_roofline": block_roof,
            "cpu_baseline": cpu, "eager_baseline": eager, "loss": float(loss),
            "max_mem_gb": peak_mem_gb,
        }
        if narr is not None:
            line["narrator"] = narr
        if inp is not None:
            line["input_pipeline"] = inp
        if eager and eager.get("value"):
            line["vs_eager"] = {"device_timed": round(value / eager["value"], 3),
                                "e2e": round(e2e["value"] / eager["value"], 3) if e2e else None,
                                "note": "ours at batch %d vs the unmodified reference eager at batch %d (clips/s over clips/s)"
                                        % (B, eager["batch"])}
        if loss_check is not None:
            line["loss_check"] = loss_check
        if ddp_exposed is not None:
            line["ddp"] = ddp_exposed
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if loss_check is not None and not loss_check["ok"]:
        sys.exit(3)


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
        rec.append((e0, e1, 2.0 * M * N * K, (int(kw.get("a_mn", 0)), int(kw.get("b_mn", 0)), int(kw.get("flags", 0)), M, N, K)))
        return r

    ops.gemm = timed_gemm
    try:
        step_fn()
        rec.clear()
        step_fn()
        torch.cuda.synchronize()
    finally:
        ops.gemm = orig
    tot_ms = sum(a.elapsed_time(b) for a, b, _, _ in rec)
    tot_fl = sum(f for _, _, f, _ in rec)
    ach = tot_fl / (tot_ms / 1e3) / 1e12
    cls = {}
    for a, b, f, key in rec:
        c = cls.setdefault(key, [0, 0.0, 0.0])
        c[0] += 1
        c[1] += a.elapsed_time(b)
        c[2] += f
    by_class = [{"a_mn": k[0], "b_mn": k[1], "flags": k[2], "M": k[3], "N": k[4], "K": k[5], "launches": v[0], "ms": round(v[1], 3),
                 "tflops": round(v[2] / (v[1] / 1e3) / 1e12, 1)} for k, v in sorted(cls.items(), key=lambda kv: -kv[1][1])[:14]]
    traffic, traffic_note, traffic_extra = None, None, {}
    try:   # DRAM bytes per launch from the committed ncu --set full capture of the same kernel (profiles/)
        cand = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.startswith("gemm_traffic_r") and f.endswith(".json"))
        t = json.load(open(os.path.join(ROOT, "profiles", cand[-1])))
        traffic, traffic_note = t["mean_dram_bytes_per_launch"], t["source"]
        traffic_extra = {k: t[k] for k in ("kernel", "algorithmic_bytes_per_launch", "ratio") if k in t}
    except Exception:
        pass
    return {"bound": "tensor", "kernel": "lv::gemm2::gemm2_bf16_kernel (tcgen05 cta_group::2, all %d GEMM launches of one step)" % len(rec),
            "achieved": round(ach, 1), "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 4), "traffic": traffic,
            "traffic_source": traffic_note, "traffic_of": traffic_extra, "flops_per_launch": tot_fl / max(1, len(rec)),
            "peak_source": which, "flops_per_step": tot_fl, "gemm_ms_per_step": round(tot_ms, 3), "by_class": by_class}



def narrator_leg(dev, batch=32, frames=4, max_len=77):
    """BASELINE config 4 in the driver-run record: VCLM_OPENAI_TIMESFORMER_LARGE_GPT2_XL (TSF-L/14 224 px, 4 frames, GPT-2 XL with
    gated cross-attention every 2nd layer), batch 32, nucleus sampling p = 0.95, T = 0.7, 77 tokens, early_stopping off (fixed 76
    decoding steps), random weights.  One warm-up call (builds the KV buffers and captures the decoding step), one timed call,
    for 1 and 10 (the script default, main_infer_narrator.py:61) sequences per clip."""
    import contextlib
    import time as _t
    from types import SimpleNamespace
    import torch
    from lavila_b200.models import models as M
    tok = SimpleNamespace(bos_token_id=50256, eos_token_id=50256, pad_token_id=0)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(sys.stderr):
        model = M.VCLM_OPENAI_TIMESFORMER_LARGE_GPT2_XL(gated_xattn=True, num_frames=frames).to(dev).eval()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "alpha" in n:
                p.fill_(0.5)
            if "timeattn" in n or "temporal_embed" in n:
                p.normal_(0, 0.02)
    clips = torch.randn(batch, 3, frames, 224, 224, device=dev)
    out = {"workload": "VCLM_OPENAI_TIMESFORMER_LARGE_GPT2_XL encode_image + generate(top_p=0.95, temperature=0.7, max_text_length=%d, "
                       "early_stopping=False), %d frames x 224^2, batch %d, random weights" % (max_len, frames, batch), "runs": []}
    for R in (1, 10):
        def run():
            t = model.encode_image(clips)
            return model.generate(t, tok, max_text_length=max_len, top_p=0.95, temperature=0.7, num_return_sequences=R,
                                  early_stopping=False)
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = _t.time()
        run()                        # encode (encoder + pooling) then generate: the split is timed below
        e1.record()
        torch.cuda.synchronize()
        dt = e0.elapsed_time(e1) / 1e3
        # encoder share
        e0.record()
        tk = model.encode_image(clips)
        e1.record()
        torch.cuda.synchronize()
        enc = e0.elapsed_time(e1) / 1e3
        out["runs"].append({"num_return_sequences": R, "seconds": round(dt, 3), "clips_per_s": round(batch / dt, 2),
                            "sequences_per_s": round(batch * R / dt, 1), "encode_image_s": round(enc, 3),
                            "ms_per_decoding_step": round((dt - enc) / (max_len - 1) * 1e3, 2)})
    out["max_mem_gb"] = round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)
    del model
    torch.cuda.empty_cache()
    return out


def input_pipeline_leg(dev, batch=64, frames=16, src_hw=(288, 384), crop=224, cpu_budget_s=8.0):
    """the train transform of main_pretrain.py:263-272 for one batch of BASELINE config 2 (64 clips x 16 decoded
    uint8 frames of 288 x 384 -> fp32 [64, 3, 16, 224, 224]) as ONE lv_clip_transform launch, sources resident in HBM; `e2e` adds
    the H2D copy of the pinned uint8 frames and the crop-box table.  Algorithmic bytes per clip: 3*T*S*S*4 written + the crop box
    (<= T*H*W*3 source bytes) read once.  `cpu_baseline`: the reference's own chain (lavila Permute + torchvision + NormalizeVideo,
    baseline/_ref when installed) on ONE host core -- the reference spends 10 DataLoader workers per GPU on it."""
    import time as _t
    import torch
    from lavila_b200.data import GpuClipTransform
    H, W = src_hw
    g = torch.Generator().manual_seed(0)
    src_h = torch.randint(0, 256, (batch, frames, H, W, 3), generator=g, dtype=torch.uint8).pin_memory()
    src_d = src_h.to(dev)
    tf = GpuClipTransform(crop, "train", device=dev)
    torch.manual_seed(0)
    tf(src_d)
    boxes = list(tf.last_boxes)
    box_bytes = sum(b[2] * b[3] for b in boxes) * 3 * frames
    out_bytes = batch * 3 * frames * crop * crop * 4
    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=dev)          # > the 126 MB L2
    ms = []
    prepared = tf.prepare(src_d, boxes=boxes)
    for _ in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tf.run(*prepared)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    k_ms = sorted(ms[1:])[len(ms[1:]) // 2]
    t0 = _t.time()
    for _ in range(3):
        out = tf(src_h, boxes=boxes)
        float(out[0, 0, 0, 0, 0])
    e2e_s = (_t.time() - t0) / 3
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks = {}
    peak = float(peaks.get("hbm_gbs", 7700.0))
    res = {"workload": "train transform (RandomResizedCrop scale 0.5-1 + NormalizeVideo), %d clips x %d uint8 frames %dx%d -> %d^2 fp32"
                       % (batch, frames, H, W, crop),
           "value": round(batch / (k_ms / 1e3), 1), "unit": "clips/s", "kernel_ms": round(k_ms, 3), "gpu_launches": 1,
           "roofline": {"bound": "hbm", "achieved": round((box_bytes + out_bytes) / k_ms / 1e6, 1), "peak": peak, "unit": "GB/s",
                        "frac": round((box_bytes + out_bytes) / k_ms / 1e6 / peak, 3), "bytes_read": box_bytes,
                        "bytes_written": out_bytes, "l2": "256 MB flushed between launches"},
           "e2e": {"value": round(batch / e2e_s, 1), "unit": "clips/s", "h2d_bytes_per_step": src_h.numel() + batch * 96,
                   "d2h_bytes_per_step": 4}}
    # CPU: the reference's chain, one clip at a time like a DataLoader worker, on one core
    try:
        from baseline import ref_shim
        ref_ok = ref_shim.install()
    except Exception:
        ref_ok = False
    try:
        from torchvision import transforms
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            from torchvision.transforms import _transforms_video as tvv
        if ref_ok:
            from lavila.data.video_transforms import Permute
        else:
            from lavila_b200.data import Permute
        from lavila_b200.data import video_transforms as VT
        chain = transforms.Compose([Permute([3, 0, 1, 2]), transforms.RandomResizedCrop(crop, scale=(0.5, 1.0), antialias=False),
                                    tvv.NormalizeVideo(mean=list(VT.OPENAI_MEAN), std=list(VT.OPENAI_STD))])
        nthr = torch.get_num_threads()
        torch.set_num_threads(1)
        n, t0 = 0, _t.time()
        while _t.time() - t0 < cpu_budget_s and n < batch:
            chain(src_h[n].float())           # video_loader hands over fp32 frames (datasets.py:74-75)
            n += 1
        dt = _t.time() - t0
        torch.set_num_threads(nthr)
        res["cpu_baseline"] = {"value": round(n / dt, 2), "unit": "clips/s", "cores": 1, "kind": "reference" if ref_ok else "port",
                               "sample": "%d clips of the same batch, lavila Permute + torchvision %s RandomResizedCrop(antialias=False) "
                                         "+ NormalizeVideo, fp32 frames" % (n, __import__("torchvision").__version__)}
    except Exception as e:
        res["cpu_baseline"] = {"error": repr(e)[:200]}
    return res


def block_roofline(step_fn, engine, torch, B, frames, flop_kw):
    """CUDA events around every SpaceTimeBlock autograd node (forward and backward) of one extra step: the unit
    BASELINE.json's '>= 0.5x tensor-pipe roofline on SpaceTimeBlock fwd+bwd' is defined on.  FLOPs are the algorithmic
    3 x F_blk per clip per block for the blocks evaluated in full; the last block (only its CLS row is consumed,
    LastBlockClsFn skips 62 % of its FLOPs exactly) is timed separately and left out of the fraction."""
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", FALLBACK_PEAK_TFLOPS))
    rec = {"full": [], "last": []}

    def wrap(cls, name, key):
        orig = getattr(cls, name)

        def timed(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig(*a, **k)
            e1.record()
            rec[key].append((name, e0, e1))
            return r
        setattr(cls, name, staticmethod(timed))
        return lambda: setattr(cls, name, staticmethod(orig))

    undo = [wrap(engine.SpaceTimeBlockFn, "forward", "full"), wrap(engine.SpaceTimeBlockFn, "backward", "full"),
            wrap(engine.LastBlockClsFn, "forward", "last"), wrap(engine.LastBlockClsFn, "backward", "last")]
    NSTEPS = 3
    try:
        step_fn()
        rec["full"].clear()
        rec["last"].clear()
        for _ in range(NSTEPS):
            step_fn()
        torch.cuda.synchronize()
    finally:
        for u in undo:
            u()
    T, n, D = frames, flop_kw["n"], flop_kw["D"]
    N = 1 + T * n
    f_blk = 32 * N * D * D + 4 * D * ((T * n) * (T + n + 2) + 2 * N)
    nfull = sum(1 for nm, _, _ in rec["full"] if nm == "forward") // NSTEPS
    ms_f = sum(a.elapsed_time(b) for nm, a, b in rec["full"] if nm == "forward") / NSTEPS
    ms_b = sum(a.elapsed_time(b) for nm, a, b in rec["full"] if nm == "backward") / NSTEPS
    ms_last = sum(a.elapsed_time(b) for _, a, b in rec["last"]) / NSTEPS
    flops = 3.0 * f_blk * B * nfull
    ach = flops / max(1e-9, (ms_f + ms_b) / 1e3) / 1e12
    return {"unit": "TFLOP/s", "blocks_timed": nfull, "flops": flops, "gflop_per_clip_per_block_fwd_bwd": round(3 * f_blk / 1e9, 1),
            "ms_fwd": round(ms_f, 3), "ms_bwd": round(ms_b, 3), "ms": round(ms_f + ms_b, 3), "achieved": round(ach, 1),
            "peak": peak, "frac": round(ach / peak, 4), "last_block_cls_only_ms": round(ms_last, 3),
            "how": "CUDA events around SpaceTimeBlockFn.forward/.backward (engine.py), mean of %d instrumented steps" % NSTEPS}


def multi_gpu_loss_check(model, crit, frames_d, text_d, rank, world, dev):
    """N > 1 parity of the one thing that differs from N = 1: the fused NVLink gather + CLIPLoss kernel
    (clip_loss_fwd_gather_kernel).  On the SAME embeddings every rank evaluates (a) the fused path, (b) NCCL all_gather +
    the single-GPU loss kernel, (c) fp32 torch on the all_gathered batch (loss.py:76-79,107-116 restated inline), and the
    embedding gradients of (a) against autograd through (c).  Max over ranks; above 1e-5 the bench exits non-zero."""
    import torch
    import torch.distributed as dist
    from lavila_b200.models.loss import CLIPLoss
    with torch.no_grad():
        out = model(frames_d, text_d, norm_embed=True)
    img, txt = out["image_embed"].detach().float(), out["text_embed"].detach().float()
    scale = out["logit_scale"].detach().float()

    def run(crit_):
        i, t = img.clone().requires_grad_(True), txt.clone().requires_grad_(True)
        ld = crit_({"image_embed": i, "text_embed": t, "logit_scale": scale})
        ld["loss"].backward()
        return float(ld["loss"]), float(ld["clip_acc"]), i.grad, t.grad

    l_a, acc_a, gi_a, gt_a = run(crit)
    path = crit.gather_path
    crit.check_peer_error()
    os.environ["LAVILA_B200_P2P_LOSS"] = "0"
    try:
        nccl_crit = CLIPLoss(use_vissl=True, cache_labels=True, rank=rank, world_size=world)
        l_b, acc_b, gi_b, gt_b = run(nccl_crit)
    finally:
        os.environ["LAVILA_B200_P2P_LOSS"] = "1"
    # fp32 torch on the concatenated batch
    both = torch.cat((img, txt), 1).contiguous()
    buf = [torch.empty_like(both) for _ in range(world)]
    dist.all_gather(buf, both)
    allb = torch.cat(buf, 0)
    E = img.shape[1]
    B = img.shape[0]
    ai = allb[:, :E].clone().requires_grad_(True)
    at = allb[:, E:].clone().requires_grad_(True)
    logits = (scale * ai) @ at.t()
    lab = torch.arange(logits.shape[0], device=dev)
    l_c = (torch.nn.functional.cross_entropy(logits, lab) + torch.nn.functional.cross_entropy(logits.t(), lab)) / 2
    l_c.backward()
    acc_c = 100.0 * float((logits.argmax(-1) == lab).float().mean())
    sl = slice(rank * B, (rank + 1) * B)
    gi_c, gt_c = world * ai.grad[sl], world * at.grad[sl]          # GatherLayer semantics (distributed_utils.py:64-67)
    d = torch.tensor([abs(l_a - l_b), abs(l_a - float(l_c)), float((gi_a - gi_b).abs().max()), float((gt_a - gt_b).abs().max()),
                      float((gi_a - gi_c).abs().max()), float((gt_a - gt_c).abs().max()), abs(acc_a - acc_c)], device=dev)
    dist.all_reduce(d, op=dist.ReduceOp.MAX)
    d = [float(x) for x in d]
    # a numerical mismatch fails the run; the route taken is reported ("nccl" = symmetric memory unavailable on this box or
    # the global batch above the cooperative kernel's row limit -- a slower but equally exact path, not an error)
    ok = max(d[:6]) <= 1e-5 and d[6] == 0.0
    return {"path": path, "fused_gather": path == "p2p", "ok": bool(ok), "loss": l_a, "abs_diff": d[0], "vs_fp32_torch": d[1],
            "grad_abs_diff_vs_nccl": max(d[2], d[3]), "grad_abs_diff_vs_fp32_torch": max(d[4], d[5]), "acc_diff": d[6],
            "global_batch": int(logits.shape[0]), "tol": 1e-5}


def eager_baseline(args, dev, amp="bf16", steps=4, warmup=2):
    """The >= 6x target's denominator: the UNMODIFIED reference (baseline/_ref -- its own factory, CLIP, CLIPLoss and the
    body of train(), main_pretrain.py:486-530) as PyTorch-eager on this GPU.  Batch 64 if it fits, else the largest of
    (48, 32, 16) that does (torch raises OutOfMemoryError, caught).  Returns None when baseline/_ref is not installed."""
    import torch
    from baseline import ref_shim
    if not ref_shim.available():
        return {"unavailable": "baseline/_ref not installed (python baseline/install_ref.py needs /root/reference)"}
    from baseline import ref_arms
    import contextlib
    import gc
    res = None
    for B in [b for b in (args.batch, 48, 32, 16) if b <= args.batch]:
        step = None
        try:
            with contextlib.redirect_stdout(sys.stderr):
                step = ref_arms.reference_step_fn(dev, amp=amp, num_frames=args.frames)
            x, text = make_batch(B, args.frames, 1234)
            x, text = x.to(dev), text.to(dev)
            for _ in range(warmup):
                step(x, text)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = step(x, text)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            res = {"value": round(B / (ms / 1e3), 3), "unit": "clips/s", "batch": B, "ms_per_step": round(ms, 2),
                   "dtype": "bf16 autocast" if amp == "bf16" else "fp16 autocast + GradScaler", "kind": "reference",
                   "steps": steps, "warmup": warmup, "loss": float(loss),
                   "max_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
                   "what": "baseline/_ref lavila.models.models.CLIP_OPENAI_TIMESFORMER_BASE + get_loss (CLIPLoss) + AdamW, "
                           "train() body of main_pretrain.py:486-530, inputs resident, 1 GPU"}
        except torch.OutOfMemoryError:
            res = None
        finally:
            del step
            x = text = None
            gc.collect()
            torch.cuda.empty_cache()
        if res is not None:
            break
    torch.cuda.reset_peak_memory_stats()
    return res if res is not None else {"unavailable": "out of memory at every batch size tried"}


# --------------------------------------------------------------------------------------------- CPU arms
def cpu_step_fn(frames):
    """(step(batch, seed) -> loss, kind): the unmodified reference on the host cores in fp32 (kind 'reference') when
    baseline/_ref is installed, else the oracle port of the same algorithm (kind 'port')."""
    from baseline import ref_shim
    if ref_shim.available():
        import contextlib
        from baseline import ref_arms
        with contextlib.redirect_stdout(sys.stderr):
            ref_step = ref_arms.reference_step_fn("cpu", amp="off", num_frames=frames)

        def step(batch, seed):
            x, text = make_batch(batch, frames, seed)
            return ref_step(x, text)
        return step, "reference"

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

    return step, "port"


def host_threads():
    """Threads the CPU arms may use: scheduler affinity, capped by the cgroup CPU quota and by 32 (the fp32 GEMMs stop
    scaling past that, and oversubscribing a shared host is catastrophic: 250 s/clip was observed)."""
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
    step, kind = cpu_step_fn(frames)
    t0 = time.time()
    step(2, 1)                                  # calibrate on two clips (the reference cannot train at batch 1: SURVEY B.16)
    t1 = (time.time() - t0) / 2
    batch = int(max(2, min(4, budget_s // max(t1, 1e-3) - 1)))
    t0 = time.time()
    step(batch, 2)
    dt = time.time() - t0
    what = "the unmodified reference (baseline/_ref)" if kind == "reference" else "the oracle port"
    return {"value": round(batch / dt, 4), "unit": "clips/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": "1 fp32 training step (fwd+CLIPLoss+bwd+AdamW) of %s at batch %d, %d frames, after a "
                      "2-clip calibration step (%.1f s)" % (what, batch, frames, 2 * t1)}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path (baseline/_ref through its public API: factory,
    CLIP.forward, CLIPLoss, the train() body) in fp32 on the host cores, same metric; each step is a bounded sample (small
    batch) so the whole run ends within minutes.  Rank 0 only; the other ranks exit without work."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch
    cores = host_threads()
    torch.set_num_threads(cores)
    step, kind = cpu_step_fn(args.frames)
    t0 = time.time()
    step(2, 0)
    t1 = (time.time() - t0) / 2
    total = args.steps + args.warmup
    batch = int(max(2, min(4, 150.0 / max(total * t1, 1e-3))))
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
            "cpu_baseline": {"value": round(v, 4), "unit": "clips/s", "cores": torch.get_num_threads(), "kind": kind,
                             "sample": "%d steps of batch %d" % (args.steps, batch)},
            "e2e": {"value": round(v, 4), "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_eager(args):
    """Not part of the driver contract: the eager_baseline leg on its own (the unmodified reference as PyTorch-eager on one
    B200, bf16 autocast or --amp fp16 = its literal fp16 + GradScaler mode) -- BASELINE.md B1/B2."""
    import torch
    torch.cuda.set_device(0)
    res = eager_baseline(args, torch.device("cuda", 0), amp=args.amp, steps=max(2, args.steps), warmup=max(2, args.warmup))
    res = dict(res or {})
    res.update({"impl": "eager", "metric": METRIC})
    print(json.dumps(res))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "eager":
        run_eager(a)
    else:
        run_ours(a)
