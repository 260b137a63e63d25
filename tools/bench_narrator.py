"""Narrator inference throughput (BASELINE.json config 4 shape, TSF-B video encoder + gated GPT-2 XL decoder, random
weights): encode_image + generate(top_p=0.95, temperature=0.7, max_text_length=77, early_stopping=False).

    python tools/bench_narrator.py --batch 32 --returns 1 [--impl eager]   # eager = oracle port on the GPU (fp32)
"""
import argparse
import json
import os
import sys
import time
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--returns", default="1", help="num_return_sequences; comma list = several runs with one model")
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--max-len", type=int, default=77)
    ap.add_argument("--impl", default="ours", choices=["ours", "eager"])
    ap.add_argument("--decoder", default="xl", choices=["xl", "base"])
    ap.add_argument("--no-graph", action="store_true", help="ours: KV-cached decoding without the CUDA-graph replay of the step")
    ap.add_argument("--no-kv-cache", action="store_true", help="ours: re-run the whole prefix every step (reference algorithm)")
    ap.add_argument("--encoder", default="base", choices=["base", "large"],
                    help="large = TSF-L/14 224px: VCLM_OPENAI_TIMESFORMER_LARGE_GPT2_XL, the model BASELINE config 4 names")
    a = ap.parse_args()
    if a.no_graph:
        os.environ["LAVILA_B200_DECODE_GRAPH"] = "0"
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    tok = SimpleNamespace(bos_token_id=50256, eos_token_id=50256, pad_token_id=0)
    frames = torch.randn(a.batch, 3, a.frames, 224, 224, device=dev)
    if a.impl == "ours":
        from lavila_b200.models import models as M
        f = {("base", "xl"): M.VCLM_OPENAI_TIMESFORMER_BASE_GPT2_XL, ("base", "base"): M.VCLM_OPENAI_TIMESFORMER_BASE_GPT2,
             ("large", "xl"): M.VCLM_OPENAI_TIMESFORMER_LARGE_GPT2_XL, ("large", "base"): M.VCLM_OPENAI_TIMESFORMER_LARGE_GPT2}[
            (a.encoder, a.decoder)]
        model = f(gated_xattn=True, num_frames=a.frames).to(dev).eval()
        with torch.no_grad():
            for n, p in model.named_parameters():
                if "alpha" in n:
                    p.fill_(0.5)
                if "timeattn" in n or "temporal_embed" in n:
                    p.normal_(0, 0.02)

        def run():
            t = model.encode_image(frames)
            return model.generate(t, tok, max_text_length=a.max_len, top_p=0.95, temperature=0.7,
                                  num_return_sequences=R, early_stopping=False, use_kv_cache=not a.no_kv_cache)
    else:
        from oracle import narrator as ON
        H, Ld, nh, freq = (1600, 48, 25, 2) if a.decoder == "xl" else (768, 12, 12, 1)
        vis = dict(img_size=224, patch_size=16, embed_dim=768, depth=12, num_heads=12) if a.encoder == "base" else \
            dict(img_size=224, patch_size=14, embed_dim=1024, depth=24, num_heads=16)
        cfg = dict(visual=dict(num_frames=a.frames, ln_pre=True, **vis),
                   n_embd=H, n_head=nh, n_layer=Ld, cross_attn_freq=freq, vocab_size=50257, n_positions=1024,
                   num_img_queries=256, pool_heads=nh)
        p = {k: v.to(dev) for k, v in ON.init_narrator_params(cfg, seed=0).items()}
        # oracle helpers build zeros on the CPU: make default tensors land on the GPU for this arm
        torch.set_default_device(dev)

        def run():   # the reference algorithm: no caches, full re-forward and full LM head every step
            with torch.no_grad():
                t = ON.vclm_encode_image(frames, p, cfg).repeat_interleave(R, 0)
                ids = torch.full((t.shape[0], 1), 50256, dtype=torch.int64, device=dev)
                for _ in range(a.max_len - 1):
                    lg = ON.gpt2_lm_logits(ids, t, p, cfg)[:, -1]
                    pr = torch.softmax(ON.warp_logits(lg, 0.7, 0.95), dim=-1)
                    ids = torch.cat((ids, torch.multinomial(pr, 1)), 1)
                return ids, None
    for R in [int(x) for x in str(a.returns).split(",")]:
        run()
        torch.cuda.synchronize()
        t0 = time.time()
        ids, _ = run()
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(json.dumps({"impl": a.impl, "encoder": a.encoder, "decoder": a.decoder, "kv_cache": (a.impl == "ours" and not a.no_kv_cache), "cuda_graph": (a.impl == "ours" and not a.no_kv_cache and not a.no_graph),
                          "batch": a.batch, "returns": R, "frames": a.frames,
                          "tokens": int(ids.shape[1]), "seconds": round(dt, 3), "clips_per_s": round(a.batch / dt, 3),
                          "sequences_per_s": round(a.batch * R / dt, 3),
                          "max_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}), flush=True)

if __name__ == "__main__":
    main()
