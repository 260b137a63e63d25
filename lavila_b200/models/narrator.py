"""VCLM_HF -- mirror of lavila/models/narrator.py:32-147,368-389 (encode_image, forward, generate, _get_logits_warper).

generate() keeps the reference's sampling algorithm (temperature / top-k / top-p warpers from `transformers`,
torch.multinomial, entropy-based "ppl") but (a) projects the cross-attention K/V of the 256 video tokens once per clip
instead of once per decoding step and layer, and (b) evaluates the LM head on the last position only -- both leave
every returned value unchanged (SURVEY.md 8(a) a16/a19 note the waste).  beam_sample / group_beam_search (narrator.py:149-366)
run the reference's candidate selection on the same decoder with the scorer of beam_search.py."""
import os

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops

from .coca import CrossAttention, LayerNorm
from .timesformer import SpaceTimeTransformer


class VCLM_HF(nn.Module):
    def __init__(self, vision_width: int, vision_model: nn.Module, text_width: int, text_decoder: nn.Module,
                 num_img_queries=256, dim_head=64, heads=8, **kwargs):
        super().__init__()
        self.vision_width = vision_width
        self.visual = vision_model
        self.text_width = text_width
        self.text_decoder = text_decoder
        self.img_queries = nn.Parameter(torch.empty(num_img_queries, text_width))
        self.img_attn_pool = CrossAttention(dim=text_width, context_dim=vision_width, dim_head=dim_head, heads=heads,
                                            norm_context=True)
        self.img_attn_pool_norm = LayerNorm(text_width)
        self.initialize_parameters()

    def half(self):
        """`--use-half` (main_infer_narrator.py:155-156, eval_zeroshot.py:142): a no-op here.  The kernels own the precision
        (bf16 operands, fp32 accumulation, fp32 parameters / residual stream / softmax -- at least the accuracy of the reference's
        fp16 module), so the parameters stay fp32 and half-precision inputs are widened on entry."""
        return self

    def bfloat16(self):
        return self

    def initialize_parameters(self):
        nn.init.normal_(self.img_queries, std=self.text_width ** -0.5)

    @torch.no_grad()
    def encode_image(self, image, use_checkpoint=False):
        """narrator.py:63-87.  image [B, C, T, H, W] -> [B, num_img_queries, text_width]."""
        if not isinstance(self.visual, SpaceTimeTransformer):
            raise NotImplementedError("only the SpaceTimeTransformer video encoder is on the B200 hot path")
        x = self.visual._features_bcthw(image, use_checkpoint=use_checkpoint, cls_at_last=False)   # [B, N, Dv]
        q = self.img_queries.unsqueeze(0).expand(x.shape[0], -1, -1)
        q = self.img_attn_pool(q, x)
        return self.img_attn_pool_norm(q)

    @torch.no_grad()
    def forward(self, image, text, mask=None, use_checkpoint=False, norm_embed=False):
        """narrator.py:89-104 (teacher-forced logits; inference only)."""
        text, labels = text[:, :-1], text[:, 1:]
        image_tokens = self.encode_image(image, use_checkpoint=use_checkpoint)
        logits = self.text_decoder(text.contiguous(), encoder_hidden_states=image_tokens).logits
        return {'text_tokens_logits': logits.permute(0, 2, 1), 'labels': labels}

    def generate(self, image_tokens, tokenizer, target=None, max_text_length=77, top_k=None, top_p=None,
                 num_return_sequences=1, temperature=1.0, teacher_forcing=False, early_stopping=False, use_kv_cache=True):
        """narrator.py:106-147.  Same sampling loop; the decoder is evaluated incrementally (use_kv_cache: self-attention
        keys/values of the prefix and the cross-attention K/V of the clip are cached, the LM head sees the last position
        only) instead of re-running the whole prefix every step -- the logits, hence the sampled ids, are the same."""
        image_tokens = image_tokens.repeat_interleave(num_return_sequences, dim=0)
        device = image_tokens.device
        generated_text_ids = torch.LongTensor([[tokenizer.bos_token_id]] * image_tokens.shape[0]).to(device)
        condition_text_ids = generated_text_ids.clone()
        logits_warper = self._get_logits_warper(top_k=top_k, top_p=top_p, typical_p=None, temperature=temperature, num_beams=1)
        # the fused temperature + top-p kernel covers the script's sampling setting (main_infer_narrator.py:57-60: top_p, no top_k)
        fused_filter = (top_p is not None and 0.0 < top_p < 1.0 and not top_k and image_tokens.is_cuda
                        and os.environ.get("LAVILA_B200_FUSED_SAMPLING", "1") == "1")
        nlls, num_tokens = torch.zeros(image_tokens.shape[0]).to(device), torch.zeros(image_tokens.shape[0]).to(device)
        is_reach_eos = torch.zeros(image_tokens.shape[0]).bool().to(device)
        use_cache = use_kv_cache and not teacher_forcing
        use_graph = use_cache and image_tokens.is_cuda and os.environ.get("LAVILA_B200_DECODE_GRAPH", "1") == "1"
        st = self._decode_state(image_tokens, max_text_length, num_return_sequences) if use_graph else None
        kv_cache = st["ctx"] if st is not None else {}
        kv_cache["_repeat"] = num_return_sequences      # rows r*R .. r*R+R-1 carry the same video tokens: shared cross K/V
        self_cache = st["self"] if st is not None else ({"max_len": max_text_length} if use_cache else None)
        with torch.no_grad():
            for i in range(max_text_length - 1):
                if st is not None and i >= 1:
                    # positions >= 1: one decoder step with static arguments (position in device memory); run eagerly once,
                    # then captured in a CUDA graph and replayed -- ~680 launches per step become one graph launch
                    st["ids"].copy_(condition_text_ids[:, -1:])
                    st["dyn"]["pos_idx"].fill_(i)
                    st["dyn"]["lk_dev"].fill_(i + 1)
                    if st["graph"] is not None:
                        st["graph"].replay()
                        out = st["out"]
                    elif not st["warm"]:
                        out = self.text_decoder(st["ids"], encoder_hidden_states=image_tokens, last_only=True,
                                                ctx_kv_cache=kv_cache, self_kv_cache=self_cache, dyn=st["dyn"])
                        st["warm"] = True
                    else:
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            st["out"] = self.text_decoder(st["ids"], encoder_hidden_states=image_tokens, last_only=True,
                                                          ctx_kv_cache=kv_cache, self_kv_cache=self_cache, dyn=st["dyn"])
                        st["graph"] = g
                        g.replay()                      # capture records, replay executes
                        out = st["out"]
                elif self_cache is not None:
                    out = self.text_decoder(condition_text_ids[:, -1:], encoder_hidden_states=image_tokens, last_only=True,
                                            ctx_kv_cache=kv_cache, self_kv_cache=self_cache, past_len=i)
                else:
                    out = self.text_decoder(condition_text_ids, encoder_hidden_states=image_tokens, last_only=True,
                                            ctx_kv_cache=kv_cache)
                next_token_logits = out.logits[:, -1, :]
                if target is not None:
                    nll = F.cross_entropy(next_token_logits, target[:, i + 1], ignore_index=tokenizer.pad_token_id, reduction='none')
                    nlls += nll
                    num_tokens += target[:, i + 1].ne(tokenizer.pad_token_id)
                else:
                    nll = torch.special.entr(F.softmax(next_token_logits, dim=1)).sum(dim=1)
                    nlls += nll * (~is_reach_eos)
                    num_tokens += (~is_reach_eos)
                if fused_filter:
                    # TemperatureLogitsWarper + TopPLogitsWarper(min_tokens_to_keep=1) as ONE kernel (csrc/sampling.cu): the same
                    # surviving token set found by a radix select instead of a [sequences x vocab] sort per step
                    if not next_token_logits.is_contiguous() and next_token_logits.stride(-1) != 1:
                        next_token_logits = next_token_logits.contiguous()
                    next_token_logits = ops.top_p_filter_(next_token_logits, temperature if temperature is not None else 1.0, top_p)
                else:
                    next_token_logits = logits_warper(generated_text_ids, next_token_logits)
                filtered_p = F.softmax(next_token_logits, dim=-1)
                next_token = torch.multinomial(filtered_p, num_samples=1)
                is_reach_eos = is_reach_eos | (next_token[:, 0] == tokenizer.eos_token_id)
                if early_stopping and torch.all(is_reach_eos):
                    break
                if teacher_forcing:
                    condition_text_ids = target[:, :i + 2]
                else:
                    condition_text_ids = torch.cat((generated_text_ids, next_token), dim=1)
                generated_text_ids = torch.cat((generated_text_ids, next_token), dim=1)
        return generated_text_ids, torch.exp(nlls / num_tokens)


    # ------------------------------------------------------------------------------------------------ beam decoding
    def _last_logprobs(self, ids, ctx, kv_cache):
        """log-softmax of the next-token logits of every row (the decoder re-reads the prefix; the clip's cross-attention
        K / V are projected once -- beams of a batch element are only permuted among themselves, so the cache stays valid)."""
        out = self.text_decoder(ids.contiguous(), encoder_hidden_states=ctx, last_only=True, ctx_kv_cache=kv_cache)
        return F.log_softmax(out.logits[:, -1, :], dim=-1)

    @torch.no_grad()
    def beam_sample(self, image_tokens, tokenizer, target=None, max_text_length=77, top_k=None, top_p=None, temperature=1.0,
                    length_penalty=1., num_beams=3, num_return_sequences=1, teacher_forcing=False, early_stopping=False):
        """narrator.py:149-241: stochastic beam search -- per step, 2 x num_beams continuations are SAMPLED (torch.multinomial)
        from the warped joint distribution over (beam, token), sorted by score and handed to the scorer."""
        from .beam_search import BeamSearchScorer
        device = image_tokens.device
        n_clips = image_tokens.shape[0]
        per_clip = num_beams * num_return_sequences
        ids = torch.full((n_clips * per_clip, 1), tokenizer.bos_token_id, device=device, dtype=torch.long)
        ctx = image_tokens.repeat_interleave(per_clip, dim=0)
        warper = self._get_logits_warper(top_k=top_k, top_p=top_p, typical_p=None, temperature=temperature, num_beams=num_beams)
        scorer = BeamSearchScorer(batch_size=n_clips * num_return_sequences, num_beams=num_beams, device=device,
                                  length_penalty=length_penalty)
        rows = n_clips * num_return_sequences                 # independent beam searches (one per returned sequence)
        beam_scores = torch.zeros(rows * num_beams, device=device)
        reached_eos = torch.zeros(ids.shape[0], dtype=torch.bool, device=device)
        kv_cache = {"_repeat": per_clip}
        cand_tokens = cand_beams = None
        for _ in range(max_text_length - 1):
            logp = self._last_logprobs(ids, ctx, kv_cache)                         # [rows * num_beams, V]
            joint = warper(ids, logp + beam_scores[:, None].expand_as(logp))       # the reference warps the JOINT scores (:199)
            V = joint.shape[-1]
            joint = joint.view(rows, num_beams * V)
            picks = torch.multinomial(F.softmax(joint, dim=-1), num_samples=2 * num_beams)
            cand_scores = torch.gather(joint, -1, picks)
            cand_scores, order = torch.sort(cand_scores, descending=True, dim=1)
            picks = torch.gather(picks, -1, order)
            cand_beams = torch.div(picks, V, rounding_mode="floor")
            cand_tokens = picks % V
            step = scorer.process(ids, cand_scores, cand_tokens, cand_beams, pad_token_id=tokenizer.pad_token_id,
                                  eos_token_id=tokenizer.eos_token_id)
            beam_scores = step["next_beam_scores"]
            ids = torch.cat([ids[step["next_beam_indices"], :], step["next_beam_tokens"].unsqueeze(-1)], dim=-1)
            reached_eos = reached_eos | (ids[:, -1] == tokenizer.eos_token_id)
            if scorer.is_done or bool(torch.all(reached_eos)):
                break
        fin = scorer.finalize(ids, beam_scores, cand_tokens, cand_beams, pad_token_id=tokenizer.pad_token_id,
                              eos_token_id=tokenizer.eos_token_id, max_length=max_text_length)
        return fin["sequences"], fin["sequence_scores"]

    @torch.no_grad()
    def group_beam_search(self, image_tokens, tokenizer, target=None, max_text_length=77, top_k=None, top_p=None,
                          temperature=1.0, length_penalty=1., num_beams=6, num_beam_groups=3, num_return_sequences=1,
                          teacher_forcing=False, early_stopping=False):
        """narrator.py:243-366: the beams of a clip are split into num_beam_groups groups that are advanced one after the other
        on the SAME decoder output of the step, each taking its top 2 x group_size continuations (the reference applies no
        diversity penalty between groups -- its logits_processor call is commented out, :309-310 -- and neither does this)."""
        from .beam_search import BeamSearchScorer
        device = image_tokens.device
        n_clips = image_tokens.shape[0]
        ids = torch.full((n_clips * num_beams, 1), tokenizer.bos_token_id, device=device, dtype=torch.long)
        ctx = image_tokens.repeat_interleave(num_beams, dim=0)
        warper = self._get_logits_warper(top_k=top_k, top_p=top_p, typical_p=None, temperature=temperature, num_beams=num_beams)
        scorer = BeamSearchScorer(batch_size=n_clips, num_beams=num_beams, num_beam_groups=num_beam_groups,
                                  num_beam_hyps_to_keep=num_return_sequences, device=device, length_penalty=length_penalty)
        gsz = num_beams // num_beam_groups
        beam_scores = torch.full((n_clips, num_beams), -1e9, dtype=torch.float, device=device)
        beam_scores[:, ::gsz] = 0                       # one live beam per group at the start
        beam_scores = beam_scores.view(-1)
        reached_eos = torch.zeros(ids.shape[0], dtype=torch.bool, device=device)
        # rows of group g across the batch, in batch order
        base = torch.arange(n_clips, device=device).view(-1, 1) * num_beams
        group_rows = [(base + torch.arange(g * gsz, min((g + 1) * gsz, num_beams), device=device).view(1, -1)).reshape(-1)
                      for g in range(num_beam_groups)]
        kv_cache = {"_repeat": num_beams}
        cand_tokens = cand_beams = None
        for _ in range(max_text_length - 1):
            logp_all = self._last_logprobs(ids, ctx, kv_cache)
            new_last = torch.zeros(n_clips * num_beams, dtype=ids.dtype, device=device)
            for rows_g in group_rows:
                g_ids = ids[rows_g]
                g_size = rows_g.numel() // n_clips
                logp = logp_all[rows_g]
                V = logp.shape[-1]
                joint = warper(ids, logp + beam_scores[rows_g].unsqueeze(-1))
                joint = joint.view(n_clips, g_size * V)
                cand_scores, picks = torch.topk(joint, 2 * g_size, dim=1, largest=True, sorted=True)
                cand_beams = torch.div(picks, V, rounding_mode="floor")
                cand_tokens = picks % V
                step = scorer.process(g_ids, cand_scores, cand_tokens, cand_beams, pad_token_id=tokenizer.pad_token_id,
                                      eos_token_id=tokenizer.eos_token_id, beam_indices=None)
                beam_scores[rows_g] = step["next_beam_scores"]
                picked = step["next_beam_indices"]
                ids[rows_g] = g_ids[picked]
                new_last[rows_g] = step["next_beam_tokens"]
            ids = torch.cat([ids, new_last.unsqueeze(-1)], dim=-1)
            reached_eos = reached_eos | (ids[:, -1] == tokenizer.eos_token_id)
            if scorer.is_done or bool(torch.all(reached_eos)):
                break
        fin = scorer.finalize(ids, beam_scores, cand_tokens, cand_beams, pad_token_id=tokenizer.pad_token_id,
                              eos_token_id=tokenizer.eos_token_id, max_length=max_text_length, beam_indices=None)
        return fin["sequences"], fin["sequence_scores"]

    def _decode_state(self, image_tokens, max_len, repeat=1):
        """Persistent buffers of the graph-captured decoding step, keyed by (sequences, max length, device, weight versions):
        per-layer self-attention KV caches, cross-attention K/V buffers (refilled in place for every new batch of clips),
        the static input ids / position scalars and the captured graph.  Weights changing (training) drop the state."""
        from ..engine import param_generation
        ver = (sum(p._version for p in self.text_decoder.parameters()), param_generation())
        key = (image_tokens.shape[0], image_tokens.shape[1], max_len, str(image_tokens.device), ver, int(repeat))
        store = self.__dict__.setdefault("_decode_states", {})
        st = store.get(key)
        if st is None:
            store.clear()      # one live configuration at a time (the caches are large)
            dev = image_tokens.device
            st = {"ctx": {}, "self": {"max_len": max_len}, "graph": None, "warm": False, "out": None,
                  "ids": torch.zeros(image_tokens.shape[0], 1, dtype=torch.int64, device=dev),
                  "dyn": {"pos_idx": torch.zeros(1, dtype=torch.int64, device=dev),
                          "lk_dev": torch.ones(1, dtype=torch.int32, device=dev)}}
            store[key] = st
        # the cross-attention K/V buffers hold the previous clips: every layer recomputes its own at step 0
        st["ctx"]["_stale"] = {k for k in st["ctx"] if isinstance(k, int)}
        return st

    def _get_logits_warper(self, top_k=None, top_p=None, typical_p=None, temperature=None, num_beams=None,
                           renormalize_logits=None):
        """narrator.py:368-389."""
        from transformers.generation.logits_process import (LogitNormalization, LogitsProcessorList, TemperatureLogitsWarper,
                                                             TopKLogitsWarper, TopPLogitsWarper, TypicalLogitsWarper)
        top_k = top_k if top_k is not None else 0
        top_p = top_p if top_p is not None else 1.0
        typical_p = typical_p if typical_p is not None else 1.
        temperature = temperature if temperature is not None else 1.
        warpers = LogitsProcessorList()
        if temperature is not None and temperature != 1.0:
            warpers.append(TemperatureLogitsWarper(temperature))
        if top_k is not None and top_k != 0:
            warpers.append(TopKLogitsWarper(top_k=top_k, min_tokens_to_keep=(2 if num_beams > 1 else 1)))
        if top_p is not None and top_p < 1.0:
            warpers.append(TopPLogitsWarper(top_p=top_p, min_tokens_to_keep=(2 if num_beams > 1 else 1)))
        if typical_p is not None and typical_p < 1.0:
            warpers.append(TypicalLogitsWarper(mass=typical_p, min_tokens_to_keep=(2 if num_beams > 1 else 1)))
        if renormalize_logits is True:
            warpers.append(LogitNormalization())
        return warpers
