"""VCLM_HF -- mirror of lavila/models/narrator.py:32-147,368-389 (encode_image, forward, generate, _get_logits_warper).

generate() keeps the reference's sampling algorithm (temperature / top-k / top-p warpers from `transformers`,
torch.multinomial, entropy-based "ppl") but (a) projects the cross-attention K/V of the 256 video tokens once per clip
instead of once per decoding step and layer, and (b) evaluates the LM head on the last position only -- both leave
every returned value unchanged (SURVEY.md 8(a) a16/a19 note the waste).  beam_sample / group_beam_search are "next"."""
import torch
import torch.nn.functional as F
from torch import nn

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
        nlls, num_tokens = torch.zeros(image_tokens.shape[0]).to(device), torch.zeros(image_tokens.shape[0]).to(device)
        is_reach_eos = torch.zeros(image_tokens.shape[0]).bool().to(device)
        use_cache = use_kv_cache and not teacher_forcing
        import os
        use_graph = use_cache and image_tokens.is_cuda and os.environ.get("LAVILA_B200_DECODE_GRAPH", "1") == "1"
        st = self._decode_state(image_tokens, max_text_length) if use_graph else None
        kv_cache = st["ctx"] if st is not None else {}
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

    def _decode_state(self, image_tokens, max_len):
        """Persistent buffers of the graph-captured decoding step, keyed by (sequences, max length, device, weight versions):
        per-layer self-attention KV caches, cross-attention K/V buffers (refilled in place for every new batch of clips),
        the static input ids / position scalars and the captured graph.  Weights changing (training) drop the state."""
        from ..engine import param_generation
        ver = (sum(p._version for p in self.text_decoder.parameters()), param_generation())
        key = (image_tokens.shape[0], image_tokens.shape[1], max_len, str(image_tokens.device), ver)
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
