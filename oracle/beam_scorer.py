"""TEST INFRASTRUCTURE: plain restatement of `transformers==4.27`'s BeamSearchScorer (generation/beam_search.py), the class the
reference drives from VCLM_HF.beam_sample / group_beam_search (lavila/models/narrator.py:16-24,166-170,260-265,210-241).

transformers 4.27 is a pinned dependency of the reference (requirements.txt:8) that is NOT installed here (5.5 dropped the class)
and not vendored, so this is a restatement of its published algorithm from its documented behaviour -- the scorer ITSELF is
"parity unpinned"; what is pinned (tests/golden/make_golden_beam.py) is the reference's own beam_sample / group_beam_search
code executed with this scorer injected as `transformers.BeamSearchScorer`.  Deliberately written with the per-element loops and
`.item()` calls of the original so that it reads like the algorithm, not like the product's vectorised copy.
"""
import torch


class _Hyps:
    def __init__(self, num_beams, length_penalty, early_stopping):
        self.num_beams, self.length_penalty, self.early_stopping = num_beams, length_penalty, early_stopping
        self.beams, self.worst_score = [], 1e9

    def add(self, hyp, sum_logprobs):
        score = sum_logprobs / (hyp.shape[-1] ** self.length_penalty)
        if len(self.beams) < self.num_beams or score > self.worst_score:
            self.beams.append((score, hyp))
            if len(self.beams) > self.num_beams:
                order = sorted(range(len(self.beams)), key=lambda i: self.beams[i][0])
                second_worst = self.beams[order[1]][0]
                del self.beams[order[0]]
                self.worst_score = second_worst
            else:
                self.worst_score = min(score, self.worst_score)

    def is_done(self, best_sum_logprobs, cur_len):
        if len(self.beams) < self.num_beams:
            return False
        if self.early_stopping:
            return True
        return self.worst_score >= best_sum_logprobs / cur_len ** self.length_penalty


class BeamSearchScorer:
    def __init__(self, batch_size, num_beams, device, length_penalty=1.0, do_early_stopping=False, num_beam_hyps_to_keep=1,
                 num_beam_groups=1, **kwargs):
        assert num_beams > 1 and num_beams % num_beam_groups == 0
        self.num_beams, self.device, self.length_penalty = num_beams, device, length_penalty
        self.num_beam_hyps_to_keep, self.num_beam_groups = num_beam_hyps_to_keep, num_beam_groups
        self.group_size = num_beams // num_beam_groups
        self._beam_hyps = [_Hyps(num_beams, length_penalty, do_early_stopping) for _ in range(batch_size)]
        self._done = torch.tensor([False] * batch_size, dtype=torch.bool, device=device)

    @property
    def is_done(self):
        return self._done.all()

    def process(self, input_ids, next_scores, next_tokens, next_indices, pad_token_id=None, eos_token_id=None, beam_indices=None):
        cur_len = input_ids.shape[-1]
        B = len(self._beam_hyps)
        assert B == input_ids.shape[0] // self.group_size
        nbs = torch.zeros((B, self.group_size), dtype=next_scores.dtype, device=input_ids.device)
        nbt = torch.zeros((B, self.group_size), dtype=next_tokens.dtype, device=input_ids.device)
        nbi = torch.zeros((B, self.group_size), dtype=next_indices.dtype, device=input_ids.device)
        for b in range(B):
            if self._done[b]:
                nbs[b, :], nbt[b, :], nbi[b, :] = 0, pad_token_id, 0
                continue
            beam_idx = 0
            for rank in range(next_tokens.shape[1]):
                tok, sc, idx = next_tokens[b, rank], next_scores[b, rank], next_indices[b, rank]
                row = b * self.group_size + idx
                if eos_token_id is not None and tok.item() == eos_token_id:
                    if rank >= self.group_size:
                        continue
                    self._beam_hyps[b].add(input_ids[row].clone(), sc.item())
                else:
                    nbs[b, beam_idx], nbt[b, beam_idx], nbi[b, beam_idx] = sc, tok, row
                    beam_idx += 1
                if beam_idx == self.group_size:
                    break
            assert beam_idx == self.group_size
            self._done[b] = self._done[b] or self._beam_hyps[b].is_done(next_scores[b].max().item(), cur_len)
        return {"next_beam_scores": nbs.view(-1), "next_beam_tokens": nbt.view(-1), "next_beam_indices": nbi.view(-1)}

    def finalize(self, input_ids, final_beam_scores, final_beam_tokens, final_beam_indices, max_length, pad_token_id=None,
                 eos_token_id=None, beam_indices=None):
        B, keep = len(self._beam_hyps), self.num_beam_hyps_to_keep
        for b in range(B):
            if self._done[b]:
                continue
            for j in range(self.num_beams):
                row = b * self.num_beams + j
                self._beam_hyps[b].add(input_ids[row], final_beam_scores[row].item())
        lengths = input_ids.new(B * keep)
        best, scores = [], torch.zeros(B * keep, device=self.device, dtype=torch.float32)
        for b in range(B):
            ranked = sorted(self._beam_hyps[b].beams, key=lambda x: x[0])
            for j in range(keep):
                s, h = ranked.pop()
                lengths[keep * b + j] = len(h)
                best.append(h)
                scores[b * keep + j] = s
        out_len = min(lengths.max().item() + 1, max_length) if max_length is not None else lengths.max().item() + 1
        decoded = input_ids.new(B * keep, out_len)
        if lengths.min().item() != lengths.max().item():
            decoded.fill_(pad_token_id)
        for i, h in enumerate(best):
            decoded[i, :lengths[i]] = h
            if lengths[i] < out_len:
                decoded[i, lengths[i]] = eos_token_id
        return {"sequences": decoded, "sequence_scores": scores}
