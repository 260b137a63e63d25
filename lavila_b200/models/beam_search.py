"""BeamSearchScorer for VCLM_HF.beam_sample / group_beam_search (lavila/models/narrator.py:149-366).

The reference imports `BeamSearchScorer` from `transformers==4.27` (requirements.txt:8, narrator.py:16-24) and drives it with
`process()` once per step (per beam group) and `finalize()` at the end.  transformers 5.x no longer ships the class, so the
bookkeeping is restated here from the published 4.27 algorithm (`transformers/generation/beam_search.py`): one hypothesis heap
per batch element holding up to `num_beams` finished sequences scored `sum_logprobs / len ** length_penalty`; a candidate
that ends in EOS is moved to the heap if it ranks inside the first `group_size` candidates, the others refill the beam; a
batch element is done when the heap is full and its worst kept score beats the best score still attainable.
Index arithmetic only -- nothing here touches the GPU kernels.
"""
from collections import UserDict

import torch


class BeamHypotheses:
    def __init__(self, num_beams, length_penalty, early_stopping, max_length=None):
        self.length_penalty = length_penalty
        self.early_stopping = early_stopping
        self.max_length = max_length
        self.num_beams = num_beams
        self.beams = []
        self.worst_score = 1e9

    def __len__(self):
        return len(self.beams)

    def add(self, hyp, sum_logprobs, beam_indices=None):
        score = sum_logprobs / (hyp.shape[-1] ** self.length_penalty)
        if len(self) < self.num_beams or score > self.worst_score:
            self.beams.append((score, hyp, beam_indices))
            if len(self) > self.num_beams:
                ranked = sorted([(s, idx) for idx, (s, _, _) in enumerate(self.beams)])
                del self.beams[ranked[0][1]]
                self.worst_score = ranked[1][0]
            else:
                self.worst_score = min(score, self.worst_score)

    def is_done(self, best_sum_logprobs, cur_len):
        if len(self) < self.num_beams:
            return False
        if self.early_stopping:
            return True
        return self.worst_score >= best_sum_logprobs / cur_len ** self.length_penalty


class BeamSearchScorer:
    def __init__(self, batch_size, num_beams, device, length_penalty=1.0, do_early_stopping=False, num_beam_hyps_to_keep=1,
                 num_beam_groups=1, max_length=None):
        if not isinstance(num_beams, int) or num_beams <= 1:
            raise ValueError("`num_beams` has to be an integer strictly greater than 1, but is %r" % (num_beams,))
        if not isinstance(num_beam_groups, int) or num_beam_groups > num_beams or num_beams % num_beam_groups != 0:
            raise ValueError("`num_beam_groups` has to divide `num_beams`")
        self.num_beams = num_beams
        self.device = device
        self.length_penalty = length_penalty
        self.do_early_stopping = do_early_stopping
        self.num_beam_hyps_to_keep = num_beam_hyps_to_keep
        self.num_beam_groups = num_beam_groups
        self.group_size = num_beams // num_beam_groups
        self._beam_hyps = [BeamHypotheses(num_beams, length_penalty, do_early_stopping, max_length) for _ in range(batch_size)]
        self._done = torch.zeros(batch_size, dtype=torch.bool, device=device)

    @property
    def is_done(self):
        return bool(self._done.all())

    def process(self, input_ids, next_scores, next_tokens, next_indices, pad_token_id=None, eos_token_id=None, beam_indices=None):
        cur_len = input_ids.shape[-1]
        batch_size = len(self._beam_hyps)
        if batch_size != input_ids.shape[0] // self.group_size:
            raise ValueError("input_ids holds %d rows for %d x %d beams" % (input_ids.shape[0], batch_size, self.group_size))
        device = input_ids.device
        out_scores = torch.zeros((batch_size, self.group_size), dtype=next_scores.dtype, device=device)
        out_tokens = torch.zeros((batch_size, self.group_size), dtype=next_tokens.dtype, device=device)
        out_indices = torch.zeros((batch_size, self.group_size), dtype=next_indices.dtype, device=device)
        eos = [eos_token_id] if isinstance(eos_token_id, int) else eos_token_id
        # one D2H copy per step instead of one .item() per candidate
        tok_h, sc_h, idx_h = next_tokens.tolist(), next_scores.tolist(), next_indices.tolist()
        done_h = self._done.tolist()
        for b, hyp in enumerate(self._beam_hyps):
            if done_h[b]:
                if eos is None or pad_token_id is None:
                    raise ValueError("a finished batch element needs eos_token_id and pad_token_id")
                out_scores[b, :] = 0
                out_tokens[b, :] = pad_token_id
                out_indices[b, :] = 0
                continue
            k = 0
            for rank, (tok, sc, idx) in enumerate(zip(tok_h[b], sc_h[b], idx_h[b])):
                row = b * self.group_size + idx
                if eos is not None and tok in eos:
                    if rank >= self.group_size:          # an EOS candidate outside the top group_size is dropped
                        continue
                    hyp.add(input_ids[row].clone(), sc, beam_indices=None)
                else:
                    out_scores[b, k] = sc
                    out_tokens[b, k] = tok
                    out_indices[b, k] = row
                    k += 1
                if k == self.group_size:
                    break
            if k < self.group_size:
                raise ValueError("fewer than %d non-EOS candidates for batch element %d" % (self.group_size, b))
            if hyp.is_done(max(sc_h[b]), cur_len):
                self._done[b] = True
        return UserDict({"next_beam_scores": out_scores.view(-1), "next_beam_tokens": out_tokens.view(-1),
                         "next_beam_indices": out_indices.view(-1)})

    def finalize(self, input_ids, final_beam_scores, final_beam_tokens, final_beam_indices, max_length, pad_token_id=None,
                 eos_token_id=None, beam_indices=None):
        batch_size = len(self._beam_hyps)
        eos = [eos_token_id] if isinstance(eos_token_id, int) else eos_token_id
        done_h = self._done.tolist()
        fs = final_beam_scores.tolist()
        for b, hyp in enumerate(self._beam_hyps):
            if done_h[b]:
                continue
            for j in range(self.num_beams):       # every open beam competes for a place among the kept hypotheses
                row = b * self.num_beams + j
                hyp.add(input_ids[row], fs[row], beam_indices=None)
        keep = self.num_beam_hyps_to_keep
        sent_lengths = input_ids.new_zeros(batch_size * keep)
        best, best_scores = [], torch.zeros(batch_size * keep, device=self.device, dtype=torch.float32)
        for b, hyp in enumerate(self._beam_hyps):
            ranked = sorted(hyp.beams, key=lambda x: x[0])
            for j in range(keep):
                score, seq, _ = ranked.pop()
                sent_lengths[keep * b + j] = len(seq)
                best.append(seq)
                best_scores[b * keep + j] = score
        longest = int(sent_lengths.max().item()) + 1
        out_len = min(longest, max_length) if max_length is not None else longest
        decoded = input_ids.new_zeros(batch_size * keep, out_len)
        if int(sent_lengths.min().item()) != int(sent_lengths.max().item()):
            if pad_token_id is None:
                raise ValueError("`pad_token_id` has to be defined")
            decoded.fill_(pad_token_id)
        for i, seq in enumerate(best):
            n = int(sent_lengths[i])
            decoded[i, :n] = seq
            if n < out_len:
                decoded[i, n] = eos[0]
        return UserDict({"sequences": decoded, "sequence_scores": best_scores, "beam_indices": None})
