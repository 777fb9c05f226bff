"""Token-id vocabulary of the PARSeq path: EOS=0, charset 1..N, BOS=N+1, PAD=N+2 — the id assignment
of the reference `Tokenizer` (strhub/data/utils.py:102-129) and its greedy `decode`/`_filter`
(strhub/data/utils.py:79-99,120-129), plus `CharsetAdapter` (strhub/data/utils.py:26-42).
Host-side glue (string <-> ids); the ids are inputs/outputs of the CUDA path."""
from __future__ import annotations

import re
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor


class CharsetAdapter:
    """Maps a predicted label onto the evaluation charset (case folding + dropping unknown chars)."""

    def __init__(self, target_charset: str) -> None:
        self._fold_lower = target_charset == target_charset.lower()
        self._fold_upper = target_charset == target_charset.upper()
        self._drop = re.compile("[^" + re.escape(target_charset) + "]")

    def __call__(self, label: str) -> str:
        if self._fold_lower:
            label = label.lower()
        elif self._fold_upper:
            label = label.upper()
        return self._drop.sub("", label)


class Tokenizer:
    EOS, BOS, PAD = "[E]", "[B]", "[P]"

    def __init__(self, charset: str) -> None:
        self._itos: Tuple[str, ...] = (self.EOS,) + tuple(charset) + (self.BOS, self.PAD)
        self._stoi = {s: i for i, s in enumerate(self._itos)}
        self.eos_id = 0
        self.bos_id = len(charset) + 1
        self.pad_id = len(charset) + 2

    def __len__(self) -> int:
        return len(self._itos)

    def _tok2ids(self, tokens: str) -> List[int]:
        return [self._stoi[ch] for ch in tokens]

    def _ids2tok(self, token_ids: Sequence[int], join: bool = True):
        toks = [self._itos[i] for i in token_ids]
        return "".join(toks) if join else toks

    def encode(self, labels: Sequence[str], device: Optional[torch.device] = None) -> Tensor:
        rows = [[self.bos_id] + self._tok2ids(y) + [self.eos_id] for y in labels]
        width = max(len(r) for r in rows)
        out = torch.full((len(rows), width), self.pad_id, dtype=torch.long)
        for i, r in enumerate(rows):
            out[i, : len(r)] = torch.tensor(r, dtype=torch.long)
        return out.to(device) if device is not None else out

    def _filter(self, probs: Tensor, ids: Tensor):
        id_list = ids.tolist()
        cut = id_list.index(self.eos_id) if self.eos_id in id_list else len(id_list)
        return probs[: cut + 1], id_list[:cut]      # keep the EOS probability, drop EOS and what follows

    def decode(self, token_dists: Tensor, raw: bool = False):
        """token_dists: [N, L, C] probabilities -> (labels, per-token probabilities)."""
        # one device->host transfer for the whole batch instead of one per sample
        probs_all, ids_all = token_dists.max(-1)
        probs_all, ids_all = probs_all.cpu(), ids_all.cpu()
        labels, probs_out = [], []
        for probs, ids in zip(probs_all, ids_all):
            if raw:
                labels.append(self._ids2tok(ids.tolist(), False))
                probs_out.append(probs)
            else:
                p, kept = self._filter(probs, ids)
                labels.append(self._ids2tok(kept, True))
                probs_out.append(p)
        return labels, probs_out
