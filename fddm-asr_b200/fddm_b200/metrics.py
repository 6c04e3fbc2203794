"""calculate_cer / calculate_wer -- B200 host mirror of models/evaluate.py:94-134.

Same signatures and the same empty-reference rules as the reference; the Levenshtein distances of ALL pairs handed
in are computed by one kernel launch (`fddm_edit_distance`).  The reference calls these once per utterance inside
its evaluation loops (evaluate.py:185,331,448); `batch_cer` / `batch_wer` take the whole evaluation set at once,
which is where a GPU pays.  Characters are compared as code points, words as ids of a per-call vocabulary.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import _lib as L


def _edit_distances(refs: Sequence[Sequence[int]], hyps: Sequence[Sequence[int]], device=None) -> List[int]:
    if len(refs) != len(hyps):
        raise ValueError("refs and hyps must have the same length")
    n = len(refs)
    if n == 0:
        return []
    if not torch.cuda.is_available():
        raise ValueError("fddm_b200 runs on CUDA only (no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    r_off = np.zeros(n + 1, dtype=np.int64); h_off = np.zeros(n + 1, dtype=np.int64)
    r_off[1:] = np.cumsum([len(r) for r in refs]); h_off[1:] = np.cumsum([len(h) for h in hyps])
    r_sym = np.fromiter((s for r in refs for s in r), dtype=np.int32, count=int(r_off[-1]))
    h_sym = np.fromiter((s for h in hyps for s in h), dtype=np.int32, count=int(h_off[-1]))
    max_h = int(max(len(h) for h in hyps))
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a if a.size else np.zeros(1, a.dtype))).to(dev)
    d_r, d_h, d_ro, d_ho = to(r_sym), to(h_sym), to(r_off), to(h_off)
    ws = torch.empty(int(L.lib.fddm_edit_distance_workspace_bytes(n, max_h)), dtype=torch.uint8, device=dev)
    out = torch.empty(n, dtype=torch.int32, device=dev)
    L.check(L.lib.fddm_edit_distance(d_r.data_ptr(), d_ro.data_ptr(), d_h.data_ptr(), d_ho.data_ptr(), n, max_h,
                                     ws.data_ptr(), out.data_ptr(), L.stream_ptr(dev)), "edit_distance")
    return out.cpu().tolist()


def batch_cer(refs: Sequence[str], hyps: Sequence[str]) -> List[float]:
    """calculate_cer (evaluate.py:94-118) for every (ref, hyp) pair: Levenshtein on characters / len(ref);
    an empty reference gives 0.0 for an empty hypothesis and 1.0 otherwise (evaluate.py:116-117)."""
    d = _edit_distances([[ord(c) for c in r] for r in refs], [[ord(c) for c in h] for h in hyps])
    return [(0.0 if len(h) == 0 else 1.0) if len(r) == 0 else float(x) / float(len(r)) for x, r, h in zip(d, refs, hyps)]


def batch_wer(refs: Sequence[str], hyps: Sequence[str]) -> List[float]:
    """calculate_wer (evaluate.py:120-134): whitespace-split words, Levenshtein / number of reference words;
    an empty reference gives 0.0."""
    vocab = {}
    enc = lambda s: [vocab.setdefault(w, len(vocab)) for w in s.strip().split()]
    rr = [enc(r) for r in refs]; hh = [enc(h) for h in hyps]
    d = _edit_distances(rr, hh)
    return [0.0 if len(r) == 0 else float(x) / float(len(r)) for x, r in zip(d, rr)]


def calculate_cer(ref: str, hyp: str) -> float:
    return batch_cer([ref], [hyp])[0]


def calculate_wer(ref: str, hyp: str) -> float:
    return batch_wer([ref], [hyp])[0]
