"""Post-processing and scoring right after the denoise loop (CLIP-DDPM.py:620-631, COCO_BLEU.py:256-263):
`indexes.unique_consecutive(dim=-1)` -> decode -> corpus BLEU-4. torchmetrics / torchtext are not in this image, so the BLEU
of `torchmetrics.BLEUScore()` (n_gram = 4, uniform weights, no smoothing, brevity penalty, whitespace tokens) is restated here."""
from __future__ import annotations

import math
from collections import Counter
from typing import Callable, List, Sequence

import torch


def postprocess(indexes: torch.Tensor, per_sequence: bool = False) -> List[torch.Tensor]:
    """The reference's `indexes.unique_consecutive(dim=-1)` (CLIP-DDPM.py:621) on a [B, L] tensor removes a COLUMN only when it
    equals the previous column for every sequence of the batch - per-sequence deduplication only happens with
    B = 1, as in COCO_BLEU.py. per_sequence=True applies the B = 1 behaviour to every row (what the authors intended)."""
    if per_sequence:
        return [row.unique_consecutive() for row in indexes]
    return list(indexes.unique_consecutive(dim=-1))


def decode(rows: Sequence[torch.Tensor], id_to_token: Callable[[int], str]) -> List[str]:
    """`tokenizer.decode` stand-in for a caller-supplied vocabulary: tokens joined by single spaces."""
    return [" ".join(id_to_token(int(i)) for i in row) for row in rows]


def _ngrams(tokens: Sequence[str], n: int) -> Counter:
    return Counter(tuple(tokens[i:i + n]) for i in range(len(tokens) - n + 1))


def bleu_score(candidates: Sequence[str], references: Sequence[Sequence[str]], n_gram: int = 4) -> float:
    """Corpus-level BLEU as torchmetrics.functional.bleu_score computes it: clipped n-gram counts summed over the corpus, geometric
    mean of the n precisions (0 if any is 0), brevity penalty with the closest reference length - ties go to the FIRST such reference in
    list order, as torchmetrics' `target_len_diff.index(min(target_len_diff))` does. torchmetrics itself is absent from this image (no
    network); the function is pinned on the published doc-test vectors of torchmetrics.functional.bleu_score (0.7598) and of
    nltk's sentence_bleu / corpus_bleu (0.5045666840058485, 0.5920778868801042), reproduced to 1e-12
    (tests/test_host_cpu.py::test_bleu_published_known_answers), besides hand-computed answers."""
    assert len(candidates) == len(references)
    num = [0] * n_gram
    den = [0] * n_gram
    c_len = r_len = 0
    for cand, refs in zip(candidates, references):
        c = cand.split()
        rs = [r.split() for r in refs]
        c_len += len(c)
        if rs:
            diffs = [abs(len(r) - len(c)) for r in rs]
            r_len += len(rs[diffs.index(min(diffs))])
        for n in range(1, n_gram + 1):
            cc = _ngrams(c, n)
            best: Counter = Counter()
            for r in rs:
                best |= _ngrams(r, n)
            num[n - 1] += sum((cc & best).values())
            den[n - 1] += sum(cc.values())
    if min(num) == 0 or c_len == 0:
        return 0.0
    log_p = sum(math.log(a / b) for a, b in zip(num, den)) / n_gram
    bp = 1.0 if c_len > r_len else math.exp(1.0 - r_len / c_len)
    return bp * math.exp(log_p)
