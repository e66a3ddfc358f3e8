"""Generation metrics: ROUGE-1/2/L (F-measure) and exact match, dependency-free."""
import re
from collections import Counter
from typing import List


def _tok(s: str) -> List[str]:
    return re.findall(r"[a-z0-9]+", s.lower())


def _f(p: float, r: float) -> float:
    return 0.0 if p + r == 0 else 2 * p * r / (p + r)


def _ngrams(t: List[str], n: int) -> Counter:
    return Counter(tuple(t[i:i + n]) for i in range(len(t) - n + 1))


def rouge_n(pred: str, ref: str, n: int) -> float:
    p, r = _ngrams(_tok(pred), n), _ngrams(_tok(ref), n)
    overlap = sum((p & r).values())
    if not p or not r:
        return 0.0
    return _f(overlap / sum(p.values()), overlap / sum(r.values()))


def _lcs(a: List[str], b: List[str]) -> int:
    dp = [0] * (len(b) + 1)
    for x in a:
        prev = 0
        for j, y in enumerate(b, 1):
            cur = dp[j]
            dp[j] = prev + 1 if x == y else max(dp[j], dp[j - 1])
            prev = cur
    return dp[-1]


def rouge_l(pred: str, ref: str) -> float:
    a, b = _tok(pred), _tok(ref)
    if not a or not b:
        return 0.0
    l = _lcs(a, b)
    return _f(l / len(a), l / len(b))


class MetricFactory:
    @staticmethod
    def create(name: str):
        name = name.lower()
        if name in ("rouge", "rougel", "rouge_l"):
            return lambda preds, refs: {"rouge1": sum(rouge_n(p, r, 1) for p, r in zip(preds, refs)) / max(len(preds), 1),
                                        "rouge2": sum(rouge_n(p, r, 2) for p, r in zip(preds, refs)) / max(len(preds), 1),
                                        "rougeL": sum(rouge_l(p, r) for p, r in zip(preds, refs)) / max(len(preds), 1)}
        if name in ("accuracy", "exact_match", "em"):
            return lambda preds, refs: {"exact_match": sum(p.strip().lower() == r.strip().lower() for p, r in zip(preds, refs)) / max(len(preds), 1)}
        raise ValueError(f"unknown metric {name}")
