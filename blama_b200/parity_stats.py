"""Parity bookkeeping shared by the GPU tests, smoke() and bench.py: pure numpy comparisons of two logit rows / top-10 lists
(whoever calls this brings the reference values; nothing here imports the oracle).

Terms: a decode step is CLEAN when the device row equals the reference row to fp32 summation noise, FLIPPED
when one Q8_K / f16 rounding went the other way somewhere upstream (bounded by FLIP_TOL).  Top-10 ids must be identical at every
rank whose gap to the next rank in the reference exceeds twice the row's measured deviation (an order can only change inside 2 x
deviation); on clean rows that is practically every rank."""
from __future__ import annotations

from typing import Dict, List

import numpy as np

CLEAN_TOL = 2e-4      # absolute; logits have std ~2
FLIP_TOL = 0.25       # small test models (V <= 5000): measured worst flip 0.175 (oracle against itself with the fp32 partial sums added
                      # in the opposite order), <= 0.2 GPU vs oracle.
# At the BASELINE sizes nothing stays clean: with d = 4096, 32 layers and 128 256 logits per row some Q8_K / f16 rounding flips in
# every token, and the oracle against ITSELF (ORC_MODE_GGML vs ORC_MODE_GGML_ALT, same integer arithmetic, opposite fp32 summation
# order) differs by rms 0.05-0.08 / max 0.26-0.41 per row on the 8B model from the first token on (measured, tools/noise_floor.py).
# Full-size comparisons are therefore made RELATIVE to that floor, measured on the same tokens: the device row must be no further
# from the reference than FLOOR_FACTOR x what a second faithful implementation of the reference arithmetic is.
FLOOR_FACTOR = 1.5


def top_sorted(logits: np.ndarray, k: int):
    """ids and values of the k largest logits, descending, ties by lower id (the device's order)"""
    part = np.argpartition(-logits, k)[: k + 1] if k + 1 < len(logits) else np.arange(len(logits))
    order = part[np.lexsort((part, -logits[part]))][:k]
    return order.astype(np.int32), logits[order]


class StepStats:
    def __init__(self):
        self.max_abs: List[float] = []
        self.rms: List[float] = []
        self.floor_max_abs: List[float] = []
        self.floor_rms: List[float] = []
        self.ids_equal: List[bool] = []
        self.top1_equal: List[bool] = []
        self.overlap: List[float] = []           # |top-10 set of the device row  ∩  top-10 set of the reference row| / 10
        self.floor_ids_equal: List[bool] = []    # the same three, for the second faithful implementation of the reference arithmetic
        self.floor_top1_equal: List[bool] = []
        self.floor_overlap: List[float] = []
        self.gap_ranks = 0          # ranks whose reference gap demanded an identical id ...
        self.gap_ok = 0             # ... and got it
        self.gaps: List[float] = []

    def add_floor(self, alt: np.ndarray, want: np.ndarray):
        """deviation of a second faithful implementation of the reference arithmetic on the same row (the noise floor)"""
        d = alt - want
        self.floor_max_abs.append(float(np.abs(d).max()))
        self.floor_rms.append(float(np.sqrt(np.mean(d.astype(np.float64) ** 2))))
        a10, w10 = top_sorted(alt, 10)[0], top_sorted(want, 10)[0]
        self.floor_ids_equal.append(bool(np.array_equal(a10, w10)))
        self.floor_top1_equal.append(bool(a10[0] == w10[0]))
        self.floor_overlap.append(len(set(a10.tolist()) & set(w10.tolist())) / 10.0)

    def add(self, got: np.ndarray, want: np.ndarray, got_top_ids: np.ndarray):
        d = got - want
        err = float(np.abs(d).max())
        self.max_abs.append(err)
        self.rms.append(float(np.sqrt(np.mean(d.astype(np.float64) ** 2))))
        ids11, val11 = top_sorted(want, 11)
        self.ids_equal.append(bool(np.array_equal(got_top_ids[:10], ids11[:10])))
        self.top1_equal.append(bool(int(got_top_ids[0]) == int(ids11[0])))
        self.overlap.append(len(set(int(i) for i in got_top_ids[:10]) & set(int(i) for i in ids11[:10])) / 10.0)
        gaps = val11[:10] - val11[1:11]
        self.gaps += [float(g) for g in gaps]
        bad = []
        # rank r is pinned when it is separated from both neighbours by more than 2 x the measured deviation
        for r in range(10):
            above = val11[r - 1] - val11[r] if r > 0 else np.inf
            below = val11[r] - val11[r + 1]
            if above > 2 * err + 1e-6 and below > 2 * err + 1e-6:
                self.gap_ranks += 1
                if int(got_top_ids[r]) == int(ids11[r]):
                    self.gap_ok += 1
                else:
                    bad.append(r)
        return err, bad

    def summary(self) -> Dict[str, float]:
        n = max(1, len(self.max_abs))
        floor = {}
        if self.floor_max_abs:
            floor = {"floor_max_abs": max(self.floor_max_abs), "floor_rms": max(self.floor_rms),
                     "within_floor": bool(max(self.max_abs) <= FLOOR_FACTOR * max(self.floor_max_abs) + 0.02 and
                                          max(self.rms) <= FLOOR_FACTOR * max(self.floor_rms) + 1e-3),
                     # what the reference arithmetic reproduces of ITS OWN top-10 when only the fp32 summation order changes
                     "floor_top10_id_match_frac": sum(self.floor_ids_equal) / max(1, len(self.floor_ids_equal)),
                     "floor_top1_match_frac": sum(self.floor_top1_equal) / max(1, len(self.floor_top1_equal)),
                     "floor_top10_overlap_mean": float(np.mean(self.floor_overlap)) if self.floor_overlap else 0.0}
        return {"steps": len(self.max_abs), "max_abs": max(self.max_abs) if self.max_abs else 0.0, "rms": max(self.rms) if self.rms else 0.0, **floor,
                "clean_frac": sum(e <= CLEAN_TOL for e in self.max_abs) / n,
                "top10_id_match_frac": sum(self.ids_equal) / n,
                "top1_match_frac": sum(self.top1_equal) / n,
                "top10_overlap_mean": float(np.mean(self.overlap)) if self.overlap else 0.0,
                "pinned_ranks": self.gap_ranks, "pinned_ranks_ok": self.gap_ok,
                "gap_median": float(np.median(self.gaps)) if self.gaps else 0.0,
                "gap_p10": float(np.percentile(self.gaps, 10)) if self.gaps else 0.0}
