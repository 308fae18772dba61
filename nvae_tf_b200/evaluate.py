"""The hot-path leg of the reference's evaluate.py: `neg_log_likelihood` (evaluate.py:111-123), the
importance-weighted bound on the test NLL.  Everything else in evaluate.py (FID, precision/recall, PPL,
TensorBoard) is outside SURVEY 8 and not provided.

Per batch: `n_attempts` calls of `model(batch, nll=True)` (inference-mode BN, fresh epsilons per attempt), each
giving log_iw = -recon(crop 28x28) - log_q + log_p [B]; then -mean_B(logsumexp_k log_iw - log K).  The k rows stay
on the device and ONE launch (`nvae_iwae_nll`) does the logsumexp + mean."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List

import numpy as np
import torch


@dataclass
class Metric:
    """evaluate.py's Metric.from_list: mean and standard deviation of per-batch values."""
    mean: float
    stddev: float

    @classmethod
    def from_list(cls, values: List[float]) -> "Metric":
        a = np.asarray(values, dtype=np.float64)
        return cls(float(a.mean()), float(a.std()))


def batch_neg_log_likelihood(model, batch, n_attempts: int = 10) -> torch.Tensor:
    """NLL bound of ONE batch -> device scalar [1] (the body of the loop at evaluate.py:113-122)."""
    rt = model.rt
    batch = model._as_device(batch)
    B = batch.shape[0]
    recon = rt.empty(n_attempts, B)
    log_q = rt.empty(n_attempts, B)
    log_p = rt.empty(n_attempts, B)
    for k in range(n_attempts):
        reconstruction, _, lp, lq = model(batch, nll=True)
        r = model.calculate_recon_loss(batch, reconstruction, crop_output=True)
        recon[k].copy_(r)
        log_q[k].copy_(lq)
        log_p[k].copy_(lp)
    nll = rt.empty(1)
    rt.lib.iwae_nll(recon.data_ptr(), log_q.data_ptr(), log_p.data_ptr(), n_attempts, B, None, nll.data_ptr(), rt.stream)
    return nll


def neg_log_likelihood(model, test_data: Iterable, n_attempts: int = 10) -> Metric:
    """evaluate.py:111-123.  `test_data` yields (batch, label) pairs like the reference's tf.data pipeline (a bare
    batch is accepted too)."""
    nlls = []
    for item in test_data:
        batch = item[0] if isinstance(item, (tuple, list)) else item
        nlls.append(float(batch_neg_log_likelihood(model, batch, n_attempts).item()))
    return Metric.from_list(nlls)
