"""ORACLE (test infrastructure): the reference's probe train step on cached features, with torch's own modules.

Follows probing/train/singletask_trainer.py:183-191 (``logits = head_only(x)``; ``CrossEntropyLoss(weight,
ignore_index=-1)`` built at :105-108), probing/train/base_trainer.py:166-181 (zero_grad, backward, optimizer step) and
the optimizer of singletask_trainer.py:162-168 (``AdamW(head params, lr, weight_decay)``). The head is
``make_head`` of probing/models/base_probe.py:7-12. The only liberty: nn.Dropout's random mask is replaced by an
explicit keep-mask argument so that the CUDA path can be compared element for element.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


class _MaskedDropout(nn.Module):
    def __init__(self, p):
        super().__init__()
        self.p = p
        self.keep: Optional[torch.Tensor] = None

    def forward(self, x):
        if self.keep is None or not self.training:
            return x
        return x * self.keep.to(x.dtype) / (1.0 - self.p)


def make_reference_head(dim: int, classes: int, dropout_p: float, state: Dict[str, torch.Tensor]) -> nn.Sequential:
    head = nn.Sequential(nn.BatchNorm1d(dim), _MaskedDropout(dropout_p), nn.Linear(dim, classes))
    head.load_state_dict(state, strict=True)
    return head.double()  # fp64: the oracle is the exact answer, the CUDA fp32 path is measured against it


def reference_train_steps(state: Dict[str, torch.Tensor], batches, classes: int, dropout_p: float,
                          class_weight: Optional[torch.Tensor], lr=1e-4, weight_decay=1e-4) -> Tuple[Dict, list]:
    """Run len(batches) optimisation steps; batches = [(x [B,D], y [B], keep [B,D] or None), ...].
    Returns (final head state dict in fp32, [loss per step])."""
    dim = state["2.weight"].shape[1]
    head = make_reference_head(dim, classes, dropout_p, state).train()
    crit = nn.CrossEntropyLoss(weight=None if class_weight is None else class_weight.double(), ignore_index=-1)
    opt = torch.optim.AdamW(head.parameters(), lr=lr, weight_decay=weight_decay)
    losses = []
    for x, y, keep in batches:
        head[1].keep = keep
        opt.zero_grad(set_to_none=True)
        loss = crit(head(x.double()), y)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    return {k: (v.float() if v.is_floating_point() else v) for k, v in head.state_dict().items()}, losses


# ---------------------------------------------------------------------------------------------------------------
# The three-phase split the CUDA path uses, restated in fp64 torch so the data-parallel algebra (sum over shards of
# phase outputs == full batch) can be checked on CPU with gloo.
def phase1_stats(x: torch.Tensor) -> torch.Tensor:
    x = x.double()
    return torch.cat([x.sum(0), (x * x).sum(0)])


def phase2_bucket(x, y, stats, n_global, gamma, beta, weight, bias, class_weight, keep, dropout_p, bn_eps=1e-5):
    """bucket = [dW | db | dgamma | dbeta | loss_sum | weight_sum] as un-normalised sums over the local rows."""
    x = x.double()
    d = x.shape[1]
    c = weight.shape[0]
    mu = stats[:d] / n_global
    var = (stats[d:] / n_global - mu * mu).clamp_min(0)
    rstd = 1.0 / torch.sqrt(var + bn_eps)
    xh = (x - mu) * rstd
    kp = torch.ones_like(x) if keep is None else keep.double() / (1.0 - dropout_p)
    z = (xh * gamma.double() + beta.double()) * kp
    logits = z @ weight.double().t() + bias.double()
    valid = (y >= 0) & (y < c)
    ysafe = y.clamp(0, c - 1)
    w = torch.ones(c, dtype=torch.float64) if class_weight is None else class_weight.double()
    wy = torch.where(valid, w[ysafe], torch.zeros((), dtype=torch.float64))
    logp = F.log_softmax(logits, dim=1)
    loss_b = -wy * logp.gather(1, ysafe[:, None])[:, 0]
    dl = wy[:, None] * (logp.exp() - F.one_hot(ysafe, c).double())
    dz = dl @ weight.double()
    dy = dz * kp
    return torch.cat([(dl.t() @ z).flatten(), dl.sum(0), (dy * xh).sum(0), dy.sum(0), loss_b.sum()[None], wy.sum()[None]])


# ---------------------------------------------------------------------------------------------------------------
# Multitask / deeper-head step: probing/train/multitask_trainer.py:188-263 over base_trainer.py:166-181, heads from
# probing/models/base_probe.py:7-22, run with torch's own modules and autograd in fp64.
def make_reference_head_any(dim: int, classes: int, hidden: int, dropout_p: float, state: Dict[str, torch.Tensor]):
    if hidden:
        head = nn.Sequential(nn.BatchNorm1d(dim), _MaskedDropout(dropout_p), nn.Linear(dim, hidden), nn.GELU(),
                             _MaskedDropout(dropout_p), nn.Linear(hidden, classes))
    else:
        head = nn.Sequential(nn.BatchNorm1d(dim), _MaskedDropout(dropout_p), nn.Linear(dim, classes))
    head.load_state_dict(state, strict=True)
    return head.double()


def reference_multitask_steps(states: Dict[str, Dict[str, torch.Tensor]], tasks: Dict[str, int], hidden: int, batches,
                              dropout_p: float, class_weights: Dict[str, Optional[torch.Tensor]],
                              task_weights: Dict[str, float], lr=1e-4, weight_decay=1e-4):
    """batches = [(x [B,D], y [B,T], masks [(keep1, keep2 or None) per task]), ...]; the loss of a task is the
    class-weighted CE over its labelled rows (``criterions[t](logit[mask], target[mask])``), the step minimises
    ``sum_t w_t loss_t`` (``reduce_losses``) with one AdamW over all heads. Returns ({task: state}, [{task: loss}])."""
    dim = next(iter(states.values()))["0.weight"].shape[0]
    heads = nn.ModuleDict({t: make_reference_head_any(dim, c, hidden, dropout_p, states[t]) for t, c in tasks.items()}).train()
    crit = {t: nn.CrossEntropyLoss(weight=None if class_weights.get(t) is None else class_weights[t].double())
            for t in tasks}
    opt = torch.optim.AdamW(heads.parameters(), lr=lr, weight_decay=weight_decay)
    history = []
    for x, y, masks in batches:
        opt.zero_grad(set_to_none=True)
        total, losses = None, {}
        for i, t in enumerate(tasks):
            heads[t][1].keep = masks[i][0]
            if hidden:
                heads[t][4].keep = masks[i][1]
            logits = heads[t](x.double())
            m = y[:, i] != -1
            if m.any():
                lt = crit[t](logits[m], y[m, i])
                losses[t] = float(lt.detach())
                total = task_weights[t] * lt if total is None else total + task_weights[t] * lt
            else:
                losses[t] = None
        if total is not None:
            total.backward()
            opt.step()
        history.append(losses)
    out = {t: {k: (v.float() if v.is_floating_point() else v) for k, v in heads[t].state_dict().items()} for t in tasks}
    return out, history


# ---- dropout keep mask -----------------------------------------------------------------------------------------------
# The reference draws nn.Dropout's mask from torch's global generator (probing/models/base_probe.py:7-22 builds
# ``nn.Dropout(p)``); nothing downstream depends on WHICH mask is drawn, only on it being Bernoulli(1 - p) per element.
# The CUDA path draws it with a counter-based generator (csrc/probe.cu dropout_mask_kernel) so that it is the same for
# any sharding of the batch; this is that generator in numpy, element for element, for the bit-exact parity test.
def _splitmix64(z):
    import numpy as np

    z = (z + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def keep_mask_reference(rows: int, cols: int, row0: int, key: int, dropout_p: float) -> torch.Tensor:
    """uint8 [rows, cols]: keep[r, c] = (high 32 bits of splitmix64(splitmix64(key) ^ ((row0 + r) * cols + c))) >= p * 2^32."""
    import numpy as np

    with np.errstate(over="ignore"):
        k = _splitmix64(np.array([key & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0]
        idx = (np.uint64(row0) * np.uint64(cols) + np.arange(rows * cols, dtype=np.uint64))
        u = (_splitmix64(k ^ idx) >> np.uint64(32)).astype(np.uint64)
    t = float(np.float32(dropout_p)) * 4294967296.0
    thr = 0xFFFFFFFF if t >= 4294967295.0 else int(t)
    return torch.from_numpy((u >= np.uint64(thr)).astype(np.uint8).reshape(rows, cols))
