"""Training-loop utilities: `ModelEMA` (ultralytics/utils/torch_utils.py:495-530) with the update as ONE CUDA
launch (`specyolo_ema_update`, csrc/ema.cu) instead of a Python loop of ~1 500 ATen launches per step; bit-identical to
the reference loop (same three fp32 roundings per element)."""
from __future__ import annotations

import math
from copy import deepcopy
from typing import Optional

import numpy as np
import torch

from .. import _lib
from .._lib import check

_CHUNK = 16384


def _de_parallel(model):
    return model.module if isinstance(model, (torch.nn.parallel.DataParallel, torch.nn.parallel.DistributedDataParallel)) else model


class _EmaPlan:
    """Device-resident pointer / chunk tables for one (ema, model) pair; rebuilt when a tensor is re-allocated."""

    def __init__(self, ema_sd, model_sd):
        self.ema_items = [(k, v) for k, v in ema_sd.items() if v.dtype.is_floating_point]
        pairs = [(v, model_sd[k]) for k, v in self.ema_items]
        self.fallback = [(v, m) for v, m in pairs if not (v.dtype == m.dtype == torch.float32 and v.is_cuda and m.is_cuda
                                                           and v.is_contiguous() and m.is_contiguous())]
        fast = [(v, m) for v, m in pairs if v.dtype == m.dtype == torch.float32 and v.is_cuda and m.is_cuda
                and v.is_contiguous() and m.is_contiguous() and v.numel() > 0]
        self.key = tuple((v.data_ptr(), m.data_ptr(), v.numel()) for v, m in pairs)
        self.n = len(fast)
        if not fast:
            return
        dev = fast[0][0].device
        numel = np.asarray([v.numel() for v, _ in fast], dtype=np.int64)
        ct, co = [], []
        for t, n in enumerate(numel):
            offs = np.arange(0, n, _CHUNK, dtype=np.int64)
            ct.append(np.full(len(offs), t, dtype=np.int32))
            co.append(offs)
        self.ema_ptrs = torch.tensor([v.data_ptr() for v, _ in fast], dtype=torch.int64, device=dev)
        self.model_ptrs = torch.tensor([m.data_ptr() for _, m in fast], dtype=torch.int64, device=dev)
        self.numel = torch.from_numpy(numel).to(dev)
        self.chunk_tensor = torch.from_numpy(np.concatenate(ct)).to(dev)
        self.chunk_off = torch.from_numpy(np.concatenate(co)).to(dev)
        self.nchunks = int(self.chunk_tensor.numel())


def ema_update(ema_sd, model_sd, d: float, plan: Optional[_EmaPlan] = None) -> _EmaPlan:
    """v = v * d + (1 - d) * m over every floating-point entry of the state_dicts (torch_utils.py:520-524).  `ema_sd` may be
    a callable returning the state_dict: it is then only evaluated when the cached plan has to be rebuilt."""
    if plan is not None and callable(ema_sd):       # the EMA copy is private: its tensors are checked through the plan's own list
        key = tuple((v.data_ptr(), model_sd[k].data_ptr(), v.numel()) for k, v in plan.ema_items)
    else:
        ema_sd = ema_sd() if callable(ema_sd) else ema_sd
        key = tuple((v.data_ptr(), model_sd[k].data_ptr(), v.numel()) for k, v in ema_sd.items() if v.dtype.is_floating_point)
    if plan is None or plan.key != key:
        ema_sd = ema_sd() if callable(ema_sd) else ema_sd
        plan = _EmaPlan(ema_sd, model_sd)
    if plan.n:
        _lib.init_device()
        d32 = float(np.float32(d))
        omd32 = float(np.float32(1 - d))             # ATen converts the Python scalar (1 - d), a double, to fp32
        check(_lib.load().specyolo_ema_update(plan.ema_ptrs.data_ptr(), plan.model_ptrs.data_ptr(), plan.numel.data_ptr(),
                                              plan.chunk_tensor.data_ptr(), plan.chunk_off.data_ptr(), plan.nchunks, _CHUNK,
                                              d32, omd32, _lib.stream_ptr()))
    for v, m in plan.fallback:                       # fp16 / CPU / non-contiguous entries: the reference's own arithmetic
        v *= d
        v += (1 - d) * m.detach()
    return plan


class ModelEMA:
    """Exponential moving average of everything in the model state_dict (ultralytics/utils/torch_utils.py:495-530)."""

    def __init__(self, model, decay=0.9999, tau=2000, updates=0):
        self.ema = deepcopy(_de_parallel(model)).eval()
        self.updates = updates
        self.decay = lambda x: decay * (1 - math.exp(-x / tau))
        for p in self.ema.parameters():
            p.requires_grad_(False)
        self.enabled = True
        self._plan = None

    def update(self, model):
        if self.enabled:
            self.updates += 1
            d = self.decay(self.updates)
            self._plan = ema_update(self.ema.state_dict, _de_parallel(model).state_dict(), d, self._plan)

    def update_attr(self, model, include=(), exclude=("process_group", "reducer")):
        if self.enabled:
            for k, v in model.__dict__.items():      # copy_attr, torch_utils.py:437-443
                if (len(include) and k not in include) or k.startswith("_") or k in exclude:
                    continue
                setattr(self.ema, k, v)
