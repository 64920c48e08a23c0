"""Serving ``rank()`` batch path: one user's features against N candidate items -> N scores.

The reference (2.FM/OnlineServer.py:77-101) builds a dict of Python lists -- the user's feature ids repeated N times,
each candidate's feature ids appended one by one -- turns every list into a ``tf.constant`` and calls the ranking
model.  Here the item profiles live in ONE pinned int64 matrix, a request is assembled into a pinned ``[F, N]`` id block
with numpy indexing (no per-item Python loop over features), moved with one async copy and scored by the fused
gather + interaction kernel; the scores come back with one D2H copy.  Same signature and result shape
(``{item_id: score}``) as the reference method.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .runtime import IdsBatch


class Ranker:
    def __init__(self, layer, user_profile: Dict[str, Sequence[int]], item_profile: Dict[str, Sequence[int]],
                 feature_names: Optional[Sequence[str]] = None, user_feature_num: Optional[int] = None, max_items: int = 4096):
        """``user_profile`` / ``item_profile``: id -> encoded feature ids, as the reference's user_profile.json /
        item_profile.json (2.FM/DataGenerator.py:95-102); ``feature_names`` = user features then item features
        (2.FM/OnlineServer.py:83-84), default: the layer's."""
        self.layer = layer
        self.rt = layer.rt
        names = list(feature_names or layer.feature_names)
        assert list(layer.feature_names) == names, "feature order must match the ranking layer's"
        some_user = next(iter(user_profile.values()))
        self.nu = int(user_feature_num if user_feature_num is not None else len(some_user))
        self.ni = len(names) - self.nu
        self.user_profile = {str(k): np.asarray(v, dtype=np.int64) for k, v in user_profile.items()}
        self.item_index = {str(k): i for i, k in enumerate(item_profile)}
        items = np.asarray([list(v) for v in item_profile.values()], dtype=np.int64).reshape(len(item_profile), self.ni)
        self.items = torch.from_numpy(np.ascontiguousarray(items.T))                  # [ni, n_items]: gather columns
        self.max_items = int(max_items)
        F = len(names)
        self._host = torch.zeros((F, self.max_items), dtype=torch.int64).pin_memory()
        self._dev = torch.zeros((F, self.max_items), dtype=torch.int64, device=self.rt.device)
        self._out = torch.zeros((self.max_items,), dtype=torch.float32).pin_memory()

    def rank(self, user_id, retrieval_result: Sequence) -> Dict[str, float]:
        n = len(retrieval_result)
        if n == 0:
            return {}
        assert n <= self.max_items, "more candidates than max_items"
        try:
            u = self.user_profile[str(user_id)]
        except KeyError:
            raise KeyError(f"unknown user_id {user_id!r}") from None
        try:
            idx = np.fromiter((self.item_index[str(i)] for i in retrieval_result), dtype=np.int64, count=n)
        except KeyError as e:
            raise KeyError(f"unknown item_id {e.args[0]!r}") from None
        host = self._host.numpy()
        host[: self.nu, :n] = u[:, None]                                   # the user's features, repeated per candidate
        host[self.nu:, :n] = self.items.numpy()[:, idx]                    # every candidate's features
        F = host.shape[0]
        self._dev[:, :n].copy_(self._host[:, :n], non_blocking=True)
        ids = IdsBatch(self.rt, self._dev, n, F, 1, 1, self.max_items, 1)  # field-major view of the first n columns
        prob = self.layer(ids)["output"]                                    # [n, 1]
        self._out[:n].copy_(prob.reshape(-1), non_blocking=True)
        torch.cuda.current_stream(self.rt.device).synchronize()
        scores = self._out[:n].tolist()
        return {str(item): s for item, s in zip(retrieval_result, scores)}
