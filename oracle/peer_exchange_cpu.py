"""CPU restatement of the peer-memory form of the row-sharded tables (no reference type) --
TEST INFRASTRUCTURE ONLY.  It states, in plain torch-CPU over a gloo group, what the device kernels of
``shard_kernels.cu`` / ``fm_fused_apply.cu`` compute, so that the protocol itself (de-duplicated request / serve
exchange, virtual ids, deferred FM gradient, rank-ordered stamped accumulation, owner-side finish) is checked with
world_size 2 on CPU against the unsharded oracle:

  request   etr_shard_request            unique ids -> (owner = id mod G, local row = id div G), one slot each
  serve     etr_shard_serve              the owner copies the requested rows into the requester's response buffer
  vid       etr_shard_vid_map            occurrence -> response-buffer row
  export    etr_fm_fused_backward_push   per unique row the DEFERRED gradient [P, sum_g], P = sum g S + sum dflat
  apply     etr_shard_mailbox_accumulate the G source regions are added in rank order into an accumulator whose rows
            + etr_shard_touched_adam     carry a stamp (first touch of a step overwrites), then dv = P - v * sum_g
  apply'    etr_shard_owner_prep         the same sums from the REQUESTS alone: every entry claims its row (any one entry
            + etr_shard_owner_apply      wins = the leader, the others register their slot with it); after the gradient
                                         barrier each leader adds its row's entries in ascending source order
"""
from typing import List, Tuple

import torch
import torch.distributed as dist


def _a2a(chunks: List[torch.Tensor], like: torch.Tensor, group=None) -> List[torch.Tensor]:
    """variable-size all-to-all of row blocks (counts first, then the payload)"""
    world = dist.get_world_size(group)
    counts = torch.tensor([c.shape[0] for c in chunks], dtype=torch.int64)
    rc = torch.empty_like(counts)
    dist.all_to_all_single(rc, counts, group=group)
    width = tuple(like.shape[1:])
    send = torch.cat(chunks) if chunks else like.new_empty((0,) + width)
    recv = like.new_empty((int(rc.sum()),) + width)
    dist.all_to_all_single(recv, send, rc.tolist(), counts.tolist(), group=group)
    out, o = [], 0
    for g in range(world):
        out.append(recv[o:o + int(rc[g])])
        o += int(rc[g])
    return out


class PeerExchangeCpu:
    def __init__(self, world: int, rank: int, shard: torch.Tensor, k: int, group=None):
        """``shard`` [local_rows, k+1]: rows rank, rank+G, ... of the global [V, k+1] table (v_0..v_{k-1}, w)."""
        self.world, self.rank, self.k, self.group = world, rank, k, group
        self.shard = shard
        self.gacc = torch.zeros((shard.shape[0], k + 2), dtype=shard.dtype)     # [P.., sum_g, stamp]
        self.epoch = 0

    # -- forward ---------------------------------------------------------------
    def exchange_forward(self, X: torch.Tensor) -> Tuple[torch.Tensor, dict]:
        """X [B,F] global ids -> rows [B,F,k+1] (bit-exact copies of the owners' rows) + the routing state."""
        G = self.world
        uids, inverse = torch.unique(X.reshape(-1), sorted=True, return_inverse=True)      # the sorted plan
        owner, lrow = uids % G, uids // G
        req = [lrow[owner == g] for g in range(G)]                         # request regions, sorted order kept
        slot_of_u = torch.empty_like(uids)
        base = 0
        for g in range(G):
            n = int((owner == g).sum())
            slot_of_u[owner == g] = base + torch.arange(n)
            base += n
        got = _a2a(req, lrow.new_empty((0,)), self.group)                  # owner side: one region per source
        served = [self.shard[r] for r in got]                              # serve
        resp = torch.cat(_a2a(served, self.shard, self.group))             # response buffer, owner-major
        vid = slot_of_u[inverse]                                           # virtual ids
        rows = resp[vid].reshape(X.shape + (self.k + 1,))
        return rows, {"uids": uids, "inverse": inverse, "owner": owner, "regions": got, "X": X}

    # -- backward ----------------------------------------------------------------
    def export_deferred(self, st: dict, rows: torch.Tensor, g: torch.Tensor, dflat: torch.Tensor) -> torch.Tensor:
        """per unique id the deferred gradient row [P (k), sum_g]; g [B] = dL/dz, dflat [B,F,k] = dL/d(flat emb)"""
        B, F = st["X"].shape
        S = rows[..., : self.k].sum(1)                                     # [B,k]
        per_occ = g[:, None, None] * S[:, None, :] + dflat                 # [B,F,k]
        out = torch.zeros((st["uids"].numel(), self.k + 1), dtype=rows.dtype)
        out[:, : self.k].index_add_(0, st["inverse"], per_occ.reshape(B * F, self.k))
        out[:, self.k].index_add_(0, st["inverse"], g[:, None].expand(B, F).reshape(-1))
        return out

    def push_and_apply(self, st: dict, deferred: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """rows travel back through the request regions; the owner adds the regions in RANK order into the stamped
        accumulator and finishes dv = P - v * sum_g.  Returns (touched local rows, their finished gradients)."""
        G, k = self.world, self.k
        sent = [deferred[st["owner"] == g_] for g_ in range(G)]
        recv = _a2a(sent, deferred, self.group)
        self.epoch += 1
        touched: List[int] = []
        for src in range(G):                                               # rank order
            for lr, row in zip(st["regions"][src].tolist(), recv[src]):
                if self.gacc[lr, k + 1] != self.epoch:                     # first touch of this step: overwrite
                    self.gacc[lr, : k + 1] = row
                    self.gacc[lr, k + 1] = self.epoch
                    touched.append(lr)
                else:
                    self.gacc[lr, : k + 1] += row
        t = torch.tensor(touched, dtype=torch.int64)
        acc = self.gacc[t]
        grad = torch.empty((t.numel(), k + 1), dtype=self.shard.dtype)
        grad[:, :k] = acc[:, :k] - self.shard[t, :k] * acc[:, k:k + 1]    # the owner finishes the FM gradient
        grad[:, k] = acc[:, k]
        return t, grad

    # -- the request-driven form of the same apply (etr_shard_owner_prep / etr_shard_owner_apply) ---------------------
    @staticmethod
    def owner_pairing(regions: List[torch.Tensor], rng=None):
        """From the request regions alone: leader[(src, slot)] -> {src': slot'} of the other entries of the same row.
        WHICH entry of a row leads is decided by a race on the device (a 64-bit CAS); ``rng`` picks an arbitrary one here --
        the sums below must not depend on it."""
        by_row = {}
        for src, reg in enumerate(regions):
            for slot, lr in enumerate(reg.tolist()):
                by_row.setdefault(lr, []).append((src, slot))
        leaders = {}
        for lr, ents in by_row.items():
            lead = ents[int(rng.integers(len(ents)))] if rng is not None else ents[0]
            leaders[lead] = (lr, {s: q for s, q in ents})
        return leaders

    def push_and_apply_paired(self, st: dict, deferred: torch.Tensor, rng=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """``push_and_apply`` without the accumulator: one pass over the leaders, contributors added in ascending source
        order (the order of the region-by-region accumulation, so the sums are bit-identical to it)."""
        G, k = self.world, self.k
        sent = [deferred[st["owner"] == g_] for g_ in range(G)]
        recv = _a2a(sent, deferred, self.group)
        rows, grads = [], []
        for (src, slot), (lr, ents) in self.owner_pairing(st["regions"], rng).items():
            acc = None
            for s in sorted(ents):                                         # ascending source order
                acc = recv[s][ents[s]].clone() if acc is None else acc + recv[s][ents[s]]
            gr = acc.clone()
            gr[:k] = acc[:k] - self.shard[lr, :k] * acc[k]
            rows.append(lr)
            grads.append(gr)
        return torch.tensor(rows, dtype=torch.int64), torch.stack(grads) if grads else deferred.new_empty((0, k + 1))
