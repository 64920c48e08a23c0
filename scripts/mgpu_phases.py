"""Phase timing of the peer-sharded DeepFM step (eager, CUDA events on the launch stream), under torchrun."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import etr_b200  # noqa: F401
    from etr_b200 import CustomLayers as L
    from etr_b200.runtime import bce_forward_backward
    from etr_b200._lib import check
    dev = torch.device("cuda", local)
    B, F, K, C = 65536, 26, 16, 13
    V = int(sum(bench.CRITEO_CARDS))
    names = [f"C{i + 1}" for i in range(F)]
    cont = [f"I{i + 1}" for i in range(C)]
    layer = L.DeepFMRankingLayer(names, feature_dims=V, embedding_dims=K, continuous_features=cont, seed=1,
                                 check_ids=False, mlp_precision="bf16", shard="peer")
    rt = layer.rt
    host = bench.make_batches(4, B, "zipf", seed=bench.SEED + 17 * rank)
    batches = []
    for X, Xc, y in host:
        d = {n: torch.from_numpy(np.ascontiguousarray(X[:, i])).to(dev) for i, n in enumerate(names)}
        d.update({n: torch.from_numpy(np.ascontiguousarray(Xc[:, i])).to(dev) for i, n in enumerate(cont)})
        batches.append((d, torch.from_numpy(y).to(dev)))
    tr = L.Trainer(layer, lr=1e-3)
    peer = layer.peer
    marks = []
    last = {}

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((name, e))

    def step(d, y):
        marks.clear()
        mark("start")
        out = layer(d, training=True)["output"]
        mark("forward (ids assemble + gather over NVLink + MLP)")
        loss, dlogit = bce_forward_backward(rt, out.reshape(-1), y)
        dlogit.mul_(1.0 / world)
        grads = layer.backward(dlogit)
        last["grads"] = grads
        mark("bce + MLP backward")
        for g in grads:
            g.fused.reduce()
        mark("fused backward export (deferred form, no remote reads)")
        for g in grads:
            peer.ensure_mailbox(g.fused.plan.n_slots)
            peer.push(g.fused.plan.unique_ids, g.fused.plan.counts, g.fused.plan.n_slots, g.fused.unique_grad)
        peer.allreduce_push(layer.params.grad)
        mark("push rows + dense grads")
        peer.barrier()
        mark("barrier 1")
        peer.allreduce_sum(layer.params.grad)
        check(rt.lib.etr_adam_step_begin(rt.ctx, tr.state.data_ptr(), tr.lr, tr.b1, tr.b2, rt.stream))
        layer.params.adam_step(0.0, tr.state[1:], tr.b1, tr.b2, tr.eps)
        mark("dense sum + adam")
        peer.apply_mailbox(tr.state[1:], tr.b1, tr.b2, tr.eps, 0)
        mark("owner: accumulate regions + touched-row adam")
        peer.barrier()
        mark("barrier 2")

    for i in range(6):
        step(*batches[i % 4])
    torch.cuda.synchronize()
    dist.barrier()
    acc = {}
    for i in range(8):
        step(*batches[i % 4])
        torch.cuda.synchronize()
        for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
            acc.setdefault(n1, []).append(e0.elapsed_time(e1) * 1e3)
    nu = int(last["grads"][0].fused.plan.counts[0].item())
    if rank == 0:
        tot = 0.0
        for n, v in acc.items():
            m = sorted(v)[len(v) // 2]
            tot += m
            print(f"{m:9.1f} us  {n}")
        print(f"{tot:9.1f} us  total (eager, launch gaps included); unique rows of this rank's batch: {nu} of {B * F}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
