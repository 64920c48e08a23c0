"""Phase timing of the peer-sharded DeepFM step (de-duplicated request/serve exchange) (eager, CUDA events on the launch stream), under torchrun."""
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

    import ctypes as C
    from etr_b200 import _lib
    from etr_b200.runtime import FusedFMGrad, IdsBatch, SparsePlan, cast_bf16, gather_fm_forward, gemm_bf16_tn
    from etr_b200.sharded import PeerSlotGrad, VirtualTable
    names_c = cont

    def step(d, y):
        """the fused train step of DeepFMRankingLayer.train_forward_backward + Trainer, phase by phase"""
        marks.clear()
        lay, P = layer, layer.params
        k, F = lay.embedding_dims, len(lay.feature_names)
        mark("start")
        ids = lay._ids(d, lay.feature_names)
        cont_m = lay._cont(d, names_c)
        B = ids.B
        col0 = lay.front_pad + len(names_c)
        n_in = col0 + F * k
        plan = SparsePlan(rt, ids, peer.rows)
        mark("ids assemble + sorted plan (CUB sort + unique)")
        peer.ensure_mailbox(plan.n_slots)
        peer._cur_set = 0
        mb, cap, ld, W = peer._mb, peer.cap, peer.ld, world
        slot_of_u = rt.empty((plan.n_slots,), torch.int32)
        check(rt.lib.etr_shard_request(rt.ctx, plan.unique_ids.data_ptr(), plan.counts.data_ptr(), plan.n_slots, W, cap,
                                       mb["ids_ptrs"][0], mb["counts_ptrs"][0], mb["local_cnt"][0].data_ptr(),
                                       slot_of_u.data_ptr(), rt.stream))
        mark("request: unique ids -> owners' mailboxes")
        peer.barrier()
        mark("barrier")
        if peer.owner_prep:
            peer._launch_owner_prep(0)                # side stream: pairs up the entries per row while the forward runs
        t = peer.local.desc()
        check(rt.lib.etr_shard_serve(rt.ctx, C.byref(t), mb["ids_t"][0].data_ptr(), mb["counts_t"][0].data_ptr(), W, cap,
                                     mb["resp_ptrs"], ld, rt.stream))
        mark("serve: owners write the rows into the requesters' buffers")
        peer.barrier()
        mark("barrier ")
        vid = rt.empty((B * F,), torch.int64)
        check(rt.lib.etr_shard_vid_map(rt.ctx, plan.sorted_bag.data_ptr(), plan.seg_start.data_ptr(),
                                       plan.counts.data_ptr(), plan.n_slots, slot_of_u.data_ptr(), vid.data_ptr(), rt.stream))
        mark("virtual ids (occurrence -> response row)")
        x = rt.empty((B, n_in), torch.bfloat16)
        fm_logit, sumv = rt.empty((B,)), rt.empty((B, k))
        gather_fm_forward(VirtualTable(rt, mb["resp_t"], peer.width), k, True, IdsBatch(rt, vid, B, F, 1, F, 1, 1),
                          bias=lay.bias, logit=fm_logit, sumv=sumv, flat=x, flat_col0=col0, cont=cont_m)
        mark("K1 gather + FM on the response buffer")
        m1, m2 = lay.MLP_layer1, lay.MLP_layer2
        k0 = P.full(f"{m1.name}/kernel_0")
        y1 = rt.empty((B, 32))
        gemm_bf16_tn(rt, x, cast_bf16(rt, k0, transpose=True), y1, B, 32, n_in, bias=P[f"{m1.name}/bias_0"], act="relu")
        prob, dlogit, d1, loss = rt.empty((B, 1)), rt.empty((B,)), rt.empty((B, 32)), rt.empty((1,))
        check(rt.lib.etr_deepfm_tail_train(
            rt.ctx, y1.data_ptr(), fm_logit.data_ptr(), y.data_ptr(), B, P[f"{m1.name}/kernel_1"].data_ptr(),
            P[f"{m1.name}/bias_1"].data_ptr(), P[f"{m2.name}/kernel_0"].data_ptr(), P[f"{m2.name}/bias_0"].data_ptr(),
            1.0 / world, prob.data_ptr(), dlogit.data_ptr(), d1.data_ptr(), loss.data_ptr(), P.g("bias").data_ptr(),
            P.g(f"{m1.name}/kernel_1").data_ptr(), P.g(f"{m1.name}/bias_1").data_ptr(),
            P.g(f"{m2.name}/kernel_0").data_ptr(), P.g(f"{m2.name}/bias_0").data_ptr(), rt.stream))
        dx = rt.empty((B, n_in), torch.bfloat16)
        check(rt.lib.etr_mlp_skinny_backward(rt.ctx, x.data_ptr(), n_in, d1.data_ptr(), k0.data_ptr(), B, n_in, 32,
                                             dx.data_ptr(), n_in, P.gfull(f"{m1.name}/kernel_0").data_ptr(),
                                             P.g(f"{m1.name}/bias_0").data_ptr(), rt.stream))
        mark("MLP layer 1 (tcgen05) + tail/loss fwd+bwd (K7c) + layer-1 backward (K7b)")
        g = PeerSlotGrad(peer, FusedFMGrad(lay.table, ids, k, dlogit, sumv, dx, col0, plan=plan), slot_of_u)
        last["grads"] = [g]
        g.push()
        mark("fused FM backward, rows written straight into the owners' mailbox slots")
        peer.allreduce_push(P.grad)
        mark("dense grads -> peers")
        peer.barrier()
        mark("barrier  ")
        peer.allreduce_sum(P.grad)
        check(rt.lib.etr_adam_step_begin(rt.ctx, tr.state.data_ptr(), tr.lr, tr.b1, tr.b2, rt.stream))
        P.adam_step(0.0, tr.state[1:], tr.b1, tr.b2, tr.eps)
        mark("dense sum + adam")
        peer.apply_mailbox(tr.state[1:], tr.b1, tr.b2, tr.eps, 0)
        mark("owner apply (one pass over the request entries; ETR_PEER_OWNER_PREP=0: accumulate regions + touched-row adam)")
        peer.barrier()
        mark("barrier   ")

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
