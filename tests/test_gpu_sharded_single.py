"""The peer-memory sharding machinery on ONE GPU (world = 1: every id is owned locally, the peer buffers are the
rank's own): request / serve / virtual ids / mailbox export / stamped accumulator / touched-row Adam / device barrier /
peer all-reduce all run, and the result must equal the unsharded layer.  (N = 2, 4, 8 runs: tests/mgpu_sharded_check.py
under torchrun, profiles/r01_mgpu.md.)"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from etr_b200 import CustomLayers
    return CustomLayers


def _batches(n, B, F, C_, V, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        X = (rng.random((B, F)) ** 3 * V).astype(np.int64)
        Xc = rng.normal(size=(B, C_)).astype(np.float32)
        y = (rng.random(B) < 0.3).astype(np.float32)
        out.append((X, Xc, y))
    return out


@pytest.mark.parametrize("mode,prec,graph", [("peer", "bf16", False), ("peer", "bf16", True), ("peer-pull", "fp32", False),
                                             ("peer-pull", "fp32", True)])
def test_peer_sharded_world1_equals_unsharded(L, mode, prec, graph):
    F, k, V, C_, B = 26, 16, 50021, 13, 2048
    names, cont = [f"f{i}" for i in range(F)], [f"c{i}" for i in range(C_)]
    sh = L.DeepFMRankingLayer(names, V, k, continuous_features=cont, seed=3, shard=(mode, 1, 0), check_ids=False,
                              mlp_precision=prec)
    full = L.DeepFMRankingLayer(names, V, k, continuous_features=cont, seed=3, check_ids=False, mlp_precision=prec)
    sh.peer.load_global(full.table.data[:, : k + 1])
    sh.params.value.copy_(full.params.value)
    tr_s, tr_f = L.Trainer(sh, lr=1e-2, graph=graph), L.Trainer(full, lr=1e-2)
    data = _batches(6 if graph else 3, B, F, C_, V, 11)

    def feed(X, Xc):
        d = {n: torch.tensor(X[:, i]).cuda() for i, n in enumerate(names)}
        d.update({n: torch.tensor(Xc[:, i]).cuda() for i, n in enumerate(cont)})
        return d

    X, Xc, _ = data[0]
    assert torch.equal(sh(feed(X, Xc))["output"], full(feed(X, Xc))["output"])          # forward bit-exact
    for step, (X, Xc, y) in enumerate(data):
        ls = float(tr_s.train_step(feed(X, Xc), torch.tensor(y).cuda()).item())
        lf = float(tr_f.train_step(feed(X, Xc), torch.tensor(y).cuda()).item())
        assert abs(ls - lf) < (1e-5 if prec == "fp32" else 1e-4), (step, ls, lf)
        if step == 0:
            torch.cuda.synchronize()
            assert (sh.table.data[:V] - full.table.data).abs().max().item() < 2e-6
            assert (sh.params.value - full.params.value).abs().max().item() < 2e-6
    sh.rt.poll_error()
    tol = 2e-5 if prec == "fp32" else 5e-3          # bf16 tower: weight-rounding flips amplify 1e-7 differences
    assert (sh.table.data[:V] - full.table.data).abs().max().item() < tol
    assert (sh.params.value - full.params.value).abs().max().item() < tol


def test_peer_barrier_and_allreduce_world1(L):
    from etr_b200.runtime import Runtime
    from etr_b200.sharded import PeerShardedTable
    rt = Runtime.get()
    T = PeerShardedTable(rt, 1000, 17, 1, 0)
    v = torch.randn(1001, device=rt.device)
    ref = v.clone()
    for _ in range(3):
        T.allreduce_push(v)
        T.barrier()
        T.allreduce_sum(v)
    rt.poll_error()
    assert torch.equal(v, ref) and int(T._epoch.item()) == 3


@pytest.mark.parametrize("model", ["dcn_vec", "dcn_matrix", "pnn", "fm"])
def test_a2a_sharded_world1_equals_unsharded(L, model):
    """row-sharded tables (all-to-all form) for every model family that gathers from one table: at world 1 the route is
    the identity, so forward is bit-exact and 3 train steps give the same weights as the unsharded layer."""
    F, k, V, C_, B = 10, 16, 6007, 3, 512
    names, cont = [f"f{i}" for i in range(F)], [f"c{i}" for i in range(C_)]

    def make(shard):
        if model.startswith("dcn"):
            return L.DeepCrossNetworkLayer(names, cont, feature_dims=V, embedding_dims=k, type="vec" if model == "dcn_vec" else "matrix",
                                           seed=5, shard=shard, check_ids=False)
        if model == "pnn":
            return L.PNNRankingLayer(names, V, k, seed=5, shard=shard, check_ids=False)
        return L.FMRankingLayer(names, V, k, seed=5, shard=shard, check_ids=False)

    sh, full = make((1, 0)), make(None)
    assert sh.shard is not None and full.shard is None
    sh.shard.load_global(full.table.data[:, : full.table.width])
    sh.params.value.copy_(full.params.value)
    tr_s, tr_f = L.Trainer(sh, lr=1e-2), L.Trainer(full, lr=1e-2)
    for step, (X, Xc, y) in enumerate(_batches(3, B, F, C_, V, 5)):
        d = {n: torch.tensor(X[:, i]).cuda() for i, n in enumerate(names)}
        if model.startswith("dcn"):
            d.update({n: torch.tensor(Xc[:, i]).cuda() for i, n in enumerate(cont)})
        if step == 0:
            assert torch.equal(sh(d)["output"], full(d)["output"])
        ls = float(tr_s.train_step(d, torch.tensor(y).cuda()).item())
        lf = float(tr_f.train_step(d, torch.tensor(y).cuda()).item())
        assert abs(ls - lf) < 1e-5, (step, ls, lf)
    w = full.table.width
    assert (sh.table.data[:V, :w] - full.table.data[:, :w]).abs().max().item() < 2e-5
    assert (sh.params.value - full.params.value).abs().max().item() < 2e-5


@pytest.mark.parametrize("record", [True, False])
def test_owner_prep_apply_equals_region_accumulation(L, record):
    """Owner side of the peer step on ONE GPU with a synthetic 4-source mailbox whose regions overlap heavily: the
    request-driven form (etr_shard_owner_prep on the requests, then the one-pass etr_shard_owner_apply) must be
    BIT-identical to the region-by-region stamped accumulator + touched-row Adam, and match a torch statement of
    'sum the sources in rank order, finish dv = P - v * sum_g, row-wise Adam' -- two steps (the pass cleans its masks)."""
    import ctypes as C
    from etr_b200 import _lib
    from etr_b200._lib import check
    from etr_b200.runtime import Runtime, EmbeddingTable
    rt = Runtime.get()
    W, cap, rows, ld, k = 4, 512, 1000, 20, 16
    g = torch.Generator(device="cpu").manual_seed(5)

    def table():
        if record:
            t = EmbeddingTable(rt, rows, k + 1, record=True)
        else:
            t = EmbeddingTable(rt, rows, k + 1)
            assert t.stride == 20
        return t

    A, Bt = table(), table()
    init = torch.rand((rows, k + 1), generator=g) * 0.1 - 0.05
    for t in (A, Bt):
        t.data[:, : k + 1] = init.cuda()
        _ = t.m, t.v
    own = {"map": rt.zeros((rows,), torch.int64), "step": rt.zeros((1,), torch.int32),
           "mask": rt.zeros((W * cap,), torch.int32), "others": rt.empty((W * cap, W), torch.int32)}
    gacc, epoch = rt.zeros((rows, ld)), rt.zeros((1,), torch.int32)
    touched, n_touched = rt.empty((W * cap,), torch.int32), rt.zeros((1,), torch.int32)
    lr_t = torch.tensor([1e-2], device=rt.device)
    ref_var, ref_m, ref_v = init.double().clone(), torch.zeros(rows, k + 1).double(), torch.zeros(rows, k + 1).double()
    for step in range(2):
        req = torch.zeros((W, cap), dtype=torch.int64)
        grads = torch.zeros((W, cap, ld))
        counts = torch.tensor([300, 0, 512, 170][::1 if step == 0 else -1], dtype=torch.int32)
        for s in range(W):
            n = int(counts[s])
            req[s, :n] = torch.randperm(rows, generator=g)[:n]
            grads[s, :n, : k + 1] = torch.randn((n, k + 1), generator=g) * 0.1
        req_d, grads_d, counts_d = req.cuda(), grads.cuda(), counts.cuda()
        epoch += 1
        # A: the request-driven form
        check(rt.lib.etr_shard_owner_prep(rt.ctx, req_d.data_ptr(), counts_d.data_ptr(), W, cap, rows, own["map"].data_ptr(),
                                          own["step"].data_ptr(), own["mask"].data_ptr(), own["others"].data_ptr(), rt.stream))
        ta = A.desc()
        check(rt.lib.etr_shard_owner_apply(rt.ctx, C.byref(ta), A.m.data_ptr(), A.v.data_ptr(), req_d.data_ptr(),
                                           counts_d.data_ptr(), grads_d.data_ptr(), W, cap, ld, own["mask"].data_ptr(),
                                           own["others"].data_ptr(), k, lr_t.data_ptr(), 0.9, 0.999, 1e-7, rt.stream))
        # B: region-by-region accumulation + touched-row Adam
        tb = Bt.desc()
        check(rt.lib.etr_shard_mailbox_accumulate(rt.ctx, req_d.data_ptr(), grads_d.data_ptr(), counts_d.data_ptr(), W, cap, ld,
                                                  gacc.data_ptr(), epoch.data_ptr(), touched.data_ptr(), n_touched.data_ptr(),
                                                  W * cap, rt.stream))
        check(rt.lib.etr_shard_touched_adam(rt.ctx, C.byref(tb), Bt.m.data_ptr(), Bt.v.data_ptr(), gacc.data_ptr(), ld,
                                            touched.data_ptr(), n_touched.data_ptr(), W * cap, k, lr_t.data_ptr(), 0.9, 0.999,
                                            1e-7, rt.stream))
        torch.cuda.synchronize()
        rt.poll_error()
        assert int(own["mask"].abs().sum().item()) == 0                 # the pass cleaned up behind itself
        for a, b in ((A.data, Bt.data), (A.m, Bt.m), (A.v, Bt.v)):
            assert torch.equal(a[:, : k + 1], b[:, : k + 1])
        # torch statement (fp64)
        acc = torch.zeros(rows, ld).double()
        hit = torch.zeros(rows, dtype=torch.bool)
        for s in range(W):
            n = int(counts[s])
            acc.index_add_(0, req[s, :n], grads[s, :n].double())
            hit[req[s, :n]] = True
        gr = acc[:, : k + 1].clone()
        gr[:, :k] -= ref_var[:, :k] * acc[:, k: k + 1]
        m2 = 0.9 * ref_m + 0.1 * gr
        v2 = 0.999 * ref_v + 0.001 * gr * gr
        var2 = ref_var - 1e-2 * m2 / (v2.sqrt() + 1e-7)
        ref_m, ref_v, ref_var = torch.where(hit[:, None], m2, ref_m), torch.where(hit[:, None], v2, ref_v), \
            torch.where(hit[:, None], var2, ref_var)
        assert (A.data[:, : k + 1].cpu().double() - ref_var).abs().max().item() < 1e-5
        assert (A.m[:, : k + 1].cpu().double() - ref_m).abs().max().item() < 1e-6
