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
