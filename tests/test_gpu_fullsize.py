"""Full BASELINE.json sizes (c2: B = 65 536, F = 26, k = 16, V = 33 762 577; c4: F = 39, k = 8, bags <= 50,
B = 8 192), where the CPU oracle would take too long: size-independent properties of the CUDA path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CRITEO_CARDS = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992,
                5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]


@pytest.fixture(scope="module")
def c2():
    from etr_b200 import CustomLayers as L
    V, B, F = int(sum(CRITEO_CARDS)), 65536, 26
    rng = np.random.default_rng(20261)
    cards = np.asarray(CRITEO_CARDS)
    offs = np.concatenate([[0], np.cumsum(cards)[:-1]])
    X = (offs[None, :] + np.floor(cards[None, :] * rng.random((B, F)) ** 3)).astype(np.int64)
    lay = L.FMRankingLayer([f"C{i}" for i in range(F)], feature_dims=V, embedding_dims=16, seed=1, check_ids=True)
    return L, lay, torch.tensor(X).cuda(), V, B, F


def test_c2_gather_rows_bit_exact(c2):
    L, lay, X, V, B, F = c2
    from etr_b200.runtime import embedding_gather
    rows = embedding_gather(lay.table, X[:4096])
    ref = lay.table.data[X[:4096].reshape(-1), :17].reshape(4096, F, 17)
    assert torch.equal(rows, ref)                                    # gathered rows bit-exact at full table size


def test_c2_forward_batch_split_invariance_and_bag_equivalence(c2):
    L, lay, X, V, B, F = c2
    full = lay(X)["output"]
    halves = torch.cat([lay(X[: B // 2])["output"], lay(X[B // 2:])["output"]])
    assert torch.equal(full, halves)                                 # a sample's output does not depend on the batch
    assert torch.equal(full, lay(X.unsqueeze(-1))["output"])         # L = 1 bags == single-hot (bit-exact)
    fm = X.t().contiguous()                                          # field-major ids == row-major ids
    from etr_b200.runtime import IdsBatch
    assert torch.equal(full, lay(IdsBatch(lay.rt, fm, B, F, 1, 1, B, 1))["output"])
    assert torch.isfinite(full).all() and float(full.min()) > 0 and float(full.max()) < 1


def test_c2_plan_and_gradient_checksums(c2):
    L, lay, X, V, B, F = c2
    from etr_b200.runtime import IdsBatch, SparsePlan
    ids = IdsBatch.from_matrix(lay.rt, X)
    plan = SparsePlan(lay.rt, ids, V)
    u = plan.n_unique
    uid = plan.unique_ids[:u]
    seg = plan.seg_start[: u + 1].long()
    assert torch.all(uid[1:] > uid[:-1])                             # strictly ascending unique ids
    assert int(seg[0]) == 0 and int(seg[-1]) == B * F and torch.all(seg[1:] > seg[:-1])
    assert u == torch.unique(X).numel()
    # every sorted occurrence points back at a slot holding that run's id
    flat = X.reshape(-1)
    run_of_pos = torch.repeat_interleave(torch.arange(u, device=X.device), seg[1:] - seg[:-1])
    assert torch.equal(flat[plan.sorted_bag[: B * F].long()], uid[run_of_pos])
    # gradient checksums: linear in dz, and the w column sums to F * sum(dz)
    g = torch.Generator(device="cuda").manual_seed(3)
    dz = torch.randn(B, device="cuda", generator=g) * 1e-3
    lay(X, training=True)
    ids1, rows1 = lay.backward(dz)[0].indexed_slices()
    lay(X, training=True)
    ids2, rows2 = lay.backward(2 * dz)[0].indexed_slices()
    assert torch.equal(ids1, ids2) and torch.equal(ids1, uid)
    assert torch.allclose(rows2, 2 * rows1, rtol=1e-5, atol=1e-9)    # linearity
    w_sum = rows1[:, 16].double().sum().item()
    assert abs(w_sum - F * dz.double().sum().item()) <= 1e-4 * F * dz.abs().double().sum().item()


def test_c2_train_step_touches_only_batch_rows(c2):
    L, lay, X, V, B, F = c2
    before = lay.table.data.clone()
    tr = L.Trainer(lay, lr=1e-3)
    y = (torch.rand(B, device="cuda") < 0.25).float()
    loss = tr.train_step(X, y)
    assert torch.isfinite(loss).all()
    changed = (lay.table.data != before).any(dim=1)
    touched = torch.zeros(V, dtype=torch.bool, device="cuda")
    touched[X.reshape(-1)] = True
    assert torch.equal(changed, touched)                             # row-wise Adam: exactly the unique rows moved
    assert torch.all(lay.table.data[:, 17:] == 0)                    # padding columns stay zero


def test_c4_ffm_bags_full_size_properties():
    """c4: 39 fields x 100 000 ids, k = 8, bags of 1..50 ids, B = 8 192."""
    from etr_b200 import CustomLayers as L
    rng = np.random.default_rng(20263)
    B, F, k, Lm = 8192, 39, 8, 50
    V = F * 100000
    lens = rng.integers(1, Lm + 1, size=(B, F))
    X = np.zeros((B, F, Lm), dtype=np.int64)
    base = (np.arange(F) * 100000)[None, :, None]
    draw = 1 + (rng.random((B, F, Lm)) ** 3 * 99999).astype(np.int64) + base
    mask = np.arange(Lm)[None, None, :] < lens[:, :, None]
    X[mask] = draw[mask]
    Xd = torch.tensor(X).cuda()
    ffm = L.FFMLayer([f"f{i}" for i in range(F)], feature_dims=V, embedding_dims=k, pad_id=0, seed=2)
    fw = L.FwFMLayer([f"f{i}" for i in range(F)], feature_dims=V, embedding_dims=k, pad_id=0, seed=2)
    fw.table.data.copy_(ffm.table.data)
    fw.params.set("bias", ffm.bias.cpu())
    fw.params.value[fw.params._views["interaction_weights/kernel"][0]:][: fw.P] = 1.0      # r = 1
    out_ffm = ffm(Xd)["output"]
    out_fw = fw(Xd)["output"]
    assert torch.isfinite(out_ffm).all()
    assert torch.allclose(out_ffm, out_fw, rtol=1e-6, atol=1e-7)      # FwFM with r = 1, r0 = 0 is FFM
    # bag order does not matter for the membership of pads: moving the pads to the front changes nothing
    Xr = torch.flip(Xd, dims=[2])
    assert torch.allclose(ffm(Xr)["output"], out_ffm, rtol=1e-5, atol=1e-7)
    # batch-split invariance (bit-exact)
    assert torch.equal(torch.cat([ffm(Xd[:4096])["output"], ffm(Xd[4096:])["output"]]), out_ffm)


def test_c2_fused_train_step_equals_layerwise_step():
    """c2 at full size, bf16 tower: ONE train step through the fused path (K7c tail + loss, K7b layer-1 backward)
    and one through the layer-by-layer path start from identical weights and must agree to fp32 rounding:
    loss, every dense variable, and the touched table rows."""
    from etr_b200 import CustomLayers as L
    V, B, F, C_ = int(sum(CRITEO_CARDS)), 65536, 26, 13
    rng = np.random.default_rng(20262)
    cards = np.asarray(CRITEO_CARDS)
    offs = np.concatenate([[0], np.cumsum(cards)[:-1]])
    X = torch.tensor((offs[None, :] + np.floor(cards[None, :] * rng.random((B, F)) ** 3)).astype(np.int64)).cuda()
    Xc = torch.tensor(rng.normal(size=(B, C_)).astype(np.float32)).cuda()
    y = torch.tensor((rng.random(B) < 0.25).astype(np.float32)).cuda()
    names, cont = [f"C{i}" for i in range(F)], [f"I{i}" for i in range(C_)]
    d = {n: X[:, i].contiguous() for i, n in enumerate(names)}
    d.update({n: Xc[:, i].contiguous() for i, n in enumerate(cont)})
    res = []
    for fused in (True, False):
        lay = L.DeepFMRankingLayer(names, V, 16, continuous_features=cont, seed=1, check_ids=False, mlp_precision="bf16",
                                   fused_tail=fused)
        tr = L.Trainer(lay, lr=1e-3)
        loss = float(tr.train_step(d, y).item())
        torch.cuda.synchronize()
        rows = lay.table.data[X[:2048].reshape(-1)].clone()
        res.append((loss, lay.params.value.clone(), rows))
        del lay, tr
        torch.cuda.empty_cache()
    (l1, p1, r1), (l0, p0, r0) = res
    assert abs(l1 - l0) <= 1e-5 * abs(l0)
    assert (p1 - p0).abs().max().item() < 2e-6
    assert (r1 - r0).abs().max().item() < 2e-6
