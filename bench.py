#!/usr/bin/env python
"""Headline benchmark: DeepFM train step (fwd + bwd + Adam) on synthetic
Criteo-shaped data (BASELINE.json configs[1]): 13 dense + 26 sparse fields,
33.76 M-row shared table, k = 16, batch 65 536, fp32.

  python bench.py --gpus N --steps K --warmup W        # this framework (CUDA)
  python bench.py --impl reference ...                 # CPU restatement of the reference

Prints ONE JSON line (see the task contract).  ``value`` = samples/s with the
inputs resident in HBM; ``e2e`` = the same step through the public layer API
from pinned HOST buffers (H2D of the 39 input columns + labels and a D2H read
of the loss inside the timed region); ``roofline`` = the fused gather + FM
kernel against the measured HBM copy peak; ``cpu_baseline`` = the oracle port
timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CRITEO_CARDS = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992,
                5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]      # sum = 33 762 577
F, K_EMB, C_DENSE = 26, 16, 13
BATCH = 65536
SEED = 20261          # 20260 + config index
NCU_K1_DRAM_BYTES = 117.0e6   # dram read 105.6 MB + write 11.4 MB per launch (profiles/r01_prof_gather_fwd.md)


def c5_cards(total: int = 100_000_000):
    """BASELINE configs[4]: 26 fields, V = 1e8, cardinalities proportional to the Criteo list."""
    base = np.asarray(CRITEO_CARDS, dtype=np.float64)
    cards = np.maximum(np.floor(base * total / base.sum()), 1).astype(np.int64)
    cards[int(np.argmax(cards))] += total - int(cards.sum())
    return [int(c) for c in cards]


def make_batches(n_batches: int, B: int, dist: str, seed: int = SEED, cards=None):
    """SURVEY 8d id space: field f owns [offset_f, offset_f + card_f)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cards = np.asarray(CRITEO_CARDS if cards is None else cards, dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(cards)[:-1]])
    out = []
    for _ in range(n_batches):
        u = rng.random((B, F))
        r = np.floor(cards[None, :] * (u ** 3 if dist == "zipf" else u)).astype(np.int64)
        X = offs[None, :] + np.minimum(r, cards[None, :] - 1)
        Xc = rng.standard_normal((B, C_DENSE)).astype(np.float32)
        y = (rng.random(B) < 0.25).astype(np.float32)
        out.append((X, Xc, y))
    return out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = f"/tmp/etr_clocks_{os.getpid()}.csv"
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load": the upper half of the samples (idle samples at the ends drag the median down)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:]
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline_run(budget_s: float, B_cpu: int, n_batches: int, dist: str):
    import torch
    from oracle.cpu_baseline import DeepFMCpuStep, time_cpu_steps
    V = int(sum(CRITEO_CARDS))
    stepper = DeepFMCpuStep(V, F, K_EMB, C_DENSE, mode="rowwise")
    batches = [(torch.from_numpy(X), torch.from_numpy(Xc), torch.from_numpy(y))
               for X, Xc, y in make_batches(n_batches, B_cpu, dist, seed=SEED + 1)]
    sps, steps, secs = time_cpu_steps(stepper, batches, budget_s=budget_s)
    return {"value": sps, "unit": "samples/s", "cores": stepper.threads, "kind": "port",
            "sample": f"{steps} DeepFM train steps of batch {B_cpu} ({steps * B_cpu} samples, {secs:.1f} s) on the "
                      f"same 33.76M-row Criteo-shaped workload; torch-CPU fp32 restatement of the reference "
                      f"(TensorFlow unavailable), row-wise Adam"}


def run_reference(args):
    """--impl reference: the reference's CPU path for this step.  TensorFlow 2.8
    cannot be installed in this image, so this times the op-for-op CPU
    restatement (oracle port) with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B_cpu = 8192
    budget = 12.0
    import torch
    from oracle.cpu_baseline import DeepFMCpuStep
    V = int(sum(CRITEO_CARDS))
    stepper = DeepFMCpuStep(V, F, K_EMB, C_DENSE, mode="rowwise")
    batches = [(torch.from_numpy(X), torch.from_numpy(Xc), torch.from_numpy(y))
               for X, Xc, y in make_batches(4, B_cpu, args.dist, seed=SEED + 1)]
    for i in range(max(args.warmup, 1)):
        stepper.step(*batches[i % 4])
    # a "step" here is a bounded sample (one 8192-sample batch) of the 65 536-sample workload step
    steps = max(1, min(args.steps, 200))
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        stepper.step(*batches[i % 4])
        done += 1
        if time.perf_counter() - t0 > 120:
            break
    dt = time.perf_counter() - t0
    sps = done * B_cpu / dt
    line = {
        "impl": "reference", "metric": "train samples/sec DeepFM (Criteo-shape)", "value": sps, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": done, "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * dt / done,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, B_cpu, graph=False),
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": stepper.threads, "kind": "port",
                         "sample": f"each step = one batch of {B_cpu} samples of the 65536-sample workload step"},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, B, graph):
    c5 = getattr(args, "config", "c2") == "c5"
    return {"workload": ("c5: DeepFM train step (fwd+bwd+Adam), 13 dense + 26 sparse, one shared 100 000 000-row table "
                         "(Criteo cardinalities rescaled) row-sharded over the ranks, k=16, MLP [429->32->8->1]" if c5 else
                         "c2: DeepFM train step (fwd+bwd+Adam), Criteo shape: 13 dense + 26 sparse, one shared "
                         "33 762 577-row table, k=16, MLP [429->32->8->1]"),
            "global_batch": B * max(args.gpus, 1), "per_gpu_batch": B,
            "table_rows": 100_000_000 if c5 else int(sum(CRITEO_CARDS)),
            "embedding_dims": K_EMB, "table_dtype": "f32",
            "mlp": ("layer 1 on tcgen05 (bf16 operands, fp32 accumulate), tail layers fp32" if getattr(args, "mlp", "bf16") == "bf16"
                    else "fp32 SIMT"), "id_distribution": args.dist,
            "apply_mode": "rowwise Adam",
            "parallelism": ((f"dp{args.gpus} batch x row-sharded table (id mod {args.gpus}), "
                             + {"peer": "CUDA-IPC peer memory: unique ids requested from / rows served by the owners as "
                                        "sequential peer stores, gradient rows returned through the same slots, "
                                        "device-side barriers (no NCCL in the step)",
                                "peer-pull": "rows pulled by the gather kernel from NVLink peer memory, gradient rows "
                                             "pushed to the owners' mailboxes, device-side barriers (no NCCL in the step)",
                                "a2a": "NCCL all-to-all"}[getattr(args, "shard", "peer")])
                            if args.gpus > 1 else "dp1"),
            "l2": "L2 flushed (512 MiB write) before every timed step", "cuda_graph": graph}


def run_c3(args):
    """BASELINE configs[2]: DCN-matrix (3 cross layers, k=64 -> D = 13 + 26*64 = 1677) + DenseLayer [64,8] +
    Dense(1), bf16 tensor-core cross (tcgen05), batch 65 536, train step fwd+bwd+row-wise Adam."""
    import torch
    import etr_b200  # noqa: F401
    from etr_b200 import CustomLayers as L
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    B, K3, LAYERS = args.batch, 64, 3
    V = int(sum(CRITEO_CARDS))
    names = [f"C{i + 1}" for i in range(F)]
    cont = [f"I{i + 1}" for i in range(C_DENSE)]
    layer = L.DeepCrossNetworkLayer(names, cont, feature_dims=V, embedding_dims=K3, units=[64, 8], layer_num=LAYERS,
                                    type="matrix", precision="bf16", check_ids=False, seed=1)
    rt = layer.rt
    host = make_batches(3, B, args.dist, seed=SEED + 1)
    dev_batches = [(torch.from_numpy(np.ascontiguousarray(X.T)).to(dev),
                    torch.from_numpy(np.ascontiguousarray(Xc.T)).to(dev), torch.from_numpy(y).to(dev))
                   for X, Xc, y in host]
    use_graph = not args.no_graph
    trainer = L.Trainer(layer, lr=1e-3, graph=use_graph)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def stage(i):
        ids, xc, y = dev_batches[i % 3]
        return trainer.stage({**{n: ids[f] for f, n in enumerate(names)}, **{n: xc[c] for c, n in enumerate(cont)}}, y)

    for i in range(max(args.warmup, 8 if use_graph else 3)):
        trainer.train_step(stage(i))
    torch.cuda.synchronize(dev)
    clocks = ClockSampler(dev.index)
    clocks.start()
    l0 = rt.launches
    ev = []
    for i in range(args.steps):
        b = stage(i)
        flush.zero_()
        torch.cuda.current_stream(dev).wait_event(b._slot.copy_done)
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        trainer.train_step(b)
        e.record()
        ev.append((a, e))
    torch.cuda.synchronize(dev)
    clk = clocks.stop()
    ms = [a.elapsed_time(e) for a, e in ev]
    total = sum(ms)
    # one cross layer forward timed alone (the dominant tensor kernel)
    import ctypes as C
    from etr_b200._lib import check
    from etr_b200.runtime import cast_bf16
    Di = layer.front_pad + layer.D
    x = (torch.randn(B, Di, device=dev) * 0.1).to(torch.bfloat16)
    Wb = cast_bf16(rt, layer.params["cross/W"][0])
    out = torch.empty_like(x)
    u = torch.empty_like(x)
    kt = []
    for i in range(12):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(rt.lib.etr_cross_mat_layer_bf16(rt.ctx, x.data_ptr(), x.data_ptr(), Di, B, Di, Wb.data_ptr(), Di,
                                              layer.params["cross/b"][0].data_ptr(), out.data_ptr(), Di, u.data_ptr(),
                                              Di, rt.stream))
        e.record()
        kt.append((a, e))
    torch.cuda.synchronize(dev)
    k_ms = statistics.mean(a.elapsed_time(e) for a, e in kt[2:])
    D = layer.D
    flops = 2.0 * B * D * D                                    # unpadded D = 1677
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}
    peak = float(peaks["bf16_tflops"])                         # burst figure: the kernel is timed alone
    ach = flops / (k_ms * 1e-3) / 1e12
    step_flops = 3.0 * LAYERS * flops                          # fwd + dgrad + wgrad of the cross layers
    line = {
        "metric": "train samples/sec DCN-matrix (Criteo-shape)", "value": B * args.steps / (total * 1e-3),
        "unit": "samples/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 tensor-core cross / dense GEMMs (fp32 accumulate), fp32 tables + Adam", "data": "synthetic",
        "config": {"workload": "c3: DCN-matrix train step, 13 dense + 26 sparse (k=64, D=1677), 3 cross layers, "
                               "DenseLayer [64,8], Dense(1); 33 762 577-row table", "global_batch": B,
                   "id_distribution": args.dist, "cuda_graph": use_graph, "l2": "L2 flushed before every timed step",
                   "apply_mode": "rowwise Adam"},
        "clocks": clk, "gpu_launches": rt.launches - l0,
        "roofline": {"bound": "tensor", "kernel": "gemm_bf16_tn_kernel<240,2> EPI_CROSS (one cross layer forward)",
                     "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                     "kernel_ms": k_ms, "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)"},
        "cross_gemm_share_of_step_at_peak": step_flops / (float(peaks.get("bf16_tflops_sustained", peak)) * 1e12)
        / (total / args.steps * 1e-3),
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="etr", choices=["etr", "reference"])
    ap.add_argument("--dist", default="zipf", choices=["zipf", "uniform"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--mlp", default="bf16", choices=["bf16", "fp32"],
                    help="first MLP layer: bf16 tcgen05 tensor cores (fp32 accumulate) or the fp32 SIMT exact-parity path")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--shard", default="peer", choices=["peer", "peer-pull", "a2a"],
                    help="N > 1: 'peer' = CUDA-IPC peer memory, de-duplicated request/serve row exchange; 'peer-pull' = "
                         "rows pulled by the gather kernel over NVLink; 'a2a' = NCCL all-to-all exchange")
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c5"],
                    help="c2 = DeepFM (the headline, BASELINE configs[1]); c3 = DCN-matrix bf16 tensor-core cross; "
                         "c5 = DeepFM with a 1e8-row table row-sharded over the ranks, global batch 262 144")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "c3":
        return run_c3(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries exactly ONE JSON line: no NCCL version banner on it (INFO / TRACE requested by the caller
        # are respected and sent to stderr)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "NONE"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import etr_b200  # noqa: F401
    from etr_b200 import CustomLayers as L
    from etr_b200.runtime import IdsBatch, gather_fm_forward

    dev = torch.device("cuda", local_rank)
    cards = c5_cards() if args.config == "c5" else CRITEO_CARDS
    if args.config == "c5" and args.batch == BATCH:
        args.batch = 262144 // world                   # BASELINE configs[4]: global batch 256K
    B = args.batch
    V = int(sum(cards))
    names = [f"C{i + 1}" for i in range(F)]
    cont = [f"I{i + 1}" for i in range(C_DENSE)]
    # N > 1: the shared table is ROW-SHARDED over the ranks (owner = id mod N) and the batch is
    # data-parallel (per-GPU batch fixed: weak scaling); ids and rows cross NVLink in two NCCL
    # all-to-alls per direction, dense gradients are all-reduced.
    layer = L.DeepFMRankingLayer(names, feature_dims=V, embedding_dims=K_EMB, continuous_features=cont,
                                 seed=1, check_ids=False, mlp_precision=args.mlp,
                                 shard=(args.shard if world > 1 else None))
    rt = layer.rt
    n_batches = 6
    host = make_batches(n_batches, B, args.dist, seed=SEED + 17 * rank, cards=cards)
    # host side of the e2e path: the reference's dict of per-feature columns, living in pinned memory as the
    # columns of one column-major block per dtype (what a data loader's pinned staging arena looks like); the
    # Trainer recognises back-to-back columns and moves each block with one async copy
    pinned = []
    for X, Xc, y in host:
        idb = torch.from_numpy(np.ascontiguousarray(X.T)).pin_memory()           # [F, B] int64
        cb = torch.from_numpy(np.ascontiguousarray(Xc.T)).pin_memory()           # [C, B] fp32
        d = {n: idb[i] for i, n in enumerate(names)}
        d.update({n: cb[i] for i, n in enumerate(cont)})
        pinned.append((d, torch.from_numpy(y).pin_memory()))
    dev_batches = []
    for X, Xc, y in host:
        ids = torch.from_numpy(np.ascontiguousarray(X.T)).to(dev)            # field-major [F,B]
        dev_batches.append((ids, torch.from_numpy(np.ascontiguousarray(Xc.T)).to(dev), torch.from_numpy(y).to(dev)))

    # the all-to-all sharded step syncs split sizes on the host (no graph); the peer-memory step does not
    use_graph = (not args.no_graph) and (world == 1 or args.shard != "a2a")
    trainer = L.Trainer(layer, lr=1e-3, apply_mode="rowwise", graph=use_graph)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def stage_from_device(i):
        ids, xc, y = dev_batches[i % n_batches]
        batch = trainer.stage({**{n: ids[f] for f, n in enumerate(names)}, **{n: xc[c] for c, n in enumerate(cont)}}, y)
        return batch

    # ---- warm-up (also captures the graph)
    try:
        for i in range(max(args.warmup, 8 if use_graph else 0)):      # 2 buffer sets x (2 eager + capture) first
            b = stage_from_device(i)
            trainer.train_step(b)
        torch.cuda.synchronize(dev)
    except Exception as e:  # graph capture failed on this box: fall back to eager launches, say so
        if not use_graph:
            raise
        sys.stderr.write(f"[bench] CUDA-graph capture failed ({e!r}); falling back to eager launches\n")
        torch.cuda.synchronize(dev)
        use_graph = False
        trainer = L.Trainer(layer, lr=1e-3, apply_mode="rowwise", graph=False)
        for i in range(args.warmup):
            b = stage_from_device(i)
            trainer._eager_step(b)
        torch.cuda.synchronize(dev)
    step_fn = trainer.train_step if use_graph else trainer._eager_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- timed region 1: inputs resident in HBM (value)
    clocks = ClockSampler(local_rank)
    launches0 = rt.launches
    barrier()
    clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        b = stage_from_device(args.warmup + i)
        flush.zero_()                                   # evict L2 (untimed)
        # inputs are resident in HBM when the timed region starts: the (device-to-device) staging
        # into the graph's static buffers must have landed before the start event
        torch.cuda.current_stream(dev).wait_event(b._slot.copy_done)
        ev[i][0].record()
        step_fn(b)
        ev[i][1].record()
    barrier()
    step_ms = [a.elapsed_time(b_) for a, b_ in ev]
    total_ms = sum(step_ms)
    eager_launches_per_step = None
    if not use_graph:
        eager_launches_per_step = (rt.launches - launches0) / args.steps

    # ---- timed region 2: end to end from pinned host buffers through the public API
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record()
    last = 0.0
    if use_graph:
        # pipelined: H2D of step i+1 (copy stream) overlaps the compute of step i; the loss of
        # every step is read back on the host (one step late, through pinned memory)
        handle = None
        for i in range(args.steps):
            d, y = pinned[(args.warmup + i) % n_batches]
            h = trainer.train_step_async(d, y)
            if handle is not None:
                last = handle.result()                  # D2H read of the previous step's result
            handle = h
        last = handle.result()
    else:
        for i in range(args.steps):
            d, y = pinned[(args.warmup + i) % n_batches]
            loss = trainer._eager_step(d, y)
            last = float(loss.item())                   # D2H read of the step's result
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t_wall0))
    clk = clocks.stop()
    h2d = F * B * 8 + C_DENSE * B * 4 + B * 4
    d2h = 4

    # ---- launches per step (count one eager step; the graph replays exactly these)
    l0 = rt.launches
    trainer._eager_step(stage_from_device(0))
    torch.cuda.synchronize(dev)
    launches_per_step = rt.launches - l0

    # ---- roofline of the fused gather + FM kernel, timed alone, L2 flushed
    ids0 = IdsBatch(rt, dev_batches[0][0], B, F, 1, 1, B, 1)
    col0 = layer.front_pad + C_DENSE
    x = rt.empty((B, col0 + F * K_EMB), torch.bfloat16 if args.mlp == "bf16" else torch.float32)
    xc_dev = dev_batches[0][1].t()
    logit = rt.empty((B,))
    # one L2 flush, then GROUP launches over GROUP different id batches inside one event pair: the rows
    # touched by a group (4 x 163 MB) are far larger than the 126 MB L2, and the ~5 us event/launch
    # latency of timing a single ~50 us launch is amortised
    GROUP = 4
    id_batches = []
    # N > 1: the exchange forms run the gather on rows that are local when the kernel starts (response buffer /
    # received rows): time it on this rank's shard with local row numbers; --shard peer-pull times the pull over NVLink
    k1_table = layer.table
    if world > 1 and args.shard == "peer":
        k1_table = layer.peer.local
    for i in range(n_batches):
        ids_t = dev_batches[i][0]
        if world > 1 and args.shard != "peer-pull":
            ids_t = ids_t // world
        id_batches.append(IdsBatch(rt, ids_t, B, F, 1, 1, B, 1))
    kt = []
    for i in range(max(args.steps // 2, 8)):
        flush.zero_()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for j in range(GROUP):
            gather_fm_forward(k1_table, K_EMB, True, id_batches[(i * GROUP + j) % n_batches], bias=layer.bias,
                              logit=logit, flat=x, flat_col0=col0, cont=xc_dev)
        b_.record()
        kt.append((a, b_))
    torch.cuda.synchronize(dev)
    k_ms = statistics.mean(a.elapsed_time(b_) for a, b_ in kt[2:]) / GROUP
    peak, peak_src = load_peaks()
    # algorithmic bytes per sample: F*(k*4 + 4 [w] + 8 [id]) + 4 [logit]  (+ F*k*4 flat written for the MLP)
    alg_fm = F * (K_EMB * 4 + 4 + 8) + 4
    alg_flat = F * K_EMB * (2 if args.mlp == "bf16" else 4)
    achieved = (alg_fm + alg_flat) * B / (k_ms * 1e-3) / 1e9
    del ids0
    # DRAM bytes of one launch of this kernel from the committed ncu --set full capture of this very
    # command (dram__bytes_read.sum + dram__bytes_write.sum); only valid for the default workload
    traffic, traffic_src = None, None
    if world == 1 and args.dist == "zipf" and args.mlp == "bf16" and B == BATCH and args.config == "c2":
        traffic = NCU_K1_DRAM_BYTES
        traffic_src = ("profiles/r01_prof_gather_fwd.md (ncu --set full, per launch): below the algorithmic bytes because "
                       "the Zipf head and the 16 small fields are L2 hits and the bf16 operand is still in L2 when the "
                       "kernel ends")

    # ---- second roofline: the largest kernel group of the step by time, the fused FM backward + segment
    # reduction + Adam, timed alone (plan precomputed, L2 flushed); algorithmic bytes = 408 B per unique row
    apply_roof = None
    if world == 1:
        from etr_b200.runtime import FusedFMGrad, SparsePlan
        plans = [SparsePlan(rt, id_batches[i], V) for i in range(4)]
        dl_ = torch.randn(B, device=dev) * 1e-7
        sumv_ = torch.randn(B, K_EMB, device=dev) * 0.1
        dx_ = (torch.randn(B, col0 + F * K_EMB, device=dev) * 1e-7).to(torch.bfloat16)
        lr_ = torch.tensor([1e-3], device=dev)
        at = []
        for i in range(10):
            g_ = FusedFMGrad(layer.table, id_batches[i % 4], K_EMB, dl_, sumv_, dx_, col0, plan=plans[i % 4])
            flush.zero_()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g_.apply(lr_, 0.9, 0.999, 1e-7)
            b_.record()
            at.append((a, b_))
        torch.cuda.synchronize(dev)
        a_ms = statistics.median(a.elapsed_time(b_) for a, b_ in at[2:])
        n_u = statistics.mean(p_.n_unique for p_ in plans)
        a_bytes = n_u * 6 * (K_EMB + 1) * 4
        apply_roof = {"bound": "hbm", "kernel": "fm_fused_short/chunk/combine_kernel<4> (FM backward + sorted-run reduction + "
                      "row-wise Adam, 3 launches)", "achieved": a_bytes / (a_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                      "frac": a_bytes / (a_ms * 1e-3) / 1e9 / peak, "kernel_ms": a_ms, "unique_rows_per_launch": n_u,
                      "algorithmic_bytes_per_launch": a_bytes,
                      "note": "408 B per unique row (3 reads + 3 writes of a 68-byte row, SURVEY 8d); request-bound, see "
                              "profiles/r01_mb_apply.md"}

    # ---- max over ranks
    t = torch.tensor([total_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = B * world * args.steps / (total_ms * 1e-3)
    e2e_value = B * world * args.steps / (e2e_ms * 1e-3)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_run(args.cpu_budget, 8192, 4, args.dist)
    line = {
        "metric": "train samples/sec DeepFM (Criteo-shape)", "value": value, "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (tables, FM terms, Adam)" + (" + bf16 tensor-core MLP layer 1" if args.mlp == "bf16" else ""),
        "data": "synthetic",
        "config": workload_config(args, B, use_graph),
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps, "last_loss": last},
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step,
        "roofline": {"bound": "hbm", "kernel": ("gather_fm_fwd_stream_kernel<float,4,13,2,true,true> (gather + FM terms + Flatten, "
                                                f"one launch); rows of the {world - 1} other shards come over NVLink"
                                                if world > 1 and args.shard == "peer-pull" else
                                                "gather_fm_fwd_tile_kernel<float,4> (gather + FM terms + Flatten, one launch)"),
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel_ms": k_ms,
                     "algorithmic_bytes_per_launch": (alg_fm + alg_flat) * B,
                     "note": "algorithmic bytes = B*(F*(4k+4+8)+4) FM terms + B*F*k*osize flattened operand written "
                             "for the MLP (osize 2 for the bf16 tensor-core MLP, 4 for fp32); timed alone with CUDA "
                             "events: L2 flushed, then 4 launches over 4 different id batches (rows touched >> L2) "
                             "per event pair"},
        "step_ms_min_median_max": [min(step_ms), statistics.median(step_ms), max(step_ms)],
    }
    if apply_roof:
        line["roofline_apply"] = apply_roof
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
