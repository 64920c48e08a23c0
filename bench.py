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
import math
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
    """SM clock + throttle reasons sampled IN-PROCESS through NVML every 2 ms from before the warm-up to the end of
    the timed regions (round 1 started nvidia-smi after the warm-up and a 8 ms timed region ended before its first
    sample).  ``busy(True/False)`` brackets the timed regions: the reported median is over samples taken inside
    them.  Falls back to an ``nvidia-smi -lms`` child process when pynvml is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, gpu_index: int, uuid=None):
        self.idx, self.uuid = gpu_index, uuid
        self.samples = []            # (busy, sm_mhz, reasons bitmask)
        self._busy = False
        self._stop = False
        self.thread = self.proc = None
        self.max_mhz = None
        self.how = None

    def busy(self, flag: bool):
        self._busy = flag

    def _loop(self, nv, h):
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop:
            try:
                self.samples.append((self._busy, float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                     int(get_reasons(h))))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            h = None
            if self.uuid:
                try:
                    u = str(self.uuid)
                    h = nv.nvmlDeviceGetHandleByUUID(u if u.startswith("GPU-") else "GPU-" + u)
                except Exception:
                    h = None
            if h is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                idx = self.idx
                if vis and all(x.strip().isdigit() for x in vis.split(",")):
                    idx = int(vis.split(",")[self.idx])
                h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._loop, args=(nv, h), daemon=True)
            self.thread.start()
            self.how = "pynvml, 2 ms period, in-process"
            return
        except Exception:
            self.thread = None
        try:
            self.path = f"/tmp/etr_clocks_{os.getpid()}.csv"
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
            self.how = "nvidia-smi -lms 20"
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            busy = [s for s in self.samples if s[0]] or self.samples
            if not busy:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"], "how": self.how}
            mask = 0
            for _, _, r in busy:
                mask |= r
            return {"sm_mhz": statistics.median(s[1] for s in busy), "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(n for b, n in self.BITS.items() if mask & b), "samples": len(busy),
                    "samples_total": len(self.samples), "sm_mhz_min": min(s[1] for s in busy), "how": self.how}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi / NVML unavailable"]}
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
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "how": self.how}
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:]
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "how": self.how}


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
    B_cpu = args.batch          # the workload's own step (same_config): one 65 536-sample batch per step
    budget = 12.0
    import torch
    from oracle.cpu_baseline import DeepFMCpuStep
    V = int(sum(CRITEO_CARDS))
    stepper = DeepFMCpuStep(V, F, K_EMB, C_DENSE, mode="rowwise")
    batches = [(torch.from_numpy(X), torch.from_numpy(Xc), torch.from_numpy(y))
               for X, Xc, y in make_batches(4, B_cpu, args.dist, seed=SEED + 1)]
    for i in range(max(args.warmup, 1)):
        stepper.step(*batches[i % 4])
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
                         "sample": f"each step = one full {B_cpu}-sample train step of the workload (fwd+bwd+row-wise Adam), "
                                   f"{done} steps, torch-CPU fp32 restatement of the reference (TensorFlow unavailable)"},
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
            "l2": "L2 flushed (512 MiB write, then 256 MiB read so that the cache holds clean lines) before every timed step",
            "cuda_graph": graph}


def run_c3(args):
    """BASELINE configs[2]: DCN-matrix (3 cross layers, k=64 -> D = 13 + 26*64 = 1677) + DenseLayer [64,8] +
    Dense(1), bf16 tensor-core cross (tcgen05), batch 65 536, train step fwd+bwd+row-wise Adam."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "NONE"
        dist.init_process_group("nccl", device_id=dev)
    import etr_b200  # noqa: F401
    from etr_b200 import CustomLayers as L
    B, K3, LAYERS = args.batch, 64, 3
    V = int(sum(CRITEO_CARDS))
    names = [f"C{i + 1}" for i in range(F)]
    cont = [f"I{i + 1}" for i in range(C_DENSE)]
    # N > 1 (BASELINE configs[2] is 1/2/4/8 x B200): batch data-parallel (weak scaling, 65 536 samples per GPU), the table
    # row-sharded with the NCCL all-to-all exchange (ids out, rows back; gradient rows to the owners), cross / tower
    # weights replicated and all-reduced -- the cross GEMMs are per-sample work and do not communicate
    layer = L.DeepCrossNetworkLayer(names, cont, feature_dims=V, embedding_dims=K3, units=[64, 8], layer_num=LAYERS,
                                    type="matrix", precision="bf16", check_ids=False, seed=1,
                                    shard=("a2a" if world > 1 else None))
    rt = layer.rt
    host = make_batches(3, B, args.dist, seed=SEED + 1 + 17 * rank)
    dev_batches = [(torch.from_numpy(np.ascontiguousarray(X.T)).to(dev),
                    torch.from_numpy(np.ascontiguousarray(Xc.T)).to(dev), torch.from_numpy(y).to(dev))
                   for X, Xc, y in host]
    use_graph = (not args.no_graph) and world == 1          # the all-to-all exchange reads split sizes on the host
    trainer = L.Trainer(layer, lr=1e-3, graph=use_graph)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    clocks = ClockSampler(dev.index, getattr(torch.cuda.get_device_properties(dev), "uuid", None))
    clocks.start()

    def stage(i):
        ids, xc, y = dev_batches[i % 3]
        return trainer.stage({**{n: ids[f] for f, n in enumerate(names)}, **{n: xc[c] for c, n in enumerate(cont)}}, y)

    for i in range(max(args.warmup, 8 if use_graph else 3)):
        trainer.train_step(stage(i))
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    clocks.busy(True)
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
    clocks.busy(False)
    ms = [a.elapsed_time(e) for a, e in ev]
    total = sum(ms)
    if world > 1:
        t_ = torch.tensor([total], device=dev, dtype=torch.float64)
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        total = float(t_[0])
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
    clk = clocks.stop()
    if rank != 0:
        dist.destroy_process_group()
        return
    D = layer.D
    flops = 2.0 * B * D * D                                    # unpadded D = 1677
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}
    peak = float(peaks["bf16_tflops"])                         # burst figure: the kernel is timed alone
    ach = flops / (k_ms * 1e-3) / 1e12
    step_flops = 3.0 * LAYERS * flops                          # fwd + dgrad + wgrad of the cross layers
    line = {
        "metric": "train samples/sec DCN-matrix (Criteo-shape)", "value": B * world * args.steps / (total * 1e-3),
        "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 tensor-core cross / dense GEMMs (fp32 accumulate), fp32 tables + Adam", "data": "synthetic",
        "config": {"workload": "c3: DCN-matrix train step, 13 dense + 26 sparse (k=64, D=1677), 3 cross layers, "
                               "DenseLayer [64,8], Dense(1); 33 762 577-row table", "global_batch": B * world, "per_gpu_batch": B,
                   "parallelism": "dp1" if world == 1 else f"dp{world} batch x row-sharded table (id mod {world}), NCCL all-to-all exchange; "
                                  "cross / tower weights replicated, gradients all-reduced",
                   "id_distribution": args.dist, "cuda_graph": use_graph, "l2": "L2 flushed before every timed step",
                   "apply_mode": "rowwise Adam"},
        "clocks": clk, "gpu_launches": rt.launches - l0,
        "roofline": {"bound": "tensor", "kernel": {"0": "gemm_bf16_tn_kernel<240,2> (one tile per CTA)",
                                                   "1": "gemm_bf16_persist_kernel<240,4,1> (persistent, TMEM double-buffered)"}.get(
                         os.environ.get("ETR_GEMM_PERSIST", "2"), "gemm_bf16_persist_kernel<240,5,2> (persistent CTA pairs, "
                         "tcgen05 cta_group::2, TMEM double-buffered)") + " EPI_CROSS: one cross layer forward, x_{l+1} = x0 (.) (W x_l + b) + x_l",
                     "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                     "traffic": (ncu_traffic("gemm_bf16_persist") or [None])[0],
                     "kernel_ms": k_ms, "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)"},
        "cross_gemm_share_of_step_at_peak": step_flops / (float(peaks.get("bf16_tflops_sustained", peak)) * 1e12)
        / (total / args.steps * 1e-3),
    }
    print(json.dumps(line))


def _line_common(metric, value, args, B, total_ms, dtype, workload, extra_cfg):
    return {"metric": metric, "value": value, "unit": "samples/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic",
            "config": {"workload": workload, "global_batch": B, "id_distribution": args.dist,
                       "l2": "L2 flushed (512 MiB write) before every timed step", **extra_cfg}}


def run_c1(args):
    """BASELINE configs[0]: the reference's own CPU-runnable case -- FMRankingLayer fwd + bwd + Adam, batch 4096, 26
    sparse fields, k = 16, one shared table of V = 160 000 rows (3.DCN/ModelManager.py:68), equal split.  The working
    set (13.6 MB of rows) is L2-resident and the step is launch-latency-bound: a parity / CPU-comparison config,
    NOT a roofline claim."""
    import torch
    import etr_b200  # noqa: F401
    from etr_b200 import CustomLayers as L
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    clocks = ClockSampler(dev.index, getattr(torch.cuda.get_device_properties(dev), "uuid", None))
    clocks.start()
    B, V = 4096, 160000
    cards = [V // F] * F
    cards[-1] += V - sum(cards)
    names = [f"C{i + 1}" for i in range(F)]
    layer = L.FMRankingLayer(names, feature_dims=V, embedding_dims=K_EMB, seed=1, check_ids=False)
    rt = layer.rt
    rng = np.random.Generator(np.random.PCG64(20260))
    offs = np.concatenate([[0], np.cumsum(cards)[:-1]])
    host = []
    for _ in range(6):
        u = rng.random((B, F))
        X = offs[None, :] + np.minimum(np.floor(np.asarray(cards)[None, :] * (u ** 3 if args.dist == "zipf" else u)).astype(np.int64),
                                       np.asarray(cards)[None, :] - 1)
        host.append((X, (rng.random(B) < 0.25).astype(np.float32)))
    pinned = []
    for X, y in host:
        idb = torch.from_numpy(np.ascontiguousarray(X.T)).pin_memory()
        pinned.append(({n: idb[i] for i, n in enumerate(names)}, torch.from_numpy(y).pin_memory()))
    devb = [({n: torch.from_numpy(np.ascontiguousarray(X[:, i])).to(dev) for i, n in enumerate(names)},
             torch.from_numpy(y).to(dev)) for X, y in host]
    trainer = L.Trainer(layer, lr=1e-3, graph=not args.no_graph)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for i in range(max(args.warmup, 8)):
        trainer.train_step(trainer.stage(*devb[i % 6]))
    torch.cuda.synchronize(dev)
    reps = []
    clocks.busy(True)
    for r in range(20):
        ev = []
        for i in range(args.steps):
            b = trainer.stage(*devb[(r * args.steps + i) % 6])
            flush.zero_()
            torch.cuda.current_stream(dev).wait_event(b._slot.copy_done)
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); trainer.train_step(b); e.record()
            ev.append((a, e))
        torch.cuda.synchronize(dev)
        reps.append(sum(a.elapsed_time(e) for a, e in ev))
    total = statistics.median(reps)
    e2e = []
    for r in range(20):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        handle = None
        for i in range(args.steps):
            d_, y_ = pinned[(r * args.steps + i) % 6]
            h = trainer.train_step_async(d_, y_) if trainer.use_graph else None
            if h is None:
                last = float(trainer._eager_step(d_, y_).item())
            else:
                if handle is not None:
                    last = handle.result()
                handle = h
        if handle is not None:
            last = handle.result()
        torch.cuda.synchronize(dev)
        e2e.append(1e3 * (time.perf_counter() - t0))
    clocks.busy(False)
    e2e_ms = statistics.median(e2e)
    l0 = rt.launches
    trainer._eager_step(trainer.stage(*devb[0]))
    torch.cuda.synchronize(dev)
    lps = rt.launches - l0
    clk = clocks.stop()
    line = _line_common("train samples/sec FM (reference CPU-runnable case)", B * args.steps / (total * 1e-3), args, B, total,
                        "f32", "c1: FMRankingLayer train step (fwd+bwd+row-wise Adam), 26 sparse fields, k=16, one shared "
                        "160 000-row table (L2-resident: launch-latency-bound, not a roofline config)",
                        {"table_rows": V, "embedding_dims": K_EMB, "apply_mode": "rowwise Adam", "cuda_graph": trainer.use_graph})
    line.update({"clocks": clk, "gpu_launches": lps * args.steps, "gpu_launches_per_step": lps,
                 "e2e": {"value": B * args.steps / (e2e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": F * B * 8 + B * 4,
                         "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps, "last_loss": last},
                 "roofline": None})
    if not args.no_cpu_baseline:
        import torch as _t
        from oracle.cpu_baseline import DeepFMCpuStep, time_cpu_steps
        stepper = DeepFMCpuStep(V, F, K_EMB, 0, mode="rowwise", with_mlp=False)
        batches = [(_t.from_numpy(X), None, _t.from_numpy(y)) for X, y in host[:4]]
        sps, steps, secs = time_cpu_steps(stepper, batches, budget_s=args.cpu_budget)
        line["cpu_baseline"] = {"value": sps, "unit": "samples/s", "cores": stepper.threads, "kind": "port",
                                "sample": f"{steps} FM train steps of batch {B} ({secs:.1f} s), same config; torch-CPU fp32 "
                                          f"restatement of the reference (TensorFlow unavailable), row-wise Adam"}
    print(json.dumps(line))


def run_c4(args):
    """BASELINE configs[3]: FFM / FwFM field-pair interaction, 39 fields x 100 000 ids, k = 8, multi-hot bags of
    1..50 ids (padded to 50, pad id 0, sum pooling), batch 8 192.  Train step = fused pooling + pair
    kernel forward, BCE, pair backward, sorted-id segment reduction, row-wise Adam.  roofline = the forward kernel
    alone at 1 256 algorithmic bytes per looked-up id."""
    import torch
    import etr_b200  # noqa: F401
    from etr_b200 import CustomLayers as L
    from etr_b200.runtime import IdsBatch
    import ctypes as C
    from etr_b200._lib import check
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    clocks = ClockSampler(dev.index, getattr(torch.cuda.get_device_properties(dev), "uuid", None))
    clocks.start()
    F4, K4, LMAX, CARD = 39, 8, 50, 100000
    B = args.batch if args.batch != BATCH else 8192
    V = F4 * CARD
    names = [f"S{i + 1}" for i in range(F4)]
    cls = L.FwFMLayer if os.environ.get("ETR_C4_MODEL", "ffm") == "fwfm" else L.FFMLayer
    layer = cls(names, feature_dims=V, embedding_dims=K4, pad_id=0, pooling="sum", seed=1, check_ids=False)
    rt = layer.rt
    rng = np.random.Generator(np.random.PCG64(20263))
    host = []
    for _ in range(3):
        lens = rng.integers(1, LMAX + 1, size=(B, F4))
        u = rng.random((B, F4, LMAX))
        r = np.floor(CARD * (u ** 3 if args.dist == "zipf" else u)).astype(np.int64)
        X = np.maximum(np.arange(F4)[None, :, None] * CARD + np.minimum(r, CARD - 1), 1)     # id 0 is the pad id
        X[np.arange(LMAX)[None, None, :] >= lens[:, :, None]] = 0
        host.append((X, (rng.random(B) < 0.25).astype(np.float32), int(lens.sum())))
    devb = [(torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev)) for X, y, _ in host]
    pinned = [(torch.from_numpy(X).pin_memory(), torch.from_numpy(y).pin_memory()) for X, y, _ in host]
    trainer = L.Trainer(layer, lr=1e-3, graph=False)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for i in range(args.warmup):
        trainer._eager_step(*devb[i % 3])
    torch.cuda.synchronize(dev)
    clocks.busy(True)
    ev = []
    l0 = rt.launches
    for i in range(args.steps):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); trainer._eager_step(*devb[i % 3]); e.record()
        ev.append((a, e))
    torch.cuda.synchronize(dev)
    lps = (rt.launches - l0) / args.steps
    ms = [a.elapsed_time(e) for a, e in ev]
    total = sum(ms)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for i in range(args.steps):
        Xp, yp = pinned[i % 3]
        loss = trainer._eager_step(Xp.to(dev, non_blocking=True), yp.to(dev, non_blocking=True))
        last = float(loss.item())
    torch.cuda.synchronize(dev)
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    # forward kernel alone
    ids = [IdsBatch.from_matrix(rt, X, 0, "sum") for X, _ in devb]
    prob = rt.empty((B, 1))
    kt = []
    for i in range(10):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_, d_ = layer.table.desc(), ids[i % 3].desc()
        a.record()
        check(rt.lib.etr_field_pair_forward(rt.ctx, C.byref(t_), K4, 1, C.byref(d_), layer.bias.data_ptr(), None, None,
                                            None, None, None, prob.data_ptr(), None, rt.stream))
        e.record()
        kt.append((a, e, host[i % 3][2]))
    torch.cuda.synchronize(dev)
    clocks.busy(False)
    k_ms = statistics.median(a.elapsed_time(e) for a, e, _ in kt[2:])
    nnz = statistics.mean(h[2] for h in host)
    peak, peak_src = load_peaks()
    alg = nnz * (F4 * K4 * 4 + 8) + B * 4
    clk = clocks.stop()
    line = _line_common("train samples/sec FFM (field-pair, multi-hot bags)", B * args.steps / (total * 1e-3), args, B, total,
                        "f32", f"c4: {cls.__name__} train step (fwd+bwd+row-wise Adam), 39 fields x 100 000 ids, k=8, padded bags of "
                        f"1..50 ids (mean 25.5), sum pooling", {"table_rows": V, "embedding_dims": K4, "apply_mode": "rowwise Adam",
                                                                 "cuda_graph": False, "lookups_per_step": nnz})
    tr_ = ncu_traffic("field_pair_fwd")
    line.update({"clocks": clk, "gpu_launches": int(lps * args.steps), "gpu_launches_per_step": lps,
                 "e2e": {"value": B * args.steps / (e2e_ms * 1e-3), "unit": "samples/s",
                         "h2d_bytes_per_step": B * F4 * LMAX * 8 + B * 4, "d2h_bytes_per_step": 4,
                         "ms_per_step": e2e_ms / args.steps, "last_loss": last},
                 "step_ms_min_median_max": [min(ms), statistics.median(ms), max(ms)],
                 "roofline": {"bound": "hbm", "kernel": "field_pair_fwd_kernel (bag pooling + 741 pair dots + linear term + sigmoid, one launch)",
                              "achieved": alg / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": alg / (k_ms * 1e-3) / 1e9 / peak, "traffic": tr_[0] if tr_ else None,
                              "traffic_source": tr_[1] if tr_ else None, "peak_source": peak_src, "kernel_ms": k_ms,
                              "algorithmic_bytes_per_launch": alg,
                              "note": "F*k*4 + 8 = 1 256 B per looked-up id (SURVEY 8d) x the valid ids of one batch + 4 B/sample out"}})
    print(json.dumps(line))


def ncu_traffic(kernel_substr: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of a kernel from the newest committed
    ``ncu --set full`` summary under profiles/ (rNN_prof_*.md, written by scripts/summarize_ncu.py); None when no
    summary names the kernel.  Read at run time -- nothing is hard-coded."""
    import glob
    import re
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_prof_*.md"))):
        txt = open(path).read()
        for sec in txt.split("\n## ")[1:]:
            head, _, body = sec.partition("\n")
            if kernel_substr not in head:
                continue
            def grab(key):
                m = re.search(r"\| " + re.escape(key) + r" \| ([0-9.eE+-]+) \| (\w+) \|", body)
                if not m:
                    return None
                return float(m.group(1)) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(m.group(2), 1.0)
            r_, w_ = grab("dram__bytes_read.sum"), grab("dram__bytes_write.sum")
            if r_ is not None and w_ is not None:
                best = (r_ + w_, os.path.relpath(path, ROOT))
    return best


def sharded_check(L, world, rank, shard_mode, dist):
    """N > 1 only (the only correctness evidence the driver's 2/4/8-GPU runs carry): one small DeepFM, row-sharded
    exactly like the benchmarked model, against an UNSHARDED replica on the same GPU fed the all-gathered global
    batch -- bit-exact forward, and table / dense weights after train steps (untimed)."""
    import torch
    Fs, k, V, C, B = 26, 16, 100003, 13, 4096
    names, cont = [f"f{i}" for i in range(Fs)], [f"c{i}" for i in range(C)]
    out = {}
    try:
        sharded = L.DeepFMRankingLayer(names, V, k, continuous_features=cont, seed=3, shard=shard_mode, check_ids=False,
                                       mlp_precision="bf16")
        full = L.DeepFMRankingLayer(names, V, k, continuous_features=cont, seed=3, check_ids=False, mlp_precision="bf16")
        g = torch.Generator(device="cuda").manual_seed(1234)
        table = torch.empty(V, k + 1, device="cuda").uniform_(-0.05, 0.05, generator=g)
        full.table.data[:, : k + 1] = table
        (sharded.peer if shard_mode.startswith("peer") else sharded.shard).load_global(table)
        sharded.params.value.copy_(full.params.value)
        torch.cuda.synchronize()
        dist.barrier()
        rng = np.random.default_rng(10 + rank)

        def batch():
            X = (rng.random((B, Fs)) ** 3 * V).astype(np.int64)
            Xc = rng.normal(size=(B, C)).astype(np.float32)
            y = (rng.random(B) < 0.3).astype(np.float32)
            d = {n: torch.tensor(X[:, i]).cuda() for i, n in enumerate(names)}
            d.update({n: torch.tensor(Xc[:, i]).cuda() for i, n in enumerate(cont)})
            return d, torch.tensor(y).cuda()

        d, y = batch()
        out["forward_bit_exact"] = bool(torch.equal(sharded(d)["output"], full(d)["output"]))
        tr_s, tr_f = L.Trainer(sharded, lr=1e-2), L.Trainer(full, lr=1e-2)
        loss_err = 0.0
        for step in range(2):
            d, y = batch()
            ls = tr_s.train_step(d, y).clone()
            gd = {}
            for n, t in d.items():
                parts = [torch.empty_like(t) for _ in range(world)]
                dist.all_gather(parts, t)
                gd[n] = torch.cat(parts)
            ys = [torch.empty_like(y) for _ in range(world)]
            dist.all_gather(ys, y)
            lf = tr_f.train_step(gd, torch.cat(ys))
            dist.all_reduce(ls)
            loss_err = max(loss_err, abs(float(ls.item()) / world - float(lf.item())))
            if step == 0:
                torch.cuda.synchronize()
                mine0 = full.table.data[rank::world, : k + 1]
                out["first_step_max_table_err"] = float((sharded.table.data[: mine0.shape[0], : k + 1] - mine0).abs().max().item())
                out["first_step_max_dense_err"] = float((sharded.params.value - full.params.value).abs().max().item())
        torch.cuda.synchronize()
        mine = full.table.data[rank::world, : k + 1]
        e = torch.tensor([float((sharded.table.data[: mine.shape[0], : k + 1] - mine).abs().max().item()),
                          float((sharded.params.value - full.params.value).abs().max().item()), loss_err,
                          out["first_step_max_table_err"], out["first_step_max_dense_err"],
                          0.0 if out["forward_bit_exact"] else 1.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        ref = sharded.params.value.clone()
        dist.broadcast(ref, 0)
        same = torch.tensor([1.0 if torch.equal(ref, sharded.params.value) else 0.0], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        out = {"max_table_err": float(e[0]), "max_dense_err": float(e[1]), "max_loss_err": float(e[2]),
               "first_step_max_table_err": float(e[3]), "first_step_max_dense_err": float(e[4]),
               "forward_bit_exact": bool(e[5] == 0.0), "dense_replicas_bit_identical": bool(same.item() == 1.0),
               "what": f"V={V}, per-rank batch {B}, 2 Adam steps (lr 1e-2, bf16 tower) of the {shard_mode}-sharded layer vs an "
                       f"unsharded replica fed the all-gathered batch; errors are max over ranks; first-step errors "
                       f"are fp32 rounding only, later ones include bf16 re-rounding of drifted weights",
               "ok": bool(e[5] == 0.0 and e[3] < 2e-6 and e[4] < 2e-6 and e[0] < 5e-3 and e[1] < 5e-3)}
        del sharded, full
        torch.cuda.empty_cache()
    except Exception as ex:            # the check must never take the bench line down with it
        out = {"ok": False, "error": repr(ex)[:300]}
    return out


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
    ap.add_argument("--no-plan-ahead", action="store_true",
                    help="N > 1: sort every batch inside its own step (round-1 behaviour) instead of one step early")
    ap.add_argument("--no-extras", action="store_true", help="skip value_fp32 / value_keras_dense / sharded_check")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--min-time", type=float, default=1.0,
                    help="repeat the K-step timed regions until this many seconds of timed work (median repetition is reported)")
    ap.add_argument("--shard", default="peer", choices=["peer", "peer-pull", "a2a"],
                    help="N > 1: 'peer' = CUDA-IPC peer memory, de-duplicated request/serve row exchange; 'peer-pull' = "
                         "rows pulled by the gather kernel over NVLink; 'a2a' = NCCL all-to-all exchange")
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5"],
                    help="c2 = DeepFM (the headline, BASELINE configs[1]); c1 = FM B=4096 V=160000 (the reference's CPU-runnable "
                         "case); c3 = DCN-matrix bf16 tensor-core cross; c4 = FFM/FwFM 39 fields, k=8, bags <= 50; "
                         "c5 = DeepFM with a 1e8-row table row-sharded over the ranks, global batch 262 144")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "c3":
        return run_c3(args)
    if args.config == "c4":
        return run_c4(args)
    if args.config == "c1":
        return run_c1(args)

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
    clocks = ClockSampler(local_rank, getattr(torch.cuda.get_device_properties(dev), "uuid", None))
    clocks.start()                                     # BEFORE the warm-up: never "no samples" again
    cards = c5_cards() if args.config == "c5" else CRITEO_CARDS
    if args.config == "c5" and args.batch == BATCH:
        args.batch = 262144 // world                   # BASELINE configs[4]: global batch 256K
    B = args.batch
    V = int(sum(cards))
    names = [f"C{i + 1}" for i in range(F)]
    cont = [f"I{i + 1}" for i in range(C_DENSE)]
    # N > 1: the shared table is ROW-SHARDED over the ranks (owner = id mod N) and the batch is
    # data-parallel (per-GPU batch fixed: weak scaling), dense gradients are all-reduced.
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
    flush_wr = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    flush_rd = torch.ones(64 << 20, dtype=torch.float32, device=dev)

    class flush:
        """L2 flush between timed iterations: a 512 MiB WRITE evicts everything, then a 256 MiB READ replaces the dirty
        flush lines by clean ones -- otherwise the timed kernel pays the write-back of up to 126 MB of flush data
        (measured: +8 us on the 110 us apply kernel, profiles/r02_mb_apply.md)."""
        @staticmethod
        def zero_():
            flush_wr.zero_()
            flush_rd.sum()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def device_dict(i):
        ids, xc, y = dev_batches[i % n_batches]
        return {**{n: ids[f] for f, n in enumerate(names)}, **{n: xc[c] for c, n in enumerate(cont)}}, y

    def make_stepper(lay, apply_mode, graph):
        """(trainer, step_fn, graph_used) after the warm-up (which also captures the graph)"""
        # plan-ahead (the sort of batch i+1 runs on the side stream during step i; every step still does exactly one sort):
        # always at N > 1 (peer), and at N = 1 for the row-wise fused step unless --no-plan-ahead
        tr = L.Trainer(lay, lr=1e-3, apply_mode=apply_mode, graph=graph,
                       plan_ahead=(graph and apply_mode == "rowwise" and not args.no_plan_ahead))
        try:
            # every buffer set (2, or 3 with plan-ahead) needs 2 eager steps + its capture before anything is timed: with 8
            # warm-ups and 3 sets the last capture fell into the timed steps of value_fp32 (4.45 instead of 1.06 ms per step)
            for i in range(max(args.warmup, 3 * tr.depth + 2 if graph else 0)):
                d_, y_ = device_dict(i)
                tr.train_step(tr.stage(d_, y_))
            torch.cuda.synchronize(dev)
        except Exception as e:  # graph capture failed on this box: fall back to eager launches, say so
            if not graph:
                raise
            sys.stderr.write(f"[bench] CUDA-graph capture failed ({e!r}); falling back to eager launches\n")
            torch.cuda.synchronize(dev)
            graph = False
            tr = L.Trainer(lay, lr=1e-3, apply_mode=apply_mode, graph=False)
            for i in range(args.warmup):
                d_, y_ = device_dict(i)
                tr._eager_step(tr.stage(d_, y_))
            torch.cuda.synchronize(dev)
        return tr, (tr.train_step if graph else tr._eager_step), graph

    # N > 1, peer-sharded, CUDA graph: the inputs of step i+1 are staged BEFORE step i is launched, and step i sorts
    # them (side stream, inside its own timed bracket) -- every step still does one sort, one step early
    ahead = {"on": (world == 1 or args.shard == "peer") and not args.no_plan_ahead, "batch": None, "i": None}

    def timed_rep(tr, step_fn, K, base):
        """EXACTLY K steps, each bracketed by CUDA events on the launch stream, L2 flushed (untimed) before every
        step, a barrier + synchronize on both sides; returns the per-step ms of this rank."""
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        clocks.busy(True)
        look = ahead["on"] and tr is trainer_ref[0] and step_fn == tr.train_step
        for i in range(K):
            if look:
                if ahead["batch"] is None or ahead["i"] != base + i:
                    d_, y_ = device_dict(base + i)
                    ahead["batch"] = tr.stage(d_, y_)
                b = ahead["batch"]
                d_, y_ = device_dict(base + i + 1)
                nxt = tr.stage(d_, y_)
                ahead["batch"], ahead["i"] = nxt, base + i + 1
            else:
                d_, y_ = device_dict(base + i)
                b, nxt = tr.stage(d_, y_), None
            flush.zero_()                                   # evict L2 (untimed)
            # inputs are resident in HBM when the timed region starts: the (device-to-device) staging
            # into the graph's static buffers must have landed before the start event
            torch.cuda.current_stream(dev).wait_event(b._slot.copy_done)
            if nxt is not None:
                torch.cuda.current_stream(dev).wait_event(nxt._slot.copy_done)
            ev[i][0].record()
            if nxt is not None:
                step_fn(b, None, nxt)
            else:
                step_fn(b)
            ev[i][1].record()
        barrier()
        clocks.busy(False)
        return [a.elapsed_time(b_) for a, b_ in ev]

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def measure_value(tr, step_fn, min_time):
        """repetitions of the K-step region until ~min_time seconds of timed work; (median rep total ms, all rep
        totals, per-step ms of the median rep).  Every total is the max over ranks."""
        first = timed_rep(tr, step_fn, args.steps, args.warmup)
        tot0 = max_over_ranks(sum(first))
        reps = int(min(60, max(3, math.ceil(min_time * 1e3 / max(tot0, 1e-3)))))
        if world > 1:
            t = torch.tensor([reps], device=dev)
            dist.broadcast(t, 0)
            reps = int(t[0])
        totals, steps_ms = [tot0], [first]
        for r in range(1, reps):
            ms = timed_rep(tr, step_fn, args.steps, args.warmup + r * args.steps)
            totals.append(max_over_ranks(sum(ms)))
            steps_ms.append(ms)
        order = sorted(range(len(totals)), key=lambda i: totals[i])
        mid = order[len(order) // 2]
        return totals[mid], totals, steps_ms[mid]

    trainer_ref = [None]
    trainer, step_fn, use_graph = make_stepper(layer, "rowwise", use_graph)
    trainer_ref[0] = trainer
    ahead["on"] = ahead["on"] and use_graph
    if ahead["on"]:
        timed_rep(trainer, step_fn, 14, 0)           # untimed: captures the look-ahead variant of every buffer set
        torch.cuda.synchronize(dev)
    launches0 = rt.launches

    # ---- timed region 1: inputs resident in HBM (value)
    total_ms, rep_totals, step_ms = measure_value(trainer, step_fn, args.min_time)

    # ---- timed region 2: end to end from pinned host buffers through the public API
    def e2e_rep(K, base):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clocks.busy(True)
        t_wall0 = time.perf_counter()
        e0.record()
        last_ = 0.0
        if use_graph:
            # pipelined: H2D of step i+1 (copy stream) overlaps the compute of step i; the loss of
            # every step is read back on the host (one step late, through pinned memory)
            handle = None
            if ahead["on"]:
                ahead["batch"] = None                        # (the value loop's look-ahead batch is dropped)
                # three buffer sets: step i computes on set i, sorts the ids of set i+1, and the copy stream fills set i+2
                # (it waits ON THE DEVICE for the step that last used that set).  The H2D of one batch takes 0.325 ms alone
                # (53 GB/s, scripts/mb_h2d.py) -- longer than the step -- so the copies must run back to back: staged one
                # step ahead, every copy started a host round trip late.  The pipeline state survives across repetitions,
                # so a timed repetition holds exactly K copies and K steps (the two priming copies are in the untimed one).
                if ahead.get("e2e") is None:
                    ahead["e2e"] = [trainer.stage(*pinned[base % n_batches]), trainer.stage(*pinned[(base + 1) % n_batches]),
                                    base + 2]
                cur_b, nxt_b, pos = ahead["e2e"]
            for i in range(K):
                if ahead["on"]:
                    d_, y_ = pinned[pos % n_batches]
                    nn_b = trainer.stage(d_, y_)             # H2D of step i+2 (copy stream)
                    pos += 1
                    h = trainer.train_step_async(cur_b, None, nxt_b)
                    cur_b, nxt_b = nxt_b, nn_b
                else:
                    d_, y_ = pinned[(base + i) % n_batches]
                    h = trainer.train_step_async(d_, y_)
                if handle is not None:
                    last_ = handle.result()                  # D2H read of the previous step's result
                handle = h
            last_ = handle.result()
            if ahead["on"]:
                ahead["e2e"] = [cur_b, nxt_b, pos]
        else:
            for i in range(K):
                d_, y_ = pinned[(base + i) % n_batches]
                loss = trainer._eager_step(d_, y_)
                last_ = float(loss.item())                   # D2H read of the step's result
        e1.record()
        barrier()
        clocks.busy(False)
        return max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t_wall0)), last_

    # one untimed repetition first: at N > 1 the first pass through the staged path still captures the look-ahead graph of
    # every buffer set and sizes the request mailboxes (23 ms/step at N = 4 in one run: with reps = ceil(min_time / first)
    # that left 3 repetitions and a polluted median)
    e2e_rep(max(3, min(args.steps, 6)), args.warmup)
    e2e_first, last = e2e_rep(args.steps, args.warmup + 7)
    e2e_first = max_over_ranks(e2e_first)
    e2e_reps = int(min(60, max(5, math.ceil(args.min_time * 1e3 / max(e2e_first, 1e-3)))))
    if world > 1:
        t = torch.tensor([e2e_reps], device=dev)
        dist.broadcast(t, 0)
        e2e_reps = int(t[0])
    e2e_totals = [e2e_first]
    for r in range(1, e2e_reps):
        ms_, last = e2e_rep(args.steps, args.warmup + r * args.steps)
        e2e_totals.append(max_over_ranks(ms_))
    e2e_ms = statistics.median(e2e_totals)
    h2d = F * B * 8 + C_DENSE * B * 4 + B * 4
    d2h = 4

    # ---- launches per step (count one eager step; the graph replays exactly these)
    l0 = rt.launches
    d_, y_ = device_dict(0)
    trainer._eager_step(trainer.stage(d_, y_))
    torch.cuda.synchronize(dev)
    launches_per_step = rt.launches - l0

    # ---- roofline of the fused gather + FM kernel, timed alone, L2 flushed
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
    clocks.busy(True)
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
    clocks.busy(False)
    k_ms = statistics.mean(a.elapsed_time(b_) for a, b_ in kt[2:]) / GROUP
    peak, peak_src = load_peaks()
    # algorithmic bytes per sample: F*(k*4 + 4 [w] + 8 [id]) + 4 [logit]  (+ F*k*osize flat written for the MLP)
    alg_fm = F * (K_EMB * 4 + 4 + 8) + 4
    alg_flat = F * K_EMB * (2 if args.mlp == "bf16" else 4)
    achieved = (alg_fm + alg_flat) * B / (k_ms * 1e-3) / 1e9
    # DRAM bytes of one launch from the newest committed ncu --set full summary (read at run time; only
    # meaningful for the default workload it was captured on)
    traffic, traffic_src = None, None
    if world == 1 and args.dist == "zipf" and args.mlp == "bf16" and B == BATCH and args.config == "c2":
        tr_ = ncu_traffic("gather_fm_fwd")
        if tr_:
            traffic, traffic_src = tr_[0], (f"{tr_[1]} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one "
                                           f"launch of this command)")
    roof_gather = {"bound": "hbm", "kernel": ("gather_fm_fwd_stream_kernel<float,4,13,2,true,true> (gather + FM terms + Flatten, "
                                              f"one launch); rows of the {world - 1} other shards come over NVLink"
                                              if world > 1 and args.shard == "peer-pull" else
                                              "gather_fm_fwd_tile_kernel<float,4> (gather + FM terms + Flatten, one launch)"),
                   "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                   "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel_ms": k_ms,
                   "algorithmic_bytes_per_launch": (alg_fm + alg_flat) * B,
                   "note": "algorithmic bytes = B*(F*(4k+4+8)+4) FM terms + B*F*k*osize flattened operand written "
                           "for the MLP (osize 2 for the bf16 tensor-core MLP, 4 for fp32); timed alone with CUDA "
                           "events: L2 flushed, then 4 launches over 4 different id batches (rows touched >> L2) "
                           "per event pair"}

    # ---- second roofline: the fused FM backward + segment reduction + Adam, timed alone (plan precomputed, L2
    # flushed); algorithmic bytes = 408 B per unique row
    roof_apply = None
    if world == 1:
        from etr_b200.runtime import FusedFMGrad, SparsePlan
        plans = [SparsePlan(rt, id_batches[i], V) for i in range(4)]
        tiled = layer.table.record and K_EMB == 16 and FusedFMGrad.apply_kernel == "tile"
        if tiled:
            for p_ in plans:
                p_.prepare_fm()          # row descriptors + long-run items: part of the plan (ids only)
        dl_ = torch.randn(B, device=dev) * 1e-7
        sumv_ = torch.randn(B, K_EMB, device=dev) * 0.1
        dx_ = (torch.randn(B, col0 + F * K_EMB, device=dev) * 1e-7).to(torch.bfloat16)
        lr_ = torch.tensor([1e-3], device=dev)
        at = []
        clocks.busy(True)
        for i in range(12):
            g_ = FusedFMGrad(layer.table, id_batches[i % 4], K_EMB, dl_, sumv_, dx_, col0, plan=plans[i % 4])
            flush.zero_()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g_.apply(lr_, 0.9, 0.999, 1e-7)
            b_.record()
            at.append((a, b_))
        torch.cuda.synchronize(dev)
        clocks.busy(False)
        a_ms = statistics.median(a.elapsed_time(b_) for a, b_ in at[2:])
        n_u = statistics.mean(p_.n_unique for p_ in plans)
        a_bytes = n_u * 6 * (K_EMB + 1) * 4
        tr_ = ncu_traffic("fm_tile_kernel" if tiled else "fm_fused_") if (args.dist == "zipf" and B == BATCH and args.config == "c2") else None
        roof_apply = {"bound": "hbm", "kernel": ("fm_tile_kernel<1,6,0,4> (ONE launch: " if tiled else
                                                 "fm_fused_{classify,short|record,chunk,combine}_kernel (") +
                      "FM backward + sorted-run reduction + row-wise Adam; table layout: " +
                      ("256-byte [var|m|v] records" if layer.table.record else "three plain arrays") + ")",
                      "achieved": a_bytes / (a_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                      "frac": a_bytes / (a_ms * 1e-3) / 1e9 / peak, "traffic": tr_[0] if tr_ else None,
                      "traffic_source": tr_[1] if tr_ else None, "peak_source": peak_src, "kernel_ms": a_ms,
                      "unique_rows_per_launch": n_u, "algorithmic_bytes_per_launch": a_bytes,
                      "note": "408 B per unique row (3 reads + 3 writes of a 68-byte row, SURVEY 8d) x the unique rows of one "
                              "batch; timed alone, plan precomputed, L2 flushed before every call"}

    # ---- the same step in the two other arithmetic modes, so the headline can be read against them (N = 1, c2/c5):
    # value_fp32 = fp32 SIMT first MLP layer (the 1e-5-parity path); value_keras_dense = Keras-2.8 Adam semantics
    # (m, v decayed and var updated for ALL rows every step -- what the reference's own checkpoints show)
    extras = {}
    if world == 1 and not args.no_extras:
        try:
            K2 = min(args.steps, 10)
            if args.mlp == "bf16":
                lay32 = L.DeepFMRankingLayer(names, feature_dims=V, embedding_dims=K_EMB, continuous_features=cont, seed=1,
                                             check_ids=False, mlp_precision="fp32")
                tr32, fn32, g32 = make_stepper(lay32, "rowwise", use_graph)
                ms32 = timed_rep(tr32, fn32, K2, args.warmup)
                extras["value_fp32"] = {"value": B * K2 / (sum(ms32) * 1e-3), "unit": "samples/s", "ms_per_step": sum(ms32) / K2,
                                        "steps": K2, "cuda_graph": g32, "ms_each": [round(x, 3) for x in ms32],
                                        "what": "same step with the first MLP layer on the fp32 SIMT path (1e-5 parity)"}
                del tr32, fn32, lay32
                torch.cuda.empty_cache()
            trd, fnd, gd_ = make_stepper(layer, "keras_dense", False)
            msd = timed_rep(trd, fnd, min(K2, 5), args.warmup)
            extras["value_keras_dense"] = {"value": B * len(msd) / (sum(msd) * 1e-3), "unit": "samples/s",
                                           "ms_per_step": sum(msd) / len(msd), "steps": len(msd), "cuda_graph": gd_,
                                           "ms_each": [round(x, 3) for x in msd],
                                           "what": "same step with apply_mode='keras_dense': Keras-2.8 Adam on IndexedSlices "
                                                   "decays m, v and updates var for ALL V rows every step (13.8 GB of "
                                                   "table traffic per step at c2)"}
        except Exception as ex:
            extras["extras_error"] = repr(ex)[:300]

    check = None
    if world > 1 and not args.no_extras:
        check = sharded_check(L, world, rank, args.shard, dist)

    clk = clocks.stop()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = B * world * args.steps / (total_ms * 1e-3)
    e2e_value = B * world * args.steps / (e2e_ms * 1e-3)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_run(args.cpu_budget, B if args.config == "c2" else 8192, 4, args.dist)
    # the contract's ``roofline`` names the DOMINANT kernel (group) of the step by time; the other one rides along
    dominant, other, other_key = roof_gather, roof_apply, "roofline_apply"
    if roof_apply and roof_apply["kernel_ms"] > roof_gather["kernel_ms"]:
        dominant, other, other_key = roof_apply, roof_gather, "roofline_gather"
    line = {
        "metric": "train samples/sec DeepFM (Criteo-shape)", "value": value, "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (tables, FM terms, Adam)" + (" + bf16 tensor-core MLP layer 1" if args.mlp == "bf16" else ""),
        "data": "synthetic",
        "config": {**workload_config(args, B, use_graph),
                   "plan_ahead": ("the sorted-id plan of batch i+1 is built on the side stream during step i (its ids are staged one "
                                  "step early); every timed step contains exactly one sort" if ahead["on"] else
                                  "each step sorts its own ids (side stream, overlapped with forward + tower)")},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps, "last_loss": last, "repetitions": len(e2e_totals),
                "ms_per_step_min_median_max": [min(e2e_totals) / args.steps, e2e_ms / args.steps,
                                               max(e2e_totals) / args.steps]},
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step,
        "roofline": dominant,
        "step_ms_min_median_max": [min(step_ms), statistics.median(step_ms), max(step_ms)],
        "repetitions": {"count": len(rep_totals), "what": f"the {args.steps}-step timed region repeated; value = the median "
                        f"repetition (max over ranks each)", "ms_per_step_min_median_max":
                        [min(rep_totals) / args.steps, total_ms / args.steps, max(rep_totals) / args.steps]},
    }
    if other:
        line[other_key] = other
    line.update(extras)
    if check is not None:
        line["sharded_check"] = check
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
