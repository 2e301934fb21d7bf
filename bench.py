#!/usr/bin/env python
"""Benchmark of the CFFM training hot path (fwd + bwd + Adagrad update) on B200.

    python bench.py --gpus N --steps K --warmup W            # this implementation
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

Prints ONE JSON line (rank 0).
  value      training samples/s of the whole job with the batches already resident in HBM (CUDA events on the
             launching stream, max over ranks), Criteo-shaped synthetic workload (BASELINE.json configs[3]),
             in the PRIMARY arithmetic (--precision, default bf16x3 = split-bf16 tensor-core mode that meets
             the fp32 graph's 1e-4 logit tolerance, so the comparison with the fp32 CPU reference is like for like)
  e2e        the same metric through the host-buffer C-ABI call (cffm_train_submit_host: H2D of every batch and
             D2H of every loss inside the timed region)
  modes      the same workload in the other arithmetics: "bf16" (north star's 1e-2 mode) and, when asked for,
             "fp32" (SIMT contraction) -- value, ms_per_step, e2e, roofline each
  workloads  short records for the Frappe / ml-tag / Book-Crossing shapes (BASELINE.json configs[0..2]) with their
             own CPU baseline, so that the north star's 50x Frappe target is measured by the same command
  strong     (N > 1) the same global batch of 8192 split over the ranks, beside the weak-scaling `value`
  roofline   dominant kernel vs the measured peak; `frac` uses the FLOPs the kernel EXECUTES
             (the factorised layer-0 kernels execute fewer than the direct form's algorithmic count, which is
             reported as `algorithmic_*`)
  cpu_baseline  the CPU oracle (a torch-CPU port of the reference graph; TF 1.14 itself cannot run) on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train samples/sec (fwd+bwd+update)"
UNIT = "samples/s"
L2_BYTES = 126 * 1024 * 1024
PRIMARY = os.environ.get("CFFM_BENCH_PRECISION", "bf16x3")
DTYPE_NAME = {"fp32": "f32", "bf16": "bf16", "bf16x3": "bf16x3"}
DTYPE_NOTE = {
    "fp32": "fp32 SIMT contraction (the reference's arithmetic)",
    "bf16": "bf16 operands on tcgen05, fp32 accumulation / master weights (north star: logits within 1e-2)",
    "bf16x3": "split bf16 on tcgen05: operands as hi+lo bf16, hi*hi + lo*hi + hi*lo in one fp32 TMEM accumulator; "
              "logits within 1e-4 of the fp32 graph (tests/test_gpu_bf16x3.py), i.e. the reference's accuracy class",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("CFFM_BENCH_WORKLOAD", "criteo"),
                    choices=["criteo", "frappe", "ml-tag", "book-crossing"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (0: the workload's own)")
    ap.add_argument("--precision", default=PRIMARY, choices=["fp32", "bf16", "bf16x3"])
    ap.add_argument("--modes", default=os.environ.get("CFFM_BENCH_MODES", "auto"),
                    help="other arithmetics measured on the same workload: comma list of fp32,bf16,bf16x3; 'auto' = "
                         "the other tensor-core mode; 'none'")
    ap.add_argument("--workloads", default=os.environ.get("CFFM_BENCH_WORKLOADS", "auto"),
                    help="extra workload records: comma list, 'auto' (frappe,ml-tag,book-crossing at N=1 on the criteo run), 'none'")
    ap.add_argument("--tables", default=os.environ.get("CFFM_BENCH_TABLES", "auto"), choices=["auto", "replicated", "sharded"],
                    help="N > 1: embedding tables replicated (touched-row all-gather) or row-sharded (all-to-all of rows); "
                         "auto = sharded for the Criteo shape (BASELINE.json configs[3]), replicated for the small tables")
    ap.add_argument("--l2", default="auto", choices=["auto", "flush", "none"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--cpu-batch", type=int, default=0)
    ap.add_argument("--profile-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), tflops=float(d.get("bf16_tflops", 1590.0)),
                    tflops_sustained=float(d.get("bf16_tflops_sustained", 1400.0)), source="measured")
    return dict(hbm_gbs=6650.0, tflops=1590.0, tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def workload_spec(name, batch):
    from cffm_b200 import synth
    w = synth.WORKLOADS[name]
    cards = synth.field_cards(name)
    return dict(name=name, F=len(cards), M=int(sum(cards)), B=batch or w["batch"], K=w["K"], activation=w["activation"])


def l2_policy(spec, B, mode="auto"):
    """(flush?, text): the working set (tables + accumulators + activations) is either larger than L2 several times
    over -- every step misses -- or a 256 MiB buffer is written between timed steps."""
    F, K, M = spec["F"], spec["K"], spec["M"]
    P = F * (F - 1) // 2
    act_bytes = 4 * B * P * sum((K >> (l + 1)) ** 2 for l in range(int(np.log2(K)) - 1)) * 2
    ws_bytes = 4 * M * (2 * K + 1) * 2 + act_bytes
    flush = mode == "flush" or (mode == "auto" and ws_bytes < 4 * L2_BYTES)
    return flush, ("flushed between steps (256 MiB write)" if flush else "working set %.1f GB > L2, no flush" % (ws_bytes / 1e9))


def shared_config(spec, world, batch_per_gpu=None, l2_mode="auto"):
    """`config` names the workload and how it is run: byte-identical on both arms (ours / --impl reference)."""
    F, K = spec["F"], spec["K"]
    B = batch_per_gpu or spec["B"]
    return {"workload": "%s-shaped synthetic" % spec["name"], "num_field": F, "pairs": F * (F - 1) // 2, "dims": K,
            "features_M": spec["M"], "batch_per_gpu": B, "global_batch": B * world, "activation": spec["activation"],
            "loss_type": "square_loss", "optimizer": "AdagradOptimizer", "lr": 0.05,
            "parallelism": "dp%d" % world, "l2": l2_policy(spec, B, l2_mode)[1]}


def algorithmic_work(spec, tag, B, world):
    """(kind, amount per launch): FLOPs for the conv contractions, bytes for gather / update kernels
    (per-sample figures x samples per launch)."""
    F, K, P = spec["F"], spec["K"], spec["F"] * (spec["F"] - 1) // 2
    if tag.startswith("conv_"):
        l = int(tag.rsplit("l", 1)[1])
        Ho = K >> (l + 1)
        return "flops", 2.0 * B * Ho * Ho * 4 * P * P          # 8 P^2 per output position
    if tag == "gather_outer":
        return "bytes", B * F * (4 + 2 * 4 * K)                  # id + row read + row write
    if tag == "sparse_adagrad":
        n = B * F * world
        return "bytes", n * (4 + 4 + 4 * (2 * K + 1)) + n * 16 * (2 * K + 1) * 0.25  # grad rows + (<=) touched rows rw
    if tag == "inner_linear_fwd":
        return "bytes", B * F * (4 + 4 * K + 4) + 4 * P * K
    if tag == "inner_linear_bwd":
        return "bytes", B * F * (4 + 4 * K + 4) + B * F * (4 * K + 4) + 4 * P * K   # rows in, gradient rows out
    return "bytes", 0.0


def executed_flops(spec, tag, B, precision):
    """FLOPs the kernel really issues when that differs from the algorithmic (direct-form) figure.
    bf16: the layer-0 kernels run the convolution over the rank-one cube in factorised form:
    per 8-sample tile and channel two (forward) or four (data gradient, weight gradient: two) small MMAs.
    bf16x3: every contraction is three MMAs (hi*hi + lo*hi + hi*lo); layer 0 factorised as in bf16, layers >= 1 direct."""
    if not tag.startswith("conv_"):
        return None
    F, P = spec["F"], spec["F"] * (spec["F"] - 1) // 2
    fact = tag.endswith("_l0") and 2 * F <= 80 and F >= 16 and B >= 512
    KA, Q16, tiles = (2 * F + 15) // 16 * 16, (P + 15) // 16 * 16, (B + 7) // 8
    per = 2.0 * 128 * KA * (KA + 128)          # one K=KA step + one K=128 (block-diagonal) step of a (tile, channel)
    table = {"conv_fwd_l0": per, "conv_dgrad_l0": 2 * per}
    if FACT_WGRAD:
        table["conv_wgrad_l0"] = 2.0 * 128 * 128 * (16 + KA)   # per-sample K=16 MMAs (N=16 each) + one K=128 step (N=KA)
    if precision == "bf16x3":
        if fact and tag in SPLIT_FACT:                          # factorised form with Z / E split into hi + lo
            return 3.0 * tiles * Q16 * table[tag]
        return 3.0 * algorithmic_work(spec, tag, B, 1)[1]
    if precision != "bf16" or not fact:
        return None
    return tiles * Q16 * table[tag] if tag in table else None


FACT_WGRAD = True    # the layer-0 weight gradient runs in factorised form as well (conv0_wfact.cuh)
SPLIT_FACT = ("conv_fwd_l0", "conv_wgrad_l0", "conv_dgrad_l0")   # layer-0 kernels that run in factorised form in split (bf16x3) mode


def ncu_traffic(spec, tag, B, precision):
    """DRAM bytes of one launch of `tag` from the committed ncu --set full capture (profiles/), or None when
    the capture was taken on another workload / arithmetic."""
    for fname in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        path = os.path.join(ROOT, "profiles", fname)
        try:
            with open(path) as f:
                t = json.load(f)
        except (OSError, ValueError):
            continue
        w = t.get("workload", {})
        if (w.get("name"), w.get("F"), w.get("K"), w.get("B"), w.get("precision")) != (spec["name"], spec["F"], spec["K"], B, precision):
            continue
        k = t.get("kernels", {}).get(tag)
        if k:
            return k["traffic_bytes"]
    return None


def roofline_of(spec, kernel_table, B, world, precision):
    """Roofline record of the kernel with the largest share of the step."""
    pk = peaks()
    top = next(iter(kernel_table))
    traffic = ncu_traffic(spec, top, B, precision)
    kind, amount = algorithmic_work(spec, top, B, world)
    avg_ms = kernel_table[top]["avg_ms"]
    if kind == "flops" and amount > 0:
        ex = executed_flops(spec, top, B, precision) or amount
        peak = pk["tflops_sustained"]
        ach = ex / (avg_ms / 1e3) / 1e12
        algo = amount / (avg_ms / 1e3) / 1e12
        return {"kernel": top, "bound": "tensor", "achieved": round(ach, 3), "peak": peak, "unit": "TFLOP/s",
                "frac": round(ach / peak, 5), "traffic": traffic, "share_of_step": kernel_table[top]["share"],
                "peak_source": pk["source"] + " bf16 sustained (kernel timed inside the step)",
                "flops_counted": "executed by the tensor cores (layer 0 in factorised form; three MMAs per product in bf16x3)",
                "executed_flops_per_launch": ex, "algorithmic_flops_per_launch": amount,
                "algorithmic_achieved": round(algo, 3), "algorithmic_frac": round(algo / peak, 5)}
    ach = (amount / (avg_ms / 1e3) / 1e9) if amount else 0.0
    return {"kernel": top, "bound": "hbm", "achieved": round(ach, 3), "peak": pk["hbm_gbs"], "unit": "GB/s",
            "frac": round(ach / pk["hbm_gbs"], 5), "traffic": traffic, "share_of_step": kernel_table[top]["share"],
            "peak_source": pk["source"], "algorithmic_bytes_per_launch": amount}


# ------------------------------------------------------------------------------------------------
class Job:
    """torch / torch.distributed state shared by the measurements of one bench.py process."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        torch.cuda.set_device(self.local)
        self.flush_buf = None

    def barrier(self, use_world=True):
        if self.world > 1 and use_world:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def flush(self):
        if self.flush_buf is None:
            self.flush_buf = self.torch.empty(2 * L2_BYTES // 4, dtype=self.torch.float32, device="cuda")
        self.flush_buf.zero_()


def tables_mode(args, wl_name, world):
    if world == 1:
        return "single"
    if args.tables != "auto":
        return args.tables
    return "sharded" if wl_name == "criteo" else "replicated"


def measure(job, args, wl_name, precision, steps, warmup, batch=0, profile=True, sample_clocks=False, use_world=True, tables=None):
    """One workload in one arithmetic: device-timed value, end-to-end value, per-kernel table, roofline."""
    torch, dist = job.torch, job.dist
    from cffm_b200 import Engine, comm_unique_id, synth
    world = job.world if use_world else 1
    rank = job.rank
    spec = workload_spec(wl_name, batch)
    B, F, K, M = spec["B"], spec["F"], spec["K"], spec["M"]
    n_pool = 8
    ids, _ = synth.make_ids(wl_name, n_pool * B, seed=2021 + 17 * rank)
    labels = synth.make_labels(n_pool * B, seed=2021 + 17 * rank)
    tables = tables or tables_mode(args, wl_name, world)
    eng = Engine(M, F, K, K, activation=spec["activation"], max_batch=B, precision=precision, device=job.local, seed=2021,
                 shard=(rank, world) if (tables == "sharded" and world > 1) else None)
    if world > 1:
        uid = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(uid[0], rank, world)
    ids_d = torch.from_numpy(ids).cuda()
    y_d = torch.from_numpy(labels).cuda()
    loss_d = torch.zeros(1, device="cuda")
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    def step_dev(i):
        j = (i % n_pool) * B
        eng.train_step_dev(ids_d[j:j + B].data_ptr(), y_d[j:j + B].data_ptr(), B, loss_d.data_ptr(), sp)

    flush, l2_text = l2_policy(spec, B, args.l2)

    for i in range(warmup):
        step_dev(i)
    job.barrier(use_world)
    launches0 = eng.launch_count()
    sampler = ClockSampler(job.local) if (sample_clocks and rank == 0) else None
    if sampler:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    job.barrier(use_world)
    t_wall0 = time.perf_counter()
    for i in range(steps):
        if flush:
            job.flush()
        ev[i][0].record(stream)
        step_dev(warmup + i)
        ev[i][1].record(stream)
    job.barrier(use_world)
    t_wall = time.perf_counter() - t_wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    dev_ms = job.max_over_ranks(dev_ms) if use_world else dev_ms
    value = world * B * steps / (dev_ms / 1e3)
    final_loss = float(loss_d.item())

    # ---- end to end through the host-buffer C ABI (pipelined submit: H2D + D2H every step) ----
    ids_h = [np.ascontiguousarray(ids[j * B:(j + 1) * B]) for j in range(n_pool)]
    y_h = [np.ascontiguousarray(labels[j * B:(j + 1) * B]) for j in range(n_pool)]
    for i in range(2):
        eng.train_submit(ids_h[i % n_pool], y_h[i % n_pool])
    eng.train_flush()
    job.barrier(use_world)
    t0 = time.perf_counter()
    for i in range(steps):
        eng.train_submit(ids_h[i % n_pool], y_h[i % n_pool])
    eng.train_flush()
    job.barrier(use_world)
    e2e_s = time.perf_counter() - t0
    e2e_s = job.max_over_ranks(e2e_s) if use_world else e2e_s
    e2e_value = world * B * steps / e2e_s
    uses_graph = eng.uses_graph()

    # ---- per-kernel device time: event-bracketed eager steps; the first one is discarded (clock / cache state) ----
    roof, kernel_table = None, {}
    if profile:
        eng.profile(True)
        eng.train_step(ids_h[0], y_h[0])
        eng.profile_report(reset=True)
        nprof = max(1, args.profile_steps)
        for i in range(nprof):
            eng.train_step(ids_h[(i + 1) % n_pool], y_h[(i + 1) % n_pool])
        rep = eng.profile_report(reset=True)
        eng.profile(False)
        total_ms = sum(ms for _, ms in rep.values()) or 1e-9
        for tag, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
            kind, amount = algorithmic_work(spec, tag, B, world)
            avg_ms = ms / max(1, n)
            kernel_table[tag] = {"launches_per_step": n / nprof, "avg_ms": round(avg_ms, 5), "share": round(ms / total_ms, 4)}
            if amount > 0:
                ex = executed_flops(spec, tag, B, precision) if kind == "flops" else None
                rate = (ex or amount) / (avg_ms / 1e3)
                kernel_table[tag]["achieved"] = round(rate / (1e12 if kind == "flops" else 1e9), 3)
                kernel_table[tag]["unit"] = "TFLOP/s" if kind == "flops" else "GB/s"
                if ex:   # what the direct form would count
                    kernel_table[tag]["algorithmic"] = round(amount / (avg_ms / 1e3) / 1e12, 3)
        roof = roofline_of(spec, kernel_table, B, world, precision)
    job.barrier(use_world)
    eng.close()
    return {
        "spec": spec, "world": world, "value": round(value, 2), "ms_per_step": round(dev_ms / steps, 4), "steps": steps, "warmup": warmup,
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": int(B * F * 4 + B * 4),
                "d2h_bytes_per_step": 4, "api": "cffm_train_submit_host (pinned double buffer)"},
        "gpu_launches": int(launches), "roofline": roof, "kernels": kernel_table, "clocks": clocks,
        "wall_s_timed_region": round(t_wall, 4), "final_loss": final_loss, "cuda_graph": uses_graph,
        "l2": l2_text, "precision": precision, "tables": tables,
    }


def brief(rec, with_kernels=False):
    keep = ("value", "ms_per_step", "e2e", "gpu_launches", "roofline", "final_loss", "cuda_graph", "l2", "precision")
    out = {k: rec[k] for k in keep}
    out["unit"] = UNIT
    out["dtype_note"] = DTYPE_NOTE[rec["precision"]]
    if with_kernels:
        out["kernels"] = rec["kernels"]
    return out


def run_ours(args):
    job = Job()
    world, rank = job.world, job.rank
    primary = args.precision
    steps = args.steps if args.steps is not None else (5 if args.workload == "criteo" and primary == "fp32" else 20)
    warmup = max(3, args.warmup if args.warmup is not None else 3)

    main = measure(job, args, args.workload, primary, steps, warmup, batch=args.batch, sample_clocks=True)
    spec = main["spec"]
    B, K = spec["B"], spec["K"]

    # ---- the same workload in the other arithmetics ----
    modes = {}
    if args.modes == "auto":
        want = [p for p in ("bf16", "bf16x3") if p != primary] if K == 32 else []
    elif args.modes == "none":
        want = []
    else:
        want = [p for p in args.modes.split(",") if p and p != primary]
    for prec in want:
        st = min(steps, 5) if (prec == "fp32" and args.workload == "criteo") else steps
        modes[prec] = brief(measure(job, args, args.workload, prec, st, 3, batch=args.batch))

    # ---- strong scaling beside the weak one: the global batch of the N=1 run split over the ranks ----
    strong, other_tables = None, None
    if world > 1 and B % world == 0 and args.workload == "criteo":
        r = measure(job, args, args.workload, primary, steps, 3, batch=B // world, profile=False)
        strong = {"global_batch": B, "batch_per_gpu": B // world, "value": r["value"], "ms_per_step": r["ms_per_step"],
                  "e2e": r["e2e"], "tables": r["tables"], "note": "fixed global batch (strong scaling); `value` above is weak scaling"}
    if world > 1 and args.workload == "criteo":
        # the other table layout on the same weak-scaling workload
        alt = "replicated" if main["tables"] == "sharded" else "sharded"
        r = measure(job, args, args.workload, primary, steps, 3, batch=args.batch, profile=True, tables=alt)
        other_tables = {"tables": alt, "value": r["value"], "ms_per_step": r["ms_per_step"], "e2e": r["e2e"],
                        "kernels": {k: v for k, v in r["kernels"].items() if k.startswith(("shard_", "dp_", "sort_", "sparse_"))}}

    # ---- the reference's own dataset shapes (N = 1 only; the data-parallel path is measured on the Criteo shape) ----
    workloads = {}
    if args.workloads == "auto":
        extra = ["frappe", "ml-tag", "book-crossing"] if (world == 1 and args.workload == "criteo") else []
    elif args.workloads == "none":
        extra = []
    else:
        extra = [w for w in args.workloads.split(",") if w and w != args.workload]
    for wl in extra:
        if rank != 0:
            continue
        rec = {}
        for prec in ("bf16x3", "bf16", "fp32"):
            r = measure(job, args, wl, prec, 50, 5, profile=False, use_world=False)
            rec[prec] = {"value": r["value"], "ms_per_step": r["ms_per_step"], "e2e": r["e2e"]["value"],
                         "gpu_launches_per_step": r["gpu_launches"] / 50, "l2": r["l2"]}
        wspec = workload_spec(wl, 0)
        rec["config"] = shared_config(wspec, 1)
        if not args.no_cpu_baseline:
            cpu = cpu_reference(wspec, args, budget_s=5.0)
            rec["cpu_baseline"] = cpu
            rec["e2e_vs_cpu"] = {p: round(rec[p]["e2e"] / cpu["value"], 1) for p in ("bf16x3", "bf16", "fp32")}
        workloads[wl] = rec

    # ---- CPU baseline: the oracle on the host cores, rank 0, N=1 only ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(spec, args, budget_s=args.cpu_seconds)

    if rank == 0:
        cfg = shared_config(spec, world, l2_mode=args.l2)
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE_NAME[primary], "dtype_note": DTYPE_NOTE[primary], "data": "synthetic",
            "config": cfg,
            "run": {"l2": main["l2"], "cuda_graph": main["cuda_graph"], "precision": primary,
                    "tables": main["tables"],
                    "parallelism": ("dp%d, dense allreduce + " % world) + (
                        "row-sharded tables: all-to-all of unique rows / gradient sums" if main["tables"] == "sharded" else
                        "replicated tables: touched-row allgather" if world > 1 else "single GPU")},
            "clocks": main["clocks"], "e2e": main["e2e"], "gpu_launches": main["gpu_launches"],
            "roofline": main["roofline"], "kernels": main["kernels"], "cpu_baseline": cpu,
            "modes": modes, "workloads": workloads, "strong": strong, "other_table_layout": other_tables,
            "wall_s_timed_region": main["wall_s_timed_region"], "final_loss": main["final_loss"],
        }
        print(json.dumps(line))
        sys.stdout.flush()
    # tear down together: ncclCommDestroy must not race with a rank that is still working
    job.barrier()
    if world > 1:
        job.dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def cpu_batch_for(spec, requested=0):
    """Micro-batch of the CPU leg: the largest that fits the host's free memory with a wide margin (the oracle
    materialises the interaction cube like the reference does: P*K*K*4 bytes per sample and ~12 copies of it through
    autograd), at most the workload's own batch."""
    if requested:
        return requested
    P = spec["F"] * (spec["F"] - 1) // 2
    per_sample = 12.0 * P * spec["K"] * spec["K"] * 4 + 1e5
    try:
        import psutil
        avail = float(psutil.virtual_memory().available)
    except Exception:
        avail = 16e9
    b = int(min(spec["B"], max(8, (0.2 * avail) // per_sample)))
    if spec["name"] == "criteo":
        b = min(b, 128)            # bounded sample: ~2 s of 16-core work per step
    p2 = 1
    while p2 * 2 <= b:
        p2 *= 2
    return p2


def cpu_reference(spec, args, budget_s=15.0, steps=None, warmup=1):
    """The reference's CPU path: TensorFlow 1.14 cannot run here, so this is the oracle -- a torch-CPU
    port of the same graph (kind "port") -- on all host threads, on a bounded sample of the workload: each CPU
    step is fwd + bwd + Adagrad update of one micro-batch of the workload's batch."""
    import torch
    from cffm_b200 import synth
    from oracle.cffm_ref import CFFMRef
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    F, K, M = spec["F"], spec["K"], spec["M"]
    Bc = cpu_batch_for(spec, args.cpu_batch)
    ids, _ = synth.make_ids(spec["name"], 4 * Bc, seed=99)
    y = synth.make_labels(4 * Bc, seed=99)
    m = CFFMRef(M, F, K, K, activation=spec["activation"], dtype=torch.float32, seed=1)
    for i in range(warmup):
        m.train_step(ids[:Bc], y[:Bc])
    n, t0 = 0, time.perf_counter()
    while True:
        j = (n % 4) * Bc
        m.train_step(ids[j:j + Bc], y[j:j + Bc])
        n += 1
        el = time.perf_counter() - t0
        if (steps is not None and n >= steps) or (steps is None and (el > budget_s or n >= 200)):
            break
    el = time.perf_counter() - t0
    return {"value": round(n * Bc / el, 3), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d steps, each fwd+bwd+update of a %d-sample micro-batch of the %s-shaped workload's %d-sample batch "
                      "(torch-CPU restatement of the TF-1.14 graph, fp32)" % (n, Bc, spec["name"], spec["B"]),
            "ms_per_step": round(1e3 * el / n, 2)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    spec = workload_spec(args.workload, args.batch)
    steps = args.steps if args.steps is not None else 5
    warmup = args.warmup if args.warmup is not None else 1
    cpu = cpu_reference(spec, args, steps=steps, warmup=warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": cpu["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": shared_config(spec, world, l2_mode=args.l2),
        "run": {"note": "CPU restatement of the reference TF-1.14 graph (TensorFlow is not installable here); rank 0 only"},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
