#!/usr/bin/env python
"""Benchmark of the CFFM training hot path (fwd + bwd + Adagrad update) on B200.

    python bench.py --gpus N --steps K --warmup W            # this implementation
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

Prints ONE JSON line (rank 0).  `value` = training samples/s of the whole job with the batches
already resident in HBM (device-timed with CUDA events, max over ranks); `e2e` = the same metric
through the host-buffer C-ABI call (cffm_train_submit_host: H2D of every batch and D2H of every
loss inside the timed region); `roofline` = dominant kernel vs the measured peak; `cpu_baseline`
= the CPU oracle (a port of the reference graph, TF 1.14 itself cannot run) on the host cores.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("CFFM_BENCH_WORKLOAD", "criteo"),
                    choices=["criteo", "frappe", "ml-tag", "book-crossing"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (0: the workload's own)")
    # bf16 operands on the tensor cores, fp32 accumulation / master weights (north star: logits within
    # 1e-2 of the fp32 graph); --precision fp32 runs the SIMT contraction at reference arithmetic
    ap.add_argument("--precision", default=os.environ.get("CFFM_BENCH_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--l2", default="auto", choices=["auto", "flush", "none"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--cpu-batch", type=int, default=0)
    ap.add_argument("--profile-steps", type=int, default=2)
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
    return "bytes", 0.0


def executed_flops(spec, tag, B, precision):
    """FLOPs the kernel really issues when it differs from the algorithmic (direct-form) figure: the layer-0
    kernels of the bf16 path run the convolution over the rank-one cube in factorised form:
    per 8-sample tile and channel two (forward) or four (data gradient, weight gradient: two) small MMAs."""
    if precision != "bf16" or not tag.endswith("_l0") or 2 * spec["F"] > 80:
        return None
    F, P = spec["F"], spec["F"] * (spec["F"] - 1) // 2
    KA, Q16, tiles = (2 * F + 15) // 16 * 16, (P + 15) // 16 * 16, (B + 7) // 8
    per = 2.0 * 128 * KA * (KA + 128)          # one K=KA step + one K=128 (block-diagonal) step of a (tile, channel)
    table = {"conv_fwd_l0": per, "conv_dgrad_l0": 2 * per}
    if FACT_WGRAD:
        table["conv_wgrad_l0"] = 2.0 * 128 * 128 * (16 + KA)   # per-sample K=16 MMAs (N=16 each) + one K=128 step (N=KA)
    return tiles * Q16 * table[tag] if tag in table else None


FACT_WGRAD = True    # the layer-0 weight gradient runs in factorised form as well (conv0_wfact.cuh)


def ncu_traffic(spec, tag, B, precision):
    """DRAM bytes of one launch of `tag` from the committed ncu --set full capture (profiles/), or None when
    the capture was taken on another workload."""
    path = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
    except (OSError, ValueError):
        return None
    w = t.get("workload", {})
    if (w.get("name"), w.get("F"), w.get("K"), w.get("B"), w.get("precision")) != (spec["name"], spec["F"], spec["K"], B, precision):
        return None
    k = t.get("kernels", {}).get(tag)
    return k["traffic_bytes"] if k else None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from cffm_b200 import Engine, comm_unique_id, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    spec = workload_spec(args.workload, args.batch)
    B, F, K, M = spec["B"], spec["F"], spec["K"], spec["M"]
    steps = args.steps if args.steps is not None else (5 if args.workload == "criteo" and args.precision == "fp32" else 20)
    warmup = args.warmup if args.warmup is not None else 3
    warmup = max(3, warmup)

    # ---- synthetic batches: a pool of distinct batches per rank (the global batch is world*B) ----
    n_pool = 8
    ids, _ = synth.make_ids(args.workload, n_pool * B, seed=2021 + 17 * rank)
    labels = synth.make_labels(n_pool * B, seed=2021 + 17 * rank)
    eng = Engine(M, F, K, K, activation=spec["activation"], max_batch=B, precision=args.precision, device=local, seed=2021)
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

    # working set: tables + accumulators + activations.  Larger than L2 -> no flush needed.
    P = F * (F - 1) // 2
    act_bytes = 4 * B * P * sum((K >> (l + 1)) ** 2 for l in range(int(np.log2(K)) - 1)) * 2
    ws_bytes = 4 * M * (2 * K + 1) * 2 + act_bytes
    flush = args.l2 == "flush" or (args.l2 == "auto" and ws_bytes < 4 * L2_BYTES)
    flush_buf = torch.empty(2 * L2_BYTES // 4, dtype=torch.float32, device="cuda") if flush else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step_dev(i)
    barrier()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(steps):
        if flush:
            flush_buf.zero_()
        ev[i][0].record(stream)
        step_dev(warmup + i)
        ev[i][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = world * B * steps / (dev_ms / 1e3)
    final_loss = float(loss_d.item())

    # ---- end to end through the host-buffer C ABI (pipelined submit: H2D + D2H every step) ----
    ids_h = [np.ascontiguousarray(ids[j * B:(j + 1) * B]) for j in range(n_pool)]
    y_h = [np.ascontiguousarray(labels[j * B:(j + 1) * B]) for j in range(n_pool)]
    for i in range(2):
        eng.train_submit(ids_h[i % n_pool], y_h[i % n_pool])
    eng.train_flush()
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        eng.train_submit(ids_h[i % n_pool], y_h[i % n_pool])
    eng.train_flush()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * steps / float(t.item())

    # ---- per-kernel device time: event-bracketed eager steps after the timed region ----
    roof = None
    kernel_table = {}
    if rank == 0 or world > 1:
        eng.profile(True)
        eng.profile_report(reset=True)
        for i in range(max(1, args.profile_steps)):
            eng.train_step(ids_h[i % n_pool], y_h[i % n_pool])
        rep = eng.profile_report(reset=True)
        eng.profile(False)
        pk = peaks()
        total_ms = sum(ms for _, ms in rep.values()) or 1e-9
        for tag, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
            kind, amount = algorithmic_work(spec, tag, B, world)
            avg_ms = ms / max(1, n)
            kernel_table[tag] = {"launches_per_step": n / max(1, args.profile_steps), "avg_ms": round(avg_ms, 5),
                                 "share": round(ms / total_ms, 4)}
            if amount > 0:
                rate = amount / (avg_ms / 1e3)
                kernel_table[tag]["achieved"] = round(rate / (1e12 if kind == "flops" else 1e9), 3)
                kernel_table[tag]["unit"] = "TFLOP/s" if kind == "flops" else "GB/s"
                ex = executed_flops(spec, tag, B, args.precision)
                if ex:   # factorised kernel: what the tensor cores really do (the figure to hold against the peak)
                    kernel_table[tag]["executed"] = round(ex / (avg_ms / 1e3) / 1e12, 3)
        top = next(iter(kernel_table))
        traffic = ncu_traffic(spec, top, B, args.precision)
        kind, amount = algorithmic_work(spec, top, B, world)
        avg_ms = kernel_table[top]["avg_ms"]
        if kind == "flops" and amount > 0:
            ex = executed_flops(spec, top, B, args.precision)
            ach = amount / (avg_ms / 1e3) / 1e12            # algorithmic FLOPs, as the contract says
            peak = pk["tflops_sustained"]
            roof = {"kernel": top, "bound": "tensor", "achieved": round(ach, 3), "peak": peak, "unit": "TFLOP/s",
                    "frac": round(ach / peak, 5), "traffic": traffic, "share_of_step": kernel_table[top]["share"],
                    "peak_source": pk["source"] + " bf16 sustained (kernel timed inside the step)",
                    "algorithmic_flops_per_launch": amount, "executed_flops_per_launch": ex or amount,
                    # factorised kernels issue fewer FLOPs than the direct form the algorithmic figure counts, so
                    # `frac` can exceed 1; what the tensor cores really sustain is executed_achieved / executed_frac
                    "executed_achieved": round((ex or amount) / (avg_ms / 1e3) / 1e12, 3),
                    "executed_frac": round((ex or amount) / (avg_ms / 1e3) / 1e12 / peak, 5)}
        else:
            ach = (amount / (avg_ms / 1e3) / 1e9) if amount else 0.0
            roof = {"kernel": top, "bound": "hbm", "achieved": round(ach, 3), "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": round(ach / pk["hbm_gbs"], 5), "traffic": traffic, "share_of_step": kernel_table[top]["share"],
                    "peak_source": pk["source"], "algorithmic_bytes_per_launch": amount}

    # ---- CPU baseline: the oracle on the host cores, rank 0, N=1 only ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(spec, args, budget_s=args.cpu_seconds)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": round(dev_ms / steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": "%s-shaped synthetic" % spec["name"], "num_field": F, "pairs": P, "dims": K,
                       "features_M": M, "batch_per_gpu": B, "global_batch": B * world, "activation": spec["activation"],
                       "loss_type": "square_loss", "optimizer": "AdagradOptimizer", "lr": 0.05,
                       "parallelism": "dp%d (replicated tables, dense allreduce + touched-row allgather)" % world,
                       "l2": "flushed between steps (256 MiB write)" if flush else
                             "working set %.1f GB > L2, no flush" % (ws_bytes / 1e9),
                       "cuda_graph": True, "precision": args.precision},
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": int(B * F * 4 + B * 4),
                    "d2h_bytes_per_step": 4, "api": "cffm_train_submit_host (pinned double buffer)"},
            "gpu_launches": int(launches),
            "roofline": roof, "kernels": kernel_table, "cpu_baseline": cpu,
            "wall_s_timed_region": round(t_wall, 4), "final_loss": final_loss,
        }
        print(json.dumps(line))
        sys.stdout.flush()
    # tear down together: ncclCommDestroy must not race with a rank that is still working
    barrier()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def cpu_reference(spec, args, budget_s=15.0, steps=None, warmup=1):
    """The reference's CPU path: TensorFlow 1.14 cannot run here, so this is the oracle -- a torch-CPU
    port of the same graph (kind "port") -- on all host threads, on a bounded sample of the workload."""
    import torch
    from cffm_b200 import synth
    from oracle.cffm_ref import CFFMRef
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    F, K, M = spec["F"], spec["K"], spec["M"]
    # bounded sample: the materialised cube is P*K*K*4 bytes per sample (3 MB at the Criteo shape)
    Bc = args.cpu_batch or min(spec["B"], 32 if spec["name"] == "criteo" else 256)
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
            "sample": "%d steps of %d-sample batches of the %s-shaped workload (torch-CPU restatement of the TF graph)"
                      % (n, Bc, spec["name"]), "ms_per_step": round(1e3 * el / n, 2)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    spec = workload_spec(args.workload, args.batch)
    steps = args.steps if args.steps is not None else 5
    warmup = args.warmup if args.warmup is not None else 1
    steps_c = min(steps, 20)
    cpu = cpu_reference(spec, args, steps=steps_c, warmup=min(warmup, 2))
    F, K = spec["F"], spec["K"]
    line = {
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": world, "steps": steps_c,
        "warmup": min(warmup, 2), "ms_per_step": cpu["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s-shaped synthetic" % spec["name"], "num_field": F, "pairs": F * (F - 1) // 2, "dims": K,
                   "features_M": spec["M"], "note": "CPU restatement of the reference TF-1.14 graph (TensorFlow not installable)"},
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
