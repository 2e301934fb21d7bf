#!/usr/bin/env python
"""Inference-only scoring sweep (BASELINE.json configs[4]): forward latency and samples/s of
sess.run(self.out) (CFFM.py:596) for batch 64..65536, k = 16/32/64, Frappe and Criteo shapes.
Device-timed (CUDA events on the launching stream), ids resident in HBM.  One JSON line per case."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    import torch
    from cffm_b200 import Engine, synth, CffmError
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="frappe,criteo")
    ap.add_argument("--batches", default="64,256,1024,4096,16384,65536")
    ap.add_argument("--dims", default="16,32,64")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--precisions", default="bf16,bf16x3,fp32")
    a = ap.parse_args()
    stream = torch.cuda.current_stream()
    for shape in a.shapes.split(","):
        cards = synth.field_cards(shape)
        F, M = len(cards), int(sum(cards))
        act = synth.WORKLOADS[shape]["activation"]
        for K in [int(k) for k in a.dims.split(",")]:
            for prec in a.precisions.split(","):
                for B in [int(b) for b in a.batches.split(",")]:
                    P = F * (F - 1) // 2
                    S = sum((K >> (l + 1)) ** 2 for l in range(int(np.log2(K)) - 1))
                    flops = 2.0 * B * 4 * P * P * S
                    if prec == "fp32" and flops > 4e13:   # keep the SIMT cases to a few seconds
                        continue
                    # activations of a batch: B * S * Pp * 2 B (x2 in split mode): stay inside the GPU
                    if B * S * ((P + 63) // 64 * 64) * (4 if prec == "bf16x3" else 2 if prec == "bf16" else 4) > 120e9:
                        continue
                    try:
                        eng = Engine(M, F, K, K, activation=act, max_batch=B, precision=prec, seed=1)
                    except CffmError as e:
                        print(json.dumps({"shape": shape, "k": K, "batch": B, "precision": prec, "skipped": str(e)[:80]}))
                        continue
                    ids, _ = synth.make_ids(shape, B, seed=3)
                    ids_d = torch.from_numpy(ids).cuda(); out_d = torch.empty(B, device="cuda")
                    for _ in range(3):
                        eng.forward_dev(ids_d.data_ptr(), B, out_d.data_ptr(), stream.cuda_stream)
                    torch.cuda.synchronize()
                    reps = a.reps if flops < 5e12 else 3
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    for _ in range(reps):
                        eng.forward_dev(ids_d.data_ptr(), B, out_d.data_ptr(), stream.cuda_stream)
                    e1.record(stream); torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / reps
                    print(json.dumps({"shape": shape, "num_field": F, "k": K, "batch": B, "precision": prec,
                                      "fwd_ms": round(ms, 4), "samples_per_s": round(B / ms * 1e3, 1),
                                      "conv_tflops": round(flops / ms / 1e9, 2)}), flush=True)
                    eng.close()


if __name__ == "__main__":
    main()
