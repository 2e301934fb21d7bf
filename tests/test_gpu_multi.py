"""Data parallelism over NCCL: N ranks x B/N samples == 1 rank x B samples (needs >= 2 GPUs)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.timeout(280)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_data_parallel_equals_single_gpu(tmp_path, precision):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    out = tmp_path / "dp.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(HERE, "dp_worker.py"), str(out), precision]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=260)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.load(open(out))
    assert res["replicas_identical"]
    tol = 1e-4 if precision == "fp32" else 2e-2
    for a, b in zip(res["losses"], res["ref_losses"]):
        assert abs(a - b) <= tol * max(1.0, abs(b)), (res["losses"], res["ref_losses"])
    _check_weight_drift(res, precision)


def _check_weight_drift(res, precision):
    """Adagrad with acc0 = 1e-8 turns the first steps into +-lr*sign(g): an entry whose gradient changes sign between the
    two runs ends 0.05 - 0.1 apart.  In fp32 only summation order differs and that happens to < 2 % of the entries.  In bf16
    the runs also tile the batch differently (B/N samples per rank vs B on one GPU), the conv gradients carry percent-level
    noise, and for filters whose gradient is a cancelling sum most signs are noise (measured: 20 - 66 % of layer 0's
    entries): there the figure is reported and only bounded by what a few +-lr steps can do."""
    if precision == "fp32":
        bad = {k: v for k, v in res["frac_over_2e-3"].items() if v > 0.02}
        assert not bad, bad
    else:
        print("bf16 weight drift (fraction of entries > 2e-3):", {k: round(v, 3) for k, v in res["frac_over_2e-3"].items() if v > 0.05})
        bad = {k: v for k, v in res["max_abs_diff"].items() if v > 0.6}
        assert not bad, bad


@pytest.mark.timeout(280)
@pytest.mark.parametrize("precision,optimizer", [("fp32", "AdagradOptimizer"), ("bf16", "AdagradOptimizer"), ("fp32", "AdamOptimizer")])
def test_row_sharded_tables_equal_single_gpu(tmp_path, precision, optimizer):
    """BASELINE.json configs[3] "row-sharded data-parallel": rows move by all-to-all, tables are never replicated."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    out = tmp_path / "shard.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29543", os.path.join(HERE, "shard_worker.py"), str(out), precision, optimizer]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=260)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.load(open(out))
    assert res["local_rows"] == 351                     # 701 rows over two ranks: 351 + 350
    assert res["untouched_rows_bit_identical"] and res["init_identical"]
    assert res["resume_identical"] and res["shard_rows_in_state"] == 351      # per-rank checkpoint: save -> new handle -> same next step
    tol = 1e-4 if precision == "fp32" else 2e-2
    for a, b in zip(res["losses"], res["ref_losses"]):
        assert abs(a - b) <= tol * max(1.0, abs(b)), (res["losses"], res["ref_losses"])
    assert res["pred_max_abs_diff"] <= (2e-2 if precision == "fp32" else 0.2) * max(1.0, res["pred_scale"]), res
    _check_weight_drift(res, precision)
