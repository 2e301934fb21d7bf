"""Host-side pieces of bench.py that can be checked without a GPU: the work figures the roofline is computed
from (BASELINE.md section 4 / DESIGN.md section 3) and the committed ncu traffic table."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _criteo():
    return bench.workload_spec("criteo", 0)


def test_workload_is_the_one_baseline_json_names():
    spec = _criteo()
    assert (spec["F"], spec["K"], spec["B"]) == (39, 32, 8192)
    assert spec["M"] == 10_000_000
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "samples" in base["metric"].lower()


def test_algorithmic_flops_of_the_conv_layers():
    spec = _criteo()
    P = 39 * 38 // 2
    for l, ho in enumerate((16, 8, 4, 2)):
        for kind in ("fwd", "dgrad", "wgrad"):
            k, amount = bench.algorithmic_work(spec, "conv_%s_l%d" % (kind, l), 8192, 1)
            assert k == "flops" and amount == 2.0 * 8192 * ho * ho * 4 * P * P   # 8 P^2 per output position
    assert bench.algorithmic_work(spec, "gather_outer", 8192, 1) == ("bytes", 8192 * 39 * (4 + 2 * 4 * 32))


def test_executed_flops_of_the_factorised_layer0_kernels():
    spec = _criteo()
    fwd = bench.executed_flops(spec, "conv_fwd_l0", 8192, "bf16")
    dgr = bench.executed_flops(spec, "conv_dgrad_l0", 8192, "bf16")
    wgr = bench.executed_flops(spec, "conv_wgrad_l0", 8192, "bf16")
    tiles, q16, ka = 1024, 752, 80
    assert fwd == tiles * q16 * 2.0 * 128 * ka * (ka + 128)
    assert dgr == 2 * fwd
    assert wgr == tiles * q16 * 2.0 * 128 * 128 * (16 + ka)
    algo = bench.algorithmic_work(spec, "conv_fwd_l0", 8192, 1)[1]
    assert fwd < algo / 2.5 and dgr < algo and wgr < algo / 3     # fewer FLOPs than the direct form
    # direct form: other layers, fp32 mode, more than 40 fields
    assert bench.executed_flops(spec, "conv_fwd_l1", 8192, "bf16") is None
    assert bench.executed_flops(spec, "conv_fwd_l0", 8192, "fp32") is None
    # split bf16: three MMAs per product, direct form on every layer
    assert bench.executed_flops(spec, "conv_fwd_l0", 8192, "bf16x3") == 3 * fwd        # factorised, Z split
    assert bench.executed_flops(spec, "conv_dgrad_l0", 8192, "bf16x3") == (3 * dgr if "conv_dgrad_l0" in bench.SPLIT_FACT else 3 * algo)
    assert bench.executed_flops(spec, "conv_dgrad_l2", 8192, "bf16x3") == 3 * bench.algorithmic_work(spec, "conv_dgrad_l2", 8192, 1)[1]
    assert bench.executed_flops(spec, "gather_outer", 8192, "bf16x3") is None
    assert bench.executed_flops(dict(spec, F=44), "conv_fwd_l0", 8192, "bf16") is None


def test_both_arms_print_the_same_config():
    spec = _criteo()
    a = bench.shared_config(spec, 1)
    assert a["batch_per_gpu"] == 8192 and a["global_batch"] == 8192 and "l2" in a and a["parallelism"] == "dp1"
    assert a == bench.shared_config(bench.workload_spec("criteo", 0), 1)
    assert bench.shared_config(spec, 8)["global_batch"] == 65536
    flush, text = bench.l2_policy(bench.workload_spec("frappe", 0), 256)
    assert flush and "flushed" in text
    assert not bench.l2_policy(spec, 8192)[0]


def test_cpu_micro_batch_is_bounded():
    b = bench.cpu_batch_for(_criteo())
    assert 8 <= b <= 128 and b & (b - 1) == 0
    assert bench.cpu_batch_for(bench.workload_spec("frappe", 0)) <= 256


def test_ncu_traffic_table_matches_the_bench_workload():
    spec = _criteo()
    algo_bytes = 8192 * 256 * 768 * 2      # one pass over X1 / dY0 (bf16, padded channels)
    for tag in ("conv_fwd_l0", "conv_dgrad_l0", "conv_wgrad_l0"):
        t = bench.ncu_traffic(spec, tag, 8192, "bf16")
        assert isinstance(t, int) and algo_bytes <= t < 1.25 * algo_bytes, (tag, t)
    assert bench.ncu_traffic(spec, "conv_fwd_l0", 4096, "bf16") is None      # other batch: no capture
    assert bench.ncu_traffic(spec, "conv_fwd_l0", 8192, "fp32") is None
    assert bench.ncu_traffic(spec, "conv_fwd_l1", 8192, "bf16") is None
