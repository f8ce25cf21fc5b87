"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm (the oracle timed on the
host cores) prints ONE JSON line with the agreed keys, and every rank but 0 stays silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--cpu-batch", "2"]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["metric"] == "CoR2 train samples/sec (fwd+bwd)" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["steps"] == 1 and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_kernel_roofline_record():
    """bench.kernel_roofline: the dominant GEMM launch (by average duration) with algorithmic flops 2*M*N*K*groups,
    the measured peak as denominator and the committed DRAM traffic of that launch."""
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    per_kernel = {"k:tc_linear_bwd.dgrad M9216 N2048 K310 g1 s1": 0.150, "k:tc_linear_fwd M256 N310 K2400 g4 s9": 0.060}
    per_op = {"compress_v2.bwd": 0.25, "q_proj4.fwd": 0.12}
    pk = {"hbm": 6452.8, "tensor": 1414.3, "tensor_burst": 1637.0, "src": "test"}
    r = bench.kernel_roofline(per_kernel, per_op, 2.0, pk, argparse.Namespace(precision="tf32x3", model="CoR2"), 256, 36, 2000)
    assert r["kernel"] == "tc_linear_bwd.dgrad M9216 N2048 K310 g1 s1" and r["bound"] == "tensor" and r["unit"] == "TFLOP/s"
    flops = 2.0 * 9216 * 2048 * 310
    assert abs(r["achieved"] - flops / 0.150e-3 / 1e12) < 1e-6 and abs(r["frac"] - r["achieved"] / 1414.3) < 1e-12
    assert r["peak"] == 1414.3 and r["traffic"] == 45470000 and abs(r["share_of_step"] - 0.075) < 1e-12
    # a bandwidth-bound kernel dominates: its algorithmic bytes against the measured copy bandwidth
    r3 = bench.kernel_roofline({"k:pool_bwd hbm=75497472": 0.2, "k:tc_linear_fwd M256 N310 K2400 g4 s9": 0.06}, per_op, 2.0, pk,
                               argparse.Namespace(precision="tf32x3", model="CoR2"), 256, 36, 2000)
    assert r3["kernel"] == "pool_bwd" and r3["bound"] == "hbm" and r3["unit"] == "GB/s" and r3["peak"] == 6452.8
    assert abs(r3["achieved"] - 75497472 / 0.2e-3 / 1e9) < 1e-6
    # a CUDA-core kernel (ODA pairwise terms) dominates: quoted against the fp32 FMA rate
    r4 = bench.kernel_roofline({"k:oda_pair_bwd_train_e flop=2000000000": 0.3}, per_op, 2.0, pk,
                               argparse.Namespace(precision="tf32x3", model="ODA"), 256, 36, 3000)
    assert r4["kernel"] == "oda_pair_bwd_train_e" and r4["unit"] == "TFLOP/s" and 70 < r4["peak"] < 80
    # no per-kernel records (an op that is not a GEMM dominates): falls back to the per-op table
    r2 = bench.kernel_roofline({}, {"compound.bwd": 0.03}, 1.0, pk, argparse.Namespace(precision="tf32x3", model="CoR2"), 256, 36, 2000)
    assert r2["bound"] == "hbm" and r2["unit"] == "GB/s" and r2["traffic"] is None


def test_bucket_plan_properties():
    """plan_buckets: contiguous, covering, at most the requested number of buckets, never empty."""
    import random
    from vqa_playground_pytorch_b200.parallel import plan_buckets
    rng = random.Random(0)
    for _ in range(200):
        sizes = [rng.randint(1, 5000) for _ in range(rng.randint(1, 14))]
        nb = rng.randint(1, 6)
        tail = rng.randint(0, 3)
        cuts = plan_buckets(sizes, nb, tail=tail)
        assert 1 <= len(cuts) <= nb and cuts[0][0] == 0 and cuts[-1][1] == len(sizes)
        assert all(a < b for a, b in cuts) and all(cuts[i][1] == cuts[i + 1][0] for i in range(len(cuts) - 1))
        if tail and nb > 1 and len(sizes) > tail:
            assert cuts[-1] == (len(sizes) - tail, len(sizes))       # the late groups form the last bucket alone
