"""bench.py --impl reference runs entirely on the host (the oracle port on a bounded sample): check the JSON line
the driver parses -- keys, units, the cpu_baseline / e2e objects of the tier's reference arm -- on this CPU box."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-sample", "32", "--cpu-workers", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mnist_cnn_train_samples_per_s" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 2 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout, r.stderr[-500:])


def test_roofline_numerators_of_the_headline_shapes():
    """bench.alg_work: the algorithmic bytes / flops that `roofline.achieved` is computed from, against
    hand-computed values for the step's dominant kernels and the ConvVjpBench shape."""
    sys.path.insert(0, ROOT)
    import bench
    # conv2 dKrn of the CNN step: A [4096,16,14,14], pooled cotangent [4096,16,7,7], 5x5 taps
    by, fl = bench.alg_work("conv_relu_pool_bwd{f64[4096,16,14,14];[4096,16,7,7]}", 5, 5)
    assert fl == 2 * 4096 * 16 * 14 * 14 * 16 * 25 == 10276044800
    assert by == (4096 * 16 * 14 * 14 * 2 + 16 * 16 * 25) * 8 == 205572096
    # conv2 forward + bias / relu / pool: K [16,16,5,5], A [4096,16,14,14] -> pooled [4096,16,7,7] (+ 1 code byte)
    by, fl = bench.alg_work("conv_relu_pool_fwd{f64[16,16,5,5];[4096,16,14,14]}", 5, 5)
    assert fl == 10276044800
    assert by == (16 * 16 * 25 + 4096 * 16 * 14 * 14) * 8 + 4096 * 16 * 7 * 7 * 9
    # ConvVjpBench (BASELINE config 3): 205.8 MB and 14.80 GFLOP per pass
    by, fl = bench.alg_work("conv2d_preserving{f64[64,64,3,3];[256,64,28,28]}", 3, 3)
    assert fl == 14797504512 and by == 205815808
    by, fl = bench.alg_work("matmul2{f64[64,784];[784,4096]}", 5, 5)
    assert fl == 2 * 64 * 784 * 4096 and by == (64 * 784 + 784 * 4096 + 64 * 4096) * 8

