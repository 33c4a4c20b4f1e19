"""N > 1 path on CPU: two gloo ranks exercise bench.py's request partitioning and max-over-ranks aggregation
(the data path has no collective: one model replica per GPU)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_gloo_aggregation(tmp_path):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "tests", "dist_worker.py"), str(tmp_path)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    subprocess.run(cmd, check=True, timeout=240, env=env, capture_output=True)
    r = [json.load(open(tmp_path / f"rank{i}.json")) for i in range(2)]
    # 100 + 200 tokens, slowest rank took 2 s -> 150 tok/s on every rank
    assert r[0]["value"] == r[1]["value"] == 150.0
    assert r[0]["max"] == 1.0 and r[0]["sum"] == 2.0
    assert r[0]["ok_file"] and r[1]["ok_file"]
    assert r[0]["prompt"] != r[1]["prompt"]          # independent requests per replica


def test_bench_reference_arm_line_shape(tmp_path):
    """--impl reference prints one JSON line with the contract keys (tiny model so the CPU suite stays fast)"""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--shape", "tiny-llama-q8", "--steps", "1",
                          "--warmup", "0"], check=True, timeout=240, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    d = json.loads(out)
    for k in ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"]:
        assert k in d, k
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
