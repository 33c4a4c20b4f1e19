"""world_size-2 gloo worker for tests/test_dist_cpu.py: the multi-GPU plumbing of bench.py without GPUs."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from blama_b200 import gguf_synth  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
d = bench.Dist(rank, world, int(os.environ.get("LOCAL_RANK", 0)), backend="gloo")
path = bench.ensure_model("tiny-llama-q8", rank, d.barrier)            # rank 0 writes, everyone waits
ok_file = os.path.exists(path) and os.path.getsize(path) >= gguf_synth.model_bytes(gguf_synth.SHAPES["tiny-llama-q8"])
# rank r "processes" 100*(r+1) tokens in (r+1) seconds
value = bench.aggregate_throughput(d, 100.0 * (rank + 1), 1.0 * (rank + 1))
prompt = gguf_synth.synth_prompt("tiny-llama-q8", 16, bench.request_seed(rank, 0)).tolist()
out = {"rank": rank, "value": value, "ok_file": ok_file, "prompt": prompt, "max": d.max(float(rank)), "sum": d.sum(1.0)}
with open(os.path.join(sys.argv[1], f"rank{rank}.json"), "w") as f:
    json.dump(out, f)
d.close()
