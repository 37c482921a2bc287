"""The N>1 host logic of bench.py on CPU: two processes, gloo backend, rendezvous on 127.0.0.1.
Checks the shard assignment (each rank its own block range / seed), the MAX-over-ranks timing
reduction, the SUM-over-ranks unit count, and that under torchrun only rank 0 of the reference
arm prints a line."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))

WORKER = r"""
import json, os, sys
sys.path.insert(0, %r)
sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
import bench
r = bench.Ranks(backend="gloo")
x = bench.make_pcm(2, 16, 44100, 4, 4096, seed=r.shard_seed())
t = r.max(10.0 + r.rank)            # slowest rank decides the step time
units = r.sum(float(x.shape[0] * x.shape[1]))
r.barrier()
open(os.path.join(sys.argv[1], "rank%%d.json" %% r.rank), "w").write(
    json.dumps({"rank": r.rank, "world": r.world, "max": t, "sum": units, "digest": int(np.abs(x).sum())}))
r.close()
""" % (ROOT, ROOT)


def _torchrun(args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + args
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)


def test_two_rank_sharding_and_reductions(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    p = _torchrun([str(w), str(tmp_path)], 29541)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [json.loads((tmp_path / f"rank{r}.json").read_text()) for r in (0, 1)]
    assert sorted(l["rank"] for l in lines) == [0, 1]
    assert all(l["world"] == 2 for l in lines)
    assert all(l["max"] == 11.0 for l in lines)                 # MAX over ranks
    assert all(l["sum"] == 2 * 4 * 4096 * 2 for l in lines)     # whole-job units
    assert lines[0]["digest"] != lines[1]["digest"]             # ranks work on different block ranges


def test_reference_arm_only_rank0_prints():
    import reflib
    if not reflib.available("default"):
        import pytest
        pytest.skip("oracle/_ref not built")
    p = _torchrun(["bench.py", "--gpus", "2", "--impl", "reference", "--steps", "1", "--warmup", "1", "--blocks", "200"], 29542)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2
    assert lines[0]["e2e"]["h2d_bytes_per_step"] == 0 and lines[0]["cpu_baseline"]["kind"] == "reference"
