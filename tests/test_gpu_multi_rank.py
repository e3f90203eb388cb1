"""Multi-rank parity on hardware (needs >= 2 GPUs; skipped on a one-GPU box): the proof-sharded job under NCCL gives the
oracle's fold-of-folds accumulator, root challenge and verdict, also with a corrupted proof on a non-zero rank
(pcs/kzg/accumulation.rs:29-62, decider.rs:60-81; snark_verifier_axiom_b200/distributed.py)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_job_matches_oracle_fold_of_folds():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs (NCCL refuses two ranks on one device)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(ROOT, "tests", "workers", "multi_rank_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    assert res["world"] == world
    assert res["valid"] == {"accumulator_equal": True, "root_challenge_equal": True, "verdict": True, "oracle_verdict": True}
    assert res["corrupted"] == {"accumulator_equal": True, "root_challenge_equal": True, "verdict": False, "oracle_verdict": False}
