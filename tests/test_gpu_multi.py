"""Multi-GPU tier (needs >= 2 GPUs on the box, else skipped): trainer.fit with rank-sharded leaves and the fused
peer-memory all-reduce + Adam, rank-sharded encode / decode.  The checks live in scripts/dist_check.py (run under
torchrun): bit-identical weights on every rank, falling loss, sharded reconstruction == single-process one."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
           "127.0.0.1", "--master-port", "29571", os.path.join(ROOT, "scripts", "dist_check.py")]
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)


@pytest.mark.gpu
@pytest.mark.parametrize("peer", ["1", "0"])
def test_two_rank_fit_and_codec(gpu, peer):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = _run(2, {"NVF_PEER_ALLREDUCE": peer})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "dist check ok: world 2" in r.stdout
    if peer == "1":
        assert "peer all-reduce: on" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
