"""Data-parallel correctness on real GPUs (needs >= 2 on the box; skipped otherwise): see tests/dp_worker.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_dp_two_ranks_gradient_sum_and_parameter_sync(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ, YB_DP_OUT=str(tmp_path / "dp"), MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(ROOT, "tests", "dp_worker.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    for r in range(2):
        res = json.load(open(f"{tmp_path / 'dp'}.rank{r}.json"))
        print(res)
        assert res["p2p_mode_used"] == "p2p", res["p2p_error"]     # the fused peer-memory kernel really ran
        assert res["p2p_sum_bit_exact"], res
        assert res["nccl_sum_bit_exact"], res
        assert res["params_identical_across_ranks"] and res["params_changed"], res
        assert all(l == l and abs(l) < 1e6 for l in res["losses"]), res
