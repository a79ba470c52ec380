"""Row f3 on one GPU: the data-parallel loop of gsvc_b200.dp_train with world = 1 (the 2-GPU run is
examples/dp_train.py under torchrun, recorded in profiles/): loss goes down, anchors are grown and pruned through the
reference-shaped adjust_anchor, and the statistic the loop accumulates equals the one an emulated 2-rank split
accumulates."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dp_loop_trains_and_densifies(cuda_device):
    run = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "dp_train.py"), "--iters", "120", "--interval", "5"],
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr[-2000:]
    out = json.loads(run.stdout.strip().splitlines()[-1])
    assert out["ranks_agree_every_round"] and out["densification_rounds"] == 24
    assert out["loss_last10"] < 0.8 * out["loss_first10"], out
    assert out["added"] > 0 and out["anchors_end"] != out["anchors_start"], out
