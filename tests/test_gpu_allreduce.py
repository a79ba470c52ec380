"""The exchange step of the frame-sharded loop: gsvc_b200.sharding.SwitchAllReduce, the library's own
one-kernel all-reduce over NVLink / NVSwitch, on two GPUs of the node (skipped on a single-GPU box; the 8-GPU run of
the same script is recorded in profiles/)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_switch_allreduce_two_ranks(cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    run = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(ROOT, "scripts", "check_switch_allreduce.py")],
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, (run.stdout[-2000:], run.stderr[-2000:])
    out = json.loads([l for l in run.stdout.strip().splitlines() if l.startswith("{")][-1])
    assert out["ok"], out
    ran = [c for c in out["cases"] if "skipped" not in c]
    assert any(c["mode"] == "peer" for c in ran), out          # peer loads need no multicast: must have run
    assert all(c["ranks_bit_identical"] for c in ran), out


def test_switch_allreduce_rejects_bad_arguments(cuda_device):
    from gsvc_b200 import _lib
    lib = _lib.lib()
    buf = torch.zeros(16, device=cuda_device)
    st = torch.zeros(_lib.EXCHANGE_STATE_WORDS, dtype=torch.int32, device=cuda_device)
    p = buf.data_ptr()
    assert lib.gsvc_rast_switch_allreduce(None, None, p, st.data_ptr(), 0, 2, 16, 1, None) < 0     # no buffers at all
    assert lib.gsvc_rast_switch_allreduce(None, p, p, st.data_ptr(), 2, 2, 16, 1, None) < 0        # rank outside world
    assert lib.gsvc_rast_switch_allreduce(None, p, p, st.data_ptr(), 0, 2, 18, 1, None) < 0        # numel % 4
    assert lib.gsvc_rast_switch_allreduce(None, p, p, st.data_ptr(), 0, 3, 16, 1, None) < 0        # peer path: 3 ranks
    assert lib.gsvc_rast_switch_allreduce(None, p, p, None, 0, 2, 16, 1, None) < 0                 # no state words
    assert b"" != lib.gsvc_rast_last_error()


def test_backward_that_carries_the_exchange_two_ranks(cuda_device):
    """gsvc_rast_backward_views_exchange: every rank's batched-view backward sums the packed [P,14] rows over the ranks
    in its own launches; equal to backward + NCCL all-reduce, ranks bit-identical, replayable from a CUDA graph."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    run = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29534",
                          os.path.join(ROOT, "scripts", "check_fused_exchange.py"), "2"],
                         capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, (run.stdout[-2000:], run.stderr[-2000:])
    out = json.loads([l for l in run.stdout.strip().splitlines() if l.startswith("{")][-1])
    assert out["ok"] and out["ranks_bit_identical"] and out["fused_launches"] >= 3, out
    assert out["max_err_vs_nccl_rel_to_column_max"] <= 2e-5 and out["graph_replay_err"] <= 2e-5, out


def test_exchange_backward_rejects_bad_arguments(cuda_device):
    import ctypes
    from gsvc_b200 import _lib
    lib = _lib.lib()
    buf = torch.zeros(64, device=cuda_device)
    p = buf.data_ptr()
    ok = _lib.Exchange(None, p, p, p, 0, 2, 0, 0)
    args = [None, 1, None, 1, 100, 0, 1] + [None] * 11 + [0, None, None]       # up to dL_dmeans2D
    assert lib.gsvc_rast_backward_views_exchange(*args, p, None, None) < 0                       # exchange is NULL
    assert lib.gsvc_rast_backward_views_exchange(*args, None, ctypes.byref(ok), None) < 0        # no packed rows
    odd = list(args); odd[4] = 101
    assert lib.gsvc_rast_backward_views_exchange(*odd, p, ctypes.byref(ok), None) < 0            # odd P
    bad = _lib.Exchange(None, p, p, p, 2, 2, 0, 0)
    assert lib.gsvc_rast_backward_views_exchange(*args, p, ctypes.byref(bad), None) < 0          # rank outside world
    bad = _lib.Exchange(None, None, p, p, 0, 2, 0, 0)
    assert lib.gsvc_rast_backward_views_exchange(*args, p, ctypes.byref(bad), None) < 0          # neither mapping
    bad = _lib.Exchange(None, p, p, p, 0, 2, 0, 100)
    assert lib.gsvc_rast_backward_views_exchange(*args, p, ctypes.byref(bad), None) < 0          # chunk_rows % 128
