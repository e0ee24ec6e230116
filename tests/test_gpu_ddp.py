"""-m gpu: the multi-rank training path on the GPU.  Two processes share cuda:0 (the test box has one
GPU; NCCL refuses two ranks on one device, so the exchange runs over gloo's CUDA all-reduce) and train with the B200
blocks under bf16 autocast: the flat fp32 gradient buffer (``harness/train.py``: every ``p.grad`` a view of it) after the
one all-reduce equals the sum of the gradients each shard produces alone -- see tests/test_ddp_gloo.py for why that is
the identity the reference's DDP step satisfies."""
import os

import pytest
import torch
import torch.multiprocessing as mp

from test_ddp_gloo import run_rank

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(900)
def test_two_ranks_on_gpu_flat_buffer_exchange(tmp_path):
    port = 31500 + os.getpid() % 2000
    out = str(tmp_path / "sd.pt")
    mp.spawn(run_rank, args=(2, port, out, "cuda:0", "b200", True), nprocs=2, join=True)
    sd = torch.load(out)
    assert all(torch.isfinite(v).all() for v in sd.values() if v.dtype.is_floating_point)
