"""CPU: the N>1 path with world_size 2 over gloo: batch sharded across ranks, gradients all-reduced
(SUM of the local ``loss.sum()`` gradients = the reference's ``loss * world_size`` + DDP mean, trainer.py:278,386-388);
nothing else is exchanged.

What "2 ranks == 1 rank" means here: BatchNorm uses per-rank batch statistics (no SyncBN in the reference) and the v8 loss
normalises by the LOCAL target-score sum, so the 2-rank gradient is by construction the SUM of the gradients each shard
produces on its own.  The worker therefore runs each shard through an independent world_size-1 trainer (plain autograd
``.grad`` tensors, no flat buffer), ships those gradients to the other rank by ``all_gather_object`` (pickled tensors: a
different code path from the tensor all-reduce under test) and compares their sum with the flat buffer after the
exchange step."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

from util import ROOT


def shard_of(full, rank, world):
    per = full["img"].shape[0] // world
    sel = (full["batch_idx"] >= rank * per) & (full["batch_idx"] < (rank + 1) * per)
    return {"img": full["img"][rank * per:(rank + 1) * per], "batch_idx": full["batch_idx"][sel] - rank * per,
            "cls": full["cls"][sel], "bboxes": full["bboxes"][sel]}


def run_rank(rank, world, port, out, device, blocks_name, amp):
    """One rank of the check above (also used by tests/test_gpu_ddp.py with device='cuda:0' and the B200 blocks)."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from improving_yolov8_cbam_swinblock_b200.harness import graph, synthetic, train

    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if blocks_name == "oracle":
        from oracle import modules as om

        blocks = {"CBAM": om.CBAM, "SwinBlock": om.SwinBlock, "SPPF": om.make_sppf(graph.Conv)}
    else:
        import improving_yolov8_cbam_swinblock_b200 as P

        blocks = P.BLOCKS
    amp_dtype = torch.bfloat16 if amp else None
    tr = train.Trainer(blocks, "n", 4, device=device, amp_dtype=amp_dtype, world_size=world, local_rank=rank, ema=False)
    solo = train.Trainer(blocks, "n", 4, device=device, amp_dtype=amp_dtype, world_size=1, ema=False)
    solo.raw.load_state_dict(tr.raw.state_dict())
    tr.max_boxes = solo.max_boxes = 8
    full = synthetic.make_batch(4, 64, 4, seed=11)
    shard = tr.to_device(shard_of(full, rank, world))
    # --- reference: this shard alone through plain autograd (world_size 1: no flat buffer)
    solo._fwd_bwd(shard)
    mine = [p.grad.detach().float().cpu() for p in solo._params]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    want = [sum(g[i] for g in gathered) for i in range(len(mine))]
    # --- the path under test: grads are views of ONE flat fp32 buffer, one all-reduce(SUM)
    assert tr._flat is not None and all(p.grad.data_ptr() >= tr._flat.data_ptr() for p in tr._params)
    tr._fwd_bwd(shard)
    local_flat = tr._flat.detach().cpu().clone()
    tr._exchange()
    got = [p.grad.detach().float().cpu() for p in tr._params]
    scale = max(float(w.abs().max()) for w in want)
    worst = max(float((g - w).abs().max()) for g, w in zip(got, want))
    tol = (2e-3 if amp else 1e-5) * scale
    assert worst <= tol, f"rank {rank}: all-reduced flat gradients differ from the sum of the shard gradients: {worst:.3e} > {tol:.3e}"
    # before the exchange the flat buffer held exactly this rank's own gradient
    off = 0
    for p, m in zip(tr._params, mine):
        loc = local_flat[off:off + p.numel()].view(p.shape) if p.is_contiguous() else None
        off += p.numel()
        if loc is not None:
            assert float((loc - m).abs().max()) <= tol
    assert float((want[0] - mine[0]).abs().max()) > 0, "the shards must contribute different gradients"
    tr._update()
    sd = {k: v.detach().cpu().clone() for k, v in tr.raw.state_dict().items()}
    if rank == 0:
        torch.save(sd, out)
    # parameters stay identical across replicas (BN running stats are per-GPU: no SyncBN in the reference)
    sums = [None] * world
    dist.all_gather_object(sums, float(sum(p.detach().double().abs().sum() for p in tr.raw.parameters())))
    assert abs(sums[0] - sums[1]) < 1e-9 * max(1.0, abs(sums[0])), "replicas diverged after one step"
    dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_two_rank_flat_gradients_equal_sum_of_shard_gradients(tmp_path):
    port = 29500 + os.getpid() % 2000
    out = str(tmp_path / "sd.pt")
    mp.spawn(run_rank, args=(2, port, out, "cpu", "oracle", False), nprocs=2, join=True)
    assert os.path.isfile(out)
    sd = torch.load(out)
    assert all(torch.isfinite(v).all() for v in sd.values() if v.dtype.is_floating_point)
