"""-m gpu: the whole custom YOLOv8-CBAM-Swin graph with the B200 blocks vs the same graph with the oracle blocks
(identical weights, identical synthetic batch): logits, the three loss components and every parameter gradient."""
import copy

import pytest
import torch

from util import rel_err

pytestmark = pytest.mark.gpu


def _build(blocks, scale, nc, seed):
    from improving_yolov8_cbam_swinblock_b200.harness import graph

    torch.manual_seed(seed)
    return graph.DetectionGraph(blocks, scale, nc)


@pytest.mark.parametrize("scale", ["n", "s"])
def test_whole_model_fp32_vs_oracle_blocks(scale):
    import improving_yolov8_cbam_swinblock_b200 as P
    from improving_yolov8_cbam_swinblock_b200.harness import graph, loss as hl, synthetic
    from oracle import modules as om

    torch.backends.cudnn.allow_tf32 = False  # stock convs would otherwise run TF32 and drown the comparison
    torch.backends.cuda.matmul.allow_tf32 = False
    ob = {"CBAM": om.CBAM, "SwinBlock": om.SwinBlock, "SPPF": om.make_sppf(graph.Conv)}
    ref = _build(ob, scale, 80, 1)
    mine = _build(P.BLOCKS, scale, 80, 2)
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    mine.load_state_dict(ref.state_dict())
    ref, mine = ref.cuda().train(), mine.cuda().train()
    batch = synthetic.make_batch(4, 320, 80, seed=7)
    dev_batch = {k: v.cuda() for k, v in batch.items()}
    img = dev_batch["img"].float() / 255
    crit = hl.DetectionLoss(80, ref.stride)
    fr, fm = ref(img), mine(img.contiguous(memory_format=torch.channels_last))
    for a, b in zip(fm, fr):
        assert rel_err(a, b) < 1e-4
    lr, ir = crit(fr, dev_batch, max_boxes=8)
    lm, im = crit(fm, dev_batch, max_boxes=8)
    torch.testing.assert_close(im, ir, rtol=1e-4, atol=1e-5)
    lr.sum().backward()
    lm.sum().backward()
    gr, gm = dict(ref.named_parameters()), dict(mine.named_parameters())
    gmax = max(float(p.grad.abs().max()) for p in gr.values() if p.grad is not None)
    for k, p in gr.items():
        if p.grad is None:
            assert gm[k].grad is None
            continue
        err = float((gm[k].grad - p.grad).abs().max())
        assert err <= 2e-4 * gmax, f"{k}: {err:.3e} (gmax {gmax:.3e})"


def test_module_protocols_on_gpu():
    """deepcopy / pickle / half / DDP-wrap behaviours the trainer relies on."""
    import io

    import improving_yolov8_cbam_swinblock_b200 as P
    from improving_yolov8_cbam_swinblock_b200.harness import graph

    m = _build(P.BLOCKS, "n", 3, 0).cuda().eval()
    x = torch.rand(1, 3, 64, 64, device="cuda")
    y = m(x)[0]
    m2 = copy.deepcopy(m)
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m3 = torch.load(buf, weights_only=False)
    assert torch.equal(m2(x)[0], y) and torch.equal(m3(x)[0], y)
    yh = m2.half()(x.half())[0]
    assert yh.dtype == torch.float16 and rel_err(yh, y) < 5e-2


@pytest.mark.parametrize("scale", ["s", "m"])
def test_scaled_variants_train_step_bf16(scale):
    """BASELINE configs[3]: the s / m width-depth variants (SwinBlock [256] / [384]; m has head dim 192, served by the
    SIMT attention kernels) take a finite bf16 training step with the B200 blocks and the fused Conv epilogue."""
    import improving_yolov8_cbam_swinblock_b200 as P
    from improving_yolov8_cbam_swinblock_b200.harness import synthetic, train

    tr = train.Trainer(P.BLOCKS, scale, 80, device="cuda:0", amp_dtype=torch.bfloat16)
    tr.max_boxes = synthetic.BOXES_PER_IMAGE
    batch = tr.to_device(synthetic.make_batch(2, 320, 80, seed=5))
    before = [p.detach().clone() for p in tr.raw.parameters()]
    items = tr.step(batch)
    assert torch.isfinite(items).all()
    moved = sum(int(not torch.equal(a, b)) for a, b in zip(before, tr.raw.parameters()))
    assert moved > 0.9 * len(before)
