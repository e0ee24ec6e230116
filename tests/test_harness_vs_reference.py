"""CPU, dev container only (skipped where /root/reference is absent): the harness' caller-side graph and loss
compute the same function as the reference's DetectionModel + v8DetectionLoss on identical weights."""
import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")


def test_graph_and_loss_match_reference():
    from improving_yolov8_cbam_swinblock_b200.harness import graph, loss as hl, synthetic
    from oracle import modules as om

    ref_loader.import_ultralytics()
    from ultralytics.cfg import get_cfg
    from ultralytics.nn.tasks import DetectionModel, yaml_model_load
    from ultralytics.utils.loss import v8DetectionLoss

    d = yaml_model_load(ref_loader.REF_ROOT + "/ultralytics/cfg/models/v8/yolov8n.yaml")
    d["backbone"][7][3] = [128]
    d["head"][3][3] = [128]
    torch.manual_seed(1)
    ref = DetectionModel(d, ch=3, nc=80, verbose=False)
    mine = graph.DetectionGraph({"CBAM": om.CBAM, "SwinBlock": om.SwinBlock, "SPPF": om.make_sppf(graph.Conv)}, "n", 80)
    assert list(ref.state_dict()) == list(mine.state_dict())
    mine.load_state_dict(ref.state_dict())
    ref.args = get_cfg()
    batch = synthetic.make_batch(2, 256, 80, seed=5)
    img = batch["img"].float() / 255
    ref.train(), mine.train()
    pr, pm = ref(img), mine(img)
    for a, b in zip(pr, pm):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-4)
    lr, ir = v8DetectionLoss(ref)(pr, batch)
    lm, im = hl.DetectionLoss(80, mine.stride)(pm, batch, max_boxes=8)
    torch.testing.assert_close(im, ir, rtol=1e-4, atol=1e-5)
    lr.sum().backward(), lm.sum().backward()
    gmax = max(float(p.grad.abs().max()) for p in ref.parameters() if p.grad is not None)
    gm = dict(mine.named_parameters())
    for k, p in ref.named_parameters():
        if p.grad is not None:
            assert float((p.grad - gm[k].grad).abs().max()) < 1e-4 * gmax, k


def test_plugin_rebinds_reference_namespaces():
    import improving_yolov8_cbam_swinblock_b200.ultralytics_plugin as plugin

    ref_loader.import_ultralytics()
    from ultralytics.nn import tasks
    from ultralytics.nn.tasks import DetectionModel, yaml_model_load

    table = plugin.install()
    try:
        assert tasks.CBAM is table["CBAM"] and tasks.SwinBlock is table["SwinBlock"] and tasks.SPPF is table["SPPF"]
        d = yaml_model_load(ref_loader.REF_ROOT + "/ultralytics/cfg/models/v8/yolov8s.yaml")  # the yaml as shipped
        m = DetectionModel(d, ch=3, nc=1, verbose=False)  # CPU stride pass runs through our shape probes
        assert type(m.model[7]) is table["SwinBlock"] and type(m.model[10]) is table["CBAM"]
        assert type(m.model[11]) is table["SPPF"] and m.model[12].k == 7
        assert sum(p.numel() for p in m.parameters()) == 13405269  # SURVEY D4 probe
        m.fuse()  # SPPF.cv1/cv2 are the reference's Conv: BN folded
        assert not hasattr(m.model[11].cv1, "bn")
    finally:
        plugin.uninstall()
    assert tasks.CBAM is not table["CBAM"]


def test_plugin_seams_patch_is_value_neutral_and_reversible():
    import torch

    import improving_yolov8_cbam_swinblock_b200.ultralytics_plugin as plugin

    ref_loader.import_ultralytics()
    from ultralytics.nn.modules.block import C2f
    from ultralytics.nn.modules.conv import Concat

    torch.manual_seed(0)
    blk = C2f(16, 16, n=2, shortcut=True).eval()
    x = torch.randn(2, 16, 8, 8)
    ref = blk(x)
    ref_cat = Concat(1)([x, x * 2])
    orig = C2f.forward
    plugin.install(seams=True)
    try:
        assert C2f.forward is not orig
        assert torch.equal(blk(x), ref)                      # CPU tensors: the stock ops behind the same call sites
        assert torch.equal(Concat(1)([x, x * 2]), ref_cat)
        assert torch.equal(Concat(2)([x, x]), torch.cat([x, x], 2))
    finally:
        plugin.uninstall()
    assert C2f.forward is orig


def test_reference_pickles_unpickle_as_plugin_classes_with_reference_state_only():
    """SPPF.forward / SwinBlock.forward must not read attributes the reference classes never store (k,
    num_heads): a module pickled by the reference and loaded after install() has run no __init__ of ours."""
    import pickle

    import improving_yolov8_cbam_swinblock_b200.ultralytics_plugin as plugin

    ref_loader.import_ultralytics()
    from ultralytics.nn.modules.block import SPPF as RSPPF
    from ultralytics.nn.modules.swin_block import SwinBlock as RSwin

    blobs = [pickle.dumps(RSwin(32, 2, 7)), pickle.dumps(RSPPF(32, 32, 7))]
    table = plugin.install()
    try:
        ms, mp_ = (pickle.loads(b) for b in blobs)
        assert type(ms) is table["SwinBlock"] and type(mp_) is table["SPPF"]
        assert ms.attn.num_heads == 2 and ms.window_size == 7 and mp_.m.kernel_size == 7
        x = torch.zeros(1, 32, 9, 9)   # shape probe path
        assert ms(x).shape == x.shape and mp_(x).shape == x.shape
    finally:
        plugin.uninstall()
