"""Host-side logic of the seam wrappers (functional.py: nhwc_concat / nhwc_chunk / nhwc_upsample_nearest), CPU only: the stride
classifier that decides whether a tensor can be handed to b200_nhwc_concat in place, and the stock-op routes taken by tensors the
kernels do not tile (no compute calls into the library here)."""
import torch

from improving_yolov8_cbam_swinblock_b200 import functional as Fb


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def test_row_stride_classifier():
    dense = _cl(torch.zeros(2, 16, 5, 7))
    assert Fb._row_strided(dense) == 16
    a, b = dense.chunk(2, 1)                                   # channel slices of a channels_last map: row stride = parent C
    assert Fb._row_strided(a) == 16 and Fb._row_strided(b) == 16
    assert Fb._row_strided(dense[:, 3:11]) == 16
    assert Fb._row_strided(torch.zeros(2, 16, 5, 7)) is None   # NCHW-dense
    assert Fb._row_strided(dense[:, :, ::2]) is None           # spatially strided view
    assert Fb._row_strided(dense[:, ::2]) is None              # channel stride 2
    assert Fb._row_strided(torch.ones(1).expand(2, 16, 5, 7)) is None   # broadcast gradient (stride 0)
    one = _cl(torch.zeros(1, 8, 1, 1))
    assert Fb._row_strided(one) == 8


def test_cpu_tensors_take_the_stock_ops():
    torch.manual_seed(0)
    x = torch.randn(2, 8, 4, 6, requires_grad=True)
    y0, y1 = Fb.nhwc_chunk(x, 2)
    z = Fb.nhwc_concat([y0, y1, y1 * 2])
    up = Fb.nhwc_upsample_nearest(z, 2, 2)
    ref = torch.nn.functional.interpolate(torch.cat([x[:, :4], x[:, 4:], x[:, 4:] * 2], 1), scale_factor=2.0, mode="nearest")
    assert torch.equal(up, ref)
    up.sum().backward()
    assert x.grad is not None and torch.isfinite(x.grad).all()


def test_concat_eligibility():
    a = torch.zeros(1, 8, 2, 2)
    assert not Fb._concat_ok([a])                              # CPU tensor
    assert not Fb._concat_ok([a.to(torch.int32)])


def test_fork_placement_in_the_graph_matches_plain_autograd():
    """The fork that routes a multi-consumer map's gradients through the fan-in kernel (functional.nhwc_fork) is placed by host logic
    in harness/graph.py: inside C2f (bottleneck outputs, the second chunk half), inside shortcut bottlenecks, and at the saved layers
    of `_predict_once` (consumer counts from the yaml `from` column).  With a pure-torch stand-in for the fork (n aliases whose
    gradients are summed) the graph's outputs and every parameter gradient must equal those of the fork-less graph, and the fork
    must be asked for exactly the consumer counts of the yaml."""
    import copy

    from improving_yolov8_cbam_swinblock_b200.harness import graph
    from oracle import modules as om

    calls = []

    class _Fork(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, n):
            return tuple(x.view_as(x) for _ in range(n))

        @staticmethod
        def backward(ctx, *gs):
            out = None
            for g in gs:
                if g is not None:
                    out = g if out is None else out + g
            return out, None

    def fork(x, n):
        calls.append(n)
        return _Fork.apply(x, n) if x.requires_grad else (x,) * n

    torch.manual_seed(0)
    blocks = {"CBAM": om.CBAM, "SwinBlock": om.SwinBlock, "SPPF": om.make_sppf(graph.Conv)}   # CPU blocks (test infrastructure)
    plain = graph.DetectionGraph(dict(blocks), "n", 8)
    forked = copy.deepcopy(plain)
    for m in forked.modules():
        if isinstance(m, (graph.C2f, graph.Bottleneck)):
            m.fork = fork
    forked._fork = fork
    # yaml `from` column: layers 4, 7, 10, 16, 19, 22 feed the next row AND a later Concat / Detect
    assert [i for i, n in enumerate(plain._consumers) if n == 2] == [4, 7, 10, 16, 19, 22]
    assert plain._consumers[-1] == 0 and max(plain._consumers) == 2
    x = torch.randn(2, 3, 64, 64)
    outs_p = plain(x)
    outs_f = forked(x)
    assert calls and all(n == 2 for n in calls)
    for a, b in zip(outs_p, outs_f):
        assert torch.equal(a, b)
    sum(o.square().mean() for o in outs_p).backward()
    sum(o.square().mean() for o in outs_f).backward()
    for (k, p), q in zip(plain.named_parameters(), forked.parameters()):
        if p.grad is None:
            assert q.grad is None, k
        else:
            torch.testing.assert_close(q.grad, p.grad, rtol=1e-4, atol=1e-6, msg=k)
