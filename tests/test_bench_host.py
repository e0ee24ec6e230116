"""Host-side pieces of bench.py that need no GPU: the bounded child-process form of the host-core baseline leg (what rank 0 runs at
world_size > 1) and the roofline entries of the shape-tagged launches added in round 2."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def test_bounded_cpu_baseline_gives_up_at_its_limit_instead_of_stalling_the_record():
    import bench

    t0 = time.time()
    out = bench.cpu_baseline_bounded(timeout_s=2)
    assert time.time() - t0 < 30
    assert set(out) == {"skipped"} and "2 s limit" in out["skipped"]


def test_roofline_entries_of_the_round_2_launch_tags():
    from improving_yolov8_cbam_swinblock_b200.harness import sweep

    n = 64 * 320 * 320 * 16 + 64 * 160 * 160 * 32
    for tag in ("b200_conv3x3_dgrad_s2[64x320x320x16<-32]", "b200_conv3x3_fwd_s2[64x320x320x16->32]",
                "b200_conv3x3_wgrad[64x320x320x16->32,s2]"):
        w = sweep.seam_work(tag)
        assert w["bound"] == "hbm" and w["amount"] == 2.0 * n, tag
    w = sweep.seam_work("b200_nhwc_add[409600x64x2]")
    assert w["amount"] == 3.0 * 409600 * 64 * 2          # two gradient maps read, their sum written, 2 bytes per element
    w = sweep.seam_work("b200_stem_conv_fwd[64x640x640x3->16]")
    assert w["amount"] == (64 * 640 * 640 * 3 + 64 * 320 * 320 * 16) * 2.0
    assert sweep.seam_work("b200_nhwc_concat[409600x64]")["amount"] == 2.0 * 409600 * 64 * 2
    assert sweep.seam_work("b200_unknown[1x2]") is None
