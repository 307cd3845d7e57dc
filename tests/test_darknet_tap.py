"""-m "not gpu": SURVEY 8(f) row 4 -- DarknetTap runs only the backbone blocks the three taps depend on.  (1) liveness on a
hand-built block list (no reference needed); (2) against the UNMODIFIED reference Darknet on the real yolov3.cfg: identical taps,
the three 255-channel head convolutions and the three YOLO decodes are skipped (skipped where /root/reference is absent)."""
import os

import pytest
import torch
import torch.nn as nn

from dcnet_b200.model.darknet import DarknetTap
from oracle import ref_loader


class _Yolo(nn.Module):
    calls = 0

    def forward(self, x):
        _Yolo.calls += 1
        return x.flatten(1)[:, None, :8].expand(-1, 3, -1)


def _toy():
    """conv conv shortcut conv | yoloconv yolo | route(-3) conv upsample route(-1, 1) conv | yoloconv yolo"""
    defs = [dict(type="convolutional"), dict(type="convolutional"), dict(type="shortcut", **{"from": "-2"}), dict(type="convolutional"),
            dict(type="yoloconvolutional"), dict(type="yolo"),
            dict(type="route", layers="-3"), dict(type="convolutional"), dict(type="upsample"), dict(type="route", layers="-1, 1"),
            dict(type="convolutional"), dict(type="yoloconvolutional"), dict(type="yolo")]
    torch.manual_seed(0)
    mods = nn.ModuleList([nn.Sequential(nn.Conv2d(3, 8, 3, 2, 1)), nn.Sequential(nn.Conv2d(8, 8, 3, 1, 1)), nn.Sequential(),
                          nn.Sequential(nn.Conv2d(8, 16, 3, 2, 1)), nn.Sequential(nn.Conv2d(16, 255, 1)), nn.Sequential(_Yolo()),
                          nn.Sequential(), nn.Sequential(nn.Conv2d(16, 8, 1)), nn.Sequential(nn.Upsample(scale_factor=2)), nn.Sequential(),
                          nn.Sequential(nn.Conv2d(16, 8, 3, 1, 1)), nn.Sequential(nn.Conv2d(8, 255, 1)), nn.Sequential(_Yolo())])

    class Net(nn.Module):
        obj_out = False

        def __init__(self):
            super().__init__()
            self.module_defs, self.module_list = defs, mods

        def forward(self, x):                 # the reference's control flow (model/darknet.py:397-431), obj_out=False
            out, lo = [], []
            for i, (d, m) in enumerate(zip(self.module_defs, self.module_list)):
                t = d["type"]
                if t in ("convolutional", "upsample"):
                    x = m(x)
                elif t == "route":
                    x = torch.cat([lo[int(v)] for v in d["layers"].split(",")], 1)
                elif t == "shortcut":
                    x = lo[-1] + lo[int(d["from"])]
                elif t == "yoloconvolutional":
                    out.append(x); x = m(x)
                elif t == "yolo":
                    x = m(x)
                lo.append(x)
            return out
    return Net()


def test_liveness_and_equal_taps_on_a_toy_block_list():
    net = _toy().eval()
    tap = DarknetTap(net).eval()
    assert tap.skipped_blocks() == [4, 5, 11, 12]          # both head convolutions and both decodes; everything else feeds a tap
    x = torch.randn(2, 3, 32, 32)
    _Yolo.calls = 0
    want = net(x)
    assert _Yolo.calls == 2
    _Yolo.calls = 0
    got = tap(x)
    assert _Yolo.calls == 0
    assert len(got) == len(want) == 2
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    assert set(tap.state_dict()) == {"darknet." + k for k in net.state_dict()}


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not mounted")
def test_against_the_reference_darknet_on_yolov3_cfg():
    import importlib
    import sys
    ref_loader.load()
    D = sys.modules["model.darknet"]
    torch.manual_seed(3)
    net = D.Darknet(config_path=os.path.join(ref_loader.REF_ROOT, "model", "yolov3.cfg")).eval()      # the real (un-stubbed) class
    tap = DarknetTap(net).eval()
    types = [d["type"] for d in net.module_defs]
    skipped = tap.skipped_blocks()
    # yolov3.cfg: per scale the tapped tensor enters `yoloconvolutional` (1x1), then a 3x3 conv, the 255-channel head conv and the
    # decode follow; the FPN route (-4) reads the yoloconvolutional output, so at strides 32 and 16 the dead blocks are
    # [3x3 conv, head conv, yolo], and at stride 8 (nothing follows) also the yoloconvolutional itself: 10 blocks, 7 convolutions
    assert sorted(types[i] for i in skipped) == ["convolutional"] * 6 + ["yolo"] * 3 + ["yoloconvolutional"]
    dead_params = sum(p.numel() for i in skipped for p in net.module_list[i].parameters())
    assert dead_params > 6_000_000                                          # 3x3 512->1024 and 256->512 head convolutions
    x = torch.randn(1, 3, 64, 64)
    with torch.no_grad():
        want = net(x)
        got = tap(x)
    assert [tuple(t.shape) for t in got] == [(1, 1024, 2, 2), (1, 512, 4, 4), (1, 256, 8, 8)]
    for a, b in zip(got, want):
        assert torch.equal(a, b)
