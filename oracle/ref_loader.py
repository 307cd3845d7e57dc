"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference (mengcaopku/DCNet).

Imports /root/reference in-process with the runtime shims listed in SURVEY.md
Appendix C (no reference source is copied or edited).  It only works in the
build container (where /root/reference is mounted); it is used by
  * tests/golden/make_golden.py   (to generate the committed golden vectors)
  * tests/test_oracle_vs_reference.py (-m "not gpu", skipped when the mount is absent)
Nothing on the product path, in bench.py or in the -m gpu tests imports this.
"""
import collections
import collections.abc
import os
import sys
import types

REF_ROOT = os.environ.get("DCNET_REFERENCE_ROOT", "/root/reference")

ANCHORS_FULL = [(373.0, 326.0), (156.0, 198.0), (116.0, 90.0), (59.0, 119.0), (62.0, 45.0),
                (30.0, 61.0), (33.0, 23.0), (16.0, 30.0), (10.0, 13.0)]  # train_DCNet.py:404-406


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "model", "DCNet_model.py"))


class _StubDarknet:  # replaced below by an nn.Module once torch is imported
    pass


_loaded = None


def load(size=256):
    """Returns (train_DCNet module, model.DCNet_model module, model.test_DCNet_model module)."""
    global _loaded
    if _loaded is not None:
        _loaded[0].args.size = size
        return _loaded
    if not available():
        raise RuntimeError("reference not mounted at %s" % REF_ROOT)
    collections.Iterable = collections.abc.Iterable                      # utils/transforms.py:10
    for n in ["pytorch_pretrained_bert", "pytorch_pretrained_bert.tokenization",
              "pytorch_pretrained_bert.modeling", "matplotlib", "matplotlib.pyplot", "scipy.misc"]:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    sys.modules["pytorch_pretrained_bert.tokenization"].BertTokenizer = object   # model/DCNet_model.py:19
    sys.modules["pytorch_pretrained_bert.modeling"].BertModel = object           # model/DCNet_model.py:20
    sys.modules["matplotlib"].use = lambda *a, **k: None                         # train_DCNet.py:14
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import torch
    import torch.nn as nn
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self                          # hard-coded .cuda()
    cwd = os.getcwd()
    os.chdir(REF_ROOT)                                                           # ./model/yolov3.cfg is cwd-relative
    sys.path.insert(0, REF_ROOT)
    try:
        import train_DCNet as T
        M = sys.modules["model.DCNet_model"]
        import model.test_DCNet_model as MT
    finally:
        os.chdir(cwd)

    class StubDarknet(nn.Module):
        """Stands in for model/darknet.py:377 Darknet: returns the synthetic feature maps
        that were attached with set_maps() (the backbone is outside the hot path)."""
        def __init__(self, config_path=None, img_size=416, obj_out=False):
            super().__init__()
            self.maps = None

        def set_maps(self, maps):
            self.maps = maps

        def load_weights(self, path):
            return None

        def forward(self, x):
            return [m for m in self.maps]

    M.Darknet = StubDarknet
    MT.Darknet = StubDarknet
    T.args = types.SimpleNamespace(size=size, anchor_imsize=416)
    T.anchors_full = list(ANCHORS_FULL)
    _loaded = (T, M, MT)
    return _loaded
