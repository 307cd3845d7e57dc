"""-m "not gpu": the reference-side patch documented in INTEGRATION.md (integration/train_DCNet.patch) applies to the
UNMODIFIED reference training script and really rebinds the hot-path names to this package (an import placed
above the script's own `def yolo_loss ...` is silently shadowed by them).  Needs /root/reference: skipped on the GPU box."""
import importlib.util
import os
import shutil
import subprocess
import sys

import pytest

from oracle import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATCH = os.path.join(ROOT, "integration", "train_DCNet.patch")

pytestmark = pytest.mark.skipif(not ref_loader.available() or shutil.which("patch") is None,
                                reason="reference checkout (or patch(1)) not available")


def test_patch_applies_and_rebinds_the_loss_functions(tmp_path):
    ref_loader.load()                                   # installs the import shims; reference stays unmodified
    shutil.copy(os.path.join(ref_loader.REF_ROOT, "train_DCNet.py"), tmp_path / "train_DCNet.py")
    r = subprocess.run(["patch", "-p1", "-i", PATCH], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert not list(tmp_path.glob("*.rej"))
    cwd = os.getcwd()
    os.chdir(ref_loader.REF_ROOT)                       # the script's imports are relative to the checkout
    try:
        spec = importlib.util.spec_from_file_location("train_DCNet_patched", tmp_path / "train_DCNet.py")
        T = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(T)
    finally:
        os.chdir(cwd)
    import dcnet_b200.losses as LS
    from dcnet_b200.model.DCNet_model import grounding_model
    for name in ("yolo_loss", "rank_loss", "loc_loss", "Interframe_contrastive_loss", "Crossmodal_constrastive_loss", "build_target",
                 "configure"):
        assert getattr(T, name) is getattr(LS, name), name
    assert T.grounding_model is grounding_model
    assert hasattr(T, "Darknet") and T.Darknet.__module__.startswith("model.darknet")     # the backbone is still the reference's
    assert T.random is __import__("random")             # the star-import chain the model relies on (utils.utils -> random) survives
    src = (tmp_path / "train_DCNet.py").read_text()
    assert "neg_feature,vit_posit,lag_posit,neg_cross = model(image, word_id, word_mask)" in src
