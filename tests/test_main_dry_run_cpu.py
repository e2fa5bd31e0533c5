"""CPU, build container only: the harness of tests/test_gpu_main_dry_run.py - the test-only yacs / matplotlib stand-ins and the
synthetic GroZi-format dataset carry the UNMODIFIED reference main.py (evaluation run) end to end on the CPU."""
import os
import subprocess
import sys

import pytest

from _util import ROOT, have_reference


def test_yacs_stand_in_covers_the_surface_the_reference_uses(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests", "shims"))
    try:
        from yacs.config import CfgNode as CN
    finally:
        sys.path.pop(0)
    cfg = CN()
    cfg.is_cuda = True
    cfg.eval = CN()
    cfg.eval.dataset_names = ["a", "b"]
    cfg.eval.nms_score_threshold = float("-inf")
    cfg.eval.batch_size = 1
    cfg.merge_from_list(["eval.dataset_names", "['grozi-val-new-cl']", "is_cuda", "False", "eval.batch_size", "4"])
    assert cfg.eval.dataset_names == ["grozi-val-new-cl"] and cfg.is_cuda is False and cfg.eval.batch_size == 4
    yml = tmp_path / "c.yml"
    yml.write_text("eval:\n  batch_size: 2\n")
    cfg.merge_from_file(str(yml))
    assert cfg.eval.batch_size == 2 and "batch_size: 2" in cfg.dump() and "batch_size" in str(cfg)
    with pytest.raises(KeyError):
        cfg.merge_from_list(["eval.no_such_key", "1"])
    clone = cfg.clone()
    cfg.freeze()
    with pytest.raises(AttributeError):
        cfg.eval.batch_size = 8
    clone.eval.batch_size = 8
    assert cfg.eval.batch_size == 2 and clone.eval.batch_size == 8
    cfg.defrost()
    cfg.eval.batch_size = 3


def test_synthetic_grozi_dataset_has_the_reference_layout(tmp_path):
    import pandas as pd
    from PIL import Image
    import _synthetic_grozi
    base = _synthetic_grozi.make(str(tmp_path / "data"))
    df = pd.read_csv(os.path.join(base, "classes", "grozi.csv"))
    assert {"imageid", "imagefilename", "classid", "classfilename", "gtbboxid", "difficult", "lx", "ty", "rx", "by", "split"} <= set(df.columns)
    assert ((df.lx >= 0) & (df.rx <= 1) & (df.lx < df.rx) & (df.ty >= 0) & (df.by <= 1) & (df.ty < df.by)).all()
    for f in set(df.imagefilename):
        assert max(Image.open(os.path.join(base, "src", "3264", f)).size) == 3264      # no resize (dataset.py:671 uses Image.ANTIALIAS)
    for f in set(df.classfilename):
        assert os.path.isfile(os.path.join(base, "classes", "images", f))


@pytest.mark.skipif(not have_reference(), reason="/root/reference not present")
def test_unmodified_main_py_evaluation_run_on_cpu(tmp_path):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "run_reference_main.py"), str(tmp_path / "ref"), "0",
                          "is_cuda", "False", "train.do_training", "False", "eval.dataset_names", "['grozi-val-new-cl']",
                          "eval.dataset_scales", "[320.0]", "eval.scales_of_image_pyramid", "[1.0]", "eval.mAP_iou_thresholds", "[0.5]"],
                         capture_output=True, text=True, timeout=900)
    log = out.stdout + out.stderr
    assert "MAIN_DRY_RUN_DONE hook=0" in out.stdout, log[-3000:]
    assert "Loaded dataset grozi-val-new-cl with 2 images, 4 boxes, 3 classes" in log and "mAP@0.50" in log
