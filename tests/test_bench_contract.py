"""CPU: the reference arm of bench.py (the one leg that may execute the oracle / the reference) keeps the JSON contract, and the
small host-side helpers added in round 2 behave."""
import json
import os
import subprocess
import sys

from _util import ROOT, have_reference


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "256",
                          "--cpu-sample-classes", "2", "--classes", "100", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout + out.stderr[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("query-classes/sec") and d["unit"] == "classes/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "classes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["cores"] >= 1 and cb["kind"] in ("reference", "port")
    if have_reference() or os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "os2d")):
        assert cb["kind"] == "reference" and "unmodified os2d.modeling.head.Os2dHead.forward" in cb["sample"]
    # the arm times a bounded sample and says so instead of claiming the full configuration
    assert d["config"]["classes_timed_per_step"] == 2 and d["config"]["same_config"] is False


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_cfg_get_reads_dicts_namespaces_and_defaults():
    from types import SimpleNamespace
    from os2d_b200.evaluate import _cfg_get
    cfg = {"eval": {"batch_size": 4, "nested": SimpleNamespace(x=7)}, "is_cuda": True}
    assert _cfg_get(cfg, "eval.batch_size") == 4 and _cfg_get(cfg, "eval.nested.x") == 7
    assert _cfg_get(cfg, "eval.missing", 3) == 3 and _cfg_get(cfg, "visualization.eval.path", "") == ""
    ns = SimpleNamespace(eval=SimpleNamespace(mAP_iou_thresholds=[0.5]))
    assert _cfg_get(ns, "eval.mAP_iou_thresholds") == [0.5]


def test_label_sharding_keeps_all_views_of_a_label_on_one_rank():
    """ClassShardedDetector shards REAL labels (first-occurrence order) in contiguous blocks; every class view follows its label."""
    from os2d_b200.dist import shard_bounds
    class_ids = [7, 3, 7, 5, 9, 3, 11]
    labels = list(dict.fromkeys(class_ids))
    owner = {}
    for world in (1, 2, 3, 4):
        seen = []
        for r in range(world):
            lo, hi = shard_bounds(len(labels), world, r)
            mine = set(labels[lo:hi])
            views = [i for i, c in enumerate(class_ids) if c in mine]
            seen += views
            for i in views:
                owner[(world, class_ids[i])] = owner.get((world, class_ids[i]), r)
                assert owner[(world, class_ids[i])] == r
        assert sorted(seen) == list(range(len(class_ids)))
