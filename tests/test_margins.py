"""CPU self-test of the margin-aware decision comparison (tests/margins.py): rows perturbed at fp16-like
and fp32-like magnitudes, decided by the oracle tail itself, must pass; a forged decision in a clean class and
an inadmissible keep in an ambiguous one must be caught."""
import numpy as np
import pytest

import margins as MG
from golden import make_golden as mg


def _geom(name):
    c = mg.CASES[name]
    shapes = [(h, w) for h, w, _ in c["levels"]]
    if c["kind"] == "ssd":
        return "ssd", {"size": c["size"], "shapes": shapes}, 20
    if c["kind"] == "retina":
        return "retina", {"data_shape": c["data_shape"], "shapes": shapes}, 20
    if c["kind"] == "yolo":
        return "yolo", {"levels": c["levels"], "priors": mg.YOLO_PRIORS}, 20
    return "fcos", {"levels": c["levels"]}, 19


@pytest.mark.parametrize("name", ["tail_ssd", "tail_retina", "tail_yolo", "tail_fcos"])
@pytest.mark.parametrize("noise", [2e-7, 3e-3])
def test_perturbed_rows_pass(name, noise):
    c = mg.CASES[name]
    kind, geom, ncls = _geom(name)
    ref = mg.make_rows(name, batch=1, seed=5)[0]
    rng = np.random.default_rng(1)
    gpu = (ref + noise * np.abs(ref).max() * rng.uniform(-1, 1, ref.shape)).astype(np.float32)
    s, b, ids, keep = mg.run_case(name, gpu)
    rep = MG.compare_image(kind, gpu, ref, geom, ids, keep, b, c["score_thr"], c["max_boxes"], c["iou_thr"], ncls)
    assert rep["clean"] + rep["ambiguous"] == ncls
    if noise < 1e-6:
        assert rep["clean"] >= ncls - 2 and rep["identical"] >= ncls - 2, rep


def test_forged_decisions_are_caught():
    name = "tail_ssd"
    c = mg.CASES[name]
    kind, geom, ncls = _geom(name)
    ref = mg.make_rows(name, batch=1, seed=5)[0]
    s, b, ids, keep = mg.run_case(name, ref)
    rep = MG.compare_image(kind, ref, ref, geom, ids, keep, b, c["score_thr"], c["max_boxes"], c["iou_thr"], ncls)
    assert rep["clean"] == ncls and rep["identical"] == ncls and rep["ds"] == 0.0
    # drop one kept box of a class
    drop = np.ones(len(ids), bool)
    drop[len(ids) // 2] = False
    with pytest.raises(AssertionError):
        MG.compare_image(kind, ref, ref, geom, ids[drop], keep[drop], b[drop], c["score_thr"], c["max_boxes"],
                         c["iou_thr"], ncls)
    # swap two keeps of one class (order matters: the reference emits descending scores)
    cls0 = np.nonzero(ids == ids[0])[0]
    if len(cls0) >= 2:
        k2 = keep.copy()
        k2[cls0[0]], k2[cls0[1]] = keep[cls0[1]], keep[cls0[0]]
        with pytest.raises(AssertionError):
            MG.compare_image(kind, ref, ref, geom, ids, k2, b, c["score_thr"], c["max_boxes"], c["iou_thr"], ncls)
    # noisy rows + a kept row far below the threshold: inadmissible even in an ambiguous class
    rng = np.random.default_rng(3)
    gpu = (ref + 3e-3 * np.abs(ref).max() * rng.uniform(-1, 1, ref.shape)).astype(np.float32)
    s, b, ids, keep = mg.run_case(name, gpu)
    S_o = MG.per_row(kind, ref, geom)[0]
    worst = int(np.argmin(S_o[:, ids[0]]))
    k3 = keep.copy()
    k3[0] = worst
    with pytest.raises(AssertionError):
        MG.compare_image(kind, gpu, ref, geom, ids, k3, b, c["score_thr"], c["max_boxes"], c["iou_thr"], ncls)
