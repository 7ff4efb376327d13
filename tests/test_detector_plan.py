"""not-gpu: the launch plan DetectorW6 builds (every conv descriptor, the op order, the head levels), recorded on CPU by
tests/plan_dryrun.py, against (a) the committed plan of the build that passed the round-1 B200 parity tests
(tests/golden/w6_plan.json) and (b) for non-square inputs, the tensor shapes of the torch oracle."""
import json
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from plan_dryrun import dry_run_plan  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "w6_plan.json")


def _norm(conv):
    return {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in conv.items()}


@pytest.mark.parametrize("idx", [0, 1])
def test_square_plan_equals_gpu_verified_golden(idx):
    g = json.load(open(GOLDEN))[idx]
    det, plan = dry_run_plan(g["batch"], g["size"])
    assert det.n_total == g["n_total"] and [n for _, _, n in det.ops] == g["ops"]
    assert len(plan) == len(g["convs"]) == 96
    for k, (a, b) in enumerate(zip(plan, g["convs"])):
        assert _norm(a) == b, "conv %d (%s) differs from the verified plan" % (k, g["ops"][k])
    for lv, ref in zip(det.head_levels, g["levels"]):
        assert (lv.h, lv.w, lv.stride, lv.level_off, lv.raw_pitch, list(lv.anchors)) == (ref["h"], ref["w"], ref["stride"], ref["level_off"],
                                                                                          ref["raw_pitch"], ref["anchors"])
