import json
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden_index():
    with open(os.path.join(GOLDEN, "index.json")) as f:
        return json.load(f)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


GENERATOR_CASES = [
    "gen_yaml_b1", "gen_yaml_b1_nospk", "gen_yaml_b1_wn", "gen_5442_b1", "gen_yaml_b2_f51",
    "gen_yaml_nospkmodule_b1", "gen_odd_b3", "gen_yaml_b1_f1", "gen_yaml_b1_f500", "gen_yaml_b32",
    # round 2 (tests/golden/make_golden_r2.py): other multiple-of-8 channel sets, un-fused level 0, B > 1 at 500 / 33
    # frames, (1, S) speaker broadcast, a 3-stage generator
    "gen_c64_b2", "gen_c256_b1", "gen_c48last_b2", "gen_c40last_b1", "gen_yaml_b3_f500", "gen_yaml_b2_f33",
    "gen_yaml_b4_spk1", "gen_3stage_b2",
]
BIG_CASES = ("gen_yaml_b32", "gen_yaml_b3_f500")   # the numpy oracle skips these (minutes of pure-numpy conv)


def case_inputs(meta):
    """Regenerate (params, ppg, sine, lft, spk) of a golden generator case from its seeds."""
    from svcc23_fastsvc_b200 import synthetic as syn
    params = syn.make_params(meta["config"], seed=meta["wseed"], weight_norm=meta["weight_norm"])
    ppg, sine, lft, spk = syn.make_inputs(meta["B"], meta["frames"], meta["config"],
                                          seed=meta["iseed"], with_spk=meta["with_spk"])
    if meta.get("spk_rows"):
        spk = spk[:meta["spk_rows"]]    # (1, S) target speaker broadcast over the batch (decode_fastsvc.py:156-158)
    return params, ppg, sine, lft, spk
