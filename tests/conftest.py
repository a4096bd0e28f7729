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
]


def case_inputs(meta):
    """Regenerate (params, ppg, sine, lft, spk) of a golden generator case from its seeds."""
    from svcc23_fastsvc_b200 import synthetic as syn
    params = syn.make_params(meta["config"], seed=meta["wseed"], weight_norm=meta["weight_norm"])
    ppg, sine, lft, spk = syn.make_inputs(meta["B"], meta["frames"], meta["config"],
                                          seed=meta["iseed"], with_spk=meta["with_spk"])
    return params, ppg, sine, lft, spk
