"""Access to the committed fixtures under tests/golden/ (see tests/golden/make_golden.py for how they were made)."""
import copy
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sphere32():
    """Reader::ms for test/microstructures/sphere32.h5: memory order [x][y][z] after the zyx->xyz transpose."""
    zyx = np.load(os.path.join(GOLDEN, "sphere32_ms.npz"))["ms_zyx"]
    return np.ascontiguousarray(zyx.transpose(2, 1, 0)).astype(np.uint16)


def reference_input(name, load_cases=None):
    cfg = copy.deepcopy(json.load(open(os.path.join(GOLDEN, "reference_inputs.json")))[name])
    if load_cases is not None:
        cfg["macroscale_loading"] = [cfg["macroscale_loading"][i] for i in load_cases]
    return cfg


def oracle_scenario(name):
    return json.load(open(os.path.join(GOLDEN, "oracle_scenarios.json")))[name]
