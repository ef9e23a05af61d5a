import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def mat(d):
    return (np.array(d["re"]) + 1j * np.array(d["im"])).reshape(d["shape"])


def z(pair):
    return complex(pair[0], pair[1])
