import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: larger CPU case")


def golden_cases():
    return sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "case_*.npz")))


def load_case(path):
    from voicesplit_b200 import synth
    z = np.load(path)
    d = [int(v) for v in z["dims"]]
    dims = synth.make_dims(d[0], d[1], d[2], d[3], d[4])
    case = dict(name=os.path.basename(path)[5:-4], model_name=str(z["model_name"]),
                flavour=str(z["flavour"]), wseed=int(z["wseed"]), iseed=int(z["iseed"]),
                B=int(z["B"]), T=int(z["T"]), dims=dims)
    for k in ("mask", "masked", "conv_out", "lstm_out", "act3_sample"):
        case[k] = z[k]
    case["state_dict"] = synth.make_state_dict(dims, case["wseed"], case["flavour"])
    case["x"], case["emb"] = synth.make_inputs(case["B"], case["T"], dims, case["iseed"])
    return case


@pytest.fixture(params=golden_cases(), ids=lambda p: os.path.basename(p)[5:-4])
def golden(request):
    return load_case(request.param)
