"""Shared fixtures.  Tests marked `gpu` need a B200 (run with `-m gpu`); everything else runs on CPU."""
import json
import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")
WEIGHT_SEED = 7


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name: str):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def t(a) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a))


@pytest.fixture(scope="session")
def state_keys():
    with open(os.path.join(GOLDEN, "state_keys.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def weights(state_keys):
    """(encoder_state_dict, decoder_state_dict): the synthetic weights the golden fixtures were made with."""
    from tinyvc_b200.weights import synth_state_dict
    tmpl = {kind: {k: torch.empty(shape) for k, shape in state_keys[kind]} for kind in ("encoder", "decoder")}
    return synth_state_dict(tmpl["encoder"], WEIGHT_SEED), synth_state_dict(tmpl["decoder"], WEIGHT_SEED)


@pytest.fixture(scope="session")
def cuda_models(weights):
    """Encoder / Decoder of this package on cuda:0 carrying the golden weights."""
    from tinyvc_b200.tinyvc import Decoder, Encoder
    enc, dec = Encoder().eval(), Decoder().eval()
    enc.load_state_dict(weights[0], strict=True)
    dec.load_state_dict(weights[1], strict=True)
    return enc.to("cuda"), dec.to("cuda")


def rmse(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double().cpu() - b.double().cpu()).pow(2).mean().sqrt())


def max_abs(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double().cpu() - b.double().cpu()).abs().max())


class _Report:
    """Collects measured parity numbers; dumped to gpurun_out/parity_report.json at session end."""

    def __init__(self):
        self.rows = {}

    def add(self, key: str, **vals):
        self.rows[key] = {k: (float(v) if isinstance(v, (int, float)) else v) for k, v in vals.items()}
        print(f"[parity] {key}: " + ", ".join(f"{k}={v:.3e}" if isinstance(v, float) else f"{k}={v}" for k, v in vals.items()))


@pytest.fixture(scope="session")
def report():
    r = _Report()
    yield r
    if r.rows:
        out = os.path.join(REPO, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_report.json"), "w") as f:
            json.dump(r.rows, f, indent=1, sort_keys=True)
