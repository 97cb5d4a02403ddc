"""pytest configuration: registers the ``gpu`` marker and shared fixtures."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden():
    arrays = dict(np.load(os.path.join(GOLDEN_DIR, "oracle_golden.npz")))
    with open(os.path.join(GOLDEN_DIR, "oracle_golden.json")) as f:
        meta = json.load(f)
    return arrays, meta


@pytest.fixture(scope="session")
def golden_rows():
    return dict(np.load(os.path.join(GOLDEN_DIR, "ddpm_rows_golden.npz")))


@pytest.fixture(scope="session")
def pkg():
    """The product package (hyphenated directory name -> importlib)."""
    import importlib
    return importlib.import_module("latent-diffusion-models-for-shape-sdfs_b200")


@pytest.fixture(scope="session")
def cuda_decoder(pkg):
    import oracle
    dec = pkg.Decoder(oracle.flatten_params(oracle.decoder_weights()), device="cuda:0", precision="bf16")
    yield dec
    dec.close()


@pytest.fixture(scope="session")
def cuda_ddpm(pkg):
    import oracle
    m = pkg.LatentDDPM(oracle.flatten_params(oracle.ddpm_weights()), device="cuda:0", precision="fp32")
    yield m
    m.close()
