import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def model_dir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("gguf"))


@pytest.fixture(scope="session")
def gguf_path(model_dir):
    """factory: shape name -> path of a freshly generated random-init GGUF (cached per session)"""
    from blama_b200 import gguf_synth

    cache = {}

    def get(name: str) -> str:
        if name not in cache:
            p = os.path.join(model_dir, name + ".gguf")
            gguf_synth.write_gguf(p, name)
            cache[name] = p
        return cache[name]

    return get


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle

    pyoracle.lib()
    return pyoracle
