import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT / "othello-alphazero_b200", ROOT / "oracle", ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import json
    with open(ROOT / "tests" / "golden" / "golden.json") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def ref_shim(oracle):
    r = oracle.ref_shim()
    if r is None:
        pytest.skip("oracle/_ref/libref_shim.so not built (needs /root/reference)")
    return r


@pytest.fixture(scope="session")
def ref_module():
    """The unmodified reference extension, when oracle/_ref holds it."""
    ref_dir = ROOT / "oracle" / "_ref"
    if not (ref_dir / "othello_mcts" / "__init__.py").exists():
        pytest.skip("reference module not built")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_othello_mcts", ref_dir / "othello_mcts" / "__init__.py",
                                                  submodule_search_locations=[str(ref_dir / "othello_mcts")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_othello_mcts"] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception as exc:  # e.g. a torch ABI mismatch on another box
        pytest.skip(f"reference module not loadable: {exc}")
    return mod
