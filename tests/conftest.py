import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_package():
    """The package directory is `hifiles-solver_b200` (hyphen): import it by path under a legal module name."""
    name = "hifiles_solver_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "hifiles-solver_b200", "__init__.py"),
                                                  submodule_search_locations=[os.path.join(ROOT, "hifiles-solver_b200")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def hb():
    mod = load_package()
    if not os.path.exists(mod.LIB_PATH):
        mod.build()
    return mod


@pytest.fixture(scope="session")
def meshgen(hb):
    import importlib
    return importlib.import_module("hifiles_solver_b200.meshgen")
