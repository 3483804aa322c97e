import os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_c():
    """The C restatement of the reference arithmetic (test infrastructure only)."""
    from oracle import c_oracle
    c_oracle.lib()
    return c_oracle


@pytest.fixture(scope="session")
def ctx():
    """A libbzhalo2 context on cuda:0 -- the product path.  Fails loudly without the CUDA library / a GPU."""
    import battlezips_halo2_b200 as bz
    c = bz.Context(0)
    yield c
    c.close()
