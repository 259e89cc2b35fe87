import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


# A rank that never publishes a phase (a bug, a dead peer) must fail a test, not hang the suite: the
# mailbox gives up after a minute here (default: half an hour; read once, at the first wait).
os.environ.setdefault("OPKB_MAILBOX_TIMEOUT", "60")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure only)."""
    import opk_oracle

    opk_oracle.build()
    opk_oracle.set_sum_mode(opk_oracle.SUM_ACCURATE)
    return opk_oracle
