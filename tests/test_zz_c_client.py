"""The plain-C client of the boundary (tests/abi_client.c) on a GPU: frames added, matched through
the CUDA kNN kernel and read back, with no Python or torch in that process.  Written after this
round's GPU minutes were spent (the CPU half runs in tests/test_abi.py), hence sorted last."""
import pytest

pytestmark = pytest.mark.gpu


def test_c_client_round_trip_on_gpu(tmp_path):
    from test_abi import run_c_client
    out = run_c_client(tmp_path)
    assert "64 matches through the CUDA matcher" in out
