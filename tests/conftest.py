import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kimera-multi_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    import kml_oracle
    kml_oracle.build()
    return kml_oracle


@pytest.fixture(scope="session")
def small_world():
    """2 robots x 400 keyframes, 100 places (50 aliased pairs), F = 500."""
    from kml import synth
    world = synth.World(100, F=500)
    chunks = list(synth.build_database(world, range(2), 400, chunk=400))
    queries = synth.make_queries(world, 12, 400, 2)
    return world, chunks, queries


def fill(det, chunks, bulk=False):
    for ch in chunks:
        if bulk:
            det.addBowVectors(ch["robot"], ch["poses"], ch["bow_off"], ch["bow_ids"], ch["bow_vals"])
            det.addVLCFrames(ch["robot"], ch["poses"], ch["desc"], ch["bearings"], ch["points"])
            continue
        for i, p in enumerate(ch["poses"]):
            o0, o1 = ch["bow_off"][i], ch["bow_off"][i + 1]
            det.addBowVector(ch["robot"], int(p), ch["bow_ids"][o0:o1], ch["bow_vals"][o0:o1])
            det.addVLCFrame(ch["robot"], int(p), ch["desc"][i], ch["bearings"][i], ch["points"][i])


@pytest.fixture(scope="session")
def oracle_lcd(oracle, small_world):
    world, chunks, queries = small_world
    lcd = oracle.LoopClosureDetector()
    fill(lcd, chunks)
    return lcd


@pytest.fixture(scope="session")
def gpu_lcd(small_world):
    import kml
    world, chunks, queries = small_world
    det = kml.LoopClosureDetector()
    fill(det, chunks, bulk=True)
    return det
