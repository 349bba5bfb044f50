import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_fixture_mesh(name: str, fast: bool = True):
    import edgefem_oracle as orc

    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return orc.mesh_from_arrays(z["xyz"], z["tet_conn"], z["tet_phys"], z["tri_conn"], z["tri_phys"], node_ids=z["node_ids"], fast=fast)


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def wr90():
    import edgefem_oracle as orc

    mesh = load_fixture_mesh("rect_waveguide")
    pec = orc.build_edge_pec(mesh, 1)
    return mesh, pec


@pytest.fixture(scope="session")
def cube():
    return load_fixture_mesh("cube_cavity")
