"""Oracle pinning (CPU): oracle/msda_ref.c against golden vectors produced by the reference's own
ms_deform_attn_core_pytorch (tests/golden/make_msda_golden.py), tolerances from the reference's
test script (ops/test.py: fp64 allclose, fp32 rtol 1e-2 / atol 1e-3 — we hold much tighter)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import msda as omsda

G = np.load(os.path.join(GOLDEN, "msda_golden.npz"))
CASES = sorted({k.split("/")[0] for k in G.files})


def case(name):
    return {k.split("/")[1]: G[k] for k in G.files if k.startswith(name + "/")}


@pytest.mark.parametrize("name", CASES)
def test_oracle_forward_matches_reference_golden(name):
    c = case(name)
    out = omsda.msda_forward(c["value"], c["shapes"], c["lsi"], c["loc"], c["attn"])
    if c["value"].dtype == np.float64:
        np.testing.assert_allclose(out, c["out"], rtol=1e-10, atol=1e-14)
    else:
        np.testing.assert_allclose(out, c["out"], rtol=1e-4, atol=1e-8)


@pytest.mark.parametrize("name", CASES)
def test_oracle_backward_matches_reference_autograd(name):
    c = case(name)
    gv, gl, ga = omsda.msda_backward(c["value"], c["shapes"], c["lsi"], c["loc"], c["attn"], c["gout"])
    if c["value"].dtype == np.float64:
        tol = dict(rtol=1e-9, atol=1e-13)
    else:
        tol = dict(rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(gv, c["gvalue"], **tol)
    np.testing.assert_allclose(ga, c["gattn"], **tol)
    np.testing.assert_allclose(gl, c["gloc"], **tol)


def test_oracle_empty_queries():
    v = np.zeros((2, 6, 2, 4), np.float32)
    out = omsda.msda_forward(v, [(2, 3)], [0], np.zeros((2, 0, 2, 1, 4, 2), np.float32),
                             np.zeros((2, 0, 2, 1, 4), np.float32))
    assert out.shape == (2, 0, 8)
