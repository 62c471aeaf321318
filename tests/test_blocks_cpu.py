"""Block-level host-logic tests on the CPU box: every block of pmf_b200.net through the numpy C-ABI model in EXACT
mode against the fp32 oracle + autograd, max-norm tolerances (forward 1e-5, gradients 1e-4)."""
import numpy as np
import pytest

from tests import block_harness as bh
from tests import cabi_mock

_CASES = bh.block_cases(np.random.RandomState(7))


@pytest.mark.parametrize("case", _CASES, ids=[c[0] for c in _CASES])
def test_block_matches_oracle_exact_model(monkeypatch, case):
    cabi_mock.install(monkeypatch, exact=True)
    name, mod, build, oracle_fn, inputs, masks, multi, in_kw = case
    res = bh.run_block("cpu", mod, build, oracle_fn, inputs, masks=masks, multi=multi, in_kw=in_kw)
    assert max(res["fwd"]) < 1e-5, res["fwd"]
    assert max(res["dinput"] + [0.0]) < 1e-4, res["dinput"]
    bad = {k: v for k, v in res["dparam"].items() if v > 1e-3}
    assert not bad, bad
    bad = {k: v for k, v in res["stats"].items() if v > 1e-4}
    assert not bad, bad
