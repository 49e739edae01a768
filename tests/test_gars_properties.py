"""Property-based checks of the host aggregation rules (hypothesis): invariances every rule must have, for random shapes and values."""

import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from aggregathor_b200.aggregators import _ops

_RULES = {
  "average": lambda G, f: _ops.host_average(G),
  "average-nan": lambda G, f: _ops.host_average_nan(G),
  "median": lambda G, f: _ops.host_median(G),
  "averaged-median": lambda G, f: _ops.host_averaged_median(G, G.shape[0] - f),
  "krum": lambda G, f: _ops.host_krum(G, f, G.shape[0] - f - 2),
}


@st.composite
def gradients(draw):
  f = draw(st.integers(0, 2))
  n = draw(st.integers(2 * f + 3, 12))
  d = draw(st.integers(1, 70))
  seed = draw(st.integers(0, 2 ** 31 - 1))
  G = torch.from_numpy(np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32))
  return G, f, seed


@settings(max_examples=40, deadline=None)
@given(gradients(), st.sampled_from(sorted(_RULES)))
def test_translation_and_scaling_equivariance(case, rule):
  G, f, _ = case
  base = _RULES[rule](G, f)
  shift = torch.linspace(-1.0, 1.0, G.shape[1])
  assert torch.allclose(_RULES[rule](G + shift, f), base + shift, atol=2e-5)
  assert torch.allclose(_RULES[rule](G * 4.0, f), base * 4.0, atol=2e-5)      # power-of-two scaling: exact selections, exact products


@settings(max_examples=40, deadline=None)
@given(gradients(), st.sampled_from(["average", "average-nan", "median", "averaged-median"]))
def test_coordinate_rules_ignore_worker_order(case, rule):
  G, f, seed = case
  order = torch.from_numpy(np.random.default_rng(seed + 1).permutation(G.shape[0]))
  assert torch.allclose(_RULES[rule](G[order], f), _RULES[rule](G, f), atol=1e-5)   # distinct values almost surely: ties do not decide


@settings(max_examples=30, deadline=None)
@given(gradients())
def test_identical_gradients_are_a_fixed_point(case):
  G, f, _ = case
  same = G[:1].repeat(G.shape[0], 1)
  for rule in _RULES:
    assert torch.allclose(_RULES[rule](same, f), G[0], atol=1e-6), rule


@settings(max_examples=30, deadline=None)
@given(gradients())
def test_outputs_stay_in_the_coordinate_wise_hull_and_nan_rows_are_survived(case):
  G, f, _ = case
  lo, hi = G.min(dim=0).values - 1e-5, G.max(dim=0).values + 1e-5
  for rule in _RULES:
    out = _RULES[rule](G, f)
    assert bool(((out >= lo) & (out <= hi)).all()), rule
  if f >= 1:
    bad = G.clone()
    bad[-1] = float("nan")                                           # one fully lost gradient
    for rule in ("average-nan", "median", "krum"):
      out = _RULES[rule](bad, f)
      assert bool(torch.isfinite(out).all()), rule
      assert bool(((out >= lo) & (out <= hi)).all()), rule
