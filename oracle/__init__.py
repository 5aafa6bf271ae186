"""CPU oracle for the vaemolsim hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, in NumPy, the arithmetic the reference executes through
TensorFlow / TensorFlow-Probability for the path named in BASELINE.json
(`north_star`).  It is the checker for the CUDA kernels in `vaemolsim_b200/csrc`.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it.  The product package (`vaemolsim_b200`) never
does: it fails loudly when the CUDA library is missing.

Parity status
-------------
* `mcmc.py` acceptance arithmetic: PINNED -- `tests/golden/make_goldens.py` runs the
  reference's own `vaemolsim/mcmc.py` (NumPy only) on top of this oracle's VAE and
  commits the decisions; `oracle.mcmc.single_step` is checked against them bit for bit.
* MAF input orders, Keras parameter counts, `make_param_transform` known answers,
  loss identities, `DistanceSelection` identities: PINNED by the reference's own tests /
  notebooks, re-checked in `tests/test_oracle_*.py`.
* RQS values / log-dets, MADE outputs, distribution log-probs, gradients: the arithmetic
  lives in tensorflow-probability 0.23 / tensorflow 2.15 (`pyproject.toml:27-28`), neither
  vendored nor installable here => **parity unpinned** against the real TFP numbers.  The
  restatement follows the published TFP v0.23.0 algorithms and is validated by analytic
  properties (round trip, log-det vs finite differences, normalisation, float64 autograd).
"""
