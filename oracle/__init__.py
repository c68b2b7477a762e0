"""CPU oracle for the VPD student hot path (TEST INFRASTRUCTURE ONLY).

This package is a CPU restatement of the algorithms on the reference's student
path (jhong93/vpd: models/rgb.py, train_vpd_model.py, vpd_dataset/*,
apply_vpd_model.py) used to check the CUDA implementation in `vpd_b200/`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it. The product path (`vpd_b200/`) never
imports, calls or falls back to anything in here.

Parity pinning: the reference ships no tests, fixtures or golden vectors
(SURVEY.md §4), so the oracle is pinned by executing the *unmodified reference*
in the build container (`oracle/gen_golden.py`, which imports /root/reference)
and committing its outputs as small fixtures under `tests/golden/`;
`tests/test_oracle_golden.py` checks the restatement against them on any box.
"""
