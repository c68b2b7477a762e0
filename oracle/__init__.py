"""CPU oracle for the VPD student hot path (TEST INFRASTRUCTURE ONLY).

This package is a CPU restatement of the algorithms on the reference's student
path (jhong93/vpd: models/rgb.py, train_vpd_model.py, vpd_dataset/*,
apply_vpd_model.py) used to check the CUDA implementation in `vpd_b200/`.

Files: `assemble_ref.py` (K1 rows A1-A4, A13), `augment_ref.py` (the stochastic half of A3:
ColorJitter / RandomResizedCrop arithmetic with the draws given), `student_ref.py` (A5-A10),
`keypoint_ref.py` / `keypoint_train_ref.py` (the VIPE* teacher's eval forward and training step),
`ref_shim.py` + `gen_golden.py` (import the unmodified reference, write tests/golden/*).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it. The product path (`vpd_b200/`) never
imports, calls or falls back to anything in here.

Parity pinning: the reference ships no tests, fixtures or golden vectors
(SURVEY.md §4), so the oracle is pinned by executing the *unmodified reference*
in the build container (`oracle/gen_golden.py`, which imports /root/reference)
and committing its outputs as small fixtures under `tests/golden/`;
`tests/test_oracle_golden.py` checks the restatement against them on any box.
"""
