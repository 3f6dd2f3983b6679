"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's algorithm for the hot path (OHDSI/bayes-bridge v0.2.6), used as the
checker by tests/, __graft_entry__.smoke() and bench.py's cpu baseline. Nothing under
bayesbridge_b200/ may import this package.

Pinning ("Oracle"): every module here is checked in tests/test_oracle_*.py against
fixtures generated from the reference itself (tests/golden/make_golden.py, run in the build container
against oracle/_ref) and against the reference's own golden vectors
(tests/regression_tests/saved_outputs/{linear_cg,logit_cg}_samples.npy, copied to tests/golden/ref_saved/).
"""
