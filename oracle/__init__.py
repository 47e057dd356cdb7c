"""ORACLE — CPU fp32 restatement of the reference's diffusion training step.

TEST INFRASTRUCTURE ONLY.  Nothing under neurosis_b200/ imports this package; only tests/,
__graft_entry__.smoke() and bench.py (cpu_baseline leg / --impl reference arm) do.
Pinned against the real reference by tests/test_oracle_vs_reference.py (run where /root/reference
exists) and the golden vectors in tests/golden/ generated from the reference by make_golden.py.
"""
