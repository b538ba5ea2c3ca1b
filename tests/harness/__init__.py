"""Measurement / diagnostic harnesses that use the oracle (test infrastructure, like everything under
tests/): full-size parity report, C4 strong scaling with a per-rank parity sample, precision study,
timing of the unmodified reference, two-device check.  Stand-alone scripts: ``python tests/harness/x.py``."""
