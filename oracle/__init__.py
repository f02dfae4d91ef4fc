"""CPU oracle for the SMPL hot path -- test infrastructure only (see smpl_oracle.py header).

PARITY STATUS: smplx arithmetic = "parity unpinned" (third-party, absent, no reference tests);
in-tree projection / rot6d / loss / index tables = pinned against the reference files via
tests/golden/intree_golden.npz.
"""
