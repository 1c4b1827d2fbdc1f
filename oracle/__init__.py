"""CPU oracle for the UNetSCN hot path: test infrastructure only (see scn_oracle.py). PARITY UNPINNED."""
