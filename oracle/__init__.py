"""CPU oracle for the dsp/conv hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package (see oracle/conv_oracle.c header)."""
