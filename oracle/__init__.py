"""CPU oracle for the burst-DSP path -- test infrastructure only (see oracle/oracle.py)."""
