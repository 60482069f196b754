"""CPU oracle for the CSM training step — TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  See oracle/csm_oracle.py for the parity status.
"""
