"""ORACLE / TEST INFRASTRUCTURE ONLY.

CPU restatement of the korentomas/mlx-mcmc sampling path (see oracle/refport) on a torch-CPU
stand-in for mlx.core (oracle/mlx_shim).  Parity status: PINNED -- `oracle/make_golden.py`
runs the *unmodified* reference source from /root/reference on the same stand-in and asserts
bit-equal draws / values against this restatement, then writes `tests/golden/*.json`.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (mlx_mcmc_b200) never does.
"""
