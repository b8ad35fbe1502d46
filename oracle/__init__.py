"""CPU oracle for the DISGAT hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product path
(``edgedisentangle_ssl_b200``) never imports it and fails loudly when the CUDA
extension is missing.

Parity status: PINNED.  Every function here is checked (tests/test_oracle_golden.py)
against golden vectors produced by importing the unmodified reference from
/root/reference (tests/golden/make_golden.py, committed with its outputs).
"""
from . import graph, disgat  # noqa: F401
