"""CPU checkers for the line-by-line path.  TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arms
may import this package.  Nothing under ``pylbl_b200/`` does.

Two checkers, same call surface (``absorption_coefficient`` mirroring
pyLBL/c_lib/gas_optics.py:46-92):

* :class:`ReferenceGas` -- the unmodified reference C sources compiled from
  ``/root/reference`` into ``oracle/_ref/libabsorption_ref.so`` (``oracle/Makefile``),
  called with the argtypes of pyLBL/c_lib/gas_optics.py:68-73.
* :class:`OracleGas` -- the restatement in ``oracle/lbl_oracle.c``, which is pinned
  bit-for-bit against the former in tests/test_oracle.py and tests/golden/.
"""
from .oracle import (OracleGas, ReferenceGas, build, have_reference, oracle_regions,  # noqa: F401
                     oracle_scale_line, oracle_tips, oracle_voigt, read_molecule,
                     reference_voigt)
