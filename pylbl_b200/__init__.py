"""pylbl_b200 -- a B200-native lines backend for pyLBL.

Importing this package registers :class:`Gas` as the lines backend ``"b200"`` in pyLBL's
plugin registry when pyLBL is importable, so that

    Spectroscopy(atmosphere, grid, database, lines_backend="b200").compute_absorption()

runs the line-by-line calculation on the GPU with everything else unchanged
(pyLBL/plugins.py:9-15, pyLBL/spectroscopy.py:117-118).
"""
from .gas_optics import Gas, grid_to_ints, pack_database, pack_info  # noqa: F401
from .continuum import Continuum, continua_of  # noqa: F401
from .mixture import Mixture, number_density  # noqa: F401
from .spectroscopy import Spectroscopy  # noqa: F401

BACKEND_NAME = "b200"


def register(name: str = BACKEND_NAME) -> bool:
    """Adds ``Gas`` to ``pyLBL.plugins.molecular_lines`` in this process.

    pyLBL builds that dict from the entry points of its *own* distribution
    (pyLBL/plugins.py:7-15), so a third-party package cannot join it through metadata;
    ``Spectroscopy.__init__`` reads the same dict object at construction time
    (pyLBL/spectroscopy.py:12,118), so inserting the class here is enough and modifies no
    reference file.  Returns False when pyLBL is not importable.
    """
    try:
        from pyLBL import plugins
    except Exception:
        return False
    plugins.molecular_lines[name] = Gas
    return True


registered = register()
