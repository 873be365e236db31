"""qgcm_b200 -- host-side mirror of the Q-GCM per-timestep interface over libqgcm_b200.so.

The directory is named ``q-gcm_b200`` (not importable by name); load it with
``_pkg.load()`` from the repo root, which registers it as module ``qgcm_b200``.

Only the hot path lives here: the C-ABI CUDA library (``csrc/``), its ctypes binding
(``abi.py``, ``model.py``) whose methods carry the reference's subroutine names
(src/q-gcm.F:1222-1269), and the host-side restatement of the *inputs* the Fortran
main program prepares before the loop (``params.py``: src/q-gcm.F:377-452,
src/eigmode.f:41-440; ``synth.py``: SURVEY.md section 8d synthetic states).
There is no CPU fallback: constructing a Model without the CUDA library raises.
"""
from .abi import QgcmConfig, QgcmScalars, QgcmValidsReport, QgcmMonitorOcean, QgcmMonitorAtmos, FLAGS, NLMAX  # noqa: F401
from .params import Params, named_config, build_config  # noqa: F401
from .model import Model, CModel, SlabGroup, slab_config, slab_bounds, load_library, library_path  # noqa: F401
from . import synth  # noqa: F401
