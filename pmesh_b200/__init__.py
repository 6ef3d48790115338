"""pmesh_b200 -- a B200-native particle-mesh engine behind the pmesh Python API.

Public interface, as in the reference (pmesh/__init__.py:1-2): ``__version__`` and
``ParticleMesh``; users import ``pmesh_b200.pm``, ``pmesh_b200.window`` and
``pmesh_b200.domain`` directly.  Importing this package loads
``pmesh_b200/csrc/libpmesh_b200.so`` and fails loudly if it is missing.
"""
from .version import __version__
from .pm import ParticleMesh
