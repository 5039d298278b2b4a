"""tauray_b200: B200-native path-tracing hot path behind Tauray's renderer/stage surface.

Only what the hot path needs lives here: the C ABI binding (``library``), the host mirror of the
reference's stage/renderer interface (``renderer``), pixel distribution (``distribution``) and the
scene-table helpers used by tests and the bench (``scene``, ``synthetic``, ``glb``).
"""
from . import abi, distribution, scene  # noqa: F401
