"""B200 engine for HIRM's collapsed-Gibbs entity move: `engine` is the ctypes binding of the C ABI
(include/hirm_engine.h), `model` the host-side mirror of the reference's Python interface."""
from .model import HIRM, IRM  # noqa: F401
