"""B200 engine for HIRM's collapsed-Gibbs entity move: `engine` is the ctypes binding of the C ABI
(include/hirm_engine.h), `model` the host-side mirror of the reference's Python interface (IRM, HIRM),
`sharded` the HIRM with its relation-cluster IRMs spread over GPUs (ShardedHIRM, ChainSet), `util_io` / `cli` the
reference's text formats and command line."""
from .model import HIRM, IRM  # noqa: F401
