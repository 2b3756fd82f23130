"""`hirm.hirm`: the module name the reference's classes live under (src/hirm.py); here they are the engine-backed
mirror classes of hirm.model."""
from .model import CRP, HIRM, IRM, BetaBernoulli, Domain, Relation, logp_observations  # noqa: F401
