"""Host-side numeric helpers with the names and meaning of the reference's src/util_math.py (callers import
`hirm.util_math.logsumexp` etc., tests/test_basic.py:18).  Nothing here is on the entity-move path: the engine has its
own fp64 device code (csrc/device_math.cuh)."""
import math
import random

inf = float("inf")


def linspace(start, stop, num=50, endpoint=True):
    """`num` evenly spaced points from start (src/util_math.py:9-12)."""
    step = (stop - start) / (num - (1 if endpoint else 0))
    return [start + k * step for k in range(num)]


def log_linspace(a, b, n):
    """n points from a to b, evenly spaced in the logarithm (src/util_math.py:14-17; the alpha grid of
    CRP.transition_alpha)."""
    return [math.exp(v) for v in linspace(math.log(a), math.log(b), num=n)]


def logsumexp(array):
    """log(sum(exp(a))) without overflow; -inf for an empty input, +/-inf propagate as in src/util_math.py:30-47."""
    array = list(array)
    if not array:
        return -inf
    top = max(array)
    if math.isinf(top) and min(array) != -top and not any(math.isnan(a) for a in array):
        return top
    return top + math.log(sum(math.exp(a - top) for a in array))


def log_normalize(log_weights):
    total = logsumexp(log_weights)
    return [w - total for w in log_weights]


def log_choices(population, log_weights, k=1, prng=None):
    """k draws from `population` with the given log-weights (src/util_math.py:24-28: Random.choices on exp of the
    normalised weights)."""
    weights = [math.exp(w) for w in log_normalize(log_weights)]
    return (prng or random).choices(population, weights, k=k)


def logmeanexp(array):
    """log(mean(exp(a))); -inf entries count towards the mean but not the sum (src/util_math.py:49-72)."""
    array = list(array)
    if not array:
        return -inf
    return logsumexp([a for a in array if not a == -inf]) - math.log(len(array))
