"""Latin-hypercube sampling with the reference's random-number consumption (`autooed/utils/sampling/lhs.py:126-144`,
the default "classic" design): one `np.random.random((samples, n))` draw for the positions inside the strata, then one
`np.random.permutation(samples)` per factor for the pairing — so that a seeded run draws the same design as AutoOED."""
import numpy as np


def lhs(n, samples=None):
    samples = n if samples is None else samples
    jitter = np.random.random((samples, n))
    edges = np.linspace(0, 1, samples + 1)  # stratum i covers [edges[i], edges[i+1])
    lo, width = edges[:-1, None], np.diff(edges)[:, None]
    points = jitter * width + lo
    design = np.empty_like(points)
    for j in range(n):
        design[:, j] = points[np.random.permutation(samples), j]
    return design
