"""Synthetic ensembles of BASELINE.json's configs.  Counter-based RNG so that every process, rank and
the CPU oracle generate identical inputs: u(i, d) = top 53 bits of splitmix64(0x5EED + 3*i + d) / 2^53."""
import numpy as np

SEED = 0x5EED
LORENZ_PARAMS = np.array([10.0, 28.0, 8.0 / 3.0])      # problem.rs:985-987 (SIGMA, RHO, BET)
PENDULUM_PARAMS = np.array([0.5, 1.2, 2.0 / 3.0])      # gamma, A, Omega


def splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform01(i, d):
    """u(i, d) in [0, 1) for trajectory index array i and component d."""
    i = np.asarray(i, dtype=np.uint64)
    with np.errstate(over="ignore"):
        ctr = np.uint64(SEED) + np.uint64(3) * i + np.uint64(d)
    return (splitmix64(ctr) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def lorenz_ics(begin, end):
    """Config 2/5: x, y ~ U[-20, 20], z ~ U[0, 50] for global trajectory indices [begin, end)."""
    i = np.arange(begin, end, dtype=np.uint64)
    y0 = np.empty((len(i), 3))
    y0[:, 0] = -20.0 + 40.0 * uniform01(i, 0)
    y0[:, 1] = -20.0 + 40.0 * uniform01(i, 1)
    y0[:, 2] = 50.0 * uniform01(i, 2)
    return y0


def vdp_sweep(begin, end, ntotal):
    """Config 3: mu_i = 0.1 + 9.9 * i / (N - 1), y0 = [2, 0]."""
    i = np.arange(begin, end, dtype=np.float64)
    mu = 0.1 + 9.9 * i / float(max(ntotal - 1, 1))
    y0 = np.tile(np.array([2.0, 0.0]), (len(i), 1))
    return y0, mu[:, None].copy()


def pendulum_ics(begin, end):
    """Config 4: theta0 ~ U[-pi, pi], omega0 ~ U[-2, 2]."""
    i = np.arange(begin, end, dtype=np.uint64)
    y0 = np.empty((len(i), 2))
    y0[:, 0] = -np.pi + 2.0 * np.pi * uniform01(i, 0)
    y0[:, 1] = -2.0 + 4.0 * uniform01(i, 1)
    return y0
