"""Synthetic input builders for the benchmark configurations of BASELINE.json.

Host-side numpy only (set-up, never timed). Restates the reference's recipes:
  bcc lattice        utility/lattice.hpp:50-109 motif_to_lattice: motif {(0,0,0),(1/2,1/2,1/2)}*a tiled x-fastest,
                     every coordinate shifted by +pi (:106)
  hydrogen cell      apps/hydrogen.cpp:111-139: remove listed Fe sites, append one H at the quoted position
  perturbation       saddle::Master::perturb, src/saddle/find.cpp:1147-1214: stddev ~ N(0.6, 0.2) redrawn until >= 0,
                     atoms within r_pert of the centre displaced by N(0, stddev)(1 - r/r_pert) and given a random
                     axis component; the centre gets a full kick; the axis is normalised
The reference seeds from std::random_device, so there are no golden perturbations; numpy's PCG64 is used here.
"""
from __future__ import annotations

import numpy as np


def bcc_lattice(a: float, n: int) -> np.ndarray:
    k, j, i = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")  # x fastest
    cells = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1).astype(np.float64)
    pos = np.empty((2 * n ** 3, 3), dtype=np.float64)
    pos[0::2] = a * 0.0 + a * cells
    pos[1::2] = a * 0.5 + a * cells
    return pos + np.pi


def iron_hydrogen_cell(n: int, a: float, remove=(1,), hydrogen: bool = True):
    """(positions [N,3], extents [3], type ids [N]) of an n^3 bcc Fe cell with vacancies and one H."""
    pos = np.delete(bcc_lattice(a, n), list(remove), axis=0)
    ids = np.zeros(len(pos), dtype=np.uint32)
    if hydrogen:
        pos = np.vstack([pos, [[2.857 / 2 + 3.14, 2.857 / 2 + 3.14, 2.857 / 4 + 3.14]]])
        ids = np.append(ids, np.uint32(1))
    return np.ascontiguousarray(pos), np.array([n * a] * 3), ids


def vacancy_cell(n: int, a: float, vacancy_fraction: float, jitter: float, seed: int):
    """C4: n^3 bcc Fe with a fraction of sites removed at random and Gaussian jitter."""
    rng = np.random.default_rng(seed)
    pos = bcc_lattice(a, n)
    kill = rng.choice(len(pos), int(round(vacancy_fraction * len(pos))), replace=False)
    pos = np.delete(pos, kill, axis=0)
    pos = pos + rng.normal(0.0, jitter, pos.shape)
    return np.ascontiguousarray(pos), np.array([n * a] * 3), np.zeros(len(pos), dtype=np.uint32)


def perturb(minimum: np.ndarray, extents: np.ndarray, centre: int, rng: np.random.Generator, r_pert: float = 4.0,
            stddev: float = 0.6, frozen=None):
    """One perturbed start state (positions, unit axis) around `centre`."""
    sd = -1.0
    while sd < 0:
        sd = rng.normal(stddev, stddev / 3.0)
    d = minimum - minimum[centre]
    d -= extents * np.round(d / extents)  # minimum image (fully periodic cubic cells)
    r = np.sqrt((d * d).sum(1))
    near = np.nonzero((r < r_pert) & (np.arange(len(r)) != centre))[0]
    if frozen is not None:
        near = near[~np.asarray(frozen, dtype=bool)[near]]
    pos = minimum.copy()
    axis = np.zeros_like(minimum)
    pos[near] += rng.normal(0.0, sd, (len(near), 3)) * (1.0 - r[near] / r_pert)[:, None]
    axis[near] += rng.normal(0.0, 1.0, (len(near), 3))
    pos[centre] += rng.normal(0.0, sd, 3)
    axis[centre] += rng.normal(0.0, 1.0, 3)
    axis /= np.sqrt((axis * axis).sum())
    return pos, axis


def perturbed_batch(minimum: np.ndarray, extents: np.ndarray, centre: int, replicas: int, seed: int):
    """[R,N,3] start positions and axes; replica r uses the independent stream (seed, r)."""
    P = np.empty((replicas,) + minimum.shape)
    A = np.empty_like(P)
    for r in range(replicas):
        P[r], A[r] = perturb(minimum, extents, centre, np.random.default_rng([seed, r]))
    return P, A
