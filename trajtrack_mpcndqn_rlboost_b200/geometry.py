"""Small geometry helpers on the host side of the path.

``polygon_halfspace_representation`` returns the same H-representation the
reference computes with scipy's ConvexHull
(/root/reference/src/util/utils_geo.py:33-59): rows (b_i, a0_i, a1_i) with
b_i - a0_i x - a1_i y > 0 inside the polygon, each row scaled so that
a_i . (v - centre) = 1 on its edge.  Implemented with a monotone-chain hull so
the path does not depend on scipy at run time.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np


def convex_hull(points: np.ndarray) -> np.ndarray:
    """Counter-clockwise convex hull (Andrew's monotone chain) of an [n,2] array."""
    pts = sorted(set(map(tuple, np.asarray(points, dtype=np.float64).tolist())))
    if len(pts) <= 2:
        return np.array(pts)

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])
    lower: list = []
    for p in pts:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], p) <= 0:
            lower.pop()
        lower.append(p)
    upper: list = []
    for p in reversed(pts):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], p) <= 0:
            upper.pop()
        upper.append(p)
    return np.array(lower[:-1] + upper[:-1])


def polygon_halfspace_representation(polygon_points: np.ndarray) -> Tuple[List[float], List[float], List[float]]:
    hull = convex_hull(polygon_points)
    centre = hull.mean(axis=0)
    V = hull - centre
    b, a0, a1 = [], [], []
    m = len(hull)
    for i in range(m):
        F = np.array([V[i], V[(i + 1) % m]])
        if np.linalg.matrix_rank(F) == 2:
            a = np.linalg.solve(F, np.ones(2))
            a0.append(float(a[0]))
            a1.append(float(a[1]))
            b.append(float(a @ centre + 1.0))
    return b, a0, a1


def pad_polygon_round(nodes: np.ndarray, radius: float, resolution: int = 4) -> np.ndarray:
    """Round-join outward offset of a convex polygon (what shapely's
    ``buffer(radius, join_style=round, resolution=4)`` yields for the reference's
    obstacles, obstacle.py:158-163): each corner becomes an arc sampled every
    90/resolution degrees.  Returns the CCW ring."""
    hull = convex_hull(nodes)
    m = len(hull)
    out = []
    for i in range(m):
        p_prev, p, p_next = hull[i - 1], hull[i], hull[(i + 1) % m]
        a_in = math.atan2(p[1] - p_prev[1], p[0] - p_prev[0]) - math.pi / 2
        a_out = math.atan2(p_next[1] - p[1], p_next[0] - p[0]) - math.pi / 2
        while a_out < a_in:
            a_out += 2 * math.pi
        steps = max(1, int(math.ceil((a_out - a_in) / (math.pi / 2 / resolution) - 1e-9)))
        for s in range(steps + 1):
            a = a_in + (a_out - a_in) * s / steps
            out.append((p[0] + radius * math.cos(a), p[1] + radius * math.sin(a)))
    return np.array(out, dtype=np.float64)
