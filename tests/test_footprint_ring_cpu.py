"""CPU: the geometric argument behind k_footprints' circle parts (img_env_b200/csrc/host_tables.h lattice_circle_ring,
foot.cuh), restated in numpy.  The reference covers a cell iff one of the 0.01 m lattice points of the disc (agent.cpp:18-62)
rounds into it.  Claim: (a) every cell whose centre lies within r_in = r - 0.01*sqrt(2)/2 - eps of the disc centre is covered
(the lattice's covering radius is smaller than half a cell), and (b) every point that falls into any OTHER cell lies farther
than r_in - res*sqrt(2)/2 from the centre -- so "fill the sure interior + evaluate the rim points" gives the same cell set as
evaluating the whole lattice.  The GPU suite checks the kernel itself word for word (imgenv_debug_check_footprints)."""
import math

import numpy as np
import pytest

COVER = 0.01 * 0.70710678118654757


def _lattice(r):
    bb = int(math.ceil(r / 0.01))
    m, n = np.meshgrid(np.arange(-bb, bb + 1), np.arange(-bb, bb + 1), indexing="ij")
    rad = np.sqrt(m * 0.01 * m * 0.01 + n * 0.01 * n * 0.01)
    keep = rad <= r
    return m[keep] * 0.01, n[keep] * 0.01, rad[keep]


def _cells(px, py, x, y, yaw, res):
    c, s = math.cos(yaw), math.sin(yaw)
    wx, wy = c * px - s * py + x, s * px + c * py + y
    return set(zip(np.floor(wx / res + 0.5).astype(int).tolist(), np.floor(wy / res + 0.5).astype(int).tolist()))


@pytest.mark.parametrize("r", [0.07, 0.1, 0.17, 0.25, 0.45])
@pytest.mark.parametrize("res", [0.015, 0.02, 0.025])
def test_interior_fill_plus_rim_points_equals_whole_lattice(r, res):
    assert res * 0.5 - COVER >= 2e-4
    r_in = r - COVER - 5e-5
    thr = r_in - res * 0.70710678118654757 - 5e-5
    px, py, rad = _lattice(r)
    rim = rad > thr
    assert rim.sum() < len(rad), "the rim must be a proper subset for the shortcut to pay"
    rng = np.random.default_rng(int(r * 1000) + int(res * 1e4))
    for _ in range(120):
        x, y, yaw = rng.uniform(1.0, 9.0), rng.uniform(1.0, 9.0), rng.uniform(-math.pi, math.pi)
        s0, s1 = rng.uniform(-0.05, 0.05, 2)              # disc centre in the part's frame (sizes_[0], sizes_[1])
        whole = _cells(px + s0, py + s1, x, y, yaw, res)
        c, s = math.cos(yaw), math.sin(yaw)
        wcx, wcy = c * s0 - s * s1 + x, s * s0 + c * s1 + y
        fill = set()
        for X in range(int(math.floor((wcx - r) / res)) - 1, int(math.ceil((wcx + r) / res)) + 2):
            h2 = r_in * r_in - (X * res - wcx) ** 2
            if h2 <= 0:
                continue
            half = math.sqrt(h2)
            y_lo, y_hi = int(math.ceil((wcy - half) / res + 1e-9)), int(math.floor((wcy + half) / res - 1e-9))
            fill |= {(X, Y) for Y in range(y_lo, y_hi + 1)}
        assert fill <= whole, "a sure-interior cell holds no lattice point"
        assert whole == fill | _cells(px[rim] + s0, py[rim] + s1, x, y, yaw, res), "a cell outside the interior is reached by a non-rim point only"
