"""TEST INFRASTRUCTURE ONLY (oracle/): numpy restatement of rigid2d::CircleFitting
(/root/reference/rigid2d/src/circle_fitting.cpp), with the same LAPACK drivers the reference reaches through
Armadillo: svd -> dgesdd (np.linalg.svd), eig_gen -> dgeev (np.linalg.eig), solve -> dgesv (np.linalg.solve).

Pinned against the reference's own four known-answer tests (nuslam/tests/circle_tests.cpp:8-76) in
tests/test_circle_oracle.py, and against oracle/_ref (the reference source itself) where that was built.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may import this module.
"""
import math

import numpy as np

PI = 3.14159265358979323846  # rigid2d.hpp:13
THRESH = 0.2                 # circle_fitting.cpp:17


def normalize_angle(rad):
    """rigid2d.cpp:336-345"""
    a = math.fmod(rad, 2 * PI)
    a = math.fmod(a + 2 * PI, 2 * PI)
    if a > PI:
        a -= 2 * PI
    return a


def cluster_ranges(ranges):
    """circle_fitting.cpp:11-90.  Returns (point_cluster, xy_cluster, index_cluster): lists of per-cluster
    range lists, [x, y] lists and beam-index lists.  The reference indexes point_cluster[0] unconditionally
    (:54, undefined behaviour when nothing survives); here that case returns no clusters."""
    ranges = [float(r) for r in ranges]
    n = len(ranges)
    res = 2 * PI / float(n)
    pcs, xys, ids = [], [], []
    cur, cur_xy, cur_id = [ranges[0]], [[ranges[0] * math.cos(0.0), ranges[0] * math.sin(0.0)]], [0]
    for i in range(1, n):
        if (abs(ranges[i] - ranges[i - 1]) < THRESH) and (i != n - 1):
            pass
        else:
            if len(cur) > 6:
                pcs.append(cur)
                xys.append(cur_xy)
                ids.append(cur_id)
            cur, cur_xy, cur_id = [], [], []
        cur.append(ranges[i])
        ang = normalize_angle(i * res)
        cur_xy.append([ranges[i] * math.cos(ang), ranges[i] * math.sin(ang)])
        cur_id.append(i)
    if not pcs:
        return [], [], []
    first = pcs[0][0]
    last = pcs[-1][-1]
    if abs(first - last) < THRESH:  # :62-70 wrap merge (a single cluster merges with itself and is popped)
        lp, lx, li = list(pcs[-1]), list(xys[-1]), list(ids[-1])
        pcs[0] = lp + pcs[0]
        xys[0] = lx + xys[0]
        ids[0] = li + ids[0]
        pcs.pop()
        xys.pop()
        ids.pop()
    return pcs, xys, ids


def circle_regression(xy_cluster):
    """circle_fitting.cpp:104-232 for one cluster ([n,2] array-like).  Returns (cx, cy, r)."""
    p = np.asarray(xy_cluster, dtype=np.float64)
    n = p.shape[0]
    x_sum = y_sum = 0.0
    for k in range(n):
        x_sum += p[k, 0]
        y_sum += p[k, 1]
    x_mean, y_mean = x_sum / n, y_sum / n
    x = p[:, 0] - x_mean
    y = p[:, 1] - y_mean
    z = x * x + y * y
    z_sum = 0.0
    for k in range(n):
        z_sum += z[k]
    z_mean = z_sum / n
    Z = np.stack([z, x, y, np.ones(n)], axis=1)
    H_inv = np.zeros((4, 4))
    H_inv[0, 3] = 0.5
    H_inv[1, 1] = 1.0
    H_inv[2, 2] = 1.0
    H_inv[3, 0] = 0.5
    H_inv[3, 3] = -2.0 * z_mean
    U, s, Vt = np.linalg.svd(Z, full_matrices=True)
    V = Vt.T
    if s[3] < 1e-12:
        A = V[:, 3]
    else:
        Y = V @ np.diag(s) @ V.T
        Q = Y @ H_inv @ Y
        w, vec = np.linalg.eig(Q)
        idx, small = 0, 1000.0
        for i in range(4):
            if w[i].real > 0 and w[i].real < small:
                small = w[i].real
                idx = i
        A_star = vec[:, idx].real
        A = np.linalg.solve(Y, A_star)
    A1, A2, A3, A4 = A
    a = -A2 / (2 * A1)
    b = -A3 / (2 * A1)
    R_sqr = (A2 * A2 + A3 * A3 - 4 * A1 * A4) / (4 * A1 * A1)
    return a + x_mean, b + y_mean, math.sqrt(R_sqr) if R_sqr >= 0 else float("nan")


def mean_inscribed_angle(xy_cluster):
    """circle_fitting.cpp:242-263"""
    p = np.asarray(xy_cluster, dtype=np.float64)
    n = p.shape[0]
    p1, p2 = p[0], p[n - 1]
    s = 0.0
    for k in range(1, n - 1):
        a = p1 - p[k]
        b = p2 - p[k]
        top = a[0] * b[0] + a[1] * b[1]
        bot = math.sqrt(a[0] * a[0] + a[1] * a[1]) * math.sqrt(b[0] * b[0] + b[1] * b[1])
        with np.errstate(all="ignore"):
            s += float(np.arccos(np.float64(top) / np.float64(bot)))
    return s / (n - 2) if n != 2 else float("nan")


def classify(xy_cluster, radius):
    """circle_fitting.cpp:264-275"""
    m = mean_inscribed_angle(xy_cluster)
    return bool(m > 1.5708 and m < 2.3562 and radius < 0.2)


def approx_circle_positions(ranges):
    """circle_fitting.cpp:298-304.  Returns (centres [k,2], details) with details = list of per-cluster dicts."""
    pcs, xys, ids = cluster_ranges(ranges)
    out, details = [], []
    for c in range(len(xys)):
        cx, cy, r = circle_regression(xys[c])
        ok = classify(xys[c], r)
        details.append({"size": len(xys[c]), "ids": ids[c], "cx": cx, "cy": cy, "r": r, "is_circle": ok,
                        "mean_angle": mean_inscribed_angle(xys[c])})
        if ok:
            out.append([cx, cy])
    return np.array(out, dtype=np.float64).reshape(-1, 2), details
