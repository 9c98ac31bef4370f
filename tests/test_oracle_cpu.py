"""CPU suite (-m "not gpu"): pins the oracle.

* oracle/ekf_oracle.c against the golden vectors produced by the reference's own source (tests/golden, made by
  scripts/make_golden.py from oracle/_ref) and — where oracle/_ref was built — against the reference run live;
* oracle/circle_oracle.py against the reference's four known-answer tests (nuslam/tests/circle_tests.cpp, values
  only) and the golden scans;
* the helper golden values of rigid2d/tests/tests.cpp:322-331 (normalize_angle).
"""
import ctypes
import os
import sys

import numpy as np
import pytest

from _oracle import OracleEKF, RefEKF, oracle_lib, ref_lib, sigma_err, state_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import circle_oracle as co  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
# oracle (O(N^2) form, libm) vs reference (dense GEMMs, OpenBLAS): observed <= 4e-11, see DESIGN.md
TOL = 1e-9


def gold(name):
    return np.load(os.path.join(GOLD, name))


def test_oracle_known_matches_reference_golden():
    g = gold("ekf_known_n20.npz")
    o = OracleEKF(20)
    ck = list(g["checkpoints"])
    worst = 0.0
    for t in range(g["twists"].shape[0]):
        o.prediction(*g["twists"][t])
        o.measurement(g["xy"][t], g["vis"][t])
        if t in ck:
            k = ck.index(t)
            worst = max(worst, state_err(o.state, g["state"][k]), sigma_err(o.sigma, g["sigma"][k]))
    assert worst < TOL, worst
    for p, (px, py) in enumerate(g["maha_probes"]):
        for i in range(20):
            want = g["maha"][p, i]
            assert abs(o.maha(px, py, i) - want) <= 1e-9 * max(1.0, abs(want))


def test_oracle_unknown_matches_reference_golden():
    g = gold("ekf_unknown_n20.npz")
    o = OracleEKF(20)
    known = np.zeros(20, np.uint8)
    ck = list(g["checkpoints"])
    worst = 0.0
    for t in range(g["twists"].shape[0]):
        o.prediction(*g["twists"][t])
        m = int(g["count"][t])
        a, dmin, sec, cr = o.data_association(g["meas"][t, :m], known)
        assert np.array_equal(known, g["known"][t])       # reference's known_list after the call
        assert np.array_equal(a, g["assoc_from_restatement"][t, :m])
        if t in ck:
            k = ck.index(t)
            worst = max(worst, state_err(o.state, g["state"][k]), sigma_err(o.sigma, g["sigma"][k]))
    assert known.sum() >= 5
    assert worst < TOL, worst


def test_oracle_n100_matches_reference_golden():
    g = gold("ekf_known_n100.npz")
    o = OracleEKF(100)
    for t in range(g["twists"].shape[0]):
        o.prediction(*g["twists"][t])
        o.measurement(g["xy"][t], g["vis"][t])
    assert state_err(o.state, g["state"]) < TOL and sigma_err(o.sigma, g["sigma"]) < TOL


@pytest.mark.skipif(ref_lib() is None, reason="oracle/_ref not built here (needs /root/reference)")
def test_oracle_matches_reference_live_long_run():
    """2,000 steps (~9k corrections): drift between the O(N^2) form and the reference's dense form stays bounded."""
    import ekf_slam_ml_b200 as pkg
    tr = pkg.tracegen.simulate_known(pkg.tracegen.default_world(20), 1, 2000, seed=77)
    r, o = RefEKF(20), OracleEKF(20)
    worst = 0.0
    for t in range(2000):
        r.prediction(*tr["twists"][t, 0])
        o.prediction(*tr["twists"][t, 0])
        r.measurement(tr["xy"][t, 0], tr["vis"][t, 0])
        o.measurement(tr["xy"][t, 0], tr["vis"][t, 0])
        if t % 250 == 0 or t == 1999:
            worst = max(worst, state_err(o.state, r.state), sigma_err(o.sigma, r.sigma))
    assert worst < TOL, worst


@pytest.mark.skipif(ref_lib() is None, reason="oracle/_ref not built here (needs /root/reference)")
def test_oracle_quirks_match_reference():
    """Reference quirks on hand-built states: stale pose in measurement(), unwrapped bearing in the
    Mahalanobis distance, strict gates, leading-true prefix of known_list, full map."""
    rng = np.random.default_rng(3)
    n = 6
    r, o = RefEKF(n), OracleEKF(n)
    A = rng.normal(size=(15, 15))
    S = A @ A.T * 1e-3 + np.eye(15) * 1e-3
    st = np.concatenate([[3.1, 0.1, -0.2], rng.uniform(-1, 1, 12)])  # theta near pi: wrap matters
    for f in (r, o):
        f.sigma = S
        f.state = st
        f.init_flag = True
    for i in range(n):
        # a landmark almost straight behind the robot: bearing near +-pi
        assert abs(r.maha(-0.5, 0.01 * (i - 3), i) - o.maha(-0.5, 0.01 * (i - 3), i)) <= 1e-9 * max(1, abs(r.maha(-0.5, 0.01 * (i - 3), i)))
    xy = rng.uniform(-1, 1, 2 * n)
    vis = np.array([1, 0, 1, 1, 0, 1], np.uint8)
    r.measurement(xy, vis)
    o.measurement(xy, vis)
    assert state_err(o.state, r.state) < TOL and sigma_err(o.sigma, r.sigma) < TOL
    # known_list with a hole: only the leading-true prefix counts (ekf_slam.cpp:281-288)
    kr = np.array([1, 1, 0, 1, 0, 0], np.uint8)
    ko = kr.copy()
    meas = rng.uniform(-1, 1, (5, 2))
    r.data_association(meas, kr)
    a, dmin, sec, cr = o.data_association(meas, ko)
    assert np.array_equal(kr, ko)
    assert state_err(o.state, r.state) < TOL and sigma_err(o.sigma, r.sigma) < TOL
    # full map: nothing can be created any more
    kr = np.ones(n, np.uint8)
    ko = kr.copy()
    far = np.array([[5.0, 5.0], [-4.0, 3.0]])
    r.data_association(far, kr)
    a, _, _, cr = o.data_association(far, ko)
    assert not cr.any() and np.array_equal(kr, ko)
    assert state_err(o.state, r.state) < TOL and sigma_err(o.sigma, r.sigma) < TOL


def test_normalize_angle_and_twist_goldens():
    L = oracle_lib()
    g = gold("helpers.npz")
    got = np.array([L.oracle_normalize_angle(float(a)) for a in g["angles"]])
    assert np.array_equal(got, g["normalized"])  # same libm fmod: bit-exact
    # rigid2d/tests/tests.cpp:322-331
    for deg, want in ((30.0, 0.523599), (230.0, -2.26893), (-330.0, 0.523599)):
        assert abs(L.oracle_normalize_angle(float(np.deg2rad(deg))) - want) < 1.2e-5 * abs(want)
    out = np.zeros(2)
    for w, t in zip(g["wheels"], g["twists"]):
        L.oracle_body_twist(ctypes.c_double(0.16), ctypes.c_double(0.033), ctypes.c_double(w[0]), ctypes.c_double(w[1]),
                            out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        assert np.array_equal(out, t)


def test_odometry_restatement_is_pinned():
    """DiffDrive::updatePose / integrateTwist: the C restatement equals the reference build bit for bit (same libm),
    incl. the acos / asin rotation of Transform2D::operator*=, and reproduces the reference's own known answers
    (rigid2d/tests/tests.cpp:285-318 integrateTwist, :334-383 updatePose; Catch Approx = relative 1.2e-5)."""
    L, R = oracle_lib(), ref_lib()
    d, P = ctypes.c_double, ctypes.POINTER(ctypes.c_double)
    rng = np.random.default_rng(5)
    for k in range(4000):
        pose = np.array([rng.normal(), rng.normal(), rng.uniform(-7, 7)])
        l, r = (rng.uniform(-0.2, 0.2), rng.uniform(-0.2, 0.2)) if k % 3 else (rng.uniform(-40, 40), rng.uniform(-40, 40))
        if k % 7 == 0:
            r = l + rng.uniform(-1e-6, 1e-6)  # below the 1e-4 rad "straight line" threshold
        a, b = pose.copy(), pose.copy()
        L.oracle_update_pose(d(0.16), d(0.033), a.ctypes.data_as(P), d(l), d(r))
        R.ref_update_pose(d(0.16), d(0.033), b.ctypes.data_as(P), d(l), d(r))
        assert np.array_equal(a, b), (pose, l, r)
    t = np.zeros(3)
    for (w, vx, vy), want in (((0.0, 1.0, 2.0), (1.0, 2.0, 0.0)), ((0.5, 0.0, 0.0), (0.0, 0.0, 0.5)),
                              ((0.5, 1.0, 2.2), (0.4202143, 2.3543072, 0.5))):
        L.oracle_integrate_twist(d(w), d(vx), d(vy), t.ctypes.data_as(P))
        assert np.allclose(t, want, rtol=1.2e-5, atol=1e-12)
    for (wb, wr, l, r), want in (((0.2, 0.01, 0.5, 0.5), (0.005, 0.0, None)), ((0.2, 0.01, -0.5, -0.5), (-0.005, 0.0, None)),
                                 ((0.2, 0.01, -15.7, 15.7), (0.0, 0.0, 1.57)),
                                 ((0.2, 0.05, 0.0, 2 * 3.1415926), (0.1, 0.1, 1.5708))):
        a = np.zeros(3)
        L.oracle_update_pose(d(wb), d(wr), a.ctypes.data_as(P), d(l), d(r))
        for got, w_ in zip(a, want):
            assert w_ is None or abs(got - w_) <= 1.2e-5 * max(abs(w_), 1e-7) + 1e-12


# ---------------------------------------------------------------- circle fitting oracle
GOLD_RANGES = [0.713136, 0.682084, 0.668864, 0.660664, 0.65551, 0.652665, 0.651814, 0.652875, 0.655952, 0.661391,
               0.670004, 0.684042, 1.01247, 1.01543, 1.01872, 1.02234, 1.0263, 1.03061, 1.04061, 1.05061, 1.06061]


def test_circle_oracle_reference_known_answers():
    """nuslam/tests/circle_tests.cpp:8-76 (Catch Approx: relative 1.2e-5)."""
    pcs, xys, ids = co.cluster_ranges(GOLD_RANGES)
    assert len(pcs) == 2 and abs(pcs[1][0] - 1.01247) < 1e-12
    cx, cy, r = co.circle_regression([(1, 7), (2, 6), (5, 8), (7, 7), (9, 5), (3, 7)])
    assert abs(cx - 4.615482) < 1.2e-5 * 4.6 and abs(cy - 2.807354) < 1.2e-5 * 2.8 and abs(r - 4.827575) < 1.2e-5 * 4.8
    cx, cy, r = co.circle_regression([(-1, 0), (-0.3, -0.06), (0.3, 0.1), (1, 0)])
    assert abs(cx - 0.4908357) < 1.2e-5 * 0.49 and abs(cy + 22.15212) < 1.2e-5 * 22.2 and abs(r - 22.17979) < 1.2e-5 * 22.2
    centres, det = co.approx_circle_positions(GOLD_RANGES)
    assert centres.shape[0] == 0 and len(det) == 2


def test_circle_oracle_matches_reference_golden_scans():
    g = gold("circles_scans.npz")
    n_c = 0
    for s in range(g["ranges"].shape[0]):
        centres, det = co.approx_circle_positions(g["ranges"][s].astype(np.float64))
        k = int(g["counts"][s])
        assert centres.shape[0] == k
        np.testing.assert_allclose(centres, g["centers"][s, :k], rtol=0, atol=1e-12)
        sizes = g["cluster_sizes"][s]
        nc = int((sizes > 0).sum())
        assert [d["size"] for d in det] == list(sizes[:nc])
        assert [d["is_circle"] for d in det] == list(g["cluster_is_circle"][s, :nc].astype(bool))
        n_c += k
    assert n_c > 100
