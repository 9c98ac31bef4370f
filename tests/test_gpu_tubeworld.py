"""the on-device input generator against its numpy oracle (tracegen.py, which restates
nurtlesim/src/tube_world.cpp).  The hash RNG is integer-exact on both sides; sin/cos/log/atan2 differ in the last
bits, so continuous outputs are compared at 1e-9 and the discrete ones (visibility) away from the threshold."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_known_association_inputs_match_tracegen(gpu_pkg):
    tg = gpu_pkg.tracegen
    w = tg.dense_world(20)
    B, T, first = 64, 30, 1000
    ref = tg.simulate_known(w, B, T, seed=12, first_filter=first)
    sim = gpu_pkg.TubeWorld(w, B, seed=12, first_filter=first)
    for t in range(T):
        sim.step_known()
        d = sim.download()
        np.testing.assert_allclose(d["twists"], ref["twists"][t], rtol=0, atol=1e-12)
        np.testing.assert_allclose(d["truth"], ref["truth"][t], rtol=0, atol=1e-9)
        np.testing.assert_allclose(d["odom"], ref["odom"][t], rtol=0, atol=1e-9)
        np.testing.assert_allclose(d["xy"], ref["xy"][t], rtol=0, atol=1e-9)
        dist = np.hypot(w.tubes_x[None, :] - ref["truth"][t][:, 0:1], w.tubes_y[None, :] - ref["truth"][t][:, 1:2])
        clear = np.abs(dist - w.max_visible) > 1e-6
        assert np.array_equal(d["vis"][:, :20][clear], ref["vis"][t][:, :20][clear])
    assert ref["vis"][1:].sum() > 1000
    # The odometer integrates the truth's own wheel increments: it equals the truth until the first collision, and
    # the reference's DiffDrive::updatePose (device twin, acos / asin composition) follows the same path.
    free = np.all(np.abs(ref["odom"][-1] - ref["truth"][-1]) < 1e-12, axis=1)
    assert free.any()
    sim1 = tg.TubeWorldSim(w, 4, seed=12, first_filter=first)
    poses = np.zeros((4, 3))
    for _ in range(60):
        sim1.step_tick()
        poses = gpu_pkg.update_pose(w.wheel_base, w.wheel_radius, poses, sim1.dl, sim1.dr)
    np.testing.assert_allclose(poses, np.column_stack([sim1.ox, sim1.oy, sim1.oth]), rtol=0, atol=1e-9)


def test_batch_independence_and_scan(gpu_pkg):
    tg = gpu_pkg.tracegen
    w = tg.default_world(20)
    a = gpu_pkg.TubeWorld(w, 8, seed=5, first_filter=0)
    b = gpu_pkg.TubeWorld(w, 2, seed=5, first_filter=3)
    for _ in range(4):
        a.step_known()
        b.step_known()
    da, db = a.download(), b.download()
    for k in ("twists", "xy", "truth", "vis"):
        assert np.array_equal(da[k][3:5], db[k])
    # laser: same tick schedule as tracegen.simulate_scans (a scan every 21st tick)
    host = tg.TubeWorldSim(w, 6, seed=9)
    dev = gpu_pkg.TubeWorld(w, 6, seed=9)
    for _ in range(3):
        for _ in range(21):
            host.step_tick()
        want = host.laser_scan(360)
        dev.step_scan(21, 360)
        got = dev.download(ranges=True)["ranges"]
        np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)  # float32 output


def test_sim_to_filter_stays_on_device(gpu_pkg):
    """Generator -> fused EKF step with no host copy in between, against the host-fed path."""
    tg = gpu_pkg.tracegen
    w = tg.dense_world(20)
    B, T = 256, 12
    sim = gpu_pkg.TubeWorld(w, B, seed=3)
    bt = gpu_pkg.EKFBatch(B, 20)
    sim.use_stream(bt.stream)
    p = sim.device_pointers()
    for _ in range(T):
        sim.step_known()
        bt.step_known_dev(p["twists"], p["xy"], p["vis"])
    bt.sync()
    ref = tg.simulate_known(w, B, T, seed=3)
    bh = gpu_pkg.EKFBatch(B, 20)
    for t in range(T):
        bh.step_known(np.ascontiguousarray(ref["twists"][t]), np.ascontiguousarray(ref["xy"][t]),
                      np.ascontiguousarray(ref["vis"][t]))
        bh.sync()
    assert bt.update_count == bh.update_count > 0
    np.testing.assert_allclose(bt.states(), bh.states(), rtol=0, atol=1e-7)
    err = bt.pose_error(sim.download()["truth"])
    assert np.all(np.sqrt(err[:3] / B) < 0.1)


def test_accuracy_report(gpu_pkg):
    """the README's Actual / Odom / Slam comparison over a batch of seeds."""
    r = gpu_pkg.report.run(gpu_pkg, robots=512, steps=40, seed=4)
    e = r["rmse_xytheta"]
    assert r["updates"] > 512 * 40 and np.isfinite(r["landmark_rmse"]) and r["landmark_rmse"] < 0.05
    assert e["slam"][0] < 0.05 and e["slam"][1] < 0.05
    assert e["slam"][1] < e["prediction_only"][1]            # the filter beats dead reckoning on the 10 Hz twist
    assert e["wheel_odometry"][0] < 1e-9                     # the 100 Hz odometer sees the truth's own increments
    assert "Slam" in gpu_pkg.report.markdown(r)
