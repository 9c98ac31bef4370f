"""Monte-Carlo accuracy report: the reference README's "Actual / Odom / Slam" comparison
(nuslam/README.md:98-114, one run on one seed) over a whole batch of seeds, entirely on the GPU:
on-device simulator (tube_world restatement) -> batched EKF step, plus a prediction-only batch fed the same twists
(what the motion model alone makes of the 10 Hz odometer twist) and the 100 Hz wheel odometer of the node.

    python -m ekf_slam_ml_b200.report --robots 4096 --steps 140 [--json]

(`python ekf-slam-ml_b200/report.py ...` works too.)  SURVEY.md §8(f)-4.
"""
import argparse
import json
import os
import sys

import numpy as np


def _rmse(est, truth):
    d = est[:, :2] - truth[:, :2]
    dth = np.arctan2(np.sin(est[:, 2] - truth[:, 2]), np.cos(est[:, 2] - truth[:, 2]))
    return [float(np.sqrt(np.mean(d[:, 0] ** 2))), float(np.sqrt(np.mean(d[:, 1] ** 2))),
            float(np.sqrt(np.mean(dth ** 2)))]


def run(pkg, robots=4096, steps=140, seed=0, device=0, world=None):
    """Returns a dict with the batch statistics after `steps` SLAM steps (11 simulator ticks each; one lap of the
    follow_circle trajectory is about 114 steps)."""
    import torch
    tg = pkg.tracegen
    world = world or tg.dense_world(20)
    n = world.n_slots
    sim = pkg.TubeWorld(world, robots, seed=seed, device=device)
    slam = pkg.EKFBatch(robots, n, device=device)
    blind = pkg.EKFBatch(robots, n, device=device)
    sim.use_stream(slam.stream)
    p = sim.device_pointers()
    no_vis = torch.zeros(robots * n, dtype=torch.uint8, device=f"cuda:{device}")
    torch.cuda.synchronize()
    for _ in range(steps):
        sim.step_known()
        slam.step_known_dev(p["twists"], p["xy"], p["vis"])
        slam.sync()  # the prediction-only batch runs on its own stream: order it after this step's inputs
        blind.step_known_dev(p["twists"], p["xy"], no_vis.data_ptr())
        blind.sync()
    d = sim.download()
    truth, odom = d["truth"], d["odom"]
    est, pred = slam.poses()[:, [1, 2, 0]], blind.poses()[:, [1, 2, 0]]  # (theta, x, y) -> (x, y, theta)
    lm = slam.states()[:, 3:].reshape(robots, n, 2)
    nt = min(world.n_tubes, n)  # every tube's slot is initialised by the first measurement() call (ekf_slam.cpp:113-128)
    tubes = np.stack([world.tubes_x[:nt], world.tubes_y[:nt]], axis=1)[None]
    lm_err = np.linalg.norm(lm[:, :nt] - tubes, axis=2)
    out = {
        "robots": int(robots), "steps": int(steps), "seed": int(seed), "updates": int(slam.update_count),
        "actual_mean_xy": [float(truth[:, 0].mean()), float(truth[:, 1].mean())],
        "rmse_xytheta": {"slam": _rmse(est, truth), "prediction_only": _rmse(pred, truth), "wheel_odometry": _rmse(odom, truth)},
        "worst_xy_error": {"slam": float(np.abs(est[:, :2] - truth[:, :2]).max()),
                           "prediction_only": float(np.abs(pred[:, :2] - truth[:, :2]).max()),
                           "wheel_odometry": float(np.abs(odom[:, :2] - truth[:, :2]).max())},
        "landmark_rmse": float(np.sqrt(np.mean(lm_err ** 2))),
        "landmarks": int(nt),
    }
    for h in (sim, slam, blind):
        h.close()
    return out


def markdown(r):
    rows = ["Path type | x RMSE (m) | y RMSE (m) | theta RMSE (rad) | worst |x|,|y| error (m)", "--- | --- | --- | --- | ---"]
    for key, name in (("wheel_odometry", "Odom (100 Hz wheel integration)"), ("prediction_only", "Prediction only (10 Hz twist)"),
                      ("slam", "Slam")):
        e = r["rmse_xytheta"][key]
        rows.append(f"{name} | {e[0]:.5f} | {e[1]:.5f} | {e[2]:.5f} | {r['worst_xy_error'][key]:.5f}")
    head = (f"{r['robots']} robots x {r['steps']} SLAM steps (seed {r['seed']}), {r['updates']} landmark corrections; "
            f"landmark RMSE {r['landmark_rmse']:.5f} m over {r['landmarks']} landmarks per robot")
    return head + "\n\n" + "\n".join(rows)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--robots", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=140)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--json", action="store_true")
    a = ap.parse_args(argv)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import ekf_slam_ml_b200 as pkg
    r = run(pkg, a.robots, a.steps, a.seed, a.device)
    print(json.dumps(r) if a.json else markdown(r))


if __name__ == "__main__":
    main()
