"""Scenario files for tests/ros_lite/harness.cpp and a runner for the node binaries that oracle/Makefile builds from
the reference's UNMODIFIED node sources (oracle/_ref/node_*).  Test infrastructure only."""
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

ROBOT_PARAMS = {"odom_frame_id": "odom", "body_frame_id": "base_footprint", "left_wheel_joint": "wheel_left_joint",
                "right_wheel_joint": "wheel_right_joint", "wheel_base": 0.16, "wheel_radius": 0.033,
                "turtle_frame_id": "turtle"}


def node_path(name):
    p = os.path.join(REF_DIR, name)
    return p if os.path.exists(p) else None


def _fmt(v):
    return repr(float(v))


class Scenario:
    def __init__(self, params):
        self.lines = [f"P {k} {v if isinstance(v, str) else (','.join(_fmt(x) for x in v) if hasattr(v, '__len__') else _fmt(v))}"
                      for k, v in params.items()]
        self.events = []

    def add(self, t, kind, values):
        self.events.append((float(t), f"E {t:.4f} {kind} " + " ".join(values)))

    def joint(self, t, lw, rw, dl, dr):
        self.add(t, "J", [_fmt(lw), _fmt(rw), _fmt(dl), _fmt(dr)])

    def fake_sensor(self, t, xs, ys, visible):
        vals = [str(len(xs))]
        for x, y, v in zip(xs, ys, visible):
            vals += [_fmt(x), _fmt(y), "0" if v else "2"]  # Marker::ADD = 0, DELETE = 2
        self.add(t, "F", vals)

    def circles(self, t, pts):
        vals = [str(len(pts))]
        for x, y in pts:
            vals += [_fmt(x), _fmt(y)]
        self.add(t, "C", vals)

    def scan(self, t, ranges):
        self.add(t, "S", [str(len(ranges))] + [repr(float(np.float32(r))) for r in ranges])

    def cmd_vel(self, t, v, w):
        self.add(t, "V", [_fmt(v), _fmt(w)])

    def write(self, path, t_end):
        self.events.sort(key=lambda e: e[0])
        with open(path, "w") as f:
            f.write("\n".join(self.lines + [e[1] for e in self.events] + [f"END {t_end:.4f}"]) + "\n")


def run_node(binary, scenario, t_end, topics=None, timeout=600):
    """Runs one node binary on a scenario; returns {topic: [(t, [fields...]), ...]} with numeric fields as float."""
    with tempfile.TemporaryDirectory() as d:
        sc, out = os.path.join(d, "scenario.txt"), os.path.join(d, "out.txt")
        scenario.write(sc, t_end)
        env = dict(os.environ, ROS_LITE_SCENARIO=sc, ROS_LITE_OUT=out)
        if topics:
            env["ROS_LITE_TOPICS"] = ",".join(topics)
        r = subprocess.run([binary], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, timeout=timeout)
        assert r.returncode == 0, f"{binary} failed ({r.returncode}): {r.stderr[-2000:]}"
        res = {}
        with open(out) as f:
            for line in f:
                parts = line.split()
                fields = []
                for p in parts[2:]:
                    try:
                        fields.append(float(p))
                    except ValueError:
                        fields.append(p)
                res.setdefault(parts[1], []).append((float(parts[0]), fields))
        return res, r.stderr


def compare_streams(a, b, topic, tol, skip_text=True):
    """Two recordings of one topic: same number of messages, same stamps, numeric fields within tol (absolute, scaled by
    max(1, |value|)), text fields identical.  Returns the worst deviation."""
    A, B = a.get(topic, []), b.get(topic, [])
    assert len(A) == len(B) and len(A) > 0, f"{topic}: {len(A)} vs {len(B)} messages"
    worst = 0.0
    for (ta, fa), (tb, fb) in zip(A, B):
        assert ta == tb and len(fa) == len(fb), f"{topic} at t={ta}: shape differs ({len(fa)} vs {len(fb)} fields)"
        for x, y in zip(fa, fb):
            if isinstance(x, str) or isinstance(y, str):
                assert x == y, f"{topic} at t={ta}: {x} != {y}"
            elif np.isnan(x) or np.isnan(y):
                assert np.isnan(x) and np.isnan(y), f"{topic} at t={ta}: NaN on one side only"
            else:
                worst = max(worst, abs(x - y) / max(1.0, abs(y)))
    assert worst <= tol, f"{topic}: worst deviation {worst:.3e} > {tol}"
    return worst
