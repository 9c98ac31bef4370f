"""Deterministic "nurtlesim-shaped" input generator (host side, numpy, vectorised over filters).

Restates the parts of the reference simulator that define the INPUT DISTRIBUTION of the EKF hot path
(it is not on the parity path itself):

* command noise per received cmd_vel              nurtlesim/src/tube_world.cpp:193-210
* wheel slip, truth kinematics, collision         tube_world.cpp:214-250, 316-366
* fake relative-position sensor + visibility      tube_world.cpp:369-414
* 360-beam ray-cast laser with box walls + tubes  tube_world.cpp:423-577
* tick schedule (100 Hz; sensor every 11th tick, scan every 21st)   tube_world.cpp:583-600
* odometry twist handed to the filter = getBodyTwistForUpdate(10 dL, 10 dR)  nuslam/src/slam.cpp:173-176
* the reading vector keeps zeros in slots >= number of tubes, and visibility is only filled from the
  second sensor message on                         nuslam/src/slam.cpp:259, 305-333
* trajectory = follow_circle (v = 0.10 m/s, R = 0.20 m, commands at 10 Hz)  nuturtle_robot/src/follow_circle.cpp:39-87

Differences from the reference, on purpose: std::random_device is replaced by a counter-based hash RNG
keyed on (seed, filter, tick, purpose) so that filter b's trace does not depend on the batch size, and
the SLAM step is taken at the sensor tick itself (the reference's 10 Hz timer adds up to 0.1 s of lag).
"""
from dataclasses import dataclass, field

import numpy as np

PI = 3.14159265358979323846  # rigid2d.hpp:13
TWO_PI = 2 * PI

_M1 = np.uint64(0x9E3779B97F4A7C15)
_M2 = np.uint64(0xBF58476D1CE4E5B9)
_M3 = np.uint64(0x94D049BB133111EB)


def _mix(z):
    z = z.astype(np.uint64, copy=True)
    z ^= z >> np.uint64(30)
    z *= _M2
    z ^= z >> np.uint64(27)
    z *= _M3
    z ^= z >> np.uint64(31)
    return z


def hash_uniform(seed, filt, counter):
    """U(0,1) (never exactly 0) from (seed, filter index array, counter array)."""
    with np.errstate(over="ignore"):
        f = np.asarray(filt, dtype=np.uint64)
        c = np.asarray(counter, dtype=np.uint64)
        z = np.uint64(seed) * _M1 + f * _M2 + c * _M3 + np.uint64(0x1234567)
        z = _mix(_mix(z) + _M1)
    return ((z >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def hash_normal(seed, filt, counter):
    """N(0,1) by Box-Muller on two hashed uniforms (counter, counter + 2^40)."""
    u1 = hash_uniform(seed, filt, counter)
    u2 = hash_uniform(seed, filt, np.asarray(counter, dtype=np.uint64) + np.uint64(1 << 40))
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(TWO_PI * u2)


def normalize_angle(a):
    """rigid2d::normalize_angle (rigid2d.cpp:336-345), vectorised."""
    r = np.fmod(a, TWO_PI)
    r = np.fmod(r + TWO_PI, TWO_PI)
    return np.where(r > PI, r - TWO_PI, r)


@dataclass
class World:
    tubes_x: np.ndarray
    tubes_y: np.ndarray
    tube_radius: float = 0.0762          # nurtlesim/config/tube_param.yaml
    border: float = 2.0
    wheel_base: float = 0.16             # nuturtle_description/config/diff_params.yaml
    wheel_radius: float = 0.033
    vx_std: float = 0.01                 # nurtlesim/config/noise_param.yaml
    the_std: float = 0.01
    slip_min: float = 0.90
    slip_max: float = 1.10
    sensor_std: float = 0.005
    max_visible: float = 0.7
    range_std: float = 0.005
    cmd_v: float = 0.10                  # follow_circle speed / radius (odom_teleop.launch)
    cmd_radius: float = 0.20
    n_slots: int = 20                    # nuslam/src/slam.cpp:250
    extra: dict = field(default_factory=dict)

    @property
    def n_tubes(self):
        return len(self.tubes_x)


def default_world(n_slots=20):
    """The reference's default 10-tube world (tube_param.yaml:2-3)."""
    return World(
        tubes_x=np.array([0.5, 0.7, 0.7, 0.2, 0.6, -0.3, -0.7, -0.3, -0.7, 0.0]),
        tubes_y=np.array([0.1, 0.7, 0.4, -0.3, -0.8, -0.6, -0.2, 0.5, 0.7, 1.1]),
        n_slots=n_slots)


def dense_world(n_slots=20, seed=7):
    """Default world + seeded extra tubes so that all `n_slots` slots are exercised:
    rejection-sampled inside the box, >= 0.3 m from the circular path and >= 0.2 m from each other."""
    w = default_world(n_slots)
    xs, ys = list(w.tubes_x), list(w.tubes_y)
    rng = np.random.default_rng(seed)
    cx, cy, R = 0.0, w.cmd_radius, w.cmd_radius
    while len(xs) < n_slots:
        x, y = rng.uniform(-0.9, 0.9, 2)
        if abs(np.hypot(x - cx, y - cy) - R) < 0.3:
            continue
        if min(np.hypot(x - np.array(xs), y - np.array(ys))) < 0.2:
            continue
        xs.append(float(x))
        ys.append(float(y))
    w.tubes_x, w.tubes_y = np.array(xs), np.array(ys)
    return w


def grid_world(nx, ny, pitch=0.5, n_slots=None, max_visible=0.7):
    """nx*ny landmarks on a grid; the robot circles inside one cell cluster."""
    gx, gy = np.meshgrid((np.arange(nx) - (nx - 1) / 2) * pitch + 0.13, (np.arange(ny) - (ny - 1) / 2) * pitch + 0.29)
    w = default_world(n_slots or nx * ny)
    w.tubes_x, w.tubes_y = gx.ravel().copy(), gy.ravel().copy()
    w.border = max(nx, ny) * pitch + 2.0
    w.max_visible = max_visible
    return w


class TubeWorldSim:
    """B independent robots in the same world; advance tick by tick (10 ms)."""

    def __init__(self, world: World, B: int, seed: int = 0, first_filter: int = 0):
        self.w = world
        self.B = int(B)
        self.seed = int(seed)
        self.fid = np.arange(first_filter, first_filter + B, dtype=np.uint64)
        self.x = np.zeros(B)
        self.y = np.zeros(B)
        self.th = np.zeros(B)
        # dead-reckoning odometer (slam.cpp:96): integrates the same wheel increments, never sees a collision
        self.ox = np.zeros(B)
        self.oy = np.zeros(B)
        self.oth = np.zeros(B)
        self.tick = 0
        self.v = np.zeros(B)
        self.om = np.zeros(B)
        self.dl = np.zeros(B)
        self.dr = np.zeros(B)

    def _ctr(self, purpose, sub=0):
        return np.uint64(self.tick) * np.uint64(4096) + np.uint64(purpose * 512 + sub)

    def step_tick(self):
        w = self.w
        if self.tick % 10 == 0:  # a new (noisy) command arrives at 10 Hz, tube_world.cpp:193-210
            self.v = w.cmd_v + w.vx_std * hash_normal(self.seed, self.fid, self._ctr(0))
            self.om = w.cmd_v / w.cmd_radius + w.the_std * hash_normal(self.seed, self.fid, self._ctr(1))
        D, r = w.wheel_base * 0.5, w.wheel_radius
        wl = -(D / r) * self.om + (1.0 / r) * self.v  # diff_drive.cpp:24-36
        wr = (D / r) * self.om + (1.0 / r) * self.v
        span = w.slip_max - w.slip_min
        dl = (wl / 100.0) * (w.slip_min + span * hash_uniform(self.seed, self.fid, self._ctr(2)))
        dr = (wr / 100.0) * (w.slip_min + span * hash_uniform(self.seed, self.fid, self._ctr(3)))
        self.dl, self.dr = dl, dr
        # DiffDrive::updatePose, diff_drive.cpp:50-67 + integrateTwist rigid2d.cpp:304-333
        om_b = (r / (2.0 * D)) * (dr - dl)
        vx_b = (r / 2.0) * (dr + dl)
        turning = np.abs(om_b) > 0.0001
        safe = np.where(turning, om_b, 1.0)
        bx = np.where(turning, (vx_b / safe) * np.sin(safe), vx_b)
        by = np.where(turning, (vx_b / safe) * (1.0 - np.cos(safe)), 0.0)
        dth = np.where(turning, om_b, 0.0)
        c, s = np.cos(self.th), np.sin(self.th)
        self.x = self.x + c * bx - s * by
        self.y = self.y + s * bx + c * by
        self.th = self.th + dth
        co, so = np.cos(self.oth), np.sin(self.oth)
        self.ox = self.ox + co * bx - so * by
        self.oy = self.oy + so * bx + co * by
        self.oth = self.oth + dth
        # collision: first tube closer than radius + wheel_base/2 snaps the robot back, :316-366
        lim = w.tube_radius + w.wheel_base / 2
        dx = w.tubes_x[None, :] - self.x[:, None]
        dy = w.tubes_y[None, :] - self.y[:, None]
        hit = np.hypot(dx, dy) < lim
        if hit.any():
            rows = np.where(hit.any(axis=1))[0]
            first = hit[rows].argmax(axis=1)
            tx, ty = w.tubes_x[first], w.tubes_y[first]
            ang = np.arctan2(ty - self.y[rows], tx - self.x[rows])
            self.x[rows] = tx - lim * np.cos(ang)
            self.y[rows] = ty - lim * np.sin(ang)
        self.tick += 1

    def odom_twist(self):
        """(dtheta, dx) the SLAM node passes to prediction(): slam.cpp:173-176 + diff_drive.cpp:38-47."""
        D, r = self.w.wheel_base * 0.5, self.w.wheel_radius
        l, rr = self.dl * 10.0, self.dr * 10.0
        return (r / (2.0 * D)) * (rr - l), (r / 2.0) * (rr + l)

    def fake_sensor(self):
        """Relative tube positions + noise, and the visibility flags: tube_world.cpp:369-414."""
        w = self.w
        dx = w.tubes_x[None, :] - self.x[:, None]
        dy = w.tubes_y[None, :] - self.y[:, None]
        c, s = np.cos(self.th)[:, None], np.sin(self.th)[:, None]
        rx = c * dx + s * dy
        ry = -s * dx + c * dy
        j = np.arange(w.n_tubes, dtype=np.uint64)[None, :]
        nx = hash_normal(self.seed, self.fid[:, None], self._ctr(4) + j * np.uint64(2))
        ny = hash_normal(self.seed, self.fid[:, None], self._ctr(4) + j * np.uint64(2) + np.uint64(1))
        vis = np.hypot(dx, dy) <= w.max_visible
        return rx + w.sensor_std * nx, ry + w.sensor_std * ny, vis

    def laser_scan(self, n_beams=360):
        """float32 ranges [B, n_beams]: tube_world.cpp:454-577 (box walls + tubes + range noise)."""
        w = self.w
        B = self.B
        res = TWO_PI / n_beams
        i = np.arange(n_beams)
        curr = normalize_angle(res * i)[None, :]                       # beam angle in the robot frame
        x_dis = w.border / 2.0 - self.x[:, None]
        y_dis = w.border / 2.0 - self.y[:, None]
        box_ang = normalize_angle(res * i[None, :] + self.th[:, None])
        y_t = np.where(box_ang < 0, -(w.border - y_dis), y_dis)
        x_t = np.where((box_ang > PI / 2.0) | (box_ang < -PI / 2.0), -(w.border - x_dis), x_dis)
        with np.errstate(divide="ignore", invalid="ignore"):
            r = np.minimum(x_t / np.cos(box_ang), y_t / np.sin(box_ang))
        min_r = np.minimum(r, 3.5)
        # tubes, in the robot frame
        dx = w.tubes_x[None, :] - self.x[:, None]
        dy = w.tubes_y[None, :] - self.y[:, None]
        c, s = np.cos(self.th)[:, None], np.sin(self.th)[:, None]
        tx = c * dx + s * dy
        ty = -s * dx + c * dy
        half = np.arctan2(w.tube_radius, 0.12)                          # largest_tube_scan_theta / 2
        bearing = np.arctan2(ty, tx)
        sb = normalize_angle(bearing - half)[:, :, None]                # [B, T, 1]
        eb = normalize_angle(bearing + half)[:, :, None]
        ca = curr[:, None, :]                                           # [1, 1, n_beams]
        inside = (ca > sb) & (ca < eb)
        wrap = (sb > 0) & (eb < 0)
        flag = np.where(wrap, (ca > sb) | (ca < eb), inside)
        # ray from the robot (origin) towards (3.5 cos, 3.5 sin), circle centred at the tube:
        # getLineCircleIntersection in the tube frame (tube_world.cpp:423-451)
        x1, y1 = -tx[:, :, None], -ty[:, :, None]
        x2 = 3.5 * np.cos(ca) + x1
        y2 = 3.5 * np.sin(ca) + y1
        ddx, ddy = x2 - x1, y2 - y1
        dr2 = ddx * ddx + ddy * ddy
        Dd = x1 * y2 - x2 * y1
        delta = w.tube_radius ** 2 * dr2 - Dd * Dd
        ok = flag & (delta > 0)
        sq = np.sqrt(np.where(ok, delta, 0.0))
        sgn = np.where(ddy < 0, -1.0, 1.0)
        ix1 = (Dd * ddy + sgn * ddx * sq) / dr2
        iy1 = (-Dd * ddx + np.abs(ddy) * sq) / dr2
        ix2 = (Dd * ddy - sgn * ddx * sq) / dr2
        iy2 = (-Dd * ddx - np.abs(ddy) * sq) / dr2
        d1 = np.hypot(x1 - ix1, y1 - iy1)
        d2 = np.hypot(x1 - ix2, y1 - iy2)
        hitr = np.where(ok, np.minimum(d1, d2), np.inf)
        min_r = np.minimum(min_r, hitr.min(axis=1))
        noise = hash_normal(self.seed, self.fid[:, None], self._ctr(6) + i[None, :].astype(np.uint64))
        return (min_r + w.range_std * noise).astype(np.float32)


def _parallel(fn, world, B, steps, seed, first_filter, workers, **kw):
    """Run `fn` on contiguous filter chunks in threads (filters are independent and the RNG is keyed on the
    filter index, so the result is identical to a single call) and concatenate along the filter axis."""
    from concurrent.futures import ThreadPoolExecutor
    chunk = -(-B // workers)
    jobs = [(first_filter + s, min(chunk, B - s)) for s in range(0, B, chunk)]
    with ThreadPoolExecutor(max_workers=workers) as ex:
        parts = list(ex.map(lambda j: fn(world, j[1], steps, seed, j[0], **kw), jobs))
    return {k: np.concatenate([p[k] for p in parts], axis=1) for k in parts[0]}


def simulate_known(world: World, B: int, steps: int, seed: int = 0, first_filter: int = 0, workers: int = 1):
    """`steps` SLAM steps (one per fake-sensor message) for B filters, known association.

    Returns dict of time-major arrays: twists [T,B,2] (dtheta, dx), xy [T,B,2*n_slots],
    vis [T,B,n_slots] uint8 (all zero at step 0: the node's first measurement() call only initialises,
    slam.cpp:315-327), truth [T,B,3] (x, y, theta), odom [T,B,3] (the dead-reckoning pose of slam.cpp:96)."""
    if workers > 1 and B >= 4 * workers:
        return _parallel(simulate_known, world, B, steps, seed, first_filter, workers)
    sim = TubeWorldSim(world, B, seed, first_filter)
    n, nt = world.n_slots, min(world.n_tubes, world.n_slots)
    tw = np.zeros((steps, B, 2))
    xy = np.zeros((steps, B, 2 * n))
    vis = np.zeros((steps, B, n), dtype=np.uint8)
    truth = np.zeros((steps, B, 3))
    odom = np.zeros((steps, B, 3))
    for t in range(steps):
        for _ in range(11):
            sim.step_tick()
        odom[t, :, 0], odom[t, :, 1], odom[t, :, 2] = sim.ox, sim.oy, sim.oth
        rx, ry, v = sim.fake_sensor()
        dth, dx = sim.odom_twist()
        tw[t, :, 0], tw[t, :, 1] = dth, dx
        xy[t, :, 0:2 * nt:2] = rx[:, :nt]
        xy[t, :, 1:2 * nt:2] = ry[:, :nt]
        if t > 0:
            vis[t, :, :nt] = v[:, :nt]
        truth[t, :, 0], truth[t, :, 1], truth[t, :, 2] = sim.x, sim.y, sim.th
    return {"twists": tw, "xy": xy, "vis": vis, "truth": truth, "odom": odom}


def simulate_unknown(world: World, B: int, steps: int, seed: int = 0, first_filter: int = 0, m_max=None,
                     shuffle=True):
    """Unknown-association inputs without the laser: the visible fake-sensor readings, unlabelled and
    (optionally) in a seeded shuffled order.  Returns twists [T,B,2], meas [T,B,m_max,2], count [T,B] int32,
    truth [T,B,3]."""
    sim = TubeWorldSim(world, B, seed, first_filter)
    nt = world.n_tubes
    m_max = m_max or nt
    tw = np.zeros((steps, B, 2))
    meas = np.zeros((steps, B, m_max, 2))
    cnt = np.zeros((steps, B), dtype=np.int32)
    truth = np.zeros((steps, B, 3))
    for t in range(steps):
        for _ in range(11):
            sim.step_tick()
        rx, ry, v = sim.fake_sensor()
        dth, dx = sim.odom_twist()
        tw[t, :, 0], tw[t, :, 1] = dth, dx
        key = hash_uniform(seed, sim.fid[:, None], sim._ctr(5) + np.arange(nt, dtype=np.uint64)[None, :])
        if not shuffle:
            key = np.broadcast_to(np.arange(nt, dtype=np.float64)[None, :] / nt, key.shape)
        key = np.where(v, key, 2.0)  # invisible last
        me = min(m_max, nt)
        order = np.argsort(key, axis=1, kind="stable")[:, :me]
        c = np.minimum(v.sum(axis=1), me)
        sel = np.arange(me)[None, :] < c[:, None]
        meas[t, :, :me, 0] = np.where(sel, np.take_along_axis(rx, order, axis=1), 0.0)
        meas[t, :, :me, 1] = np.where(sel, np.take_along_axis(ry, order, axis=1), 0.0)
        cnt[t] = c
        truth[t, :, 0], truth[t, :, 1], truth[t, :, 2] = sim.x, sim.y, sim.th
    return {"twists": tw, "meas": meas, "count": cnt, "truth": truth}


def simulate_scans(world: World, B: int, steps: int, seed: int = 0, first_filter: int = 0, n_beams=360):
    """Laser variant (cfg2).  The simulator publishes a scan every 21st tick (tube_world.cpp:594-599); the
    landmarks node re-publishes its circle fit of the latest scan every 10 ms (landmarks.cpp:139-143), so the
    SLAM node steps at its own 10 Hz timer with whatever scan is newest (unknown_data_assoc.cpp:406-428) —
    each scan is therefore consumed by two or three consecutive steps.  Returns twists [T,B,2],
    ranges [T,B,n_beams] float32, scan_id [T] (which scan each step saw), truth [T,B,3]."""
    sim = TubeWorldSim(world, B, seed, first_filter)
    tw = np.zeros((steps, B, 2))
    ranges = np.zeros((steps, B, n_beams), dtype=np.float32)
    scan_id = np.zeros(steps, dtype=np.int32)
    truth = np.zeros((steps, B, 3))
    latest, sid, t = None, -1, 0
    while t < steps:
        sim.step_tick()
        if sim.tick % 21 == 0:
            latest = sim.laser_scan(n_beams)
            sid += 1
        if sim.tick % 10 == 0 and latest is not None:
            ranges[t] = latest
            scan_id[t] = sid
            dth, dx = sim.odom_twist()
            tw[t, :, 0], tw[t, :, 1] = dth, dx
            truth[t, :, 0], truth[t, :, 1], truth[t, :, 2] = sim.x, sim.y, sim.th
            t += 1
    return {"twists": tw, "ranges": ranges, "scan_id": scan_id, "truth": truth}
