"""Python mirror of rigid2d::CircleFitting (rigid2d/include/rigid2d/circle_fitting.hpp:18-60) over the C ABI in
include/circle_fit_b200.h: same method names and call sequence as the reference class, plus the batched entry
point the GPU path exists for.  All arithmetic runs in the CUDA library."""
import ctypes

import numpy as np

from . import _lib
from ._lib import c_double_p, c_float_p, c_i32_p, c_u8_p
from .ekf_slam import Vector2D


def _check(rc):
    if rc != 0:
        raise _lib.EkfError(rc, _lib.load().circles_last_error().decode("utf-8", "replace"))


class CircleFitting:
    def __init__(self, device=0, max_scans=1, max_circles=32):
        self._L = _lib.load()
        self.device, self.max_scans, self.max_circles = int(device), int(max_scans), int(max_circles)
        self._ctx, self._n_beams = None, None
        self._centres_only = False
        self.point_cluster, self.xy_cluster, self.r_cluster = [], [], []
        self._flags = []
        self.MAXC = self._L.circles_max_clusters()

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.circles_destroy(self._ctx)
            self._ctx = None

    __del__ = close

    def _ensure(self, n_beams, scans=1):
        if self._ctx is None or self._n_beams != n_beams or scans > self.max_scans:
            self.close()
            self.max_scans = max(self.max_scans, scans)
            h = ctypes.c_void_p()
            _check(self._L.circles_create(self.max_scans, int(n_beams), self.max_circles, self.device, ctypes.byref(h)))
            self._ctx, self._n_beams = h, int(n_beams)
            if self._centres_only:
                _check(self._L.circles_set_centres_only(self._ctx, 1))

    # ---- batched GPU entry points
    def set_centres_only(self, on=True):
        """Batched runs then return exactly what approxCirclePositions() returns and skip the fit of clusters that fail
        the inscribed-angle test (wall segments); accepted centres are unchanged."""
        self._centres_only = bool(on)  # kept across the lazy (re)creation of the device context
        if self._ctx is not None:
            _check(self._L.circles_set_centres_only(self._ctx, 1 if on else 0))

    def run_batch(self, ranges):
        """ranges [B, n_beams] float32 (LaserScan wire format) or float64 -> (centers [B,max_circles,2], counts [B])."""
        r = np.ascontiguousarray(ranges)
        B, nb = r.shape
        self._ensure(nb, B)
        centers = np.zeros((B, self.max_circles, 2))
        counts = np.zeros(B, dtype=np.int32)
        if r.dtype == np.float32:
            _check(self._L.circles_run_f32(self._ctx, r.ctypes.data_as(c_float_p), B, centers.ctypes.data_as(c_double_p),
                                           counts.ctypes.data_as(c_i32_p)))
        else:
            r = np.ascontiguousarray(r, dtype=np.float64)
            _check(self._L.circles_run_f64(self._ctx, r.ctypes.data_as(c_double_p), B, centers.ctypes.data_as(c_double_p),
                                           counts.ctypes.data_as(c_i32_p)))
        return centers, counts

    def last_clusters(self, scan=0):
        """Per-cluster detail of one scan of the last run."""
        n = ctypes.c_int32()
        segs = np.zeros((self.MAXC, 4), dtype=np.int32)
        cxr = np.zeros((self.MAXC, 4))
        flags = np.zeros(self.MAXC, dtype=np.uint8)
        xy = np.zeros((self._n_beams, 2))
        _check(self._L.circles_last_clusters(self._ctx, int(scan), ctypes.byref(n), segs.ctypes.data_as(c_i32_p),
                                             cxr.ctypes.data_as(c_double_p), flags.ctypes.data_as(c_u8_p),
                                             xy.ctypes.data_as(c_double_p)))
        k = n.value
        ids = [list(range(s1, s1 + l1)) + list(range(s2, s2 + l2)) for s1, l1, s2, l2 in segs[:k]]
        return {"n": k, "ids": ids, "cxr": cxr[:k], "is_circle": (flags[:k] & 1).astype(bool),
                "fallback": (flags[:k] & 2).astype(bool), "xy": xy}

    # ---- the reference's public surface (single scan)
    def approxCirclePositions(self, ranges):
        """circle_fitting.cpp:298-304"""
        self.clusteringRanges(ranges)
        return self.classifyCircle(self.circleRegression())

    def clusteringRanges(self, ranges):
        """circle_fitting.cpp:11-90 (clustering runs on the device; the clusters are read back)."""
        r = np.ascontiguousarray(ranges, dtype=np.float64).reshape(1, -1)
        self.run_batch(r)
        d = self.last_clusters(0)
        self.point_cluster = [[float(r[0, i]) for i in ids] for ids in d["ids"]]
        self.xy_cluster = [[Vector2D(d["xy"][i, 0], d["xy"][i, 1]) for i in ids] for ids in d["ids"]]
        self.r_cluster = []

    def circleRegression(self):
        """circle_fitting.cpp:104-232 on the current xy_cluster; appends to r_cluster like the reference."""
        if not self.xy_cluster:
            return []
        sizes = np.array([len(c) for c in self.xy_cluster], dtype=np.int32)
        flat = np.array([[p.x, p.y] for c in self.xy_cluster for p in c], dtype=np.float64)
        k = len(sizes)
        cxr = np.zeros((k, 4))
        flags = np.zeros(k, dtype=np.uint8)
        self._ensure(self._n_beams or 360)
        _check(self._L.circles_fit_clusters(self._ctx, flat.ctypes.data_as(c_double_p), sizes.ctypes.data_as(c_i32_p), k,
                                            cxr.ctypes.data_as(c_double_p), flags.ctypes.data_as(c_u8_p)))
        self.r_cluster.extend(float(v) for v in cxr[:, 2])
        self._flags = [bool(f & 1) for f in flags]
        self._last_flags_raw = flags  # bit0 = circle, bit1 = eigenvalue fallback
        self._last_cxr = cxr
        return [Vector2D(cxr[i, 0], cxr[i, 1]) for i in range(k)]

    def classifyCircle(self, circle_positions):
        """circle_fitting.cpp:234-296: keeps the positions whose cluster passed the inscribed-angle / radius test."""
        return [p for p, ok in zip(circle_positions, self._flags) if ok]

    def get_point_cluster(self):
        return self.point_cluster

    def get_r_cluster(self):
        return self.r_cluster

    def set_xy_cluster(self, new_xy_cluster):
        self.xy_cluster = [[p if isinstance(p, Vector2D) else Vector2D(p[0], p[1]) for p in c] for c in new_xy_cluster]

    @property
    def launch_count(self):
        v = ctypes.c_uint64()
        _check(self._L.circles_launch_count(self._ctx, ctypes.byref(v)))
        return v.value
