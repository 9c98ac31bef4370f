"""On-device batched input generator (include/tube_world_b200.h): the GPU twin of tracegen.TubeWorldSim."""
import ctypes

import numpy as np

from . import _lib
from ._lib import c_double_p, c_float_p, c_u8_p


class _Params(ctypes.Structure):
    _fields_ = [(k, ctypes.c_double) for k in ("tube_radius", "border", "wheel_base", "wheel_radius", "vx_std", "the_std",
                                               "slip_min", "slip_max", "sensor_std", "max_visible", "range_std", "cmd_v",
                                               "cmd_radius")] + [("n_slots", ctypes.c_int32), ("pad", ctypes.c_int32)]


def _check(rc):
    if rc != 0:
        raise _lib.EkfError(rc, _lib.load().tubeworld_last_error().decode("utf-8", "replace"))


class TubeWorld:
    """B robots of one tracegen.World on the GPU; same RNG keys as tracegen (seed, first_filter + b, tick, purpose)."""

    def __init__(self, world, n_robots, seed=0, first_filter=0, device=0):
        self._L = _lib.load()
        self.B, self.n = int(n_robots), int(world.n_slots)
        p = _Params(**{k: float(getattr(world, k)) for k, _ in _Params._fields_[:13]}, n_slots=self.n, pad=0)
        tx = np.ascontiguousarray(world.tubes_x, dtype=np.float64)
        ty = np.ascontiguousarray(world.tubes_y, dtype=np.float64)
        h = ctypes.c_void_p()
        _check(self._L.tubeworld_create(self.B, ctypes.byref(p), tx.ctypes.data_as(c_double_p), ty.ctypes.data_as(c_double_p),
                                        len(tx), int(seed), int(first_filter), int(device), ctypes.byref(h)))
        self._h = h
        self.n_beams = 0

    def close(self):
        if getattr(self, "_h", None):
            self._L.tubeworld_destroy(self._h)
            self._h = None

    __del__ = close

    def use_stream(self, stream_ptr):
        _check(self._L.tubeworld_set_stream(self._h, ctypes.c_void_p(stream_ptr)))

    def step_known(self):
        """11 ticks + fake sensor + odometry twist; results stay on the device (see device_pointers)."""
        _check(self._L.tubeworld_step_known(self._h))

    def step_scan(self, ticks, n_beams=360):
        _check(self._L.tubeworld_step_scan(self._h, int(ticks), int(n_beams)))
        self.n_beams = int(n_beams)

    def device_pointers(self):
        p = [ctypes.c_void_p() for _ in range(5)]
        _check(self._L.tubeworld_outputs(self._h, *[ctypes.byref(x) for x in p]))
        return {k: v.value for k, v in zip(("twists", "xy", "vis", "truth", "ranges"), p)}

    def download(self, ranges=False):
        tw = np.empty((self.B, 2))
        xy = np.empty((self.B, 2 * self.n))
        vis = np.empty((self.B, self.n), dtype=np.uint8)
        truth = np.empty((self.B, 3))
        rg = np.empty((self.B, self.n_beams), dtype=np.float32) if ranges else None
        _check(self._L.tubeworld_download(self._h, tw.ctypes.data_as(c_double_p), xy.ctypes.data_as(c_double_p),
                                          vis.ctypes.data_as(c_u8_p), truth.ctypes.data_as(c_double_p),
                                          rg.ctypes.data_as(c_float_p) if ranges else None))
        odom = np.empty((self.B, 3))
        _check(self._L.tubeworld_odometry(self._h, None, odom.ctypes.data))
        out = {"twists": tw, "xy": xy, "vis": vis, "truth": truth, "odom": odom}
        if ranges:
            out["ranges"] = rg
        return out

    def sync(self):
        _check(self._L.tubeworld_sync(self._h))
