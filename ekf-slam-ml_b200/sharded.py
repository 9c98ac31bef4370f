"""Row-block-sharded EKF-SLAM filter (include/ekf_sharded_b200.h): the rigid2d::EKF_SLAM verbs for maps whose
covariance does not fit — or should not live — on one GPU (BASELINE.json cfg5).  One process per GPU; every rank calls
every verb with the same arguments."""
import ctypes
import os

import numpy as np

from . import _lib
from ._lib import c_double_p, c_i32_p, c_u8_p


def _check(rc):
    if rc != 0:
        raise _lib.EkfError(rc, _lib.load_sharded().ekf_sharded_last_error().decode("utf-8", "replace"))


class ShardedEKF:
    def __init__(self, handle, n, world, local):
        self._L = _lib.load_sharded()
        self._h, self.n, self.N, self.world, self.local = handle, int(n), 3 + 2 * int(n), int(world), bool(local)

    @classmethod
    def local_emulation(cls, n, world, device=0):
        """`world` shards on one device, exchanges as device copies (single-GPU parity tests of the sharding)."""
        L = _lib.load_sharded()
        h = ctypes.c_void_p()
        _check(L.ekf_sharded_create_local(int(n), int(world), int(device), ctypes.byref(h)))
        return cls(h, n, world, True)

    @classmethod
    def from_process_group(cls, n, dist, device):
        """One rank of a torch.distributed job: rank 0 draws the NCCL id and the group broadcasts it."""
        L = _lib.load_sharded()
        rank, world = dist.get_rank(), dist.get_world_size()
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            _check(L.ekf_sharded_unique_id(buf))
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=0)
        idbuf = ctypes.create_string_buffer(box[0], 128)
        h = ctypes.c_void_p()
        _check(L.ekf_sharded_create(int(n), rank, world, idbuf, int(device), ctypes.byref(h)))
        self = cls(h, n, world, False)
        self.exchange = "nccl all-reduce"
        if os.environ.get("EKF_SHARDED_EXCHANGE", "push") != "nccl":
            self._attach_push_exchange(dist, device)
        return self

    def _attach_push_exchange(self, dist, device):
        """Symmetric exchange buffers (peer-mapped, multicast where the fabric has it) from torch's symmetric memory;
        on any failure the handle simply keeps the NCCL all-reduce."""
        try:
            import torch
            import torch.distributed._symmetric_memory as symm_mem
            nbytes = ctypes.c_uint64()
            _check(self._L.ekf_sharded_exchange_bytes(self._h, ctypes.byref(nbytes)))
            buf = symm_mem.empty((nbytes.value + 7) // 8, dtype=torch.float64, device=torch.device("cuda", device))
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
            torch.cuda.synchronize(device)
            dist.barrier()
            peers = (ctypes.c_uint64 * self.world)(*[int(p) for p in hdl.buffer_ptrs])
            mc = int(hdl.multicast_ptr) if getattr(hdl, "multicast_ptr", 0) else 0
            if os.environ.get("EKF_SHARDED_EXCHANGE") == "push-unicast":
                mc = 0
            _check(self._L.ekf_sharded_attach_exchange(self._h, self.world, peers, ctypes.c_uint64(mc)))
            self._xbuf, self._xhdl = buf, hdl
            self.exchange = "push over NVLink (%s)" % ("NVSwitch multicast" if mc else "peer stores")
        except Exception as e:  # noqa: BLE001 - optional fast path
            import sys
            sys.stderr.write(f"[ekf_sharded] push exchange unavailable, keeping the NCCL all-reduce: {e!r}\n")

    def close(self):
        if getattr(self, "_h", None):
            self._L.ekf_sharded_destroy(self._h)
            self._h = None

    __del__ = close

    def prediction(self, twist):
        dth, dx = (twist.angular(), twist.linearX()) if hasattr(twist, "angular") else twist
        _check(self._L.ekf_sharded_predict(self._h, float(dth), float(dx)))

    def measurement(self, sensor_reading, visible_list, known_list=None):
        xy, vis = sensor_reading, visible_list
        if not (isinstance(xy, np.ndarray) and xy.dtype == np.float64 and xy.flags.c_contiguous):
            xy = np.ascontiguousarray(xy, dtype=np.float64)
        if not (isinstance(vis, np.ndarray) and vis.dtype == np.uint8 and vis.flags.c_contiguous):
            vis = np.ascontiguousarray(vis, dtype=np.uint8)
        if xy.size != 2 * self.n or vis.size != self.n:
            raise ValueError("measurement() takes 2n readings and n visibility flags")
        rc = self._L.ekf_sharded_measurement(self._h, xy.ctypes.data, vis.ctypes.data)
        if rc:
            _check(rc)

    def data_association(self, measures, known_list):
        xy = np.ascontiguousarray(measures, dtype=np.float64).reshape(-1)
        m = xy.size // 2
        known = np.ascontiguousarray(known_list, dtype=np.uint8).reshape(-1).copy()
        assoc = np.full(m, -2, dtype=np.int32)
        dmin, second = np.zeros(m), np.zeros(m)
        created = np.zeros(m, dtype=np.uint8)
        _check(self._L.ekf_sharded_data_association(self._h, xy.ctypes.data_as(c_double_p), m, known.ctypes.data_as(c_u8_p),
                                                    assoc.ctypes.data_as(c_i32_p), dmin.ctypes.data_as(c_double_p),
                                                    second.ctypes.data_as(c_double_p), created.ctypes.data_as(c_u8_p)))
        known_list[...] = known
        return {"assoc": assoc, "dmin": dmin, "second": second, "created": created}

    @property
    def state(self):
        out = np.zeros(self.N)
        _check(self._L.ekf_sharded_get_state(self._h, out.ctypes.data_as(c_double_p)))
        return out

    def rows(self, shard=0):
        a, b = ctypes.c_int64(), ctypes.c_int64()
        _check(self._L.ekf_sharded_rows(self._h, int(shard), ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def sigma_rows(self, shard=0):
        """The rows of Sigma owned by `shard` (NCCL mode: 0 = this rank), shape [rows, N]."""
        a, b = self.rows(shard)
        out = np.zeros((max(b - a, 0), self.N))
        if b > a:
            _check(self._L.ekf_sharded_get_sigma_rows(self._h, int(shard), out.ctypes.data_as(c_double_p), self.N))
        return out

    def sigma_row_list(self, rows, shard=0):
        """Selected GLOBAL rows owned by `shard` (NCCL mode: 0 = this rank), shape [len(rows), N]."""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        out = np.zeros((rows.size, self.N))
        if rows.size:
            _check(self._L.ekf_sharded_get_sigma_row_list(self._h, int(shard), rows.ctypes.data_as(_lib.c_i64_p), int(rows.size),
                                                          out.ctypes.data_as(c_double_p), self.N))
        return out

    def sigma_full_local(self):
        """Local-emulation mode only: the whole covariance, assembled from all shards."""
        assert self.local
        return np.concatenate([self.sigma_rows(g) for g in range(self.world)], axis=0)

    @property
    def update_count(self):
        v = ctypes.c_uint64()
        _check(self._L.ekf_sharded_update_count(self._h, ctypes.byref(v)))
        return v.value

    @property
    def launch_count(self):
        v = ctypes.c_uint64()
        _check(self._L.ekf_sharded_launch_count(self._h, ctypes.byref(v)))
        return v.value

    @property
    def sweep_count(self):
        v = ctypes.c_uint64()
        _check(self._L.ekf_sharded_sweep_count(self._h, ctypes.byref(v)))
        return v.value

    def set_carry_pending(self, on):
        """Let correction factors stay pending across prediction() / measurement() calls (default on)."""
        _check(self._L.ekf_sharded_set_carry_pending(self._h, 1 if on else 0))

    def set_max_pending(self, k):
        """Corrections per sweep, 1..20 (default 14); every rank must use the same value."""
        _check(self._L.ekf_sharded_set_max_pending(self._h, int(k)))

    def sync(self):
        _check(self._L.ekf_sharded_sync(self._h))

    def timer_start(self):
        _check(self._L.ekf_sharded_timer_start(self._h))

    def timer_stop(self):
        ms = ctypes.c_float()
        _check(self._L.ekf_sharded_timer_stop(self._h, ctypes.byref(ms)))
        return ms.value
