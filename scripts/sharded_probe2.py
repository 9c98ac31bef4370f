"""Development aid (under torchrun): device time of predictions alone and of corrections alone on the sharded filter."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import torch.distributed as dist

import ekf_slam_ml_b200 as pkg
from ekf_slam_ml_b200.sharded import ShardedEKF

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
nx = int(round(np.sqrt(n)))
tg = pkg.tracegen
w = tg.grid_world(nx, nx, pitch=0.5, n_slots=n, max_visible=0.7)
tr = tg.simulate_known(w, 1, 8, seed=77)
f = ShardedEKF.from_process_group(n, dist, local)
for t in range(2):
    f.prediction(tuple(tr["twists"][t, 0]))
    f.measurement(tr["xy"][t, 0], tr["vis"][t, 0])
f.sync()
ids_all = np.flatnonzero(tr["vis"][2, 0])
far = np.array([n - 1 - 7 * k for k in range(12)])  # landmarks owned by the last rank
for rep in range(3):
    out = []
    # predictions alone (nothing pending)
    dist.barrier()
    f.timer_start()
    for k in range(12):
        f.prediction(tuple(tr["twists"][3, 0]))
    out.append(("12 predictions", f.timer_stop()))
    for name, ids in (("12 corrections, landmarks near the robot", ids_all[:12]), ("12 corrections, last rank's landmarks", far)):
        vis = np.zeros(n, np.uint8)
        vis[ids] = 1
        dist.barrier()
        h0 = time.perf_counter()
        f.timer_start()
        f.measurement(tr["xy"][2, 0], vis)
        host = (time.perf_counter() - h0) * 1e3
        ms_a = f.timer_stop()          # corrections + their sweep
        f.measurement(tr["xy"][2, 0], vis)
        f.timer_start()
        ms_b = f.timer_stop()          # the sweep alone
        out.append((name + f" (host enqueue {host:.3f} ms)", ms_a - ms_b))
    L = pkg._lib.load_sharded()
    if hasattr(L, "ekf_sharded_debug_prof"):
        import ctypes
        buf = (ctypes.c_uint64 * 16)()
        L.ekf_sharded_debug_prof(buf)
        names = ["entry", "pdl wait done", "h + rows loaded", "partial stored", "-", "-", "lines seen", "S^-1 done"]
        for blk in range(2):
            t0 = buf[8 * blk]
            print(f"   rank {dist.get_rank()} {'first' if blk == 0 else 'last'} CTA of the last correction: " +
                  ", ".join(f"{names[k]} +{(buf[8 * blk + k] - t0) / 1e3:.1f}" for k in (1, 2, 3, 6, 7)) + " us", flush=True)
    if dist.get_rank() == 0:
        print(f"[{f.exchange}] " + "; ".join(f"{k}: {v * 1e3 / 12:.1f} us each" for k, v in out), flush=True)
f.close()
dist.barrier()
dist.destroy_process_group()
