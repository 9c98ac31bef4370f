"""Development aid (under torchrun): how a sharded correction period splits into sweep and per-correction chain."""
import os
import sys

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
tr = tg.simulate_known(w, 1, 40, seed=77)
f = ShardedEKF.from_process_group(n, dist, local)
t = 0
for t in range(2):
    f.prediction(tuple(tr["twists"][t, 0]))
    f.measurement(tr["xy"][t, 0], tr["vis"][t, 0])
f.sync()
res = []
for rep in range(3):
    # (a) corrections only: stop before the group fills (timer_stop settles: subtract the sweep measured in (b))
    dist.barrier()
    import time
    h0 = time.perf_counter()
    f.timer_start()
    k = 0
    tt = t + 1
    while k < 12:
        ids = np.flatnonzero(tr["vis"][tt, 0])[: 12 - k]
        vis = np.zeros(n, np.uint8)
        vis[ids] = 1
        f.prediction(tuple(tr["twists"][tt, 0]))
        f.measurement(tr["xy"][tt, 0], vis)
        k += len(ids)
        tt += 1
    host_ms = (time.perf_counter() - h0) * 1e3   # host time to ENQUEUE the 12 corrections
    ms_a = f.timer_stop()       # 12 corrections + predictions + ONE sweep (P = 12)
    # (b) the same again, then a lone sweep by itself
    k = 0
    while k < 12:
        ids = np.flatnonzero(tr["vis"][tt, 0])[: 12 - k]
        vis = np.zeros(n, np.uint8)
        vis[ids] = 1
        f.prediction(tuple(tr["twists"][tt, 0]))
        f.measurement(tr["xy"][tt, 0], vis)
        k += len(ids)
        tt += 1
    f.timer_start()             # event after the 12 corrections (still pending) ...
    ms_b = f.timer_stop()       # ... so this is just their sweep
    t = tt
    res.append((ms_a, ms_b, host_ms))
if dist.get_rank() == 0:
    for ms_a, ms_b, host_ms in res:
        print(f"host enqueue {host_ms:.3f} ms | 12 corrections + sweep: {ms_a:.3f} ms; sweep alone (P=12): {ms_b:.3f} ms; chain per correction: {(ms_a - ms_b) / 12 * 1e3:.1f} us; exchange: {f.exchange}")
f.close()
dist.barrier()
dist.destroy_process_group()
