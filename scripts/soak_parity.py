"""Long-run parity of the batched engine against the oracle (T >= 200 steps, both
association variants).  B filters x T steps on the GPU, `--check` of them replayed through oracle/ekf_oracle.c.

    python scripts/soak_parity.py [--filters 16384] [--steps 230] [--check 1024] > profiles/r2_soak_parity.json

Test infrastructure (it imports tests/_oracle.py); not part of the product or of bench.py."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--filters", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=230)
    ap.add_argument("--check", type=int, default=1024)
    ap.add_argument("--only-large", action="store_true")
    a = ap.parse_args()
    import torch

    import ekf_slam_ml_b200 as pkg
    from _oracle import OracleEKF, sigma_err, state_err

    tg = pkg.tracegen
    B, n, T, M = a.filters, 20, a.steps, 12
    rng = np.random.default_rng(7)
    picks = sorted(set([0, B - 1] + list(rng.integers(0, B, a.check))))
    out = {"filters": B, "steps": T, "filters_checked": len(picks), "tol": 1e-9,
           "checker": "oracle/ekf_oracle.c (pinned to the reference build), every checked filter replayed from step 0"}
    if not a.only_large:
        # ---- known association
        t0 = time.time()
        tr = tg.simulate_known(tg.dense_world(n), B, T, seed=31, workers=os.cpu_count())
        bt = pkg.EKFBatch(B, n)
        d_tw = torch.from_numpy(np.ascontiguousarray(tr["twists"])).cuda()
        d_xy = torch.from_numpy(np.ascontiguousarray(tr["xy"])).cuda()
        d_vis = torch.from_numpy(np.ascontiguousarray(tr["vis"])).cuda()
        for t in range(T):
            bt.step_known_dev(d_tw[t].data_ptr(), d_xy[t].data_ptr(), d_vis[t].data_ptr())
        bt.sync()
        states = bt.states()
        ws = wg = 0.0
        for b in picks:
            o = OracleEKF(n)
            for t in range(T):
                o.prediction(*tr["twists"][t, b])
                o.measurement(tr["xy"][t, b], tr["vis"][t, b])
            ws = max(ws, state_err(states[b], o.state))
            wg = max(wg, sigma_err(bt.sigma(int(b)), o.sigma))
        out["known"] = {"corrections_total": int(bt.update_count), "state_err": ws, "sigma_err": wg, "ok": bool(ws < 1e-9 and wg < 1e-9),
                        "seconds": round(time.time() - t0, 1)}
        bt.close()
        del d_tw, d_xy, d_vis
        # ---- unknown association
        t0 = time.time()
        tr = tg.simulate_unknown(tg.dense_world(n), B, T, seed=32, m_max=M)
        bt = pkg.EKFBatch(B, n)
        d_tw = torch.from_numpy(np.ascontiguousarray(tr["twists"])).cuda()
        d_me = torch.from_numpy(np.ascontiguousarray(tr["meas"])).cuda()
        d_ct = torch.from_numpy(np.ascontiguousarray(tr["count"])).cuda()
        for t in range(T):
            bt.step_unknown_dev(d_tw[t].data_ptr(), d_me[t].data_ptr(), d_ct[t].data_ptr(), M)
        bt.sync()
        states = bt.states()
        known_gpu = bt.known
        ws = wg = 0.0
        known_same = True
        for b in picks:
            o = OracleEKF(n)
            known = np.zeros(n, np.uint8)
            for t in range(T):
                o.prediction(*tr["twists"][t, b])
                o.data_association(tr["meas"][t, b, :tr["count"][t, b]], known)
            known_same = known_same and bool(np.array_equal(known, known_gpu[b]))
            ws = max(ws, state_err(states[b], o.state))
            wg = max(wg, sigma_err(bt.sigma(int(b)), o.sigma))
        out["unknown"] = {"corrections_total": int(bt.update_count), "state_err": ws, "sigma_err": wg, "known_list_identical": known_same,
                          "ok": bool(ws < 1e-9 and wg < 1e-9 and known_same), "seconds": round(time.time() - t0, 1)}
        bt.close()
    # ---- the large map over several sweeps: >= 45 corrections at n = 8,192 (groups of 14 carried across predictions),
    # the whole covariance against the oracle's O(N^2) form (bench.py's own check stops after the first group)
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    t0 = time.time()
    trl = tg.simulate_known(tg.grid_world(128, 64, pitch=0.5, n_slots=8192, max_visible=0.7), 1, 24, seed=99)
    par, _ = bench.parity_large_map(pkg, 0, 8192, trl, min_corrections=45)
    par["seconds"] = round(time.time() - t0, 1)
    out["large_map"] = par
    out["ok"] = bool(par["ok"] and all(out[k]["ok"] for k in ("known", "unknown") if k in out))
    print(json.dumps(out))
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
