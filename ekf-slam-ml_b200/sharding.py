"""Host-side plumbing for the two multi-GPU modes (one process per GPU, torch.distributed for the exchange).

* independent filters (cfg3): filter b lives on rank b // per_gpu; nothing is exchanged on the data path, ranks only
  all-reduce the 4-double error statistic and the timing (max over ranks);
* row-block-sharded covariance (cfg5): rank g owns rows [row_begin(g), row_begin(g+1)) of Sigma, all N columns.

Everything here is backend-agnostic (nccl on the GPUs, gloo in the CPU tests)."""
import numpy as np


def filter_range(rank, world, per_gpu):
    """Global filter indices [first, first + per_gpu) owned by `rank` (weak scaling: per-GPU work is fixed)."""
    assert 0 <= rank < world
    return rank * per_gpu, per_gpu


def row_blocks(N, world, align=1):
    """Row-block partition of an N x N covariance: returns world+1 boundaries, block sizes differing by at most
    `align` rows; boundaries are multiples of `align` (except the last, which is N)."""
    per = -(-N // world)
    per = -(-per // align) * align
    b = [min(N, g * per) for g in range(world + 1)]
    b[-1] = N
    return b


def landmark_row_blocks(n_landmarks, world):
    """The partition the row-sharded engine uses (csrc/sharded/sharded.cu::partition): rank g owns the two rows of each
    landmark in [L_g, L_g+1) with ceil(n / world) landmarks per rank, and rank 0 additionally the three robot rows.
    Returns world+1 row boundaries; boundaries other than 0 and N are odd, so a landmark's row pair is never split."""
    per = -(-n_landmarks // world)
    lm = [min(n_landmarks, g * per) for g in range(world + 1)]
    return [0] + [3 + 2 * v for v in lm[1:]]


def owner_of_row(row, bounds):
    """Rank that owns `row` under boundaries from row_blocks()."""
    return int(np.searchsorted(np.asarray(bounds), row, side="right") - 1)


def allreduce_sum(values, dist=None, device=None):
    """Sum a small float64 vector over ranks (error statistics).  `dist` = torch.distributed or None."""
    v = np.asarray(values, dtype=np.float64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return v
    import torch
    t = torch.tensor(v, dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def allreduce_max(value, dist=None, device=None):
    """Max of a scalar over ranks (multi-GPU timings are the slowest rank's)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def rmse_from_stats(stats):
    """stats = {sum dx^2, sum dy^2, sum dtheta^2, count} -> (rmse_x, rmse_y, rmse_theta)."""
    s = np.asarray(stats, dtype=np.float64)
    return tuple(float(np.sqrt(s[i] / s[3])) for i in range(3))
