"""Multi-GPU plumbing for the hot path.

Every utterance is an independent unit (its K samples, rewards, baseline, gradient and CTC lattice never
leave one GPU), so ranks take contiguous slices of the batch and the loss kernels need no collective.
The only exchanges are scalars: the loss for logging and, when a global baseline is wanted, the reward
statistics (sum R, sum R^2, count).  Upstream's only parallelism is nn.DataParallel (model.py:201).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [lo, hi) slice of n utterances for `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def balanced_assignment(lengths, world):
    """Longest-first greedy assignment: utterance indices per rank so ragged T balances across GPUs."""
    parts, loads = [[] for _ in range(world)], [0] * world
    for i in sorted(range(len(lengths)), key=lambda i: -int(lengths[i])):
        r = loads.index(min(loads))
        parts[r].append(i)
        loads[r] += int(lengths[i])
    return parts


def reward_stats(rewards):
    """(sum R, sum R^2, count) as a float64 tensor on the rewards' device."""
    r = rewards.to(torch.float64)
    return torch.stack([r.sum(), (r * r).sum(), torch.tensor(float(r.numel()), dtype=torch.float64, device=r.device)])


def allreduce_reward_stats(rewards, group=None):
    """Global reward mean and variance over all ranks (one 3-element all-reduce)."""
    s = reward_stats(rewards)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    mean = s[0] / s[2]
    var = torch.clamp(s[1] / s[2] - mean * mean, min=0.0)
    return float(mean), float(var), int(s[2])


def allreduce_mean(value, group=None):
    """Average a scalar tensor over ranks (loss logging)."""
    v = value.detach().clone().reshape(1)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
        v /= dist.get_world_size(group)
    return v[0]


class MovingBaseline:
    """Exponential moving average of the global mean reward; feed .value to baseline='value'."""

    def __init__(self, momentum=0.9):
        self.momentum, self.value, self._init = float(momentum), 0.0, False

    def update(self, rewards, group=None):
        mean, _, _ = allreduce_reward_stats(rewards, group)
        self.value = mean if not self._init else self.momentum * self.value + (1 - self.momentum) * mean
        self._init = True
        return self.value
