"""world_size-2 gloo tests (CPU) of the N > 1 host logic: sharding by global environment id and the one
collective of the job, the all-reduce of the episode-statistics vector.  The stepping itself is done by the
oracle here (there is no GPU): it is keyed by the same global ids as the CUDA kernels."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TOTAL, T, SEED = 12, 30, 5


def _rollout(g0, n):
    from oracle import draws as D
    from oracle.ballenv_oracle import OracleConfig, OracleVec
    vec = OracleVec(OracleConfig(window=5, max_episode_steps=11), D.PhiloxDraws(SEED), n, g0)
    vec.reset()
    src = D.PhiloxDraws(SEED)
    obs = []
    for t in range(T):
        acts = [D.mulhi(src.action_word(g0 + i, t), 9) for i in range(n)]
        vec.step(acts)
        obs.append(np.array(vec.observe(), dtype=np.float32))
    stats = torch.zeros(16, dtype=torch.float64)
    for i, k in enumerate(("episodes", "return_sum", "length_sum", "goals", "hits_static", "hits_dynamic", "timeouts", "steps")):
        stats[i] = vec.stats[k]
    return np.stack(obs), stats


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gym_ballenv_b200.distributed import allreduce_stats, rank_world, shard_bounds
    assert rank_world() == (rank, world, rank)
    off, cnt = shard_bounds(TOTAL, rank, world)
    obs, stats = _rollout(off, cnt)
    local = stats.clone()
    summed, work = allreduce_stats(stats, async_op=True)
    work.wait()
    assert torch.equal(stats, local)                      # the rank's own vector keeps counting locally
    gathered = [torch.zeros(T, TOTAL // world, obs.shape[2]) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(obs))
    if rank == 0:
        out.put((summed.numpy(), torch.cat(gathered, dim=1).numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.timeout(300)
def test_two_rank_sharding_and_stats_allreduce():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    summed, obs2 = out.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    obs1, stats1 = _rollout(0, TOTAL)                      # the same job on one "GPU"
    assert np.array_equal(obs2, obs1), "trajectories must not depend on the sharding"
    np.testing.assert_allclose(summed, stats1.numpy(), rtol=1e-12)
    assert summed[7] == TOTAL * T and summed[0] > 0        # steps, episodes


def test_single_process_allreduce_is_identity():
    from gym_ballenv_b200.distributed import allreduce_stats
    s = torch.arange(16, dtype=torch.float64)
    out, work = allreduce_stats(s)
    assert work is None and torch.equal(out, s) and out.data_ptr() != s.data_ptr()
    with pytest.raises(ValueError):
        allreduce_stats(torch.zeros(4, dtype=torch.float64))
