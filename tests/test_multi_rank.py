"""world_size-2 tests of the data-parallel scheme.

CPU (gloo): the product's sharding + all-reduce helpers, with the per-shard un-normalised gradients supplied by the CPU
oracle -- after ONE all-reduce of [grads | sum(mask)] and the division by the global sum(mask), every rank holds the
full-batch gradient of the reference loss.  GPU (nccl, needs >= 2 devices): two QLearner replicas on half the episodes
each end the step with the parameters of one learner trained on the whole batch."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _oracle_case():
    from golden_util import load_learner_case
    return load_learner_case("refil")


def _unnormalised_grads(c, lo_, hi_):
    """gradient of (1-l) sum((m td)^2) + l sum((m td_im)^2) over episodes [lo_, hi_), plus sum(mask) -- oracle, CPU."""
    from oracle import learner_oracle as lo
    batch = {k: v[lo_:hi_] for k, v in c.batch.items()}
    ap = {k: v.clone().requires_grad_(True) for k, v in c.agent.items() if "scale_factor" not in k}
    mp_ = {k: v.clone().requires_grad_(True) for k, v in c.mixer.items() if "scale_factor" not in k}
    loss, aux = lo.td_loss(ap, mp_, c.tagent, c.tmixer, batch, c.args, group_a=c.group_a[lo_:hi_])
    msum = aux["mask"].sum()
    (loss * msum).backward()
    g = torch.cat([p.grad.reshape(-1) for p in list(ap.values()) + list(mp_.values())])
    return g, msum


def _gloo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from refil_b200 import parallel
    parallel.init_distributed(backend="gloo")
    assert parallel.rank_world() == (rank, world)
    c = _oracle_case()
    lo_, hi_ = parallel.shard_range(c.dims[0], rank, world)
    g, msum = _unnormalised_grads(c, lo_, hi_)
    buf = torch.cat([g, msum.reshape(1), torch.zeros(7)])          # [grads | sum(mask), ...] layout of QLearner.gradbuf
    parallel.all_reduce_sum_(buf)
    flat = torch.arange(4.0) + rank
    parallel.broadcast_(flat, src=0)
    assert torch.equal(flat, torch.arange(4.0))
    # host-integer reductions the CLI loop relies on (run.py: global t_env, common truncation length)
    assert parallel.all_reduce_sum_int(10 + rank) == 21 and parallel.all_reduce_max_int(5 * rank) == 5
    if rank == 0:
        torch.save({"grad": buf[:-8] / buf[-8], "msum": buf[-8]}, out)
    dist.destroy_process_group()


def test_sharded_gradients_equal_full_batch_gloo(tmp_path):
    from refil_b200 import parallel
    assert [parallel.shard_range(7, r, 3) for r in range(3)] == [(0, 3), (3, 5), (5, 7)]
    assert parallel.shard_range(8, 0, 1) == (0, 8)
    out = str(tmp_path / "g.pt")
    mp.spawn(_gloo_worker, args=(2, 29541, out), nprocs=2, join=True)
    got = torch.load(out)
    c = _oracle_case()
    g_full, msum = _unnormalised_grads(c, 0, c.dims[0])
    ref = g_full / msum
    assert float(got["msum"]) == float(msum)
    assert torch.allclose(got["grad"], ref, rtol=1e-4, atol=1e-7), float((got["grad"] - ref).abs().max())
    # ... and that is the reference gradient (golden fixture) before clipping: grad_norm matches
    assert abs(float(ref.norm()) - c.stats["grad_norm"]) <= 1e-4 * max(1.0, c.stats["grad_norm"])


def _nccl_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    torch.cuda.set_device(rank)
    from gpu_util import build_product
    from refil_b200 import parallel
    parallel.init_distributed(backend="nccl", device="cuda:%d" % rank)
    c = _oracle_case()
    lo_, hi_ = parallel.shard_range(c.dims[0], rank, world)
    dims = (hi_ - lo_,) + tuple(c.dims[1:])
    batch, mac, learner, logger = build_product(c.args, dims, {k: v[lo_:hi_] for k, v in c.batch.items()}, "cuda:%d" % rank)
    mac.agent.load_state_dict(c.agent)
    learner.target_mac.agent.load_state_dict(c.tagent)
    learner.mixer.load_state_dict(c.mixer)
    learner.target_mixer.load_state_dict(c.tmixer)
    learner.train(batch, t_env=10, episode_num=0, group_bits=c.group_a[lo_:hi_].to("cuda:%d" % rank))
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({"flat": learner.flat.cpu(), "loss": logger.stats["loss"][0], "grad_norm": logger.stats["grad_norm"][0]}, out)
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_step_equals_single_gpu_step(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    from gpu_util import build_product
    out = str(tmp_path / "p.pt")
    mp.spawn(_nccl_worker, args=(2, 29543, out), nprocs=2, join=True)
    got = torch.load(out)
    c = _oracle_case()
    batch, mac, learner, logger = build_product(c.args, c.dims, c.batch, "cuda:0")
    mac.agent.load_state_dict(c.agent)
    learner.target_mac.agent.load_state_dict(c.tagent)
    learner.mixer.load_state_dict(c.mixer)
    learner.target_mixer.load_state_dict(c.tmixer)
    learner.train(batch, t_env=10, episode_num=0, group_bits=c.group_a.to("cuda:0"))
    torch.cuda.synchronize()
    assert abs(got["loss"] - logger.stats["loss"][0]) <= 1e-5 * max(1.0, abs(got["loss"]))
    assert abs(got["loss"] - c.stats["loss"]) <= 1e-4 * max(1.0, abs(c.stats["loss"]))
    assert abs(got["grad_norm"] - c.stats["grad_norm"]) <= 1e-4 * max(1.0, c.stats["grad_norm"])
    assert float((got["flat"] - learner.flat.cpu()).abs().max()) <= 2e-5


@pytest.mark.gpu
def test_cli_under_torchrun_two_gpus(tmp_path):
    """src/main.py under torchrun: both ranks stay in lock-step (global t_env), train ONE model (identical replicas) and only
    rank 0 writes the checkpoint."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import glob
    import subprocess
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29547", os.path.join(ROOT, "src", "main.py"), "--env-config=group_matching",
           "--config=refil_group_matching", "with", "env_args.n_agents=4", "batch_size_run=16", "batch_size=8", "t_max=3000",
           "seed=3", "test_interval=100000", "test_nepisode=16", "save_model=True", "save_model_interval=100000",
           "local_results_path=%s" % tmp_path, "cuda_graph=True"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    tokens = glob.glob(os.path.join(str(tmp_path), "models", "*"))
    assert len(tokens) == 1, tokens                     # one writer, one run directory
    steps = sorted(os.listdir(tokens[0]))
    assert steps and all(os.path.exists(os.path.join(tokens[0], s, "agent.th")) for s in steps)
