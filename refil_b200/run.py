"""Experiment orchestration: build runner / replay buffer / controller / learner and run the train loop.

Mirror of /root/reference/src/run.py:25-311 for the entity scheme (scheme construction :173-196, training loop :258-308,
checkpoint directory layout :214-241,290-301).  sacred is not a dependency: `run(config, console_logger)` takes the merged
config dict directly."""
import datetime
import os
import pprint
import time
from types import SimpleNamespace

import numpy as np
import torch

from . import parallel
from .components.episode_buffer import ReplayBuffer
from .components.transforms import OneHot
from .controllers import REGISTRY as mac_REGISTRY
from .learners import REGISTRY as le_REGISTRY
from .runners import REGISTRY as r_REGISTRY
from .utils.logging import Logger


def args_sanity_check(config, console):
    if config["use_cuda"] and not torch.cuda.is_available():
        raise RuntimeError("refil_b200 runs on CUDA devices only (sm_100a kernels); no CPU fallback exists")
    if config["test_nepisode"] < config["batch_size_run"]:
        config["test_nepisode"] = config["batch_size_run"]
    else:
        config["test_nepisode"] = (config["test_nepisode"] // config["batch_size_run"]) * config["batch_size_run"]
    return config


def run(config, console, jsonl_path=None):
    config = args_sanity_check(dict(config), console)
    args = SimpleNamespace(**config)
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.device = "cuda:%d" % local_rank
    args.rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(local_rank)
    # torchrun: one process per GPU.  Env instances and replay episodes are sharded by rank, the learner's ONE all-reduce
    # joins the gradients (parallel.py); only rank 0 writes stats and checkpoints.  gloo is used on a CUDA-less test box.
    args.rank, args.world_size = parallel.init_distributed(device=args.device)
    if args.world_size > 1 and args.batch_size % args.world_size:
        raise ValueError("batch_size=%d is not divisible by WORLD_SIZE=%d (episodes are sharded over the ranks)"
                         % (args.batch_size, args.world_size))
    logger = Logger(console, jsonl_path if args.rank == 0 else None)
    console.info("Experiment Parameters:\n\n" + pprint.pformat(config, indent=4, width=1) + "\n")
    args.unique_token = "{}__{}".format(getattr(args, "name", "run"), datetime.datetime.now().strftime("%Y-%m-%d_%H-%M-%S"))
    run_sequential(args, logger)
    console.info("Exiting Main")
    if args.world_size > 1:
        # every rank is past its last collective: leave together without tearing the communicator down (ncclCommDestroy has
        # been seen to block behind captured graphs / side streams on this stack; the reference ends with os._exit too,
        # src/main.py:37)
        import sys
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def evaluate_sequential(args, runner):
    for _ in range(max(1, args.test_nepisode // runner.batch_size)):
        runner.run(test_mode=True)
    runner.close_env()


def run_sequential(args, logger):
    args.entity_scheme = bool(args.env_args.get("entity_scheme", False))
    if not args.entity_scheme:
        raise NotImplementedError("only the entity scheme is implemented (every shipped algorithm uses it)")
    runner = r_REGISTRY[args.runner](args=args, logger=logger)
    env_info = runner.get_env_info()
    args.n_agents, args.n_actions = env_info["n_agents"], env_info["n_actions"]
    args.entity_shape, args.n_entities = env_info["entity_shape"], env_info["n_entities"]
    args.gt_mask_avail = env_info.get("gt_mask_avail", False)
    scheme = {
        "entities": {"vshape": env_info["entity_shape"], "group": "entities"},
        "obs_mask": {"vshape": env_info["n_entities"], "group": "entities", "dtype": torch.uint8},
        "entity_mask": {"vshape": env_info["n_entities"], "dtype": torch.uint8},
        "actions": {"vshape": (1,), "group": "agents", "dtype": torch.long},
        "avail_actions": {"vshape": (env_info["n_actions"],), "group": "agents", "dtype": torch.int},
        "reward": {"vshape": (1,)},
        "terminated": {"vshape": (1,), "dtype": torch.uint8},
    }
    if args.gt_mask_avail:
        scheme["gt_mask"] = {"vshape": env_info["n_entities"], "group": "agents", "dtype": torch.uint8}
    groups = {"agents": args.n_agents, "entities": args.n_entities}
    preprocess = {"actions": ("actions_onehot", [OneHot(out_dim=args.n_actions)])}

    world = int(getattr(args, "world_size", 1))
    local_bs = args.batch_size // world                       # replay is sharded by episode: each rank samples B / world
    buffer = ReplayBuffer(scheme, groups, max(args.buffer_size // world, local_bs), env_info["episode_limit"] + 1,
                          preprocess=preprocess, device=args.device)   # device-resident: fed by the env kernel, no host hop
    mac = mac_REGISTRY[args.mac](buffer.scheme, groups, args)
    runner.setup(scheme=scheme, groups=groups, preprocess=preprocess, mac=mac)
    learner = le_REGISTRY[args.learner](mac, buffer.scheme, logger, args)
    if args.use_cuda:
        learner.cuda()
    learner.sync_replicas()                                   # every replica starts from rank 0's parameters

    if args.checkpoint_path != "":
        if not os.path.isdir(args.checkpoint_path):
            logger.console_logger.info("Checkpoint directory {} doesn't exist".format(args.checkpoint_path))
            return
        timesteps = [int(n) for n in os.listdir(args.checkpoint_path)
                     if os.path.isdir(os.path.join(args.checkpoint_path, n)) and n.isdigit()]
        if not timesteps:
            logger.console_logger.info("No checkpoints under {}".format(args.checkpoint_path))
            return
        step = max(timesteps) if args.load_step == 0 else min(timesteps, key=lambda x: abs(x - args.load_step))
        model_path = os.path.join(args.checkpoint_path, str(step))
        logger.console_logger.info("Loading model from {}".format(model_path))
        learner.load_models(model_path, evaluate=args.evaluate)
        learner.sync_replicas()
        runner.t_env = step
        if args.evaluate or args.save_replay:
            evaluate_sequential(args, runner)
            return

    episode, last_test_T, last_log_T, model_save_time = 0, -args.test_interval - 1, 0, 0
    start_time = time.time()
    logger.console_logger.info("Beginning training for {} timesteps".format(args.t_max))
    while runner.t_env <= args.t_max:
        episode_batch = runner.run(test_mode=False)
        buffer.insert_episode_batch(episode_batch)
        if buffer.can_sample(local_bs):                        # same insert count on every rank -> same decision
            for _ in range(args.training_iters):
                sample = buffer.sample(local_bs)
                # truncate to the longest filled episode (run.py:269-270) -- over ALL ranks, rounded up to a multiple of 8
                # steps so that the learner sees a handful of shapes (workspaces, CUDA graphs); the extra steps are
                # unfilled and carry zero loss weight
                max_ep_t = parallel.all_reduce_max_int(int(sample.max_t_filled()), args.device)
                max_ep_t = min((max_ep_t + 7) // 8 * 8, sample.max_seq_length)
                sample = sample[:, :max_ep_t]
                learner.train(sample, runner.t_env, episode)
        n_test_runs = max(1, args.test_nepisode // runner.batch_size)
        if (runner.t_env - last_test_T) / args.test_interval >= 1.0:
            logger.console_logger.info("t_env: {} / {}  (time passed: {:.0f} s)".format(runner.t_env, args.t_max,
                                                                                      time.time() - start_time))
            last_test_T = runner.t_env
            for _ in range(n_test_runs):
                runner.run(test_mode=True)
        if args.save_model and (runner.t_env - model_save_time >= args.save_model_interval or model_save_time == 0
                                or runner.t_env > args.t_max):
            model_save_time = runner.t_env
            if args.rank == 0:                                  # replicas are identical: one writer
                save_path = os.path.join(args.local_results_path, "models", args.unique_token, str(runner.t_env))
                os.makedirs(save_path, exist_ok=True)
                logger.console_logger.info("Saving models to {}".format(save_path))
                learner.save_models(save_path)
        episode += args.batch_size_run * world
        if (runner.t_env - last_log_T) >= args.log_interval:
            logger.log_stat("episode", episode, runner.t_env)
            logger.print_recent_stats()
            last_log_T = runner.t_env
    runner.close_env()
    logger.console_logger.info("Finished Training")
    return logger
