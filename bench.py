#!/usr/bin/env python
"""Benchmark of the REFIL hot path on B200 (contract: see the task's bench.py section).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ns|ns-strong|cfg5|cfg3|gm]

One "step" = one QLearner.train pass (online + target forward, TD loss, hand-written backward, clip + RMSprop, and the
gradient all-reduce when N > 1) over one synthetic replay batch.  Default workload = BASELINE.json's north-star shape
(REFIL, B=128 episodes per GPU, T=60, 8 agents, 24 entities, d=128).  `value` = learner transitions/s with the batch
resident in HBM; `e2e` = the same through the public API with the batch in pinned HOST memory (H2D copy of every input
tensor + D2H read of the loss inside the timed region).  The Group Matching env kernel is measured in the same run and
reported under "env" (kernel alone, and full rollouts through ParallelRunner.run for BASELINE configs 2 and 4).  The headline
line is WEAK scaling (B=128 per GPU); the "strong" sub-record re-runs the named GLOBAL batch sharded B/N per rank for the
north-star shape and for BASELINE config 5 (sc2 3-8csz replay shape: B=128, T=120, 8 agents, 16 entities).
`--impl reference` times the reference's OWN QLearner.train (sources staged by oracle/stage_ref.py into oracle/_ref, all
host threads; falls back to the torch-CPU restatement in oracle/ when nothing is staged) on the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (alg, B, T, na, ne, ed, A, scaling)   weak: B episodes PER GPU; strong: B episodes in total, B / N per GPU
    "ns": ("refil", 128, 60, 8, 24, 39, 14, "weak"),
    "ns-strong": ("refil", 128, 60, 8, 24, 39, 14, "strong"),
    "cfg5": ("refil", 128, 120, 8, 16, 39, 14, "strong"),       # BASELINE.json configs[4]: sc2custom 3-8csz replay shape
    "cfg3": ("qmix_atten", 64, 60, 8, 16, 39, 14, "weak"),
    "gm": ("refil_group_matching", 4096, 51, 4, 4, 12, 3, "weak"),
}


def make_args(alg, na, ne, ed, A):
    from types import SimpleNamespace
    a = dict(gamma=0.99, lr=0.0005, optim_alpha=0.99, optim_eps=0.00001, grad_norm_clip=10, weight_decay=0,
             double_q=True, lmbda=0.5, attn_n_heads=4, attn_embed_dim=128, hypernet_embed=128, mixing_embed_dim=32,
             rnn_hidden_dim=64, softmax_mixing_weights=True, entity_last_action=True, mixer="flex_qmix",
             agent="imagine_entity_attend_rnn", gt_obs_mask=False, train_gt_factors=False, train_rand_gt_factors=False,
             test_gt_factors=False, pooling_type=None, epsilon_start=1.0, epsilon_finish=0.05,
             epsilon_anneal_time=500000, mac="entity_mac", learner="q_learner", agent_output_type="q",
             action_selector="epsilon_greedy", target_update_interval=200, learner_log_interval=10 ** 12,
             gt_mask_avail=False, n_agents=na, n_actions=A, n_entities=ne, entity_shape=ed)
    if alg == "qmix_atten":
        a.update(agent="entity_attend_rnn")
    elif alg == "refil_group_matching":
        a.update(agent="imagine_entity_attend_ff", mixer="lin_flex_qmix", attn_embed_dim=64, hypernet_embed=64,
                 entity_last_action=False)
    return SimpleNamespace(**a)


class ClockSampler:
    """SM clocks / throttle reasons / power during the timed region (B200_PROFILING.md recipe), sampled every 20 ms through NVML
    from a thread of this process (nvidia_ml_py: no fork, no nvidia-smi start-up); falls back to one long-lived
    `nvidia-smi -lms` child when NVML cannot be imported."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, enabled=True):
        self.index, self.proc, self.enabled = index, None, enabled
        self.samples, self.stop_flag, self.thread, self.max_mhz = [], False, None, None

    def _nvml_loop(self, nv, h):
        bits = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.samples.append((float(sm), pw, [n for n, b in bits if rs & b]))
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if not self.enabled:
            return
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].strip().isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def summary(self):
        if not self.enabled:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampled on rank 0 only"]}
        samples = []
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            samples = list(self.samples)
        elif self.proc is not None:
            try:
                self.proc.terminate()
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                out = ""
            for line in out.strip().splitlines():
                f = [x.strip() for x in line.split(",")]
                if len(f) >= 9:
                    try:
                        self.max_mhz = float(f[2])
                        samples.append((float(f[1]), float(f[3]), [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown",
                                        "sw_thermal_slowdown", "sw_power_cap"), f[5:9]) if v.lower().startswith("active")]))
                    except ValueError:
                        pass
        if not samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (NVML and nvidia-smi unavailable)"]}
        pmax = max(p for _, p, _ in samples)
        load = [x for x in samples if x[1] >= 0.6 * pmax] or samples          # samples taken under load
        sm = sorted(x[0] for x in load)
        reasons = sorted({r for x in load for r in x[2]})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(samples),
                "samples_under_load": len(load), "power_w_max": pmax,
                "source": "nvml" if self.thread is not None else "nvidia-smi"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


# --------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own code (oracle/_ref, staged by oracle/stage_ref.py) or, when nothing is staged, the torch-CPU
# restatement (oracle/learner_oracle.py).  bench.py is one of the three places allowed to execute oracle/.
def _syn_for(alg, B, T, na, ne, ed, A, seed):
    from refil_b200.utils.synthetic import synthetic_replay
    gm = alg == "refil_group_matching"
    return synthetic_replay(B, T, na, ne, ed, A, seed=seed, gt_mask=gm, pad=ne > na)


class CpuLearnerArm:
    """One QLearner.train step of the reference on the host cores, on B episodes of a workload."""

    def __init__(self, alg, B, T, na, ne, ed, A, seed=0):
        import torch
        from oracle import stage_ref
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.B, self.T = B, T
        syn = _syn_for(alg, B, T, na, ne, ed, A, seed)
        gm = alg == "refil_group_matching"
        if stage_ref.available() and not os.environ.get("REFIL_REF_FORCE_PORT"):
            self.kind = "reference"
            self.learner, self.batch, _ = stage_ref.build_reference_learner(
                alg, syn, (B, T, na, ne, ed, A), gt=gm, seed=seed, learner_log_interval=10 ** 12)
            self.what = "the reference's own EntityMAC + QLearner.train (oracle/_ref, staged from /root/reference/src)"
            self._i = 0
        else:
            from oracle import learner_oracle as lo
            self.kind = "port"
            self.lo, self.args = lo, make_args(alg, na, ne, ed, A)
            gen = torch.Generator().manual_seed(seed)
            ein = ed + (A if self.args.entity_last_action else 0)
            self.syn = syn
            self.ap, self.mp = lo.init_agent_params(gen, self.args, ein), lo.init_mixer_params(gen, self.args, ein)
            self.group_a = (torch.rand(B, ne, generator=gen) < 0.5).to(torch.uint8)
            self.what = "torch-CPU restatement of QLearner.train (oracle/learner_oracle.py)"

    def step(self):
        if self.kind == "reference":
            self._i += 1
            self.learner.train(self.batch, t_env=self._i, episode_num=self._i)
        else:
            self.lo.train_step(self.ap, self.mp, self.ap, self.mp, self.syn, self.args, group_a=self.group_a)

    def transitions(self):
        return self.B * (self.T - 1)


def cpu_reference_rate(alg, T, na, ne, ed, A, budget_s=20.0, sample_B=16, seed=0):
    """cpu_baseline leg of our arm: a BOUNDED sample (sample_B episodes, ~budget_s seconds) on rank 0 at N=1."""
    arm = CpuLearnerArm(alg, sample_B, T, na, ne, ed, A, seed)
    arm.step()                                                   # warm-up
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < 3 or (time.perf_counter() < t_end and len(times) < 30):
        t0 = time.perf_counter()
        arm.step()
        times.append(time.perf_counter() - t0)
    med = sorted(times)[len(times) // 2]
    return {"value": arm.transitions() / med, "unit": "transitions/s", "cores": arm.cores, "kind": arm.kind,
            "sample": "%d timed steps (median) of %s on B=%d episodes of the same (T=%d, agents=%d, entities=%d) workload, "
                      "%d torch threads" % (len(times), arm.what, sample_B, T, na, ne, arm.cores),
            "ms_per_step": med * 1e3}


def cpu_env_rate(na, budget_s=5.0):
    """Group Matching on one host core: the reference class itself when staged (step + get_avail_actions + get_masks +
    get_entities per step = what env_worker does, parallel_runner.py:253-280), else the C restatement."""
    from oracle import stage_ref
    cfg = dict(n_agents=na, n_states=6, n_groups=2, rand_trans=0.1, episode_limit=50)
    if stage_ref.available():
        import numpy as np
        env = stage_ref.group_matching_env(seed=0, **cfg)
        rng = np.random.RandomState(0)
        env.reset()
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < budget_s:
            for _ in range(200):
                _, done, _ = env.step(rng.randint(0, 3, size=na))
                env.get_avail_actions(); env.get_masks(); env.get_entities()
                n += 1
                if done:
                    env.reset()
        dt = time.perf_counter() - t0
        return {"value": n / dt, "unit": "env-steps/s", "cores": 1, "kind": "reference",
                "sample": "%d steps of the reference GroupMatching class (step + avail + masks + entities, reset on done), "
                          "1 process; the reference runs one such process per env" % n}
    from oracle.gm_env_oracle import GroupMatchingOracle
    env = GroupMatchingOracle(seed=0, **cfg)
    env.reset()
    t0 = time.perf_counter()
    done = env.bench_loop(200000)
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": "env-steps/s", "cores": 1, "kind": "port",
            "sample": "%d steps of the C oracle (step + entities + masks), 1 thread" % done}


# --------------------------------------------------------------------------------------------------------------------
def run_reference(a):
    """`--impl reference`: the reference's own CPU path on this box's host cores, same workload / metric / unit.  Each step is
    the workload's own batch when that fits the time budget, else the largest power-of-two sample of it that does (a
    calibration step decides; config.sample_B says which)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    alg, B, T, na, ne, ed, A, scaling = WORKLOADS[a.workload]
    if a.batch:
        B = a.batch
    budget_s = float(os.environ.get("REFIL_REF_BUDGET_S", "110"))
    n_steps = a.steps + max(1, min(a.warmup, 1))
    # calibration on a small sample: seconds per episode-step of this workload on this box
    cal_B = min(B, 8 if a.workload != "gm" else 128)
    cal = CpuLearnerArm(alg, cal_B, T, na, ne, ed, A)
    cal.step()
    t0 = time.perf_counter()
    cal.step()
    per_ep = (time.perf_counter() - t0) / cal_B
    sample_B = B
    while sample_B > cal_B and per_ep * sample_B * n_steps > budget_s:
        sample_B //= 2
    arm = cal if sample_B == cal_B else CpuLearnerArm(alg, sample_B, T, na, ne, ed, A)
    del cal
    for _ in range(max(1, min(a.warmup, 1))):
        arm.step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        arm.step()
    dt = (time.perf_counter() - t0) / a.steps
    v = arm.transitions() / dt
    sample = ("each step = %s on B=%d episodes of the (T=%d, agents=%d, entities=%d) workload, %d torch threads"
              % (arm.what, sample_B, T, na, ne, arm.cores))
    print(json.dumps({
        "impl": "reference", "metric": "learner transitions/sec", "value": v, "unit": "transitions/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a.workload, B, T, na, ne), "sample_B": sample_B, "same_batch": sample_B == B,
                   "note": "CPU arm runs on rank 0's host cores only, whatever --gpus says; a step is the workload's own "
                           "batch when (steps + warm-up) x step time fits %.0f s, else a power-of-two sample of it" % budget_s},
        "cpu_baseline": {"value": v, "unit": "transitions/s", "cores": arm.cores, "kind": arm.kind, "sample": sample},
        "e2e": {"value": v, "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_name(w, B, T, na, ne):
    alg, scaling = WORKLOADS[w][0], WORKLOADS[w][7]
    return "%s learner step, synthetic replay (B=%d %s, T=%d, agents=%d, entities=%d, d=%d)" % (
        alg, B, "per GPU" if scaling == "weak" else "in total, sharded over the GPUs", T, na, ne, 64 if w == "gm" else 128)


class _Log:
    class console_logger:
        @staticmethod
        def info(*x, **k):
            pass

    def log_stat(self, *x, **k):
        pass


# --------------------------------------------------------------------------------------------------------------------
def bench_learner(w, B, a, dev, rank, world, do_e2e=True, do_breakdown=False, steps=None):
    """Time `steps` QLearner.train steps of workload `w` with B episodes on THIS rank (max over ranks).  -> dict."""
    import torch
    import torch.distributed as dist
    from refil_b200 import ops
    from refil_b200.components.episode_buffer import EpisodeBatch
    from refil_b200.controllers import REGISTRY as mac_REGISTRY
    from refil_b200.learners import REGISTRY as le_REGISTRY
    from refil_b200.utils.synthetic import entity_scheme, synthetic_replay

    alg, _, T, na, ne, ed, A, scaling = WORKLOADS[w]
    steps = steps or a.steps
    args = make_args(alg, na, ne, ed, A)
    args.device = dev
    args.cuda_graph = not a.no_graph
    args.group_hypernets = False if a.no_group else (True if a.group else "auto")
    if a.crit_tiles >= 0:
        args.critical_min_tiles = a.crit_tiles
    args.critical_priority = not a.no_prio
    gm = w == "gm"
    scheme, groups, preprocess = entity_scheme(na, ne, ed, A, gt_mask=gm)
    syn = synthetic_replay(B, T, na, ne, ed, A, seed=1000 + rank, gt_mask=gm, pad=ne > na)
    args.gt_mask_avail = gm
    batch = EpisodeBatch(scheme, groups, B, T, preprocess=preprocess, device=dev)
    host = {k: v.pin_memory() for k, v in syn.items()}
    for k, v in host.items():
        batch.data.transition_data[k].copy_(v)
    torch.manual_seed(0)                       # identical initial weights on every rank
    mac = mac_REGISTRY[args.mac](batch.scheme, groups, args)
    learner = le_REGISTRY[args.learner](mac, batch.scheme, _Log(), args)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for i in range(n):
            fn(i)
        host_ms[0] = (time.perf_counter() - t0) * 1e3 / n      # host time to ENQUEUE one step (no sync inside fn)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            if os.environ.get("REFIL_BENCH_DEBUG"):
                print("rank %d: %.3f ms for %d steps" % (rank, float(ms.item()), n), file=sys.stderr)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    ep = [0]

    def step_resident(i):
        ep[0] += 1
        learner.train(batch, t_env=ep[0], episode_num=ep[0])

    h2d = sum(v.numel() * v.element_size() for v in host.values())
    for i in range(max(a.warmup, 3)):
        step_resident(i)
    l0 = ops.launch_count()
    ms = timed(step_resident, steps)
    res = {"workload": workload_name(w, B * world if scaling == "strong" else B, T, na, ne), "episodes_per_gpu": B,
           "ms_per_step": ms / steps, "host_enqueue_ms_per_step": round(host_ms[0], 3),
           "gpu_launches": ops.launch_count() - l0, "launches_per_step": (ops.launch_count() - l0) / steps,
           "transitions_per_step": world * B * (T - 1), "value": world * B * (T - 1) * steps / (ms * 1e-3),
           "activation_bytes": learner_bytes(learner, mac), "args": args, "dims": (B, T, na, ne, ed, A)}

    if do_e2e:
        # end to end through the public API: every step's inputs come from pinned host memory and its loss statistics go back
        # to the host.  Two device-side EpisodeBatch buffers: the H2D copy of step i+1 runs on a copy stream while step i trains.
        batch_b = EpisodeBatch(scheme, groups, B, T, preprocess=preprocess, device=dev)
        bufs = [batch, batch_b]
        copy_stream = torch.cuda.Stream(device=dev)
        ev_copied = [torch.cuda.Event(), torch.cuda.Event()]
        ev_trained = [torch.cuda.Event(), torch.cuda.Event()]

        def copy_async(slot):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev_trained[slot])           # the step that last read this buffer is done
                for k, v in host.items():
                    bufs[slot].data.transition_data[k].copy_(v, non_blocking=True)
                ev_copied[slot].record(copy_stream)

        def step_e2e(i):
            slot = i % 2
            if i == 0:
                copy_async(0)
            torch.cuda.current_stream().wait_event(ev_copied[slot])
            ep[0] += 1
            learner.train(bufs[slot], t_env=ep[0], episode_num=ep[0])
            ev_trained[slot].record()
            copy_async(slot ^ 1)                                   # next step's inputs, overlapped with this step's kernels
            learner.gradbuf[learner.n_params:].cpu()          # D2H read of the step's loss statistics (syncs)

        for i in range(2):
            step_e2e(i)
        torch.cuda.synchronize()
        ms_e2e = timed(step_e2e, steps)
        res["e2e"] = {"value": world * B * (T - 1) * steps / (ms_e2e * 1e-3), "unit": "transitions/s",
                      "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 32, "ms_per_step": ms_e2e / steps}

    if do_breakdown:
        # per-kernel breakdown (instrumented repeat of the step, events around every launch)
        args.concurrent_streams = False          # serialised launches: per-kernel durations are not inflated by overlap
        args.cuda_graph = False                  # ... and enqueued one by one so that events can bracket each of them
        ops.set_timing(True)
        step_resident(0)
        step_resident(1)
        tsum = ops.timing_summary()
        shapes = ops.shape_timing_summary()
        ops.set_timing(False)
        res["kern"] = {k: {"launches": v[0] // 2, "ms_per_step": v[1] / 2, "flop_per_step": v[2] / 2,
                           "bytes_per_step": v[3] / 2} for k, v in tsum.items()}
        res["shapes"] = shapes
    del learner, mac, batch
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return res


def run_ours(a):
    import torch
    import torch.distributed as dist

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")               # those levels print the NCCL banner on stdout: keep it to the JSON line
        from refil_b200 import parallel
        parallel.init_distributed(backend="nccl", device=dev)
    alg, B, T, na, ne, ed, A, scaling = WORKLOADS[a.workload]
    if a.batch:
        B = a.batch
    if scaling == "strong":
        if B % world:
            raise SystemExit("workload %s: global batch %d is not divisible by %d GPUs" % (a.workload, B, world))
        B //= world
    clocks = ClockSampler(local, enabled=(rank == 0))
    clocks.start()
    head = bench_learner(a.workload, B, a, dev, rank, world, do_e2e=True, do_breakdown=True)
    clk = clocks.summary()           # sampled across both timed regions (resident + e2e)
    args = head["args"]
    kern, shapes = head["kern"], head["shapes"]
    hbm, tf_burst, tf_sus, src = peaks()
    timing_note = "per-launch CUDA events on an instrumented repeat of the step (2 steps averaged)"

    def hbm_roof(name, label):
        k = kern.get(name)
        if not k or not k["ms_per_step"]:
            return None
        ach = k["bytes_per_step"] / (k["ms_per_step"] * 1e-3) / 1e9
        return {"kernel": label, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                "traffic": None, "launches_per_step": k["launches"], "ms_per_step": k["ms_per_step"],
                "algorithmic_bytes_per_step": k["bytes_per_step"], "algorithmic_tflops": k["flop_per_step"] / (k["ms_per_step"] * 1e-3) / 1e12,
                "peak_source": src, "timing": timing_note}

    roof_att = hbm_roof("masked_attn_fwd", "attn_fwd_kernel (K2 masked MHA; bytes = 4d(2ne+2na)+na*ne per (b,t,copy))")
    # dominant kernel of the step: the 3xTF32 tcgen05 GEMM behind every dense layer (forward + backward-data)
    roofline = hbm_roof("tc_gemm_tn", "tc_gemm_ts_kernel (K1 dense layers, tcgen05 kind::tf32 x3 with the A operand in tensor memory, fp32 accumulate in TMEM; "
                        "bytes = 4(M*K + M*N + N*K) per launch)")
    if roofline is not None:
        # DRAM traffic per launch of the same kernel from the committed ncu capture (never measured under this run)
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            roofline["traffic"] = tj["traffic_bytes_per_launch"]
            roofline["traffic_source"] = tj["source"]
            roofline["algorithmic_bytes_per_launch"] = roofline["algorithmic_bytes_per_step"] / max(roofline["launches_per_step"], 1)
        roofline["tensor_frac_of_bf16_peak"] = 3.0 * roofline["algorithmic_tflops"] / tf_sus   # 3 MMAs per product
        roofline["note"] = ("memory-bound by construction (48 flop/B at N=384,K=128); tensor-side: 3 tf32 MMAs per "
                            "product, quoted against the measured bf16 sustained peak for context")
        # whole-step view (fused K1 accounting): what the step HAS to move if every trunk were one fused kernel
        roofline["step_algorithmic_bytes_all_kernels"] = sum(k["bytes_per_step"] for k in kern.values())
        roofline["step_bytes_over_hbm_peak_ms"] = roofline["step_algorithmic_bytes_all_kernels"] / (hbm * 1e9) * 1e3
    dom = max(kern.items(), key=lambda kv: kv[1]["ms_per_step"])

    out = {
        "metric": "learner transitions/sec", "value": head["value"], "unit": "transitions/s", "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": head["workload"], "global_batch_episodes": world * B,
                   "parallelism": "dp%d (episodes sharded, one all-reduce of grads+stats)" % world,
                   "step_submission": "cuda graph replay" if not a.no_graph else "eager launches",
                   "l2": "working set per step (%.1f GB of activations) exceeds the 126 MB L2" % (head["activation_bytes"] / 1e9)},
        "e2e": head["e2e"],
        "gpu_launches": head["gpu_launches"], "launches_per_step": head["launches_per_step"],
        "host_enqueue_ms_per_step": head["host_enqueue_ms_per_step"], "clocks": clk, "roofline": roofline, "roofline_attention": roof_att,
        "roofline_wgrad": hbm_roof("tc_gemm_wgrad", "tc_wgrad_ts_kernel (weight gradients: X^T in tensor memory, MN-major Y tiles)"),
        "roofline_attention_bwd": hbm_roof("masked_attn_bwd", "attn_bwd_kernel"), "dominant_kernel": dom[0],
        "kernels_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms_per_step"])},
        # dense-layer kernels by shape [M x N x K]: (launches per step, us per launch)
        "dense_shapes_us": {k: [v[0] // 2, round(1e3 * v[1] / max(v[0], 1), 1)] for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][1])},
    }
    # ---- strong scaling of the NAMED global batches (every rank takes part; B / N episodes per GPU) -----------------
    if a.workload == "ns" and not a.no_extra:
        strong = {}
        for w in ("ns-strong", "cfg5"):
            gB = WORKLOADS[w][1]
            if gB % world:
                strong[w] = {"error": "global batch %d not divisible by %d GPUs" % (gB, world)}
                continue
            if w == "ns-strong" and world == 1:          # one GPU: the strong and the weak workload are the same run
                r = head
            else:
                r = bench_learner(w, gB // world, a, dev, rank, world, do_e2e=False, do_breakdown=False, steps=min(a.steps, 10))
            strong[w] = {"workload": r["workload"], "value": r["value"], "unit": "transitions/s", "scaling": "strong",
                         "episodes_per_gpu": r["episodes_per_gpu"], "ms_per_step": r["ms_per_step"],
                         "launches_per_step": r["launches_per_step"], "host_enqueue_ms_per_step": r["host_enqueue_ms_per_step"]}
        out["strong"] = strong
    # ---- env kernel in the same run ---------------------------------------------------------------------------
    if rank == 0 or world > 1:
        out["env"] = bench_env(dev, world, rank, a, hbm, src)
    if not a.no_extra and a.workload == "ns":
        try:          # BASELINE configs[3]: group_matching 8 agents / 24 entity slots, REFIL (RNN agent), 4096 envs PER GPU
            out["env"]["rollout_cfg4"] = bench_rollout(dev, rank, world, "refil", 4096, 8, 24)
        except Exception as e:                    # the rollout legs are informational: never lose the bench line over them
            out["env"]["rollout_cfg4"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank == 0 and world == 1 and not a.no_extra:
        try:          # BASELINE configs[1]: group_matching 4 agents, refil_group_matching, 4096 envs on one GPU
            out["env"]["rollout"] = bench_rollout(dev, rank, world, "refil_group_matching", 4096, 4, 4)
        except Exception as e:
            out["env"]["rollout"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank == 0 and world == 1 and a.workload == "ns" and not a.no_extra:
        # the other BASELINE.json learner configs, each in its own process (own workspaces): configs[1]'s learner side
        # (group_matching 4 agents, refil_group_matching, 4096 episodes of 4096 parallel envs) and configs[2] (qmix_atten)
        out["other_workloads"] = {}
        for w in ("gm", "cfg3"):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", w, "--steps", "5", "--warmup", "3",
                                    "--no-cpu", "--no-extra"], capture_output=True, text=True, timeout=240)
                line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
                d = json.loads(line)
                out["other_workloads"][w] = {"workload": d["config"]["workload"], "value": d["value"], "unit": d["unit"],
                                             "ms_per_step": d["ms_per_step"], "e2e": d["e2e"]["value"]}
            except Exception as e:
                out["other_workloads"][w] = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank == 0 and world == 1 and not a.no_cpu:
        out["cpu_baseline"] = cpu_reference_rate(alg, T, na, ne, ed, A, sample_B=16 if a.workload != "gm" else 256)
        out["env"]["cpu_baseline"] = cpu_env_rate(8)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        # every rank is past its last collective (bench_env all-reduces): leave without tearing the communicator down --
        # ncclCommDestroy has been seen to block for minutes behind captured graphs / side streams on this stack
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def learner_bytes(learner, mac):
    n = learner.ws.nbytes() + mac.agent.ws.nbytes() + learner.target_mac.agent.ws.nbytes()
    if learner.mixer is not None:
        n += learner.mixer.ws.nbytes() + learner.target_mixer.ws.nbytes()
    return n


def step_flops(args, B, T, na, ne, ed, A):
    """Dense-layer FLOPs actually executed per train step (de-duplicated schedule: fc1/in_trans once per net)."""
    N = B * T
    d, he, me, r = args.attn_embed_dim, args.hypernet_embed, args.mixing_embed_dim, args.rnn_hidden_dim
    ein = ed + (A if args.entity_last_action else 0)
    rnn = "rnn" in args.agent
    C = 3 if "imagine" in args.agent else 1

    def trunk(dm, copies):
        return 2 * ne * ein * dm + 2 * ne * dm * 3 * dm + copies * 2 * na * dm * dm

    def head(copies):
        if rnn:
            return copies * (2 * na * d * r + 2 * na * r * 3 * r + 2 * na * r * A)
        return copies * 2 * na * d * A
    agent_f = trunk(d, C) + head(C)
    tgt_f = trunk(d, 1) + head(1)
    n_h = {"flex_qmix": 4, "lin_flex_qmix": 2, "vdn": 0}[args.mixer]
    mix_f = tgt_mix_f = 0
    if n_h:
        w1_copies = 3 if C == 3 else 1
        mix_f = trunk(he, w1_copies) + w1_copies * 2 * na * he * me + (n_h - 1) * (trunk(he, 1) + 2 * na * he * me)
        tgt_mix_f = n_h * (trunk(he, 1) + 2 * na * he * me)
    return N * (3 * (agent_f + mix_f) + tgt_f + tgt_mix_f)       # fwd + 2x for backward on the online nets


def bench_env(dev, world, rank, a, hbm, src):
    import torch
    import torch.distributed as dist
    from refil_b200.envs.group_matching import GroupMatchingBatch
    na, ns, ng, lim = 8, 6, 2, 50
    E = 32768 // max(world, 1) if world > 1 else 32768
    T = lim + 1
    env = GroupMatchingBatch(E, n_agents=na, n_states=ns, n_groups=ng, rand_trans=0.1, episode_limit=lim, seed=0,
                             device=dev, first_env_index=rank * E)
    ed = env.get_entity_size()
    b = dict(entities=torch.zeros(E, T, na, ed, device=dev), gt_mask=torch.zeros(E, T, na, na, dtype=torch.uint8, device=dev),
             obs_mask=torch.zeros(E, T, na, na, dtype=torch.uint8, device=dev),
             entity_mask=torch.zeros(E, T, na, dtype=torch.uint8, device=dev),
             avail_actions=torch.zeros(E, T, na, 3, dtype=torch.int32, device=dev),
             actions=torch.randint(0, 3, (E, T, na, 1), device=dev), reward=torch.zeros(E, T, 1, device=dev),
             terminated=torch.zeros(E, T, 1, dtype=torch.uint8, device=dev),
             filled=torch.zeros(E, T, 1, dtype=torch.int64, device=dev))

    def rollout():
        env.reset(b)
        for ts in range(lim):
            env.step(b, ts)

    for _ in range(3):
        rollout()
    torch.cuda.synchronize()
    # the 51 launches of a rollout are replayed as ONE CUDA graph: enqueueing them from Python costs ~70 us each (ctypes call with
    # 27 arguments), several times the kernel itself -- the eager figure measured the host, not the kernel
    graph = None
    if not a.no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                rollout()
        except Exception:
            graph = None
            torch.cuda.synchronize()
    run = graph.replay if graph is not None else rollout
    for _ in range(2):
        run()
    env.step_counter.zero_()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    steps = env.step_counter.clone().double()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(steps)
    rate = float(steps.item()) / (float(ms.item()) * 1e-3)
    bytes_per_step = 16 * na + 22 + 4 * na * ed + 25.6 * na
    ach = rate * bytes_per_step / 1e9
    return {"metric": "GroupMatching env-steps/sec", "value": rate, "unit": "env-steps/s", "n_envs": E * world,
            "n_agents": na, "ms_per_rollout": float(ms.item()) / reps, "env_steps_per_rollout": float(steps.item()) / reps,
            "config": "group_matching 8 agents, 6 states, 2 groups, rand_trans 0.1, limit 50; random action tape; "
                      "reset + 50 step launches per rollout (%s), rollout tensors written in place"
                      % ("one CUDA graph replay per rollout" if graph is not None else "eager launches"),
            "roofline": {"kernel": "gm_step_kernel (K0)", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s",
                         "frac": ach / hbm, "traffic": None, "bytes_per_env_step": bytes_per_step, "peak_source": src}}


def bench_rollout(dev, rank, world, alg, n_envs, n_agents, n_entities):
    """Full on-device rollouts through the reference runner interface (ParallelRunner.run: agent forward + epsilon-greedy
    selection + env kernel + EpisodeBatch writes per timestep), `n_envs` instances PER GPU, env instances sharded by index with
    no collective.  BASELINE.json configs[1] = (refil_group_matching, 4 agents, FF imagine agent, d=64);
    configs[3] = (refil, 8 agents in 24 entity slots, RNN agent + flex_qmix dims of refil.yaml, d=128)."""
    import torch
    import torch.distributed as dist
    from types import SimpleNamespace
    from refil_b200.components.episode_buffer import EpisodeBatch
    from refil_b200.config import build_config
    from refil_b200.controllers import REGISTRY as mac_REGISTRY
    from refil_b200.runners import REGISTRY as r_REGISTRY
    from refil_b200.utils.synthetic import entity_scheme

    over = ["batch_size_run=%d" % n_envs, "env_args.n_agents=%d" % n_agents]
    if n_entities != n_agents:
        over.append("env_args.n_entities=%d" % n_entities)
    cfg = build_config("group_matching", alg, over)
    cfg["env_args"]["seed"] = 0
    args = SimpleNamespace(**cfg)
    args.device, args.rank = dev, rank
    args.rollout_graph = not bool(os.environ.get("REFIL_NO_ROLLOUT_GRAPH"))
    runner = r_REGISTRY[args.runner](args=args, logger=_Log())
    info = runner.get_env_info()
    args.n_agents, args.n_actions, args.entity_shape, args.n_entities = (info["n_agents"], info["n_actions"],
                                                                         info["entity_shape"], info["n_entities"])
    args.gt_mask_avail, args.entity_scheme = True, True
    scheme, groups, preprocess = entity_scheme(args.n_agents, args.n_entities, args.entity_shape, args.n_actions, gt_mask=True)
    proto = EpisodeBatch(scheme, groups, 1, 2, preprocess=preprocess, device=dev)
    torch.manual_seed(0)
    mac = mac_REGISTRY[args.mac](proto.scheme, groups, args)
    runner.setup(scheme=scheme, groups=groups, preprocess=preprocess, mac=mac)
    for _ in range(3):
        runner.run(test_mode=False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0_env = runner.t_env
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        runner.run(test_mode=False)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    steps = runner.t_env - t0_env                # under torchrun the runner already sums env steps over the ranks
    del runner, mac
    torch.cuda.empty_cache()
    return {"metric": "GroupMatching env-steps/sec through ParallelRunner.run (agent forward + epsilon-greedy + env kernel)",
            "value": steps / (ms * 1e-3), "unit": "env-steps/s", "n_envs": n_envs * world, "n_envs_per_gpu": n_envs,
            "n_agents": n_agents, "n_entities": n_entities,
            "ms_per_rollout": ms / reps, "env_steps_per_rollout": steps / reps,
            "config": "group_matching %d agents / %d entity slots, %s, %d parallel envs per GPU x %d GPUs, epsilon-greedy "
                      "training rollouts" % (n_agents, n_entities, alg, n_envs, world)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ns", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="episodes per GPU (default: the workload's)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the other-workload and rollout legs")
    ap.add_argument("--crit-tiles", type=int, default=-1, help="m-tiles per CTA on the agent chain (QLearner args.critical_min_tiles; 0 = default)")
    ap.add_argument("--no-prio", action="store_true", help="agent chain on the default-priority stream")
    ap.add_argument("--group", action="store_true", help="force the lock-step grouped hypernetwork launches (default: auto by size)")
    ap.add_argument("--no-group", action="store_true", help="run the hypernetworks net by net on their own streams instead of "
                    "in lock-step with grouped launches (QLearner args.group_hypernets)")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from Python instead of replaying the "
                    "captured CUDA graph of the step (QLearner args.cuda_graph)")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
