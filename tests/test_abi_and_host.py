"""CPU-side checks: the C-ABI library loads and exports every symbol include/refil_b200.h declares (no compute calls),
the product path refuses to run without CUDA, and the host-side containers behave like the reference's."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from refil_b200 import _lib, build
    build.build()
    return _lib


def test_library_exports_every_declared_symbol(lib):
    protos = lib.parse_header()
    assert len(protos) >= 27
    so = ctypes.CDLL(lib.SO_PATH)
    for name in protos:
        assert hasattr(so, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", lib.SO_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("refil_")}
    assert exported == set(protos), exported ^ set(protos)
    L = lib.load()
    assert L.refil_abi_version() == 1
    assert L.refil_last_error() is not None


def test_sass_is_sm100a(lib):
    out = subprocess.run(["cuobjdump", "-lelf", lib.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out[:200]


def test_no_cpu_fallback():
    from refil_b200 import _lib, ops
    with pytest.raises(_lib.RefilError):
        ops.linear_fwd(torch.zeros(2, 2), torch.zeros(2, 2), None, torch.zeros(2, 2))
    from refil_b200.envs.group_matching import GroupMatchingBatch
    with pytest.raises(_lib.RefilError):
        GroupMatchingBatch(4, device="cpu")


def test_product_path_never_imports_oracle():
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "refil_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                if "import oracle" in src or "from oracle" in src:
                    bad.append(f)
    assert not bad, bad


def test_epsilon_schedule_matches_reference_formula():
    from refil_b200.components.epsilon_schedules import DecayThenFlatSchedule
    s = DecayThenFlatSchedule(1.0, 0.05, 500000, decay="linear")
    assert s.eval(0) == 1.0 and s.eval(10 ** 9) == 0.05 and abs(s.eval(250000) - 0.525) < 1e-12
    e = DecayThenFlatSchedule(1.0, 0.05, 1000, decay="exp")
    assert abs(e.eval(1000) - 0.05) < 1e-12 and e.eval(0) == 1.0


def _scheme():
    from refil_b200.components.transforms import OneHot
    scheme = {"entities": {"vshape": 3, "group": "entities"}, "actions": {"vshape": (1,), "group": "agents", "dtype": torch.long},
              "reward": {"vshape": (1,)}, "terminated": {"vshape": (1,), "dtype": torch.uint8}}
    return scheme, {"agents": 2, "entities": 4}, {"actions": ("actions_onehot", [OneHot(out_dim=5)])}


def test_episode_batch_update_slice_and_preprocess():
    from refil_b200.components.episode_buffer import EpisodeBatch
    scheme, groups, pre = _scheme()
    b = EpisodeBatch(scheme, groups, 3, 4, preprocess=pre, device="cpu")
    assert b["entities"].shape == (3, 4, 4, 3) and b["actions_onehot"].shape == (3, 4, 2, 5) and b["filled"].dtype == torch.long
    b.update({"actions": [[1, 4], [0, 2], [3, 3]]}, ts=1)
    assert int(b["filled"][:, 1].sum()) == 3 and int(b["filled"].sum()) == 3
    assert torch.equal(b["actions_onehot"][1, 1].argmax(-1), torch.tensor([0, 2]))
    b.update({"reward": [(0.5,)], "terminated": [(True,)]}, bs=[2], ts=2, mark_filled=False)
    assert float(b["reward"][2, 2, 0]) == 0.5 and int(b["terminated"][2, 2, 0]) == 1 and int(b["filled"][2, 2, 0]) == 0
    s = b[1:3, :2]
    assert s.batch_size == 2 and s.max_seq_length == 2 and s["actions"].shape == (2, 2, 2, 1)
    g = b[[0, 2]]
    assert g.batch_size == 2 and torch.equal(g["actions"][1], b["actions"][2])
    only = b[("reward", "terminated")]
    assert set(only.data.transition_data) == {"reward", "terminated"}
    assert int(b.max_t_filled()) == 1
    with pytest.raises(ValueError):
        b.update({"reward": np.zeros((3, 4, 2))})


def test_replay_buffer_ring_and_sampling():
    from refil_b200.components.episode_buffer import EpisodeBatch, ReplayBuffer
    scheme, groups, pre = _scheme()
    rb = ReplayBuffer(scheme, groups, 5, 4, preprocess=pre, device="cpu")
    assert not rb.can_sample(1)
    for i in range(4):
        ep = EpisodeBatch(scheme, groups, 2, 4, preprocess=pre, device="cpu")
        ep.update({"reward": torch.full((2, 4, 1), float(i))})
        rb.insert_episode_batch(ep)
    assert rb.episodes_in_buffer == 5 and rb.buffer_index == 3           # 8 inserts into a ring of 5
    assert [float(rb["reward"][k, 0, 0]) for k in range(5)] == [2.0, 3.0, 3.0, 1.0, 2.0]
    np.random.seed(0)
    s = rb.sample(3)
    assert s.batch_size == 3 and s["reward"].shape == (3, 4, 1)
    assert rb.sample(5).batch_size == 5


def test_param_store_views_and_state_dict():
    from collections import OrderedDict
    from refil_b200.modules.params import ParamStore
    a = ParamStore(OrderedDict([("w", (3, 5)), ("b", (3,))]), "cpu")
    a.p["w"].fill_(2.0)
    flat, grad = torch.zeros(40), torch.zeros(40)
    a.rebind(flat[8:8 + a.padded_size], grad[8:8 + a.padded_size])
    assert float(flat[8:23].sum()) == 30.0 and a.p["w"].data_ptr() == flat[8:].data_ptr()
    a.load_state_dict({"w": torch.ones(3, 5), "b": torch.arange(3.0)})
    assert float(flat[23:26].sum()) == 3.0
    with pytest.raises(KeyError):
        a.load_state_dict({"w": torch.ones(3, 5)})


def _run_reference_arm(env_extra):
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, REFIL_REF_BUDGET_S="3", **env_extra)
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_bench_reference_arm_schema():
    """`bench.py --impl reference` prints ONE JSON line with the contract's keys: the reference's own QLearner.train when
    oracle/_ref is staged (build() does that wherever /root/reference is mounted), else the torch-CPU restatement."""
    from oracle import stage_ref
    d = _run_reference_arm({})
    assert d["impl"] == "reference" and d["unit"] == "transitions/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == ("reference" if stage_ref.available() else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and d["config"]["sample_B"] >= 1


def test_bench_reference_arm_port_fallback():
    d = _run_reference_arm({"REFIL_REF_FORCE_PORT": "1"})
    assert d["cpu_baseline"]["kind"] == "port" and d["value"] > 0


def test_python_call_sites_match_header_arity():
    """Every `_lib.call("name", ...)` / `ops._call("name", ...)` site in the host package passes as many arguments as the C
    prototype declares (ctypes only complains at run time, i.e. on the GPU box)."""
    import ast
    import glob
    from refil_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    checked = 0
    for path in glob.glob(os.path.join(root, "refil_b200", "**", "*.py"), recursive=True):
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            if not isinstance(node, ast.Call) or not node.args or not isinstance(node.args[0], ast.Constant):
                continue
            f = node.func
            fname = f.attr if isinstance(f, ast.Attribute) else getattr(f, "id", "")
            if fname not in ("call", "_call") or not isinstance(node.args[0].value, str):
                continue
            proto = _lib.PROTOS.get("refil_" + node.args[0].value)
            if proto is None:
                continue
            n = len(node.args) - 1 + (1 if fname == "_call" else 0)      # ops._call appends the stream
            star = sum(isinstance(x, ast.Starred) for x in node.args)
            if star:
                n += star * (9 - 1)                                       # ops._copies(): 3 masks + 3 strides + 3 modes
            assert n == len(proto[1]), "%s:%d refil_%s: %d args passed, header declares %d" % (
                os.path.relpath(path, root), node.lineno, node.args[0].value, n, len(proto[1]))
            checked += 1
    assert checked >= 25


def test_sass_histogram_has_blackwell_native_instructions():
    """The dense hot path is tcgen05 + TMEM + TMA, the attention kernels stage their tiles with TMA: asserted on the SASS of the
    built library (scripts/sass_histogram.py; the committed table is profiles/sass_histogram.md)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("sass_histogram", os.path.join(ROOT, "scripts", "sass_histogram.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    h = mod.histogram()
    ts, wg = h["tc_gemm_ts_kernel<0>"], h["tc_wgrad_ts_kernel<0>"]     # single-problem / grouped instantiations
    assert h["tc_gemm_ts_kernel<1>"]["UTCHMMA"] >= 12 and h["tc_wgrad_ts_kernel<1>"]["UTCHMMA"] >= 12
    for k in (ts, wg):
        assert k["UTCHMMA"] >= 12 and k["LDTM"] >= 1 and k["STTM"] >= 2 and k["UTMALDG"] >= 1 and k["UTCBAR"] >= 2, dict(k)
    assert ts["UTMASTG"] >= 1 and ts["UTMAREDG"] >= 1                      # TMA tensor store / reduce-add epilogue
    att = [c for n, c in h.items() if n.startswith("attn_fwd_kernel<32, 4, 0>") or n.startswith("attn_bwd_kernel<32, 4, 0>")]
    assert len(att) == 2 and all(c["UTMALDG"] >= 1 and c["FFMA"] > 100 for c in att)
    assert not any(c["HMMA"] for n, c in h.items() if n.startswith("tc_"))   # no legacy mma.sync on the dense path


def test_ctypes_descriptor_structs_match_the_header(tmp_path):
    """ops.py mirrors RefilGemmDesc / RefilWgradDesc / RefilAttnDesc by hand: their sizes and last-field offsets must be the C ones."""
    import ctypes
    import subprocess
    from refil_b200 import ops
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\n'
                   'int main(void) { printf("%%zu %%zu %%zu %%zu %%zu %%zu\\n", sizeof(RefilGemmDesc), offsetof(RefilGemmDesc, n_cols), '
                   'sizeof(RefilWgradDesc), offsetof(RefilWgradDesc, M), sizeof(RefilAttnDesc), offsetof(RefilAttnDesc, n_copies)); '
                   'return 0; }\n' % os.path.join(ROOT, "include", "refil_b200.h"))
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(ops._GemmDesc), ops._GemmDesc.n_cols.offset, ctypes.sizeof(ops._WgradDesc), ops._WgradDesc.M.offset,
            ctypes.sizeof(ops._AttnDesc), ops._AttnDesc.n_copies.offset]
    assert got == want, (got, want)


def test_qkv_split_policy(monkeypatch):
    """Host-side choice of the split in_trans (queries for the agent rows only): large shards by default, never when every entity is
    an agent or the agent count does not divide 32, REFIL_QKV_SPLIT forces it either way."""
    from refil_b200 import ops
    monkeypatch.delenv("REFIL_QKV_SPLIT", raising=False)
    monkeypatch.delenv("REFIL_TC_MODE", raising=False)
    assert ops.qkv_split_ok(128 * 60, 24, 8, 128)              # the benchmark shard: 184 320 entity rows
    assert not ops.qkv_split_ok(16 * 60, 24, 8, 128)           # a 16-episode shard stays on the full product
    assert not ops.qkv_split_ok(4096 * 51, 4, 4, 64)           # Group Matching: every entity is an agent
    assert not ops.qkv_split_ok(128 * 60, 24, 5, 128)          # 5 agents: row groups must divide 32
    monkeypatch.setenv("REFIL_QKV_SPLIT", "1")
    assert ops.qkv_split_ok(60, 24, 8, 128)                    # forced on where the tensor-core path applies (480 agent rows)
    assert not ops.qkv_split_ok(4, 24, 8, 128)                 # ... but never below the tensor-core row threshold
    monkeypatch.setenv("REFIL_QKV_SPLIT", "0")
    assert not ops.qkv_split_ok(128 * 60, 24, 8, 128)
    monkeypatch.setenv("REFIL_QKV_SPLIT", "1")
    monkeypatch.setenv("REFIL_TC_MODE", "ss")
    assert not ops.qkv_split_ok(128 * 60, 24, 8, 128)          # the all-shared-memory kernel has no row groups
