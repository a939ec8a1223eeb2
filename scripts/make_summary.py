#!/usr/bin/env python
"""Build profiles/<tag>_summary.md (+ profiles/traffic.json, profiles/<tag>_ncu_full.md, profiles/<tag>_ncu_launches.csv) from one
evidence run (scripts/gpu_final.sh <tag>; files under gpurun_out/).  Runs here, no GPU needed.

    python scripts/make_summary.py r2n
"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def ncu_rows(path, page="raw"):
    """raw metrics page of an .ncu-rep, or the CSV of that page exported on the GPU box"""
    if path.endswith(".csv"):
        raw = open(path).read()
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", page, "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    return rows[0], rows[2:]


def short(name):
    return name.split("(")[0].replace("void ", "").replace("(bool)", "")[:60]


def launches_table(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    rows = [r for r in rows if not short(r[4]).startswith("gm_")]     # the env leg of bench.py is not part of the learner step
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows:
        k, ns = short(r[4]), float(r[-1])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
        tot += ns
    out = ["| launches | total ms | share | kernel |", "|---|---|---|---|"]
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        out.append("| %d | %.3f | %.1f%% | `%s` |" % (n, ns / 1e6, 100 * ns / tot, k))
    return "\n".join(out), len(rows)


def traffic(path, per_step):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    by_id = collections.OrderedDict()
    for r in rows:
        by_id.setdefault(r[0], {})[r[-3]] = float(r[-1].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9,
                                                                              "ns": 1, "us": 1e3}.get(r[-2], 1)
    ids = list(by_id)[-per_step:]
    rd = sum(by_id[i].get("dram__bytes_read.sum", 0) for i in ids)
    wr = sum(by_id[i].get("dram__bytes_write.sum", 0) for i in ids)
    return {"launches": len(ids), "read": rd, "write": wr}


def full_table(path):
    cols = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"), ("gpu__time_duration.sum", "us"),
            ("dram__bytes_read.sum", "dram rd MB"), ("dram__bytes_write.sum", "dram wr MB"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor %"),
            ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1/smem %"),
            ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 %"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"), ("launch__registers_per_thread", "regs")]
    hdr, rows = ncu_rows(path)
    idx = [(hdr.index(c) if c in hdr else None, n) for c, n in cols]
    out = ["| " + " | ".join(n for _, n in idx) + " |", "|" + "---|" * len(idx)]
    seen = collections.Counter()
    for r in rows:
        name = short(r[idx[0][0]])
        key = (name, r[idx[1][0]])
        seen[key] += 1
        if seen[key] > 2:                       # two launches per (kernel, grid) are enough
            continue
        vals = []
        for i, n in idx:
            v = r[i] if i is not None else ""
            if n == "kernel":
                v = name
            else:
                try:
                    v = "%.1f" % float(v.replace(",", ""))
                except ValueError:
                    pass
            vals.append(v)
        out.append("| " + " | ".join(vals) + " |")
    return "\n".join(out)


def main(tag):
    d = json.load(open(os.path.join(OUT, tag + "_bench.json")))
    lines = ["# Round 2, evidence run `%s` (scripts/gpu_final.sh; summarised by scripts/make_summary.py)\n" % tag]
    lines.append("Commands (B200, one GPU): `python -m pytest tests -m gpu`, `python bench.py --steps 20 --warmup 5`, "
                 "`python bench.py --impl reference --steps 5 --warmup 1`, `ncu --metrics gpu__time_duration.sum` (launch list), "
                 "`ncu --set full` on the hot kernels, `ncu --metrics dram__bytes_*` on every dense launch.  ncu per-launch times are "
                 "cold-cache and serialised: compare shares.\n")
    lines.append("## bench line (%s_bench.json)\n" % tag)
    e = d["e2e"]
    lines.append("* value %.0f transitions/s (%.3f ms/step, %s), e2e %.0f transitions/s (H2D %d B + D2H %d B per step inside the timed "
                 "region), launches/step %.0f, host enqueue %.3f ms/step" % (d["value"], d["ms_per_step"], d["config"]["step_submission"],
                                                                           e["value"], e["h2d_bytes_per_step"], e["d2h_bytes_per_step"],
                                                                           d["launches_per_step"], d["host_enqueue_ms_per_step"]))
    lines.append("* clocks: %s" % json.dumps(d["clocks"]))
    if "strong" in d:
        for k, v in d["strong"].items():
            if "value" in v:
                lines.append("* strong `%s` at this N: %.0f transitions/s, %.3f ms/step, %d episodes per GPU, %.0f launches/step"
                             % (k, v["value"], v["ms_per_step"], v["episodes_per_gpu"], v["launches_per_step"]))
    env = d.get("env", {})
    if env:
        lines.append("* env kernel: %.3g env-steps/s (%d envs x %d agents, %s)" % (env["value"], env["n_envs"], env["n_agents"],
                                                                                  env["config"].split(";")[-1].strip()))
        for k in ("rollout", "rollout_cfg4"):
            if k in env and "value" in env[k]:
                lines.append("* %s: %.3g env-steps/s, %.2f ms per rollout (%s)" % (k, env[k]["value"], env[k]["ms_per_rollout"], env[k]["config"]))
        if "cpu_baseline" in env:
            lines.append("* env cpu_baseline: %.3g env-steps/s (%s)" % (env["cpu_baseline"]["value"], env["cpu_baseline"]["sample"]))
    if "cpu_baseline" in d:
        c = d["cpu_baseline"]
        lines.append("* cpu_baseline (%s, %d threads): %.0f transitions/s -- %s" % (c["kind"], c["cores"], c["value"], c["sample"]))
    refp = os.path.join(OUT, tag + "_bench_reference.json")
    if os.path.exists(refp) and os.path.getsize(refp):
        r = json.loads([ln for ln in open(refp) if ln.startswith("{")][-1])
        lines.append("* `--impl reference`: %.0f transitions/s, %.0f ms/step, sample_B=%s, kind=%s" % (
            r["value"], r["ms_per_step"], r["config"].get("sample_B"), r["cpu_baseline"]["kind"]))
    for k, v in d.get("other_workloads", {}).items():
        if "value" in v:
            lines.append("* other workload `%s`: %.3g transitions/s (%.2f ms/step) -- %s" % (k, v["value"], v["ms_per_step"], v["workload"]))
    lines.append("\n## rooflines (serialised per-launch CUDA events inside bench.py; peak = measured HBM %.1f GB/s)\n" % d["roofline"]["peak"])
    lines.append("| kernel | ms/step | launches/step | algorithmic GB/step | achieved GB/s | frac of measured HBM |\n|---|---|---|---|---|---|")
    for key, label in (("roofline", "tc_gemm_ts_kernel (dense fwd / bwd-data)"), ("roofline_wgrad", "tc_wgrad_ts_kernel (weight gradients)"),
                       ("roofline_attention", "attn_fwd_kernel"), ("roofline_attention_bwd", "attn_bwd_kernel")):
        r = d.get(key)
        if r:
            lines.append("| %s | %.3f | %d | %.2f | %.0f | %.3f |" % (label, r["ms_per_step"], r["launches_per_step"],
                                                                      r["algorithmic_bytes_per_step"] / 1e9, r["achieved"], r["frac"]))
    lines.append("\n## per-entry-point ms/step (serialised, eager launches)\n\n| entry point | ms/step | share |\n|---|---|---|")
    ks = d["kernels_ms_per_step"]
    tot = sum(ks.values())
    for k, v in ks.items():
        lines.append("| %s | %.3f | %.1f%% |" % (k, v, 100 * v / tot))
    lines.append("\nSum %.2f ms serialised vs %.2f ms for the real step (streams + graph replay overlap the launches' fixed costs)." % (tot, d["ms_per_step"]))
    lines.append("\n## dense layers by shape [M x N x K] (launches per step, us per launch)\n")
    for k, v in list(d["dense_shapes_us"].items())[:16]:
        lines.append("* %s: %d x %.1f us" % (k, v[0], v[1]))
    lp = os.path.join(OUT, tag + "_ncu_launches.csv")
    if os.path.exists(lp):
        tbl, n = launches_table(lp)
        shutil.copy(lp, os.path.join(PROF, tag + "_ncu_launches.csv"))
        lines.append("\n## ncu launch list (%s_ncu_launches.csv, %d launches of the learner steps; the env leg's gm_* kernels left out), "
                     "aggregated by kernel\n\n%s" % (tag, n, tbl))
    tp = os.path.join(OUT, tag + "_traffic.csv")
    if os.path.exists(tp) and d.get("roofline"):
        per_step = int(d["roofline"]["launches_per_step"])
        t = traffic(tp, per_step)
        if t["launches"]:
            tj = {"kernel": "tc_gemm_ts_kernel", "launches": t["launches"], "dram_bytes_read_per_step": t["read"],
                  "dram_bytes_write_per_step": t["write"], "traffic_bytes_per_launch": (t["read"] + t["write"]) / t["launches"],
                  "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the %d launches of the last step "
                            "(scripts/gpu_final.sh, capture %s)" % (t["launches"], tag)}
            json.dump(tj, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
            alg = d["roofline"]["algorithmic_bytes_per_step"] / per_step
            lines.append("\nDominant kernel: DRAM traffic measured by ncu over the %d launches of one step = %.1f MB per launch vs %.1f MB "
                         "algorithmic per launch (%.2fx)." % (t["launches"], tj["traffic_bytes_per_launch"] / 1e6, alg / 1e6,
                                                             tj["traffic_bytes_per_launch"] / alg))
    fp = os.path.join(OUT, tag + "_ncu_full_raw.csv")
    if not os.path.exists(fp):
        fp = os.path.join(OUT, tag + "_prof.ncu-rep")
    if os.path.exists(fp):
        tbl = full_table(fp)
        open(os.path.join(PROF, tag + "_ncu_full.md"), "w").write(
            "# ncu --set full, capture %s (scripts/gpu_final.sh): two launches per (kernel, grid)\n\n%s\n" % (tag, tbl))
        lines.append("\n## ncu --set full (%s_ncu_full.md)\n\n%s" % (tag, tbl))
    shutil.copy(os.path.join(OUT, tag + "_bench.json"), os.path.join(PROF, tag + "_bench.json"))
    if os.path.exists(refp) and os.path.getsize(refp):
        shutil.copy(refp, os.path.join(PROF, tag + "_bench_reference.json"))
    open(os.path.join(PROF, tag + "_summary.md"), "w").write("\n".join(lines) + "\n")
    print("wrote profiles/%s_summary.md" % tag)


if __name__ == "__main__":
    main(sys.argv[1])
