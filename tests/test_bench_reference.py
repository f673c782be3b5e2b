"""The reference arm of bench.py (`--impl reference`: the C restatement of the reference's Java path on the host cores) needs
no GPU: it must run here, print ONE JSON line with the contract's keys, and -- under a multi-rank launch -- only on rank 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1", "--loci", "6", *args],
                         capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return [ln for ln in out.stdout.splitlines() if ln.strip()]


def test_reference_arm_prints_the_contract_line():
    lines = _run({})
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["value"] > 0 and d["unit"] == "partials/s"
    assert d["config"]["workload"] == "C2" and d["config"]["loci_total"] == 6
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_runs_on_rank_zero_only_and_on_the_global_problem():
    env = {"WORLD_SIZE": "4", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29999"}
    assert _run(dict(env, RANK="1", LOCAL_RANK="1"), "--gpus", "4") == []          # other ranks: no work, no output, exit 0
    lines = _run(dict(env, RANK="0", LOCAL_RANK="0"), "--gpus", "4")
    d = json.loads(lines[0])
    assert d["n_gpus"] == 4 and d["config"]["loci_per_gpu"] == 6 and d["config"]["loci_total"] == 24   # the same global problem as the GPU arm at N = 4
