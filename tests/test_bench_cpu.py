"""bench.py contract on CPU: the reference arm (`--impl reference`: the reference's own modules staged under oracle/_ref,
or the oracle port when they are absent) prints ONE JSON line with the keys the driver reads, and - launched as a
non-zero rank - prints nothing and exits 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                           "--cpu-sample-graphs", "1"], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_line():
    r = _run({"RANK": "0"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "encoder_node_pairs_per_sec" and d["unit"] == "node-pairs/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 1
    from oracle import ref_loader as RL
    want_kind = "reference" if RL.have_ref("generator") else "port"
    assert d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "node-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""
