"""bench.py's reference arm runs without a GPU: check the JSON line it must print (one line, contract keys) on a tiny
workload, alone and under a two-rank launch (rank 0 prints, the other rank exits 0 without work)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--genome-len", "150000", "--families", "2", "--members", "3", "--steps", "1", "--warmup", "0", "--cpu-threads", "2"]


def run(env_extra, args):
    env = dict(os.environ, **env_extra)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"] + args, env=env,
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.strip()]


def test_reference_arm_line():
    lines = run({}, SMALL)
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["n_gpus"] == 1 and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 2 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["scaling"] == "strong"
    assert "pyskani_b200" not in d["cpu_baseline"]["sample"]


def test_reference_arm_does_not_load_the_product():
    """the CPU arm must not map libskb.so / the extension: it imports the workload generator and the oracle only"""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference'] + %r; runpy.run_path(%r, run_name='__main__'); "
            "import os; maps = open('/proc/self/maps').read(); "
            "assert 'libskb' not in maps and '_skani' not in maps, 'product library mapped in the reference arm'; "
            "assert 'pyskani_b200' not in sys.modules" % (SMALL, os.path.join(ROOT, "bench.py")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]


def test_reference_arm_other_ranks_stay_silent():
    assert run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, SMALL + ["--gpus", "2"]) == []
