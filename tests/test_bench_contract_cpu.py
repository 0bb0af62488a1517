"""bench.py --impl reference (the reference's CPU algorithm through the oracle port) runs without a GPU and prints ONE
JSON line with the keys the driver reads; the algorithmic-bytes table of the roofline matches DESIGN.md section 4."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_json_line(built):
    # run bench.py as __main__ and then look at what the process loaded: the reference arm must run on the oracle alone --
    # neither the product package nor librpforest.so may be in the process (the driver records the loaded .so files)
    prog = ("import sys, runpy\n"
            "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0']\n"
            "runpy.run_path(%r, run_name='__main__')\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'librpforest' not in maps, 'the reference arm loaded the product library'\n"
            "assert 'liborc' in maps\n"
            "assert not any(m.startswith('rp_tree_b200') or m.startswith('rp-tree_b200') for m in sys.modules), 'product package imported'\n"
            % os.path.join(ROOT, "bench.py"))
    out = subprocess.run([sys.executable, "-c", prog], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    import bench
    maxd = j["config"]["max_depth"]
    assert j["config"] == bench.config_dict(bench.WORKLOAD, maxd, 1), "both arms must print the same config dict"
    assert j["cpu_baseline"]["knn_queries_per_s"] > 0 and 0 <= j["cpu_baseline"]["recall_oracle"] <= 1
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in j, key
    assert j["impl"] == "reference" and j["metric"] == "forest_build_points_per_s" and j["unit"] == "points/s"
    assert j["value"] > 0 and j["e2e"]["value"] == j["value"] and j["e2e"]["h2d_bytes_per_step"] == 0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and "workload" in j["config"]


def test_other_ranks_of_the_reference_arm_do_nothing(built):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_algorithmic_bytes_table():
    import bench
    W = bench.WORKLOAD
    ab = bench.algorithmic_bytes(W, 32, 14, 10, 1968.0)
    n, d = W["n"], W["d"]
    assert ab["project"] == 8 * d * n + 8 * 32 * 14 * n
    assert ab["top_hist"] == 32 * n * 12 and ab["top_compact"] == 32 * n * 4 and ab["top_relabel"] == 32 * n * 6
    assert ab["bottom"] == 32 * n * (4 + 4 + 8 * 4 + 8)
    fused = bench.algorithmic_bytes(W, 32, 14, 10, 1968.0, fused_levels=9)       # k_top_relabel_hist on 9 of the 10 top levels
    assert fused["top_relabel"] == 32 * n * (6 * 1 + 16 * 9) // 10 and fused["top_hist"] == ab["top_hist"]
