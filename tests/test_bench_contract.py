"""bench.py's driver contract, on the CPU: the reference arm prints ONE JSON line with the keys the driver and the judge
read (same metric / config / checksum fields as the b200 arm), non-zero ranks print nothing, and the helpers both arms
share are deterministic."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT

sys.path.insert(0, ROOT)


def run_bench(*args, env=None):
    e = dict(os.environ, **(env or {}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=e, stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    lines = [ln for ln in run_bench("--impl", "reference", "--envs", "8192", "--steps", "2", "--warmup", "1").splitlines() if ln.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"].startswith("MSJ env-steps/sec") and j["unit"] == "env-steps/s"
    assert j["higher_is_better"] is True and j["scaling"] == "weak" and j["vs_baseline"] is None and j["dtype"] == "f32"
    assert j["value"] > 0 and j["steps"] == 2 and j["warmup"] == 1 and j["gpu_launches"] == 0
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "reference_python" in cb
    assert j["e2e"] == {"value": j["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["config"]["envs_per_step_in_this_arm"] == 8192 and "workload" in j["config"]
    cs = j["parity_checksum"]
    assert cs["envs"] == 8192 and cs["steps"] == 3 and cs["done_count"] > 0 and len(cs["obs_xor32"]) == 8
    # the unmodified Python reference is timed on THIS box when it is present (baseline/_ref or /root/reference)
    rp = cb["reference_python"]
    assert "single_env_steps_per_s" in rp or "unavailable_on_this_box" in rp


def test_other_ranks_of_the_reference_arm_stay_silent():
    assert run_bench("--impl", "reference", "--envs", "4096", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"}).strip() == ""


def test_shared_inputs_are_deterministic_and_match_the_oracle_population():
    import bench
    a, b = bench.gate_action_batches(1000), bench.gate_action_batches(1000)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and a[0].shape == (1000, 8) and not np.array_equal(a[0], a[1])
    assert np.array_equal(bench.gate_action_batches(5000)[0][:1000], a[0])          # a prefix is a prefix
    assert bench.episode_phases(0, 801).tolist()[:3] == [1, 2, 3] and bench.episode_phases(0, 801)[400] == 1
    assert bench.episode_phases(400, 2).tolist() == [1, 2]                           # phases follow the GLOBAL env id
    r1, _, c1 = bench.cpu_port_rate(4096, 1, 0, 2, want_checksum=True)
    r2, _, c2 = bench.cpu_port_rate(4096, 1, 0, 1, want_checksum=True)
    assert r1 > 0 and r2 > 0
    for k in ("done_count", "obs_xor32", "obs_sum64", "goal_xor32", "step_word_xor32"):
        assert c1[k] == c2[k], k                                                     # thread count does not change results


def test_traffic_json_is_the_default_step_kernel():
    """roofline.traffic comes from profiles/traffic.json: it must describe the kernel the headline times (the default
    instantiation of the step kernel at the benchmarked size), not another capture that happened to be summarised last."""
    import json
    tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert "step_kernel<0" in tj["kernel"] and "generic" not in tj["kernel"]
    assert tj["envs"] == 16777216 and tj["algorithmic_bytes_per_launch"] == 93 * 16777216
    assert 0.9 < tj["dram_bytes_per_launch"] / tj["algorithmic_bytes_per_launch"] < 1.1
