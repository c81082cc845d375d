"""bench.py's output contract: the committed GPU line (profiles/bench_r02.json, produced on a B200) and a live run of the
reference arm on this machine's cores carry every key the driver reads, with consistent values."""
import json
import subprocess
import sys
import textwrap

from conftest import ROOT


def check_common(d):
    assert d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    for k in ("value", "ms_per_step"):
        assert isinstance(d[k], float) and d[k] > 0
    for k in ("n_gpus", "steps", "warmup"):
        assert isinstance(d[k], int) and d[k] >= 1
    assert d["warmup"] >= 3 and d["scaling"] in ("weak", "strong") and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == d["unit"] and "h2d_bytes_per_step" in e and "d2h_bytes_per_step" in e
    c = d["cpu_baseline"]
    assert c["value"] > 0 and c["unit"] == d["unit"] and c["cores"] >= 1 and c["kind"] in ("port", "reference") and c["sample"]


def test_committed_gpu_line_has_the_contract_keys():
    """profiles/bench_r02.json = `python bench.py --steps 20 --warmup 5` on one B200 (tools/profile_round.sh)."""
    d = json.loads((ROOT / "profiles" / "bench_r02.json").read_text().strip().splitlines()[-1])
    check_common(d)
    assert d["n_gpus"] == 1 and d["gpu_launches"] == d["steps"]                 # one nav3d_step launch per timed step
    assert d["e2e"]["h2d_bytes_per_step"] == 8 << 20 and d["e2e"]["d2h_bytes_per_step"] == (1 << 20) * (320 + 4 + 1 + 1)
    assert d["e2e"]["value"] < d["value"]                                        # host copies inside the timed region
    assert d["e2e"]["d2h_achieved_gbs"] <= 1.1 * d["e2e"]["host_ceiling_gbs"]    # the e2e path against the measured host ceiling
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["algorithmic_bytes_per_env_step"] == 592 and r["env_steps_per_launch"] == 1 << 20
    assert abs(r["achieved"] - 592 * (1 << 20) / (r["launch_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert abs(d["value"] - (1 << 20) / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    # DRAM bytes of the SAME regime as the timed launches, from the committed ncu capture, and labelled as profiled
    assert r["regime"] == "cold" and d["config"]["regime"] == "cold" and d["config"]["preroll_steps"] == 0
    traffic = json.loads((ROOT / "profiles" / "traffic.json").read_text())
    # (the line was written by the run that also made the capture now in traffic.json; it quoted the previous capture of
    # the same kernel and regime, 0.03 % away)
    assert abs(r["traffic"] / (traffic["c4"]["cold"]["dram_bytes_per_env_step"] * (1 << 20)) - 1) < 0.01 and r["traffic"] > 592 * (1 << 20)
    assert "profiled" in r["traffic_source"] and "step_r02_cold" in r["traffic_source"]
    # the number is tied to correct output: a strided sample replayed by the oracle
    assert d["check"] == {"envs_replayed": 256, "steps_replayed": d["steps"] + d["warmup"], "obs_bit_exact": True, "state_equal": True}
    s = d["steady"]
    assert s["preroll_steps"] == 600 and s["steps"] == d["steps"] and s["check"]["obs_bit_exact"] and s["check"]["state_equal"]
    assert s["value"] > d["value"] and abs(s["traffic"] / (traffic["c4"]["steady"]["dram_bytes_per_env_step"] * (1 << 20)) - 1) < 0.01
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert {"c2", "c3", "c5_train", "simple_env", "fused_rollout_full"} <= set(d["extra"])
    f = d["extra"]["fused_rollout_full"]
    assert f["T"] == 32 and f["algorithmic_bytes_per_launch"] == 592 * (1 << 20) * 32
    for regime in ("cold", "steady"):                                          # full work: every step checked against the oracle
        assert f[regime]["check"]["obs_bit_exact"] and f[regime]["check"]["state_equal"]
        assert abs(f[regime]["roofline_frac"] - 592 * f[regime]["value"] / 1e9 / r["peak"]) < 1e-6
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1


def test_reference_arm_runs_on_host_cores():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    check_common(d)
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # the unmodified reference class when build() staged it (oracle/_ref, present wherever /root/reference was), else the port
    staged = (ROOT / "oracle" / "_ref" / "envs" / "CubicEnv.py").exists()
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["config"]["same_config"] is False
    assert 1e3 < d["value"] < 1e7                                               # a Python env: thousands of steps/s per core


def test_reference_arm_does_not_map_the_product_library():
    """the CPU arm must not import nav3d (whose __init__ dlopens libnav3d_b200.so)."""
    code = textwrap.dedent("""
        import runpy, sys
        sys.argv = ["bench.py", "--impl", "reference", "--steps", "1", "--warmup", "3"]
        try:
            runpy.run_path("bench.py", run_name="__main__")
        except SystemExit:
            pass
        assert "nav3d" not in sys.modules
        print("MAPPED", [l for l in open("/proc/self/maps").read().split() if "libnav3d" in l])
    """)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    assert "MAPPED []" in out.stdout
