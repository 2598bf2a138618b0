"""bench.py's CPU-side pieces (no GPU): the reference arm runs the oracle on the full synthetic design (measured, not
extrapolated) and prints the contract's JSON line; the parity helpers recover coefficients as DataStd does and flag
coefficient / support / iteration-count differences."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_a_measured_line(tmp_path):
    env = dict(os.environ, TMPDIR=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "20000", "--p", "256", "--nlambda", "12",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["steps_measured"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["extrapolated"] is False and cb["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    m = cb["measured_s"]
    wall = m["standardize"] + m["gram"] + m["lanczos_cholesky"] + m["iterations"]          # generation excluded: X is an input
    assert abs(line["ms_per_step"] - wall * 1e3) < 1e-6 * wall * 1e3 + 1e-9
    assert abs(line["value"] - line["niter_path"] / wall) < 1e-9 * line["value"]
    # the fit is left for the GPU arm to compare with
    saved = [f for f in os.listdir(tmp_path) if f.startswith("b200admm_ref_tall_n20000_p256_l12")]
    assert saved
    z = np.load(os.path.join(tmp_path, saved[0]))
    assert z["beta"].shape == (257, 12) and int(z["niter"].sum()) == line["niter_path"]


def test_rank_other_than_zero_prints_nothing_in_the_reference_arm():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                       timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_parity_helpers():
    import bench
    rng = np.random.default_rng(0)
    p = 50
    z = np.zeros(p, dtype=np.float32)
    z[[1, 7, 30]] = [0.5, -1.25, 2.0]
    meanX = rng.normal(size=p).astype(np.float32)
    scaleX = rng.uniform(0.5, 2, size=p).astype(np.float32)
    b = bench.recover_f32(z, meanX, scaleX, 3.0, 1.5)
    assert b.shape == (p + 1,) and (b[1:] != 0).sum() == 3
    assert np.allclose(b[1:][[1, 7, 30]], z[[1, 7, 30]] / scaleX[[1, 7, 30]] * 1.5, rtol=1e-6)
    assert np.isclose(b[0], 3.0 - (b[1:] * meanX).sum(), rtol=1e-5)
    B = np.stack([b, 0.5 * b], axis=1)
    same = bench.compare_paths(B, B.copy(), np.array([10, 12]), np.array([10, 12]), None, 2e-4, 1e-4)
    assert same["ok"] and same["max_abs_dbeta"] == 0 and same["support_mismatch"] == 0
    B2 = B.copy()
    B2[5, 0] = 0.01                                      # a spurious coefficient outside the band
    bad = bench.compare_paths(B2, B, np.array([10, 12]), np.array([10, 12]), None, 2e-4, 1e-4)
    assert not bad["ok"] and bad["support_mismatch_outside_band"] == 1
    B3 = B.copy()
    B3[5, 0] = 5e-5                                      # inside the band and the tolerance: reported, not fatal
    near = bench.compare_paths(B3, B, np.array([10, 12]), np.array([10, 12]), None, 2e-4, 1e-4)
    assert near["ok"] and near["support_mismatch"] == 1 and near["support_mismatch_outside_band"] == 0
    slow = bench.compare_paths(B, B, np.array([10, 40]), np.array([10, 12]), None, 2e-4, 1e-4)
    assert not slow["ok"]
    # a lambda whose runs stop at different iterations may differ by the stopping rule's slack, one that stops at the
    # same iteration may not
    B4 = B.copy()
    B4[2, 1] += 2e-3
    apart = bench.compare_paths(B4, B, np.array([10, 13]), np.array([10, 12]), None, 1e-4, 1e-4, p=10000)
    assert apart["ok"] and apart["max_abs_dbeta_where_niter_equal"] == 0
    same = bench.compare_paths(B4, B, np.array([10, 12]), np.array([10, 12]), None, 1e-4, 1e-4, p=10000)
    assert not same["ok"]


def test_trace_rel_diff_treats_common_zeros_as_equal():
    import bench
    a = np.array([[7.07e-3, 2.2347, 7.07e-3, 0.0, 1.0], [7.29e-3, 2.2347, 7.29e-3, 0.5, 1.0]])
    assert bench.trace_rel_diff(a, a.copy()) == 0.0                 # the first iteration's dual residual is exactly 0 in both
    b = a.copy()
    b[1, 3] *= 1 + 1e-6
    assert abs(bench.trace_rel_diff(b, a) - 1e-6) < 1e-9
    c = a.copy()
    c[0, 3] = 1e-3                                                  # zero in one trace only: a real difference
    assert bench.trace_rel_diff(c, a) > 1.0


def test_align_traces_modulo_stutter_rows():
    """A run that takes the restart branch while z stands still repeats the iteration; the traces are then shifts of
    each other (measured on the GPU at LAD n = 5e5 x p = 5e3)."""
    import bench
    base = np.array([[0.063, 27.308, 0.0316, 0.0, 1.0],
                     [0.063, 27.308, 0.0343, 0.0, 1.0],
                     [0.063, 27.300, 0.0370, 0.324, 1.0],
                     [0.063, 26.990, 0.0398, 2.661, 1.0],
                     [0.063, 25.272, 0.0425, 8.610, 1.0],
                     [0.063, 22.183, 0.0456, 15.08, 1.0]])
    same = bench.align_traces(base, base.copy(), 1e-9)
    assert same["ok"] and same["prefix"] == 6 and same["matched"] == 6 and same["skipped_a"] == [] and same["last_a"] == 5
    stut = np.insert(base, 3, [0.063, 27.300, 0.0398, 0.0, 1.0], axis=0)       # a stutter after row 2
    a = bench.align_traces(stut, base, 1e-9)
    assert a["ok"] and a["prefix"] == 3 and a["matched"] == 6 and a["skipped_a"] == [3] and a["last_a"] == 6
    b = bench.align_traces(base, stut[:6], 1e-9)                               # the other way round, tb cut after 6 rows
    assert b["ok"] and b["skipped_b"] == [3] and b["matched"] == 5 and b["last_a"] == 4
    wrong = base.copy()
    wrong[4, 1] *= 1.001
    c = bench.align_traces(wrong, base, 1e-9)
    assert not c["ok"] and c["prefix"] == 4


def test_recorded_bench_lines_carry_the_contract_keys():
    """The committed bench lines of the round (profiles/): every key the measurement contract names, parity green."""
    import glob
    import json
    need = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config", "roofline", "clocks", "gpu_launches"}
    lines = sorted(glob.glob(os.path.join(ROOT, "profiles", "r2D_bench_line.json")) + glob.glob(os.path.join(ROOT, "profiles", "r2v_bench_line_*gpu.json"))
                   + glob.glob(os.path.join(ROOT, "profiles", "r2u_config_*.json")) + glob.glob(os.path.join(ROOT, "profiles", "r2B_config_wide.json"))
                   + glob.glob(os.path.join(ROOT, "profiles", "r2z_config5_*.json")))
    assert len(lines) >= 8
    for path in lines:
        d = json.loads(open(path).read().strip().splitlines()[-1])
        assert need <= set(d), (path, need - set(d))
        assert "workload" in d["config"] and "model" not in d["config"]
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
        assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"]))
        assert d["gpu_launches"] > 0
        par = d.get("parity_vs_n1") or d.get("parity")
        assert par and par["ok"] is True, path
        if d["n_gpus"] == 1 and "consensus" not in d["config"]["workload"]:
            assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
            assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    ref = json.loads(open(os.path.join(ROOT, "profiles", "r2u_bench_reference.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["kind"] == "port"
    assert ref["config"]["workload"] == json.loads(open(os.path.join(ROOT, "profiles", "r2u_bench.json")).read().strip().splitlines()[-1])["config"]["workload"]
