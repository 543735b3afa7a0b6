"""Host-side logic of bench.py (no GPU): workload naming, the golden |e| lookup behind `parity_rel_err`, and the
fixed configuration of the CPU reference arm."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_workloads_match_baseline_configs():
    b = _bench()
    cfgs = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    wl, pg, name = b.workload(1, 5)
    assert wl == dict(mesh="cube01_hex", rs=5) and pg == (1, 1, 1)
    assert name == "cube01_hex -p 1 -rs 5 -ok 3 -ot 2 -pa" and "-rs 5 -ok 3 -ot 2 -pa" in cfgs[1]
    assert b.workload(1, 5, "cube01", 0, 3)[2] == "cube01_hex -p 0 -rs 5 -ok 3 -ot 2 -pa"      # config 3: Taylor-Green
    wl8, pg8, name8 = b.workload(8, 5)
    assert wl8 == dict(mesh="cube01_hex", rs=6) and pg8 == (2, 2, 2) and "-rs 6" in cfgs[3]   # config 4
    for ok in (2, 3, 4, 5):                                                                     # config 5
        wl, _, name = b.workload(1, 4, "box01", 3, ok)
        assert wl == dict(mesh="box01_hex", rs=4) and f"-ok {ok} -ot {ok - 1}" in name


def test_golden_lookup_covers_default_and_driver_step_counts():
    b = _bench()
    for steps in (11, 25):     # bench.py defaults (8 + 3) and the driver's --steps 20 --warmup 5
        v = b.golden_e_norm("cube01", 1, 5, 3, steps)
        assert v is not None and v > 0
    assert b.golden_e_norm("cube01", 0, 5, 3, 11) is not None          # Taylor-Green
    assert b.golden_e_norm("cube01", 1, 5, 3, 1000) is None            # beyond the stored run
    assert b.golden_e_norm("box01", 3, 4, 5, 11) is None               # no oracle run stored for that configuration


def test_reference_arm_is_a_fixed_configuration():
    b = _bench()
    assert b.REF_RS == 4
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "budget_s" not in src          # no wall-time-driven choice of the CPU sample size any more


def test_reference_arm_reports_the_native_config():
    """--impl reference prints the native arm's config (workload name, sizes); the bounded CPU sample is named
    beside it.  Sizes against the native line of profiles/bench_r2_1gpu.json and the 8-GPU job (cube01_hex -rs 6)."""
    b = _bench()
    assert b.job_sizes(1, 5, "cube01", 3) == (262144, 7189057, 7077888)
    assert b.job_sizes(8, 5, "cube01", 3) == (262144, 385 ** 3, 128 ** 3 * 27)
    assert b.job_sizes(1, 4, "box01", 2) == (65536, 129 * 65 * 65, 65536 * 8)
    native = json.load(open(os.path.join(ROOT, "profiles", "bench_r2_1gpu.json")))["config"]
    ne, h1, l2 = b.job_sizes(1, 5, "cube01", 3)
    assert (native["elements_per_gpu"], native["h1_dofs_global"], native["l2_dofs_global"]) == (ne, h1, l2)
    assert b.workload(1, 5)[2] == native["workload"]
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert '"config": {"workload": wl_name, "elements_per_gpu": ne_gpu' in src      # the reference line uses them
