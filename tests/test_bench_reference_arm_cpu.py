"""CPU check of bench.py's reference arm (`--impl reference`: the oracle port timed on the host's cores, the one place outside
tests/ and smoke() that may execute oracle/): it prints ONE JSON line with the contract's keys, uses every host thread even when
the launcher exported OMP_NUM_THREADS=1 (torchrun does), and ranks other than 0 exit without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample-div", "8192"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip()


def test_reference_arm_line_and_thread_count():
    line = json.loads(_run({"OMP_NUM_THREADS": "1"}).splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "solve_wall_s" and line["unit"] == "s" and line["higher_is_better"] is False
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] == line["e2e"]["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    nthreads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    assert cb["cores"] == min(nthreads, cb["cores"]) and (nthreads == 1 or cb["cores"] > 1)  # not pinned to one thread by OMP_NUM_THREADS
    assert cb["sample_rows"] >= 1024 and "rows 0.." in cb["sample"] and cb["extrapolated_full_size_solve_s"] == cb["value"]
    assert line["counts"]["outer"] >= 1 and line["config"]["workload"] == "cfg3"


def test_reference_arm_other_ranks_exit_without_work():
    assert _run({"RANK": "3", "WORLD_SIZE": "8"}) == ""


def test_prefix_agreement_report():
    """bench.py's `same_problem` leg states where the CPU port's own trajectory becomes noise-driven and whether the GPU trajectory
    equals it exactly up to there (the criterion of tests/parity.py): identical traces agree; a trace that differs inside the
    compared prefix does not; one that differs only after it still does."""
    import copy
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(bench)
    finally:
        sys.argv = argv
    from tests.parity import golden
    g = golden("glm_200000_1024")
    rep = bench.prefix_agreement(g, g["inner"])
    F = rep["cpu_trace_noise_driven_from_inner_record"]
    assert rep["exact_prefix_equal"] and F is not None and rep["records_compared_exactly"] == F >= 5
    bad = copy.deepcopy(g["inner"])
    bad[2]["bp_cum"] += 1
    assert not bench.prefix_agreement(g, bad)["exact_prefix_equal"]
    tail = copy.deepcopy(g["inner"])[: F + 1]
    tail[F]["bp_cum"] += 7
    assert bench.prefix_agreement(g, tail)["exact_prefix_equal"]
    assert not bench.prefix_agreement(g, g["inner"][: F - 1])["exact_prefix_equal"]
