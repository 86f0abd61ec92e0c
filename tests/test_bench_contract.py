"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (the C port of the reference's
algorithm on the host cores) prints ONE JSON line with the keys the driver reads, on the same metric / unit / config as
the CUDA arm; the CUDA arm refuses to run without a GPU instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True,
                          text=True, timeout=timeout)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--scale", "0.01", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "walk-steps/s" and d["unit"] == "steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    assert d["config"]["workload"] == "powerlaw-1M-10M-sparseotf" and d["config"]["baseline_config"] == "#3"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "rows" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="a GPU is present")
def test_cuda_arm_refuses_to_run_without_a_gpu():
    r = _run("--scale", "0.01", "--steps", "1", "--warmup", "0", "--no-extra")
    assert r.returncode != 0
    assert "needs a GPU" in r.stderr and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_gather_probe_picks_the_faster_mechanism_and_leaves_it_selected():
    """bench.choose_gather (N > 4, --gather auto): one warm + two timed passes per candidate, same order on every rank;
    the faster wins, the first wins ties."""
    sys.path.insert(0, ROOT)
    import bench
    for cost, want in (({"mirror": 4.7, "nccl": 7.3}, "mirror"), ({"mirror": 21.7, "nccl": 7.3}, "nccl"),
                       ({"mirror": 5.0, "nccl": 5.0}, "mirror")):
        calls, clock = [], [0.0]

        def run_pass(cand, seed):
            calls.append((cand, seed))
            clock[0] += cost[cand]

        def timed_ms(fn):
            t0 = clock[0]
            fn()
            return clock[0] - t0

        best, probe = bench.choose_gather(["mirror", "nccl"], run_pass, timed_ms)
        assert best == want and probe == pytest.approx(cost)
        assert calls == [("mirror", 900), ("mirror", 901), ("mirror", 902), ("nccl", 900), ("nccl", 901), ("nccl", 902)]
