"""bench.py contract checks on the CPU: the JSON line, and the N>1 launch path under torchrun with gloo (world size 2).

The kernels run through the test-only emulation build (DOT_RING_B200_BENCH_DRYRUN=1); the line is marked `dry_run`
and is never a measurement.  The real bench refuses a non-CUDA library."""

import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KA1_ROOT_SHA256_PREFIX = "963f1e262eba87d5"  # ring 1023 root of the unmodified reference (tests/golden/ring1023_reference.json)

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
            "config", "e2e", "gpu_launches", "clocks", "roofline"]  # fmt: skip


def _start(cmd):
    env = dict(os.environ, DOT_RING_B200_BENCH_DRYRUN="1")
    return subprocess.Popen(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def _finish(proc, timeout=1500):
    out, err = proc.communicate(timeout=timeout)
    assert proc.returncode == 0, err[-2000:]
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out
    return json.loads(lines[0])


def _check_line(line, n_gpus, total, steps, scaling):
    for key in REQUIRED:
        assert key in line, key
    assert line["metric"] == "ring_vrf_proofs_per_s" and line["unit"] == "proofs/s"
    assert line["n_gpus"] == n_gpus and line["steps"] == steps and line["scaling"] == scaling
    parity = line["config"]["parity"]
    assert parity["ring_root_sha256"] == KA1_ROOT_SHA256_PREFIX and parity["equal"] is True and parity["golden_proofs"] == 2 and parity["verified_sample"] >= 1
    assert line["config"]["total_proofs_per_step"] == total
    assert abs(line["value"] * line["ms_per_step"] * 1e-3 - total) < 1e-6 * total
    assert line["gpu_launches"] > 0
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "algorithmic_reduction"} <= set(line["roofline"])
    assert 0 < line["roofline"]["frac"] < 1.2 and line["roofline"]["kernel_launches_per_step"] == 3
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert line["e2e"]["value"] <= line["value"] * 1.001


def test_bench_lines_dry_run():
    """The three launch modes of bench.py, run side by side (each builds an emulated 6145-point window table, ~1.5 minutes):
    one process / one device (weak, --batch), torchrun with two gloo ranks (strong: one batch of 3 proofs sharded 2 + 1), and
    one process driving two emulated devices through EnginePool (strong)."""
    single = _start([sys.executable, "bench.py", "--steps", "1", "--warmup", "1", "--batch", "1", "--window-bits", "4", "--no-cpu-baseline"])
    ranks = _start([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29617",
                    "bench.py", "--gpus", "2", "--steps", "1", "--warmup", "1", "--total", "3", "--window-bits", "4"])  # fmt: skip
    pool = _start([sys.executable, "bench.py", "--gpus", "2", "--steps", "1", "--warmup", "1", "--total", "2", "--window-bits", "4"])
    line = _finish(single)
    _check_line(line, 1, 1, 1, "weak")
    line = _finish(ranks)
    _check_line(line, 2, 3, 1, "strong")
    assert line["config"]["launch"].startswith("torchrun")
    assert "cpu_baseline" not in line  # rank 0 at N=1 only
    line = _finish(pool)
    _check_line(line, 2, 2, 1, "strong")
    assert "EnginePool" in line["config"]["launch"]


def test_bench_refuses_cpu_library_without_dryrun():
    env = {k: v for k, v in os.environ.items() if k != "DOT_RING_B200_BENCH_DRYRUN"}
    out = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "0", "--batch", "1"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode != 0  # no GPU here: the product path must fail loudly, not fall back
