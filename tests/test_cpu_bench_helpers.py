"""Host logic of bench.py that the GPU box runs unattended: the nvidia-smi clock summary (one poller per node watching several
GPUs), the argument defaults the driver relies on, and the reference arm's JSON contract on a tiny sample."""
import datetime
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _line(ts, idx, sm, mx, pw, reasons=("Not Active",) * 4, active="0x0000000000000000"):
    stamp = datetime.datetime.fromtimestamp(ts).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
    return ", ".join([stamp, str(idx), str(sm), str(mx), "%.2f" % pw, active] + list(reasons)) + "\n"


def _sampler(gpus, text):
    s = bench.ClockSampler(gpus[0], n_gpus=len(gpus), enabled=False)
    s.f.write(text); s.f.flush()
    s.p = object()                      # "running": window() only reads the file
    return s


def test_clock_window_single_gpu():
    t0 = 1_700_000_000.0
    txt = "".join(_line(t0 + 0.02 * i, 0, 1965 if i != 3 else 1800, 1965, 400 + i) for i in range(10))
    txt += "garbage line\n"
    txt += _line(t0 + 5.0, 0, 1000, 1965, 100)                       # outside the window
    w = _sampler([0], txt).window(t0, t0 + 0.2)
    assert w["sm_mhz"] == 1965.0 and w["sm_max_mhz"] == 1965.0 and w["samples"] == 10
    assert w["power_w_max"] == 409.0 and w["reasons"] == [] and "per_gpu" not in w


def test_clock_window_several_gpus_reports_the_slowest_and_names_the_throttled_one():
    t0 = 1_700_000_100.0
    txt = ""
    for i in range(8):
        for g in range(4):
            sm = 1965 if g != 2 else 1700
            reasons = ("Not Active", "Not Active", "Not Active", "Active") if (g == 2 and i == 5) else ("Not Active",) * 4
            txt += _line(t0 + 0.05 * i, g, sm, 1965, 300 + 10 * g, reasons)
    w = _sampler([0, 1, 2, 3], txt).window(t0, t0 + 0.5)
    assert w["sm_mhz"] == 1700.0                                   # the slowest GPU's median
    assert [p["gpu"] for p in w["per_gpu"]] == [0, 1, 2, 3] and w["per_gpu"][2]["sm_mhz"] == 1700.0
    assert w["reasons"] == ["sw_power_cap"] and w["per_gpu"][2]["reasons"] == ["sw_power_cap"] and w["per_gpu"][0]["reasons"] == []
    assert w["power_w_max"] == 330.0 and w["samples"] == 32


def test_clock_window_without_samples_or_without_nvidia_smi():
    s = _sampler([0], "")
    assert "error" in s.window(0.0, 1.0)
    off = bench.ClockSampler(0, n_gpus=8, enabled=False)
    off.start()                                                     # ranks other than local rank 0: no poller at all
    assert off.p is None and off.window(0.0, 1.0)["error"] == "nvidia-smi not available"
    off.stop()


def test_bench_refuses_to_run_the_gpu_arm_without_a_gpu():
    """No CPU fallback: without a CUDA device the default arm must fail loudly, not print a number."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-extras", "--no-cpu", "--no-e2e"],
                       capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode != 0
    assert not any(l.startswith("{") and '"value"' in l for l in r.stdout.splitlines())


def test_reference_arm_prints_the_contract_line_on_a_small_sample():
    """`bench.py --impl reference` runs on the host alone: same metric / unit / config keys as the GPU arm, impl = reference, a
    cpu_baseline describing the run, e2e equal to the line's value with zero copy bytes."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--batch", "32",
                        "--batches-per-step", "1"], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"].startswith("frames/sec") and d["value"] > 0 and d["vs_baseline"] is None
    for key in ("workload", "width", "height", "leds", "mode", "batch", "batches_per_step"):
        assert key in d["config"], key
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_under_torchrun_prints_one_line_from_rank_zero():
    """The driver launches both arms the same way; with two ranks the reference arm runs on rank 0 alone and the other rank exits 0."""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29578", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--batch", "32", "--batches-per-step", "1"], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
