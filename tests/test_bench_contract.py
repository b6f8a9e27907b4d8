"""bench.py contract pieces that can be checked without a GPU: the CPU arm (`--impl reference`) prints ONE JSON line
with the contract's keys, uses every host core even under torchrun (which exports OMP_NUM_THREADS=1 to its workers), and
under torchrun only rank 0 runs and prints; the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e"}


def _check_line(out: str, n_gpus: int):
    lines = [l for l in out.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert KEYS <= set(d), sorted(KEYS - set(d))
    assert d["impl"] == "reference" and d["metric"] == "gi_ray_samples_per_s" and d["unit"] == "Gray-samples/s"
    assert d["n_gpus"] == n_gpus and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and "cube 128x128" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    return d


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")     # what torchrun would export: the arm must override it
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cube_512", "--cpu-sample-div", "4",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr
    d = _check_line(out.stdout, 1)
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_reference_arm_under_torchrun_only_rank0_works():
    import socket
    sk = socket.socket(); sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]; sk.close()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "cube_512",
                          "--cpu-sample-div", "4", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = _check_line(out.stdout, 2)
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, cwd=ROOT,
                         timeout=300)
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)


def test_reference_arm_never_maps_the_product_library():
    """The CPU arm is the oracle alone: after timing a frame through bench.cpu_oracle_frames the process has not mapped
    librc_b200.so (the arm used to import the product for scene paths and the orbit camera)."""
    code = ("import sys; sys.path.insert(0, %r); import bench\n"
            "res, _ = bench.cpu_oracle_frames('cube_512', 1, 0, 4)\n"
            "assert res['value'] > 0\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'liboracle' in maps, 'the oracle must have run'\n"
            "assert 'librc_b200' not in maps, 'the reference arm loaded the product library'\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
