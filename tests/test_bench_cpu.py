"""The reference arm of bench.py runs without a GPU: check the JSON contract of its line on tiny samples, for the
unmodified reference (oracle/_ref) and for the torch-port fallback."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "64"] + extra,
                         capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    return out


def _check_contract(line, steps):
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["higher_is_better"] is True
    assert line["metric"] == "occupancy_queries_per_s_512cubed_dense_recon" and line["steps"] == steps and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the workload both arms are quoted on is the SAME dict (the bounded sample lives in cpu_baseline.sample)
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    want = bench.make_config(argparse.Namespace(resolution=512, gpus=line["n_gpus"], size=64, precision="fp16r"))
    assert line["config"] == want


def test_reference_arm_runs_the_unmodified_reference_reconstruction():
    """With oracle/_ref (or the mount) present a step is one call of the reference's own lib.mesh_util.reconstruction."""
    sys.path.insert(0, ROOT)
    from oracle import ref_runner
    if not ref_runner.available():
        import pytest
        pytest.skip("neither /root/reference nor oracle/_ref present")
    # torchrun exports OMP_NUM_THREADS=1 for nproc > 1: the arm must still use all host threads
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    out = _run(["--gpus", "2", "--steps", "2", "--warmup", "1", "--ref-resolution", "24", "--no-config1"], env=env)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    _check_contract(line, 2)
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["value"] == line["value"] and "lib.mesh_util.reconstruction" in cb["sample"]
    assert cb["cores"] == len(os.sched_getaffinity(0))
    assert line["n_gpus"] == 2


def test_reference_arm_falls_back_to_the_port(tmp_path):
    """Without the reference's files the arm times the oracle's torch port of the query (kind "port")."""
    env = dict(os.environ, SURS_REFERENCE_ROOT=str(tmp_path / "none"), SURS_NO_REF_COPY="1")
    out = _run(["--steps", "1", "--warmup", "0", "--cpu-points", "20000"], env=env)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    _check_contract(line, 1)
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "20000" in cb["sample"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_oracle_ref_copy_is_the_unmodified_reference():
    """oracle/_ref (git-ignored, made by oracle/make_ref.py) is byte-identical to the mount it was copied from."""
    sys.path.insert(0, ROOT)
    import pytest
    from oracle import make_ref
    if not make_ref.available():
        pytest.skip("oracle/_ref not built")
    assert make_ref.check()
    if os.path.isdir(os.path.join(make_ref.SRC, "lib")):
        import filecmp
        for rel in make_ref._files(make_ref.DST):
            assert filecmp.cmp(os.path.join(make_ref.SRC, rel), os.path.join(make_ref.DST, rel), shallow=False), rel
