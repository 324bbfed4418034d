"""CLI modes: the reference's flags and output lines, exercised on CPU/gloo."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, nproc=1, timeout=240):
    env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS="1")
    if nproc == 1:
        cmd = [sys.executable, "-m"] + args
        env.update(RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
               "--master-addr", "127.0.0.1", "--master-port", "29612", "-m"] + args
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_simulator_flags_match_reference_defaults():
    from vdpp_b200.modes.simulator import build_parser
    a = build_parser().parse_args([])
    # reference src/modes/simulator.py:35-66
    assert (a.total_steps, a.latent_channels, a.latent_frames, a.latent_height, a.latent_width) == (28, 8, 8, 32, 32)
    assert (a.dtype, a.device, a.backend, a.seed, a.latent_batch) == ("fp32", "cpu", "auto", 42, 1)


def test_benchmark_flags_match_reference_defaults():
    from vdpp_b200.modes.benchmark import build_parser
    a = build_parser().parse_args([])
    # reference src/modes/benchmark.py:29-62
    assert (a.total_steps, a.num_samples, a.warmup_samples, a.hidden_channels) == (28, 10, 2, 64)
    assert (a.latent_channels, a.latent_frames, a.latent_height, a.latent_width) == (4, 14, 40, 72)
    assert a.model == "dummy" and a.model_id == "stabilityai/stable-video-diffusion-img2vid-xt" and a.guidance_scale is None


def test_simulator_two_ranks_gloo_prints_final_norm():
    r = _run(["src.modes.simulator", "--total-steps", "6", "--latent-channels", "4", "--latent-frames", "3",
              "--latent-height", "8", "--latent-width", "8"], nproc=2)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Final latent norm:" in r.stderr + r.stdout


def test_simulator_rejects_uneven_split_like_the_reference_unless_allowed():
    base = ["src.modes.simulator", "--total-steps", "5", "--latent-channels", "4", "--latent-frames", "2",
            "--latent-height", "8", "--latent-width", "8"]
    bad = _run(base, nproc=2)
    assert bad.returncode != 0 and "divisible" in (bad.stderr + bad.stdout)
    ok = _run(base + ["--allow-uneven"], nproc=2)
    assert ok.returncode == 0 and "Final latent norm:" in ok.stderr + ok.stdout


@pytest.mark.parametrize("schedule", ["fixed", "ring"])
def test_benchmark_mode_prints_benchmark_json(schedule):
    r = _run(["src.modes.benchmark", "--device", "cpu", "--total-steps", "4", "--num-samples", "4", "--warmup-samples",
              "2", "--latent-frames", "3", "--latent-height", "8", "--latent-width", "8", "--hidden-channels", "8",
              "--schedule", schedule], nproc=2)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("BENCHMARK_JSON=")]
    assert len(line) == 1
    res = json.loads(line[0].split("=", 1)[1])
    # keys of reference src/modes/benchmark.py:270-285
    for k in ("world_size", "total_steps", "steps_per_gpu", "model", "fsdp", "num_samples_measured", "warmup_samples",
              "latent_shape", "first_sample_time_s", "avg_sample_time_s", "throughput_samples_per_s",
              "per_sample_times_ms", "peak_memory_gb_per_rank", "max_peak_memory_gb"):
        assert k in res, k
    assert res["world_size"] == 2 and res["steps_per_gpu"] == 2 and len(res["per_sample_times_ms"]) == 6
    assert res["throughput_samples_per_s"] > 0


def test_data_parallel_mode_prints_benchmark_json():
    r = _run(["src.modes.benchmark_data_parallel", "--device", "cpu", "--total-steps", "3", "--num-samples", "4",
              "--warmup-samples", "1", "--latent-frames", "3", "--latent-height", "8", "--latent-width", "8",
              "--hidden-channels", "8"], nproc=2)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("BENCHMARK_JSON=")]
    assert len(line) == 1
    res = json.loads(line[0].split("=", 1)[1])
    # keys of reference src/modes/benchmark_data_parallel.py:233-248
    for k in ("mode", "world_size", "total_steps", "steps_per_gpu", "model", "num_samples_measured", "warmup_samples",
              "samples_per_rank", "latent_shape", "first_sample_time_s", "avg_sample_time_s", "throughput_samples_per_s",
              "wall_clock_s", "per_sample_times_ms"):
        assert k in res, k
    assert res["mode"] == "data_parallel" and res["num_samples_measured"] == 4 and res["samples_per_rank"] == 2


def test_production_flags_match_reference():
    from vdpp_b200.modes.production import build_parser
    # reference src/modes/production.py:20-47: --total-steps is required, everything else has these defaults
    with pytest.raises(SystemExit):
        build_parser().parse_args([])
    a = build_parser().parse_args(["--total-steps", "25", "--latent-shape", "1", "4", "25", "72", "128"])
    assert (a.num_samples, a.seed, a.backend, a.fps, a.motion_bucket_id) == (1, 0, "auto", 6, 127)
    assert a.model_id == "stabilityai/stable-video-diffusion-img2vid-xt" and a.latent_shape == [1, 4, 25, 72, 128]
    assert not a.enable_memory_opt and not a.attention_slicing and a.timesteps is None
    import src.modes.production as shim                      # the reference's import path resolves
    assert shim.main is not None


def test_comparison_sweep_writes_reference_csv(tmp_path):
    """The pipeline-vs-data-parallel sweep of reference scripts/benchmark_comparison.sh, on CPU/gloo with the dummy model."""
    import csv
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import benchmark_comparison as bc
    path, rows = bc.main(["--gpu-counts", "1,2", "--device", "cpu", "--total-steps", "4", "--num-samples", "2",
                          "--warmup-samples", "1", "--latent-frames", "2", "--latent-height", "8", "--latent-width", "8",
                          "--hidden-channels", "8", "--results-dir", str(tmp_path)])
    table = list(csv.reader(open(path)))
    assert table[0] == ["mode", "gpu_count", "total_steps", "steps_per_gpu", "num_samples", "first_sample_s",
                        "avg_sample_s", "throughput_sps"]                     # benchmark_comparison.sh:47
    assert [(r[0], r[1]) for r in table[1:]] == [("pipeline_parallel", "1"), ("data_parallel", "1"),
                                                 ("pipeline_parallel", "2"), ("data_parallel", "2")]
    assert all(float(r[7]) > 0 for r in table[1:])


def test_tools_and_entry_points_compile():
    # the measurement tools only run on the GPU box: at least keep them syntactically valid here
    import glob
    import py_compile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = glob.glob(os.path.join(root, "tools", "*.py")) + glob.glob(os.path.join(root, "scripts", "*.py")) + [
        os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")]
    assert len(files) >= 15
    for f in files:
        py_compile.compile(f, doraise=True)
