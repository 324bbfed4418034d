"""CPU baseline of BASELINE.md section 3: the step pipeline with the DummyUNet over gloo on the host cores
(BASELINE config 1 workload: DummyUNet(channels=4, hidden 16), latent [1,4,14,64,64] fp32), streamed through
``src.modes.benchmark --device cpu`` (the reference's benchmark-mode arithmetic: steady throughput = measured samples /
sum of their last-rank completion intervals).  Configurations: 25 steps x 1 rank, 25 x 5, 28 x 4, 24 x 4 (all accepted
by the reference's divisibility rule) and 25 x 4 with the uneven extension (7,6,6,6).  A reported baseline, not a
target.  One JSON line per configuration on stdout and in gpurun_out/cpu_simulator.jsonl.
   python tools/cpu_simulator_bench.py [--samples 6] [--threads-per-rank 1]
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONFIGS = [(25, 1, False), (25, 5, False), (28, 4, False), (24, 4, False), (25, 4, True)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--threads-per-rank", type=int, default=1, help="OMP threads per rank (torchrun's default is 1)")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", "cpu_simulator.jsonl"), "a")
    for i, (T, W, uneven) in enumerate(CONFIGS):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={W}", "--master-addr",
               "127.0.0.1", "--master-port", str(29700 + i), "-m", "src.modes.benchmark", "--device", "cpu", "--model", "dummy",
               "--total-steps", str(T), "--latent-channels", "4", "--hidden-channels", "16", "--latent-frames", "14",
               "--latent-height", "64", "--latent-width", "64", "--num-samples", str(a.samples), "--warmup-samples",
               str(a.warmup), "--log-level", "WARNING"]
        if uneven:
            cmd.append("--allow-uneven")
        env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS=str(a.threads_per_rank))
        t0 = time.time()
        r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1800)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("BENCHMARK_JSON=")]
        if r.returncode != 0 or not line:
            rec = {"total_steps": T, "world_size": W, "error": (r.stderr or r.stdout)[-400:]}
        else:
            res = json.loads(line[0].split("=", 1)[1])
            rec = {"workload": "DummyUNet(channels=4, hidden 16), latent [1,4,14,64,64] fp32, gloo, CPU",
                   "total_steps": T, "world_size": W, "allow_uneven": uneven, "steps_per_rank": res["steps_per_gpu"],
                   "threads_per_rank": a.threads_per_rank, "host_cores": os.cpu_count(),
                   "samples_measured": res["num_samples_measured"], "first_sample_s": res["first_sample_time_s"],
                   "steady_s_per_sample": res["avg_sample_time_s"], "samples_per_s": res["throughput_samples_per_s"],
                   "ms_per_step_per_rank": 1000.0 * res["avg_sample_time_s"] / max(-(-T // W), 1),
                   "wall_s": round(time.time() - t0, 1)}
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n")
    out.close()


if __name__ == "__main__":
    main()
