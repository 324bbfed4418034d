"""Pipeline-parallel vs data-parallel sweep over GPU counts, writing the reference's CSV.

What reference ``scripts/benchmark_comparison.sh:17-129`` does (torchrun ``src.modes.benchmark`` and
``src.modes.benchmark_data_parallel`` at 1/2/4/7 GPUs, scrape the ``BENCHMARK_JSON=`` line, append one CSV row per run
with the columns ``mode,gpu_count,total_steps,steps_per_gpu,num_samples,first_sample_s,avg_sample_s,throughput_sps``,
print the table), as one Python driver with the same defaults (28 steps, 10 + 2 samples, latent 14 x 40 x 72, seed 42).
Additions: ``--gpu-counts`` (default 1,2,4,8 on a B200 node), ``--allow-uneven`` / ``--schedule ring`` are passed to the
pipeline runs, ``--device cpu`` runs the dummy model over gloo (how the tests exercise it), and rows that fail are
recorded with the error instead of stopping the sweep.
   python tools/benchmark_comparison.py --model svd --latent-height 72 --latent-width 128 --latent-frames 25 \
        --total-steps 25 --allow-uneven --schedule ring
"""
import argparse
import csv
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLUMNS = ["mode", "gpu_count", "total_steps", "steps_per_gpu", "num_samples", "first_sample_s", "avg_sample_s",
           "throughput_sps"]


def run_one(mode_module, ngpus, a, port, extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ngpus}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), "-m", mode_module, "--total-steps", str(a.total_steps), "--num-samples",
           str(a.num_samples), "--warmup-samples", str(a.warmup_samples), "--model", a.model, "--latent-height",
           str(a.latent_height), "--latent-width", str(a.latent_width), "--latent-frames", str(a.latent_frames),
           "--hidden-channels", str(a.hidden_channels), "--seed", str(a.seed), "--device", a.device, "--log-level",
           "WARNING"] + extra
    env = dict(os.environ, PYTHONPATH=ROOT)
    if a.device != "cpu":
        env["CUDA_VISIBLE_DEVICES"] = ",".join(str(i) for i in range(ngpus))
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=a.timeout)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("BENCHMARK_JSON=")]
    if r.returncode != 0 or not line:
        return None, (r.stderr or r.stdout)[-400:]
    return json.loads(line[-1].split("=", 1)[1]), None


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu-counts", default="1,2,4,8")
    ap.add_argument("--total-steps", type=int, default=28)
    ap.add_argument("--num-samples", type=int, default=10)
    ap.add_argument("--warmup-samples", type=int, default=2)
    ap.add_argument("--model", default="dummy", choices=["dummy", "svd"])
    ap.add_argument("--latent-height", type=int, default=40)
    ap.add_argument("--latent-width", type=int, default=72)
    ap.add_argument("--latent-frames", type=int, default=14)
    ap.add_argument("--hidden-channels", type=int, default=64)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--allow-uneven", action="store_true")
    ap.add_argument("--schedule", default="fixed", choices=["fixed", "ring"])
    ap.add_argument("--results-dir", default=os.path.join(ROOT, "benchmark_results"))
    ap.add_argument("--timeout", type=int, default=1800)
    a = ap.parse_args(argv)
    os.makedirs(a.results_dir, exist_ok=True)
    stamp = time.strftime("%Y%m%d_%H%M%S")
    csv_path = os.path.join(a.results_dir, f"comparison_{stamp}.csv")
    rows, port = [], 29720
    pp_extra = (["--allow-uneven"] if a.allow_uneven else []) + ["--schedule", a.schedule]
    for n in [int(x) for x in a.gpu_counts.split(",")]:
        for mode, module, extra in (("pipeline_parallel", "src.modes.benchmark", pp_extra),
                                    ("data_parallel", "src.modes.benchmark_data_parallel", [])):
            port += 1
            res, err = run_one(module, n, a, port, extra)
            if res is None:
                print(f"  {mode} x{n}: FAILED: {err}", file=sys.stderr)
                rows.append([mode, n, a.total_steps, "error", a.num_samples, "", "", ""])
                continue
            rows.append([mode, n, a.total_steps, json.dumps(res["steps_per_gpu"]).replace(",", ";"), a.num_samples,
                         res["first_sample_time_s"], res["avg_sample_time_s"], res["throughput_samples_per_s"]])
            print(f"  {mode} x{n} -> throughput: {res['throughput_samples_per_s']} samples/s, avg: "
                  f"{res['avg_sample_time_s']} s/sample", flush=True)
    with open(csv_path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(COLUMNS)
        w.writerows(rows)
    width = [max(len(str(r[i])) for r in [COLUMNS] + rows) for i in range(len(COLUMNS))]
    for r in [COLUMNS] + rows:
        print("  ".join(str(c).ljust(width[i]) for i, c in enumerate(r)))
    print("CSV:", csv_path)
    return csv_path, rows


if __name__ == "__main__":
    main()
