"""CPU baseline with the REFERENCE'S OWN code: its unmodified ``PipelineStage`` + ``DummyUNet``
over gloo on the host cores, imported from an installed copy of the reference (``baseline/_ref``, put there by
``pip install --no-index --no-deps --target baseline/_ref <reference>``; or ``$VDPP_REFERENCE_ROOT``).  Nothing of
this repository's package is imported here: the script is run with the reference tree first on ``sys.path``.

  python tools/reference_simulator.py --total-steps 25                                   (one rank)
  python -m torch.distributed.run --nproc-per-node 4 ... tools/reference_simulator.py --total-steps 28

BASELINE config 1 workload: DummyUNet(channels=4) (hidden 16, LayerNorm on), latent [1,4,14,64,64] fp32, samples
``manual_seed(42 + idx); randn`` on rank 0, timesteps ``reversed(range(T))`` (reference simulator.py:77-92).  Every rank
seeds before building the model (the reference does not, and is not reproducible without it: SURVEY section 0 item 5).
Timing as reference benchmark.py:233-267: per-sample completion times on the last rank, steady throughput = measured
samples / sum of their intervals after the warm-up samples.  Prints one line ``REFERENCE_SIMULATOR_JSON={...}`` on the
last rank, with the SHA-256 of the final latent of sample 0 (identical at every world size for a given step count).
"""
import argparse
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def reference_root():
    for cand in (os.environ.get("VDPP_REFERENCE_ROOT"), os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "src", "pipeline", "pipeline.py")):
            return cand
    return None


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--total-steps", type=int, default=25)
    ap.add_argument("--samples", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--threads", type=int, default=0, help="intra-op threads per rank (0: torch's default)")
    a = ap.parse_args()
    ref = reference_root()
    if ref is None:
        print('REFERENCE_SIMULATOR_JSON={"error": "no reference tree (baseline/_ref or $VDPP_REFERENCE_ROOT)"}')
        return
    # the reference's `src` package must win over this repository's compatibility namespace of the same name
    sys.path[:] = [ref] + [p for p in sys.path if os.path.abspath(p or ".") not in (ROOT, HERE)]
    import torch
    import torch.distributed as dist
    import src
    assert os.path.abspath(src.__file__).startswith(os.path.abspath(ref)), src.__file__
    from src.distributed.setup import finalize_distributed, init_distributed
    from src.models.dummy_unet import DummyUNet
    from src.pipeline.pipeline import LatentSpec, PipelineConfig, PipelineStage

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if a.threads > 0:
        torch.set_num_threads(a.threads)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29611")
        init_distributed(backend="gloo", rank=rank, world_size=world)
    T = a.total_steps
    shape = torch.Size((1, 4, 14, 64, 64))
    dev = torch.device("cpu")
    torch.manual_seed(1234)
    model = DummyUNet(channels=4).to(dev).eval()
    cfg = PipelineConfig(total_steps=T, world_size=world, rank=rank, timesteps=list(reversed(range(T))),
                         latent_spec=LatentSpec(shape=shape, dtype=torch.float32, device=dev))
    stage = PipelineStage(model=model, config=cfg)

    def supplier(idx: int):
        torch.manual_seed(42 + idx)
        return torch.randn(shape)

    total = a.warmup + a.samples
    ends, first = [], None
    if world > 1:
        dist.barrier()
    start = time.perf_counter()
    with torch.no_grad():
        for idx in range(total):
            out = stage._process_single_latent(supplier(idx) if rank == 0 else None, sample_idx=idx)
            if rank == world - 1:
                ends.append(time.perf_counter())
                if idx == 0:
                    first = out.clone()
    if rank == world - 1:
        per = [e - (start if i == 0 else ends[i - 1]) for i, e in enumerate(ends)]
        meas = per[a.warmup:]
        print("REFERENCE_SIMULATOR_JSON=" + json.dumps({
            "code": "reference PipelineStage + DummyUNet, unmodified, from " + os.path.relpath(ref, ROOT),
            "workload": "DummyUNet(channels=4) latent 1x4x14x64x64 fp32, gloo, CPU",
            "total_steps": T, "world_size": world, "threads_per_rank": torch.get_num_threads(),
            "host_cores": os.cpu_count(), "samples_measured": len(meas), "warmup_samples": a.warmup,
            "first_sample_s": round(per[0], 4), "steady_s_per_sample": round(sum(meas) / len(meas), 4),
            "samples_per_s": round(len(meas) / sum(meas), 4), "videos_per_min": round(60.0 * len(meas) / sum(meas), 3),
            "ms_per_step_per_rank": round(1000.0 * sum(meas) / len(meas) / (T // world), 3),
            "sample0_sha256": hashlib.sha256(first.numpy().tobytes()).hexdigest(), "torch": torch.__version__,
        }), flush=True)
    if world > 1:
        finalize_distributed()


if __name__ == "__main__":
    main()
