"""Stage-balance sweep (BASELINE config 5): SVD-XT 25 frames, {25, 35} Euler steps on W pipeline stages
(W = the torchrun world size, 7 or 8), uneven step assignment.  For each step count, on the same model:
  * the reference's fixed placement (stage s on rank s): first-video latency and steady-state videos/min
    (steady = (n-1) / (t_last_done - t_first_done) on the last rank, as the reference's benchmark mode does);
  * the rotating (ring) placement: videos/min over the whole stream.
One JSON line per (steps, world) on rank 0; appended to gpurun_out/stage_sweep.jsonl.
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/stage_sweep.py [--steps 25 35]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import vdpp_b200  # noqa: E402,F401
from vdpp_b200.models import StableVideoUNet  # noqa: E402
from vdpp_b200.pipeline import LatentSpec, PipelineConfig, PipelineStage, stage_sizes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, nargs="+", default=[25, 35])
    ap.add_argument("--frames", type=int, default=25)
    ap.add_argument("--videos-per-rank", type=int, default=2)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    F_, H, W = a.frames, 72, 128
    shape = torch.Size((1, 4, F_, H, W))
    last = rank == world - 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for T in a.steps:
        model = StableVideoUNet.from_pretrained("random-init:0", timesteps=StableVideoUNet._default_timestep_schedule(T),
                                                device=dev)
        torch.manual_seed(43)
        model.set_dummy_conditioning(batch_size=1, num_frames=F_, height=H, width=W, device=dev)
        model.use_cuda_graph = True
        spec = LatentSpec(shape=shape, dtype=torch.float16, device=dev)
        cfg = PipelineConfig(total_steps=T, world_size=world, rank=rank, timesteps=list(range(T)), latent_spec=spec,
                             allow_uneven=True)
        stage = PipelineStage(model=model, config=cfg)
        n = a.videos_per_rank * world

        def inputs(base):
            out = []
            for i in range(n):
                g = torch.Generator(device=dev).manual_seed(1000 * base + i)
                out.append(torch.randn(shape, device=dev, generator=g, dtype=torch.float32).half() * model.init_noise_sigma)
            return out

        # warm-up: every rank runs every stage once (ring), so all T step graphs exist everywhere
        warm = inputs(1)
        if world > 1:
            stage.run_many_ring(world, input_supplier=lambda i: warm[i])
        else:
            stage.run_many(1, input_supplier=lambda i: warm[i])
        barrier()

        # ---- fixed placement (the reference's): per-sample completion times on the last rank
        lin_in = inputs(2)
        barrier()
        t0 = time.perf_counter()
        done = []
        for i in range(n):
            out = stage._process_single_latent(lin_in[i] if rank == 0 else None, sample_idx=i)
            if last:
                torch.cuda.synchronize()
                done.append(time.perf_counter() - t0)
        barrier()
        t_lin = time.perf_counter() - t0
        rec = None
        if last:
            steady = (n - 1) / (done[-1] - done[0]) * 60.0 if n > 1 else None
            rec = {"first_video_s": done[0], "steady_videos_per_min": steady, "stream_videos_per_min": n / t_lin * 60.0}
        if world > 1:
            box = [rec]
            dist.broadcast_object_list(box, src=world - 1)
            rec = box[0]

        # ---- rotating placement
        ring_rec = None
        if world > 1:
            ring_in = inputs(3)
            barrier()
            t0 = time.perf_counter()
            stage.run_many_ring(n, input_supplier=lambda i: ring_in[i])
            barrier()
            t_ring = time.perf_counter() - t0
            ring_rec = {"stream_videos_per_min": n / t_ring * 60.0, "batch_latency_s": t_ring / a.videos_per_rank}

        if rank == 0:
            sizes = stage_sizes(T, world)
            line = {"frames": F_, "denoise_steps": T, "stages": world, "stage_sizes": sizes,
                    "ideal_efficiency_fixed": T / (world * max(sizes)), "videos": n,
                    "fixed_placement": rec, "ring_placement": ring_rec, "cuda_graph": True}
            print(json.dumps(line), flush=True)
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "stage_sweep.jsonl"), "a") as f:
                f.write(json.dumps(line) + "\n")
        del stage, model
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
