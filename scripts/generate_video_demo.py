"""Entry point with the reference's path and flags (``scripts/generate_video_demo.py``):
   python scripts/generate_video_demo.py --input-image IMG [--num-frames 25] [--model-id DIR | random-init[:seed]] ...
   torchrun --nproc_per_node N scripts/generate_video_demo.py ...            (step pipeline over N GPUs)
The implementation lives in ``video-diffusion-pipeline-parallel_b200/modes/generate_video.py``."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vdpp_b200  # noqa: E402,F401
from vdpp_b200.modes.generate_video import main  # noqa: E402

if __name__ == "__main__":
    main()
