from .pipeline import (LatentSpec, PipelineConfig, PipelineStage, run_pipeline_latents,
                       run_single_latent)
from .step_assignment import StepRange, assign_steps, assign_steps_uneven, stage_sizes

__all__ = [
    "StepRange", "assign_steps", "assign_steps_uneven", "stage_sizes",
    "LatentSpec", "PipelineStage", "PipelineConfig", "run_single_latent", "run_pipeline_latents",
]
