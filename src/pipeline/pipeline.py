from vdpp_b200.pipeline.pipeline import (InputSupplier, LatentSpec, PipelineConfig, PipelineStage,  # noqa: F401
                                         run_pipeline_latents, run_single_latent)
