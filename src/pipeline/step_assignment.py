from vdpp_b200.pipeline.step_assignment import (StepRange, assign_steps, assign_steps_uneven,  # noqa: F401
                                                stage_sizes)
