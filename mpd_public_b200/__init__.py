"""mpd_public_b200 — B200-native guided-diffusion trajectory sampler (drop-in for the guided
`p_sample_loop` path of jacarvalho/mpd-public; see DESIGN.md and INTEGRATION.md).

Exports the reference's public names for this path (`mpd.models`, `mpd.models.diffusion_models.*`):
"""
from .synthetic import UNET_DIM_MULTS  # noqa: F401
from .temporal_unet import TemporalUnet  # noqa: F401
from .diffusion_model import GaussianDiffusionModel, make_timesteps  # noqa: F401
from .sample_functions import apply_hard_conditioning, extract, ddpm_sample_fn, guide_gradient_steps  # noqa: F401
from .guides import GuideManagerTrajectoriesWithVelocity, GuideManagerTrajectories  # noqa: F401
from .costs import (CostCollision, CostGPTrajectory, CostComposite, GridSDFField, WorkspaceBoundaryField,  # noqa: F401
                    SelfCollisionField)
from .normalization import LimitsNormalizer, DatasetNormalizer  # noqa: F401
from .planning import (TrajectoryDataset, PlanningTask, Robot, compute_smoothness, compute_path_length,  # noqa: F401
                       compute_variance_waypoints)

__version__ = "0.1.0"
