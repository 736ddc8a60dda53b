"""Training-loop pieces around the hot path with the reference's names (model_v1/utils/): sam.SAM, utils.ModelEma."""
from . import sam, utils  # noqa: F401
