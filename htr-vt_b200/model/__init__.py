"""Drop-in for the reference's `model` package (model_v1/model/): `from model import HTR_VT`."""
from . import HTR_VT  # noqa: F401
