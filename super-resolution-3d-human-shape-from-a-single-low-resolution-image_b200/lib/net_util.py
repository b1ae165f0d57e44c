"""Drop-in for the part of the reference's ``lib/net_util.py`` the inference path uses."""
from .train_util import gen_mesh, make_calib  # noqa: F401  (the reference keeps a duplicate here, :50-82)
