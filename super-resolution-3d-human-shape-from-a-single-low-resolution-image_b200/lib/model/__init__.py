from .SurfaceClassifier import SurfaceClassifier
from .SuRSNet import SuRSNet

BaseSuRSNet = SuRSNet   # the reference's base class is folded into SuRSNet here
