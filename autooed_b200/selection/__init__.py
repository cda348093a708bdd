from autooed_b200.selection.hvi import HypervolumeImprovement
from autooed_b200.selection.simple import Direct, Random, Uncertainty
