from autooed_b200.selection.hvi import HypervolumeImprovement
from autooed_b200.selection.uncertainty import Uncertainty
from autooed_b200.selection.direct import Direct
