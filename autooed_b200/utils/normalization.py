"""Data normalisation for fitting the surrogate — same semantics as the reference's
`autooed/utils/normalization.py` (BoundedScaler :24-35, Normalization :38-100, StandardNormalization :103-109):
X -> clip((X - xl) / (xu - xl), 0, 1); Y -> (Y - mean) / std with the population std, constant columns -> scale 1
(sklearn StandardScaler)."""
import numpy as np


class BoundedScaler:
    def __init__(self, bounds):
        self.bounds = bounds

    def fit(self, X):
        return self

    def transform(self, X):
        return np.clip((X - self.bounds[0]) / (self.bounds[1] - self.bounds[0]), 0, 1)

    def inverse_transform(self, X):
        return np.clip(X, 0, 1) * (self.bounds[1] - self.bounds[0]) + self.bounds[0]


class StandardScaler:
    """mean_/scale_/var_ as sklearn.preprocessing.StandardScaler exposes them."""

    def fit(self, Y):
        Y = np.asarray(Y, dtype=np.float64)
        self.mean_ = Y.mean(axis=0)
        self.var_ = Y.var(axis=0)
        scale = np.sqrt(self.var_)
        scale[scale < 10 * np.finfo(np.float64).eps] = 1.0
        self.scale_ = scale
        return self

    def transform(self, Y):
        return (np.asarray(Y, dtype=np.float64) - self.mean_) / self.scale_

    def inverse_transform(self, Y):
        return np.asarray(Y, dtype=np.float64) * self.scale_ + self.mean_


class Normalization:
    def __init__(self, x_scaler, y_scaler):
        self.x_scaler, self.y_scaler = x_scaler, y_scaler
        self.fitted = False

    def fit(self, x, y):
        self.x_scaler = self.x_scaler.fit(x)
        self.y_scaler = self.y_scaler.fit(y)
        self.fitted = True

    def _pair(self, fx, fy, x, y, what):
        if not self.fitted:
            raise Exception(f'Not fitted before {what}')
        if x is not None and y is not None:
            return fx(x), fy(y)
        if x is not None:
            return fx(x)
        if y is not None:
            return fy(y)
        raise Exception(f'Cannot {what} nothing')

    def do(self, x=None, y=None):
        return self._pair(self.x_scaler.transform, self.y_scaler.transform, x, y, 'normalization')

    def undo(self, x=None, y=None):
        return self._pair(self.x_scaler.inverse_transform, self.y_scaler.inverse_transform, x, y, 'unnormalization')

    def scale(self, x=None, y=None):
        return self._pair(lambda v: v / self.x_scaler.scale_, lambda v: v / self.y_scaler.scale_, x, y, 'scaling')

    def rescale(self, x=None, y=None):
        return self._pair(lambda v: self.x_scaler.scale_ * v, lambda v: self.y_scaler.scale_ * v, x, y, 'rescaling')


class StandardNormalization(Normalization):
    def __init__(self, x_bound):
        super().__init__(BoundedScaler(x_bound), StandardScaler())
