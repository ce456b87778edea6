"""`reproducibility_metric`: the reference's rollout statistics (src/pcgym/evaluation_metrics.py:182-327) over the
`reps` axis of a `policy_eval.get_rollouts()` dictionary -- mean / median performance, std / MAD dispersion and
their scalarisation `performance + w * dispersion` (Flageat et al. 2024) -- for numpy arrays or device tensors.

The `reps` axis is the environment batch of the B200 engine, so with `as_torch=True` rollouts these are device
reductions over the last axis.  Same class and method names as the reference.  One visible difference: the
reference's MAD subtracts `np.median(data, axis=-1)` without keeping the reduced axis
(evaluation_metrics.py:113-116), which only broadcasts when the trailing dimensions happen to match ("currently
only works for the reward component"); here the median keeps its axis, which gives the same numbers whenever the
reference's expression is valid and the intended statistic otherwise.
"""
from __future__ import annotations

from abc import ABC
from typing import Any, Callable, Dict, Optional

import numpy as np
import torch


def _is_t(x):
    return isinstance(x, torch.Tensor)


class metric_base(ABC):
    def __init__(self, scalarised_weight: float) -> None:
        pass

    def evaluate(self, policy_evaluator: Any) -> Any:
        raise NotImplementedError("Subclasses must implement the evaluate method.")

    def policy_dispersion_metric(self, data: Dict[str, Any]) -> Any:
        raise NotImplementedError("Subclasses must implement the policy_dispersion_metric method.")

    def policy_performance_metric(self, data: Dict[str, Any]) -> Any:
        raise NotImplementedError("Subclasses must implement the policy_performance_metric method.")

    def scalarised_performance(self, data: Dict[str, Any]) -> Any:
        raise NotImplementedError("Subclasses must implement the scalarised_performance method.")


class standard_deviation:
    def __init__(self, data) -> None:
        self.data = data

    def get_value(self):
        d = self.data
        return torch.std(d, dim=-1, unbiased=False) if _is_t(d) else np.std(d, axis=-1)     # np.std: ddof = 0


def _median(d, keepdim=False):
    if _is_t(d):
        # numpy's median averages the two middle order statistics; torch.median returns the lower one
        return torch.quantile(d, 0.5, dim=-1, keepdim=keepdim, interpolation="midpoint")
    return np.median(d, axis=-1, keepdims=keepdim)


class median_absolute_deviation:
    def __init__(self, data) -> None:
        if data.ndim < 2:
            data = data.reshape((data.shape[0], 1))
        self.data = data

    def get_value(self):
        d = self.data
        dev = (d - _median(d, keepdim=True))
        return _median(dev.abs() if _is_t(dev) else np.abs(dev))


class mean_performance:
    def __init__(self, data) -> None:
        self.data = data

    def get_value(self):
        d = self.data
        return torch.mean(d, dim=-1) if _is_t(d) else np.mean(d, axis=-1)


class median_performance:
    def __init__(self, data) -> None:
        self.data = data

    def get_value(self):
        return _median(self.data)


class reproducibility_metric(metric_base):
    def __init__(self, dispersion: str, performance: str, scalarised_weight: float) -> None:
        # the weight is in terms of the upper confidence bound; negative for the lower one
        self.scalarised_weight = scalarised_weight
        if dispersion == "std":
            self.dispersion = standard_deviation
        elif dispersion == "mad":
            self.dispersion = median_absolute_deviation
        else:
            raise ValueError("Invalid dispersion metric")
        if performance == "mean":
            self.performance = mean_performance
        elif performance == "median":
            self.performance = median_performance
        else:
            raise ValueError("Invalid performance metric")

    def evaluate(self, policy_evaluator: Any, component: Optional[str] = None):
        try:
            self.data = policy_evaluator.data          # evaluation_metrics.py:222-225
        except Exception:
            self.data = policy_evaluator.get_rollouts()
        return self.scalarised_performance(self.data, component)

    def _apply(self, stat, data, component):
        values = {k: {} for k in data.keys()}
        for policy in data.keys():                     # arrays are (n_x, T, reps): always reduced along reps
            comps = data[policy].keys() if component is None else [component]
            for comp in comps:
                values[policy][comp] = stat(self.determine_op(comp)(data[policy][comp])).get_value()
        return values

    def policy_dispersion_metric(self, data, component: Optional[str]):
        return self._apply(self.dispersion, data, component)

    def policy_performance_metric(self, data, component: Optional[str]):
        return self._apply(self.performance, data, component)

    def scalarised_performance(self, data, component: Optional[str]):
        performance = self.policy_performance_metric(data, component)
        dispersion = self.policy_dispersion_metric(data, component)
        return {k: {comp: performance[k][comp] + self.scalarised_weight * dispersion[k][comp]
                    for comp in performance[k].keys()} for k in performance.keys()}

    def determine_op(self, component: str) -> Callable:
        if component in ("x", "u", "r"):
            return lambda x: x
        if component == "g":                           # greatest violation over the n_g constraints
            return lambda x: (torch.amax(x, dim=0) if _is_t(x) else np.max(x, axis=0))
        raise ValueError("Invalid component")
