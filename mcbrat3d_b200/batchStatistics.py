"""The driver's per-batch statistics (``Drivers/monteCarloDriver.f95``), host side.

First and second moments over batches weighted by photons per batch (DRV:1023-1052) and the
final mean / standard error (DRV:1188-1228).  The Fortran driver itself is kept as the host in
a deployment (north star); this mirror exists so that the Python host and the tests form the
statistics exactly the way the reference does.
"""
from __future__ import annotations

from typing import Dict

import numpy as np


class BatchStatistics:
    def __init__(self):
        self.moment1: Dict[str, np.ndarray] = {}
        self.moment2: Dict[str, np.ndarray] = {}
        self.totalNumPhotons = 0
        self.batchesCompleted = 0

    def accumulate(self, results: Dict[str, np.ndarray], numPhotonsProcessed: int) -> None:
        """DRV:1023-1052: stats(:,1) += x*n ; stats(:,2) += n*x**2 (double precision)."""
        n = float(numPhotonsProcessed)
        for k, v in results.items():
            x = np.asarray(v, dtype=np.float64)
            if k not in self.moment1:
                self.moment1[k] = np.zeros_like(x)
                self.moment2[k] = np.zeros_like(x)
            self.moment1[k] += x * n
            self.moment2[k] += n * x ** 2.0
        self.totalNumPhotons += int(numPhotonsProcessed)
        self.batchesCompleted += 1

    def merge(self, other: "BatchStatistics") -> None:
        """sumAcrossProcesses over workers (DRV:1151-1166)."""
        for k in other.moment1:
            if k not in self.moment1:
                self.moment1[k] = other.moment1[k].copy(); self.moment2[k] = other.moment2[k].copy()
            else:
                self.moment1[k] += other.moment1[k]; self.moment2[k] += other.moment2[k]
        self.totalNumPhotons += other.totalNumPhotons
        self.batchesCompleted += other.batchesCompleted

    def finalise(self, solarFlux: float = 1.0):
        """DRV:1188-1228: returns ({name: mean}, {name: standard error})."""
        mean, err = {}, {}
        for k in self.moment1:
            m1 = solarFlux * self.moment1[k] / self.totalNumPhotons
            m2 = solarFlux * (solarFlux * self.moment2[k] / self.totalNumPhotons)
            mean[k] = m1
            err[k] = np.sqrt(np.maximum(0.0, m2 - m1 ** 2.0) / (self.batchesCompleted - 1))
        return mean, err
