"""Host-side mirror of ``src/monteCarloIllumination.f95``.

``type(photonStream)`` (ILL:35-42) pre-generates a batch of starting positions and
directions on the host.  Here the stream is only a *description* of the source; each photon
samples its own start inside the kernel (``csrc/mcb_fast.cu`` refill block, ILL:88-96 and
ILL:481-515), so nothing of size O(photons) is ever materialised or copied.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from .emissionAndBroadBandWeights import Weights
from .RandomNumbersForMC import randomNumberSequence


@dataclass
class photonStream:
    kind: str                              # "directional" | "bbemission"
    numberOfPhotons: int
    currentPhoton: int = 0                 # 1-based like the reference; 0 = not initialised
    solarMu: float = 1.0
    solarAzimuth: float = 0.0
    weights: Optional[Weights] = None
    firstPhotonId: int = 0                 # global id of this stream's first photon


def new_PhotonStream(solarMu=None, solarAzimuth=None, numberOfPhotons=0, randomNumbers: randomNumberSequence = None,
                     theseWeights: Weights = None, iLambda: int = 1) -> photonStream:
    """``new_PhotonStream``: directional/solar (ILL:62-101) or broadband emission (ILL:431-522).

    The other four constructors (random azimuth, flux, spotlight, LW emission; ILL:103-333)
    are not called by the current driver and are not provided.
    """
    if numberOfPhotons < 0:
        raise ValueError("setIllumination: must ask for non-negative number of photons.")
    if randomNumbers is None:
        raise ValueError("new_PhotonStream: randomNumbers is required")
    if theseWeights is not None:
        ps = photonStream("bbemission", int(numberOfPhotons), 1, weights=theseWeights)
    else:
        if solarAzimuth < 0.0 or solarAzimuth > 360.0:
            raise ValueError("setIllumination: solarAzimuth out of bounds")
        if abs(solarMu) > 1.0 or abs(solarMu) <= np.finfo(np.float32).tiny:
            raise ValueError("setIllumination: solarMu out of bounds")
        ps = photonStream("directional", int(numberOfPhotons), 1, float(solarMu), float(solarAzimuth))
    # the stream claims its photon ids now, as the reference draws its positions now (ILL:88-92)
    ps.firstPhotonId = randomNumbers.nextPhotonId
    randomNumbers.nextPhotonId += int(numberOfPhotons)
    return ps


def morePhotonsExist(photons: photonStream) -> bool:
    """ILL:540-546."""
    return 0 < photons.currentPhoton <= photons.numberOfPhotons


def setCurrentPhoton(photons: photonStream, currentPhoton: int) -> None:
    """ILL:548-559."""
    if photons.currentPhoton < 1:
        raise ValueError("getNextPhoton: photons have not been initialized.")
    photons.currentPhoton = int(currentPhoton)


def finalize_PhotonStream(photons: photonStream) -> None:
    photons.currentPhoton = 0
    photons.numberOfPhotons = 0
