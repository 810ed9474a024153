"""Host-side mirror of ``src/RandomNumbersForMC.f95``'s public surface.

The reference threads one sequential MT19937 state (``type(randomNumberSequence)``,
RNG:104-107) through photon-stream creation and transport.  On the GPU every photon owns a
counter-based Philox4x32-10 stream keyed by ``(seed, global photon id)`` (``csrc/mcb_device.cuh``),
so the host object only carries that key and the next unused photon id.  Results therefore do
not depend on batch size or on how photons are split over GPUs (unlike the reference, where
they depend on both).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence, Union

_MASK64 = (1 << 64) - 1


def _mix64(z: int) -> int:
    """splitmix64 finaliser: spreads the driver's small integer seeds over the 64-bit key."""
    z = (z + 0x9E3779B97F4A7C15) & _MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK64
    return z ^ (z >> 31)


@dataclass
class randomNumberSequence:
    seed: int = 0            # 64-bit Philox key
    nextPhotonId: int = 0    # first photon id not yet handed to a batch


def new_RandomNumberSequence(seed: Union[int, Sequence[int]]) -> randomNumberSequence:
    """``new_RandomNumberSequence`` (scalar RNG:171-187, vector RNG:189-241).

    The driver seeds with ``(/ iseed, thisProc, thisThread /)`` (DRV:901); distinct seed
    vectors give distinct keys.
    """
    if isinstance(seed, int):
        key = _mix64(seed & _MASK64)
    else:
        key = 0
        for s in seed:
            key = _mix64(key ^ (int(s) & _MASK64))
    return randomNumberSequence(seed=key, nextPhotonId=0)


def finalize_RandomNumberSequence(twister: randomNumberSequence) -> None:
    twister.seed = 0
    twister.nextPhotonId = 0
