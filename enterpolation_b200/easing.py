"""Built-in easing functions of the reference (src/easing/mod.rs:67-132, src/easing/plateau.rs) as
descriptors for `Linear.builder()....easing(...)`. Only these cross to the device; an arbitrary
`FuncEase` closure cannot."""
from __future__ import annotations

from dataclasses import dataclass

from . import _abi


@dataclass(frozen=True)
class Easing:
    kind: int
    n: int = 0
    param: float = 0.0

    def triple(self):
        return (self.kind, self.n, self.param)


def Identity() -> Easing:                 # easing/mod.rs:48-72
    return Easing(_abi.EASE_IDENTITY)


flip = Easing(_abi.EASE_FLIP)             # mod.rs:86-91
smoothstep = Easing(_abi.EASE_SMOOTHSTEP)  # mod.rs:114-121
smootherstep = Easing(_abi.EASE_SMOOTHERSTEP)  # mod.rs:123-132


def smoothstart(n: int) -> Easing:        # smoothstart::<R, N>, mod.rs:93-102
    return Easing(_abi.EASE_SMOOTHSTART, int(n))


def smoothend(n: int) -> Easing:          # smoothend::<R, N>, mod.rs:104-112
    return Easing(_abi.EASE_SMOOTHEND, int(n))


def Plateau(strength: float) -> Easing:   # Plateau::new(strength), plateau.rs:19-25
    return Easing(_abi.EASE_PLATEAU, 0, float(strength))
