"""Scene presets: the in-memory scene library of the reference's particle_io crate
(particle_io/src/presets.rs:84-154: `Preset::{to_frame, from_frame}`, `Presets`), mirrored for hosts that build their
scenes in Python (tests, benchmarks). A preset keeps what a scene is made of -- box size, the two species' Mie
parameters, the particle list -- and nothing of how it is stepped (dt, steps per frame, cursor stay at the defaults of
a new frame, as in the reference)."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .frame import PARTICLE_DTYPE, FrameBuffer


@dataclass
class Preset:
    name: str
    box_size: tuple[float, float]
    particles: np.ndarray       # the two MiePotentialParams of the metadata (structured, shape (2,))
    particles_list: np.ndarray  # PARTICLE_DTYPE records

    def to_frame(self) -> FrameBuffer:
        """presets.rs:92-105: a new frame with default metadata, this preset's box and species, all its particles."""
        fb = FrameBuffer(max(len(self.particles_list), 1))
        fb.metadata["box_width"], fb.metadata["box_height"] = self.box_size
        fb.metadata["particles"] = self.particles
        fb.set_particles(np.ascontiguousarray(self.particles_list, dtype=PARTICLE_DTYPE))
        return fb

    @classmethod
    def from_frame(cls, name: str, frame: FrameBuffer) -> "Preset":
        """presets.rs:107-119."""
        return cls(name, (float(frame.metadata["box_width"]), float(frame.metadata["box_height"])),
                   np.array(frame.metadata["particles"], copy=True), frame.particles.copy())


@dataclass
class Presets:
    """presets.rs:122-154."""
    presets: list[Preset] = field(default_factory=list)

    def get_presets_len(self) -> int:
        return len(self.presets)

    def get_preset(self, ind: int) -> Preset:
        return self.presets[ind]

    def add_preset(self, preset: Preset) -> None:
        self.presets.append(preset)

    def delete_preset(self, ind: int) -> None:
        del self.presets[ind]

    def change_preset(self, preset: Preset, ind: int) -> None:
        if ind >= len(self.presets):  # the reference ignores an index past the end (presets.rs:147-152)
            return
        self.presets[ind] = preset
