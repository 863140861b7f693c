/*
 * psim_scene.h -- seeded scene generators for the particle_io frame format.
 *
 * Restates the lattice generators of the reference's particle_io crate
 * (particle_io/src/presets.rs:16-82) behind a C ABI, with a seeded generator so that scenes are
 * reproducible (the reference draws velocities from the unseeded thread RNG, presets.rs:34,64).
 * Implemented in libparticle_io_c.so (particle_simulator_b200/csrc/scene.cpp); CPU only.
 * Functions return 0 on success, -1 if the frame has no room or an argument is invalid.
 */
#pragma once

#include "particle_io.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ParticleLattice::{hex_square, square, random_vel} (presets.rs:16-82).
 * They append nx*ny particles of species `ty` to `frame` (which must have room for them after
 * frame->particle_count) centred on (center_x, center_y) metres, spaced distance_factor * r0 apart
 * with speeds uniform in [v_min, v_max] and uniformly random directions.
 */
int psim_scene_hex_square(FrameHeader* frame, uint32_t capacity, uint32_t nx, uint32_t ny, double center_x,
                          double center_y, float distance_factor, float v_min, float v_max, int32_t ty,
                          uint64_t seed);
int psim_scene_square(FrameHeader* frame, uint32_t capacity, uint32_t nx, uint32_t ny, double center_x,
                      double center_y, float distance_factor, float v_min, float v_max, int32_t ty,
                      uint64_t seed);
/* Lattice rows [row_begin, row_end) of the nx * ny hex lattice centred on (center_x, center_y): the same
 * geometry as psim_scene_hex_square, emitted row by row with one velocity stream per lattice row, so that
 * every slab of a row decomposition can generate its own part of one global crystal independently. */
int psim_scene_hex_rows(FrameHeader* frame, uint32_t capacity, uint32_t nx, uint32_t ny, uint32_t row_begin,
                        uint32_t row_end, double center_x, double center_y, float distance_factor, float v_min,
                        float v_max, int32_t ty, uint64_t seed);
/* Gas: `count` particles at uniformly random positions at least `margin` metres from the walls and
 * at least `min_dist` metres from each other and from the particles already in the frame
 * (rejection sampling), speeds as above. */
int psim_scene_gas(FrameHeader* frame, uint32_t capacity, uint32_t count, double margin, double min_dist,
                   float v_min, float v_max, int32_t ty, uint64_t seed);

/* r0 = sigma * (n/m)^(1/(n-m)), the zero-force distance (particle.rs:44-49), in double like the reference. */
double psim_force0_r(MiePotentialParams p);

#ifdef __cplusplus
} /* extern "C" */
#endif
