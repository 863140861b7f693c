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

/*
 * Scene presets: `Preset` and `Presets` of particle_io/src/presets.rs:84-154, the in-memory scene library the editor
 * keeps ("User Presets", particle_editor/src/editor.rs:961-1072). A preset holds what a scene is made of -- a name, the
 * box size, the two species' Mie parameters and the particle list -- and nothing of how it is stepped.
 * The list is an opaque handle; indices are positions in it. Functions return 0 on success and -1 on a null or
 * out-of-range argument (where the reference would panic on the index), except where stated.
 */
typedef struct PsimPresets PsimPresets;

PsimPresets* psim_presets_new(void);                          /* Presets::new, presets.rs:127-129 */
void psim_presets_destroy(PsimPresets* presets);
size_t psim_presets_len(const PsimPresets* presets);          /* get_presets_len, presets.rs:131-133 */
/* add_preset(Preset::from_frame(name, frame)), presets.rs:107-119,139-141: box, species and all particle_count records
 * of `frame` (null records included, like particles().to_vec()). Returns the new preset's index, or -1. */
long psim_presets_add_from_frame(PsimPresets* presets, const char* name, const FrameHeader* frame);
/* change_preset(Preset::from_frame(name, frame), index), presets.rs:147-153: an index past the end is IGNORED (returns 0). */
int psim_presets_change_from_frame(PsimPresets* presets, size_t index, const char* name, const FrameHeader* frame);
/* add_preset(get_preset(index).clone()) under another name (the editor's "duplicate", editor.rs:998-1000); returns the new index */
long psim_presets_duplicate(PsimPresets* presets, size_t index, const char* new_name);
int psim_presets_delete(PsimPresets* presets, size_t index);  /* delete_preset, presets.rs:143-145 */
const char* psim_preset_name(const PsimPresets* presets, size_t index);      /* owned by the list; NULL if out of range */
uint32_t psim_preset_particle_count(const PsimPresets* presets, size_t index);
/* Preset::to_frame, presets.rs:92-105: a NEW frame -- frame_header_init() defaults -- with the preset's box, species and
 * particles. `dst` is a caller buffer of packet_size(capacity) bytes; -1 if the preset holds more than `capacity`. */
int psim_preset_to_frame(const PsimPresets* presets, size_t index, FrameHeader* dst, uint32_t capacity);

#ifdef __cplusplus
} /* extern "C" */
#endif
