/*
 * psim_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-file CPU restatement of the reference's per-timestep particle update, with the
 * grid size as a run-time parameter instead of the reference's compile-time 64x64
 * (cuda_simulator/src/kernel.cuh:14-20).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load it; the product libraries never do.
 *
 * Parity status: PINNED -- tests/test_oracle.py checks this restatement bit-for-bit against the
 * reference's own code compiled from /root/reference (oracle/_ref, see oracle/Makefile) and against
 * the golden vectors under tests/golden/ that were generated from that compiled reference.
 */
#pragma once
#include "particle_io.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OracleGrid {
    uint32_t lx;       /* log2 cells in x (reference BUCKETS_X_LOG2)   */
    uint32_t ly;       /* log2 cells in y (reference BUCKETS_Y_LOG2)   */
    uint32_t capacity; /* slots per cell  (reference BUCKET_CAPACITY)  */
} OracleGrid;

/* slots in a slot array for this grid = cells * capacity */
uint64_t oracle_slot_count(OracleGrid g);

/* C = n/(n-m) * (n/m)^(m/(n-m)) in fp32 (particle.cuh:53-55) */
float oracle_params_C(MiePotentialParams p);
/* scalar Mie force F(r) (particle.cuh:63-66) */
float oracle_f_force(MiePotentialParams p, float r);

/* Initial binning (kernel.cuh:210-239). Returns the number of particles that did NOT fit (a cell
 * already holding `capacity` particles): the reference has no such check and overwrites the next
 * cell's slot 0 (kernel.cuh:228-229); the oracle drops them instead and reports the count so that a
 * parity scene can assert it is zero. */
uint32_t oracle_prepare(const FrameHeader* src, Particle* slots, OracleGrid g);

/* One re-binning pass over every cell (kernel_bucket.cuh:5-39). Returns the number of live particles
 * after the move (the reference silently loses the 17th particle of a cell and anything that moved
 * more than one cell). */
uint64_t oracle_move(const Particle* src, Particle* dst, OracleGrid g);

/* One force + leapfrog step over every slot (kernel_bucket.cuh:40-94 with particle.cuh:41-47,
 * 63-71,97-144). `threads` > 1 splits the slots over that many pthreads (results are identical:
 * every slot is independent). */
void oracle_step(const Particle* src, Particle* dst, const FrameMetadata* meta, OracleGrid g, uint32_t threads);

/* EXTENSION, not in the reference (which steps every particle with metadata.particles[0], kernel_bucket.cuh:52): the same
 * step with per-species Mie parameters -- a pair uses particles[s] when both particles are of species s = min(ty, 1), the
 * Lorentz-Berthelot mix (mean sigma, geometric-mean epsilon, mean exponents) when they differ; the wall term uses the
 * particle's own. It is the specification PsimConfig.species_physics is tested against; nothing in the reference pins it. */
void oracle_step_species(const Particle* src, Particle* dst, const FrameMetadata* meta, OracleGrid g, uint32_t threads);

/* The net force on every slot before integration (same accumulation order as oracle_step), plus
 * the largest single pair-force magnitude seen by that slot; for tolerance definitions. */
void oracle_forces(const Particle* src, const FrameMetadata* meta, OracleGrid g, float* fx, float* fy,
                   float* max_pair);

/* One frame = the reference's step / re-bin schedule (kernel_bucket.cuh:181-206) starting from
 * buf[0]; three slot arrays are rotated exactly like D_BUFFER_{0,1,INTERNAL}. The result is left in
 * buf[1]. Returns the number of steps executed (may be steps_per_frame + 1); *moves gets the number
 * of re-binning passes. */
uint32_t oracle_run_frame(Particle* buf0, Particle* buf1, Particle* buf2, const FrameMetadata* meta, OracleGrid g,
                          uint32_t threads, uint32_t* moves);

/* Order-preserving compaction of a slot array into a frame (particle.rs:371-379). dst must have room
 * for every live particle. */
void oracle_compact(const Particle* slots, const FrameMetadata* meta, OracleGrid g, FrameHeader* dst);

/* Diagnostics in double precision over a slot array: kinetic energy (half-step velocities as stored),
 * pair potential energy over the 3x3-cell stencil (each pair once), wall potential energy, and total
 * momentum. out[0]=KE, out[1]=PE_pair, out[2]=PE_wall, out[3]=px, out[4]=py, out[5]=live count. */
void oracle_diagnostics(const Particle* slots, const FrameMetadata* meta, OracleGrid g, double out[6]);

#ifdef __cplusplus
}
#endif
