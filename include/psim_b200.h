/*
 * psim_b200.h -- C ABI of the B200 particle stepper (the hot path).
 *
 * This is the in-process seam the reference's main loop drives: `kernel_prepare_frame` plus
 * `Kernel::{write, write_metadata, run_async, sync, read}` (reference:
 * cuda_simulator/src/kernel.cuh:88-151,200-250, called from cuda_simulator/src/cuda_simulator.cu:7-38).
 * Plain pointers and sizes only; frames are the `particle_io` wire format of particle_io.h.
 * Implemented by libpsim_b200.so (particle_simulator_b200/csrc/stepper.cu), CUDA sm_100a only:
 * there is no CPU path behind these calls.
 *
 * All functions returning `int` return 0 on success and a negative PSIM_E* code on failure;
 * psim_last_error() gives the message of the last failure on that stepper (or of psim_create
 * when called with NULL).
 */
#pragma once

#include "particle_io.h"

#ifdef __cplusplus
extern "C" {
#endif

#define PSIM_OK 0
#define PSIM_EINVAL (-1)   /* bad argument / configuration                       */
#define PSIM_ECUDA (-2)    /* a CUDA runtime call or kernel failed               */
#define PSIM_ECAPACITY (-3)/* more live particles than PsimConfig.max_particles  */
#define PSIM_ESTATE (-4)   /* call made in the wrong state (e.g. no scene yet)   */

/* Step / re-bin schedules. */
#define PSIM_SCHEDULE_REFERENCE 0 /* `S M S*17 M S*17 ...`, countdown restarts every frame, may run
                                     steps_per_frame+1 steps: reference kernel_bucket.cuh:181-206 */
#define PSIM_SCHEDULE_NATIVE 1    /* exactly steps_per_frame steps, re-bin every `rebin_every` steps,
                                     the counter carries over between frames                     */

typedef struct PsimStepper PsimStepper; /* opaque */

typedef struct PsimConfig {
    uint32_t grid_x_log2;   /* cells in x = 1 << grid_x_log2 (reference: BUCKETS_X_LOG2 = 6, kernel.cuh:15) */
    uint32_t grid_y_log2;   /* cells in y = 1 << grid_y_log2 (reference: BUCKETS_Y_LOG2 = 6, kernel.cuh:16) */
    uint32_t max_particles; /* capacity of the device buffers (live particles)                            */
    uint32_t schedule;      /* PSIM_SCHEDULE_*                                                             */
    uint32_t rebin_every;   /* native schedule only; 0 => 17 (the reference's effective cadence)           */
    int32_t device;         /* CUDA device ordinal; -1 => current device                                   */
    uint32_t use_graph;     /* 1 => capture each frame's launches into a CUDA graph and replay it          */
    uint32_t _reserved[5];
} PsimConfig;

/* Defaults: 64x64 cells (the reference grid), 65536 particles, reference schedule, device -1. */
PsimConfig psim_default_config(void);

int psim_create(const PsimConfig* config, PsimStepper** out);
void psim_destroy(PsimStepper* s);
const char* psim_last_error(const PsimStepper* s);

/* Run all subsequent work of this stepper on an existing CUDA stream (a `cudaStream_t` passed as
 * void*; NULL restores the stepper's own stream). Lets a host time the step loop with events
 * recorded on its own stream. */
int psim_set_stream(PsimStepper* s, void* cuda_stream);

/* Scene ingest = kernel_prepare_frame (MatrixBuckets branch, kernel.cuh:210-239) + Kernel::write
 * (kernel.cuh:103-115): takes a COMPACT host frame (null particles, ty < 0, are skipped), copies it
 * to the device and bins it with a stable counting sort by cell = (x >> (32-LX)) + (y >> (32-LY)) * BX.
 * Order inside a cell = input order, exactly the reference's append order (kernel.cuh:219-229).
 * The frame's metadata becomes the stepper's metadata. Synchronous w.r.t. the host. */
int psim_upload_frame(PsimStepper* s, const FrameHeader* frame);

/* Same, for particle records that already live in device memory (AoS, 20 B each). */
int psim_upload_device(PsimStepper* s, const FrameMetadata* meta, const void* d_particles, uint32_t count);

/* = Kernel::write_metadata (kernel.cuh:96-101): takes effect for the next frame that is enqueued. */
int psim_set_metadata(PsimStepper* s, const FrameMetadata* meta);
int psim_get_metadata(const PsimStepper* s, FrameMetadata* out);

/* = Kernel::run_async (kernel.cuh:139-151) for DataStructure::MatrixBuckets: enqueues one frame,
 * i.e. metadata.steps_per_frame leapfrog steps with re-binning according to the schedule, then packs
 * a snapshot of the result for psim_download_frame. Returns immediately. */
int psim_run_frame_async(PsimStepper* s);

/* Finer-grained control (used by tests and by the driver when it needs it):
 * one fused force + kick + drift step (= bucket_step, kernel_bucket.cuh:40-94) ... */
int psim_step_async(PsimStepper* s, uint32_t steps);
/* ... one re-binning pass (= bucket_move, kernel_bucket.cuh:5-39, without its 16-per-cell and
 * one-cell-per-move losses) ... */
int psim_rebin_async(PsimStepper* s);
/* ... and packing the current state into the snapshot buffer. */
int psim_snapshot_async(PsimStepper* s);

/* = Kernel::sync (kernel.cuh:88-94). */
int psim_sync(PsimStepper* s);

/* = Kernel::read (kernel.cuh:117-129) + frame_compact (frontend.hpp:50-56): copies the last packed
 * snapshot to `dst` as a compact frame (header + live particles in cell-major order, the order the
 * reference's compacted slot array has). dst->particle_count must hold dst's capacity on entry.
 * Waits only for the snapshot, not for frames enqueued after it. */
int psim_download_frame(PsimStepper* s, FrameHeader* dst);

/* Introspection for parity checks and benchmarks. */
uint32_t psim_particle_count(const PsimStepper* s);          /* live particles on the device      */
uint64_t psim_steps_executed(const PsimStepper* s);          /* leapfrog steps enqueued so far    */
uint64_t psim_rebins_executed(const PsimStepper* s);         /* re-binning passes enqueued so far */
uint64_t psim_kernel_launches(const PsimStepper* s);         /* kernels of this library launched  */
uint32_t psim_cell_count(const PsimStepper* s);              /* BX * BY                           */
/* cell_start[0..cells] of the current binning (cells+1 entries, exclusive prefix sum of the per-cell
 * particle counts; synchronises). */
int psim_get_cell_start(PsimStepper* s, uint32_t* out);
/* Last step-kernel duration statistics gathered with CUDA events around every step launch when
 * enabled (adds two event records per step). */
int psim_enable_step_timing(PsimStepper* s, int enable);
int psim_get_step_timing(PsimStepper* s, double* total_ms, uint64_t* launches);

/* Device pointers of the live state (cell-sorted structure of arrays), for zero-copy consumers.
 * pos: uint2[n] fixed-point (x, y); vel: float2[n]; ty: int32[n]; cell_start: uint32[cells+1]. */
int psim_device_state(PsimStepper* s, const void** pos, const void** vel, const void** ty,
                      const void** cell_start);

#ifdef __cplusplus
} /* extern "C" */
#endif
