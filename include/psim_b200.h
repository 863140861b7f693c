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
#define PSIM_ENCCL (-5)    /* NCCL could not be loaded or an NCCL call failed    */
#define PSIM_EMIGRATION (-6) /* a particle moved past the adjacent slab between two re-bins, or more
                                particles changed slab in one re-bin than migrant_capacity */

/* Step / re-bin schedules. */
#define PSIM_SCHEDULE_REFERENCE 0 /* `S M S*17 M S*17 ...`, countdown restarts every frame, may run
                                     steps_per_frame+1 steps: reference kernel_bucket.cuh:181-206 */
#define PSIM_SCHEDULE_NATIVE 1    /* exactly steps_per_frame steps, re-bin every `rebin_every` steps,
                                     the counter carries over between frames                     */

typedef struct PsimStepper PsimStepper; /* opaque */
typedef struct PsimGroup PsimGroup;     /* opaque: all slabs of a decomposition inside one process */

typedef struct PsimConfig {
    uint32_t grid_x_log2;   /* cells in x = 1 << grid_x_log2 (reference: BUCKETS_X_LOG2 = 6, kernel.cuh:15) */
    uint32_t grid_y_log2;   /* cells in y = 1 << grid_y_log2 (reference: BUCKETS_Y_LOG2 = 6, kernel.cuh:16) */
    uint32_t max_particles; /* capacity of the device buffers (live particles)                            */
    uint32_t schedule;      /* PSIM_SCHEDULE_*                                                             */
    uint32_t rebin_every;   /* native schedule only; 0 => 17 (the reference's effective cadence)           */
    int32_t device;         /* CUDA device ordinal; -1 => current device                                   */
    uint32_t use_graph;     /* 1: frames on grids below 1024 cells per axis (the reference's own scenes: a frame is
                               ~150 launches of a few microseconds, launch-bound) are captured once as a CUDA graph
                               and replayed with one cudaGraphLaunch; reference schedule, single slab            */
    /* Slab decomposition over cell rows (the reference is single-device). The grid
     * above is the GLOBAL grid; this stepper owns rows [slab_rank, slab_rank + 1) * (cells in y / slab_count)
     * and keeps one ghost row of each adjacent slab. max_particles is the capacity of the slab. */
    uint32_t slab_rank;        /* 0 .. slab_count-1                                                        */
    uint32_t slab_count;       /* 0 or 1 => the whole grid                                                 */
    uint32_t ghost_capacity;   /* particles one ghost row can hold; 0 => 4x the mean row of a full slab as thin as the
                                  thinnest of this slab and its neighbours, >= 4096                        */
    uint32_t migrant_capacity; /* particles that can move to ONE neighbour slab in one re-bin, the same on every
                                  slab; 0 => 4x the mean row of a full slab of average height, >= 4096     */
    uint32_t ingest_capacity;  /* records an uploaded frame can hold (a slab is usually handed the whole
                                  scene and keeps its own rows); 0 => max_particles                        */
    uint32_t snapshot_buffers; /* 0 or 1: one snapshot buffer; 2: two that alternate, so that the snapshot of the
                                  previous frame can be downloaded (psim_download_frame_ex, age 1) while the next
                                  frame, its snapshot included, is already enqueued -- the double buffering of the
                                  reference's main loop (cuda_simulator.cu:28-37)                           */
    uint32_t slab_bounds[4];   /* all 0: slabs own equal shares of the cell rows. Otherwise global cell rows
                                  {first row of the slab below, first own row, end of the own rows = first row of
                                  the slab above, end of the slab above}: entries k-1 .. k+2 of one array of
                                  slab_count + 1 boundaries every slab takes its four from (clamped at the ends:
                                  slab 0 has {0, 0, ..}, the last slab {.., rows, rows}), e.g. from
                                  psim_balance_rows. Every slab owns at least 2 rows; adjacent slabs must meet
                                  (checked by psim_group_create / psim_comm_init).                          */
    uint32_t species_physics;  /* 0: every particle is stepped with metadata.particles[0], as the reference does
                                  (kernel_bucket.cuh:52); `ty` is a label. 1 (an extension): a
                                  pair uses the Mie parameters of its species pair -- particles[s] for two particles
                                  of species s = min(ty, 1), the Lorentz-Berthelot mix (mean sigma, geometric-mean
                                  epsilon, mean exponents) for an unlike pair -- and the wall term the particle's own.
                                  Takes effect when the two entries differ; MatrixBuckets scenes on a single slab.  */
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
 * The frame's metadata becomes the stepper's metadata. Synchronous w.r.t. the host.
 * A frame whose metadata says DataStructure::CompactArray is ingested as the reference ingests it
 * (frame_compact_into, kernel.cuh:207-209: input order, no binning) and stepped all-pairs from then on. */
int psim_upload_frame(PsimStepper* s, const FrameHeader* frame);

/* Same, for particle records that already live in device memory (AoS, 20 B each). */
int psim_upload_device(PsimStepper* s, const FrameMetadata* meta, const void* d_particles, uint32_t count);

/* = Kernel::write_metadata (kernel.cuh:96-101): takes effect for the next frame that is enqueued.
 * One deviation: the reference picks its layout from every frame's metadata (kernel.cuh:143-150); here the layout is
 * the one the scene was uploaded in, so an update whose data_structure differs from it returns PSIM_EINVAL and changes
 * nothing (upload the scene again to switch between MatrixBuckets and CompactArray). */
int psim_set_metadata(PsimStepper* s, const FrameMetadata* meta);
int psim_get_metadata(const PsimStepper* s, FrameMetadata* out);

/* = Kernel::run_async (kernel.cuh:139-151) for DataStructure::MatrixBuckets: enqueues one frame,
 * i.e. metadata.steps_per_frame leapfrog steps with re-binning according to the schedule, then packs
 * a snapshot of the result for psim_download_frame. On the reference's own grid (and any grid below 1024 cells per
 * axis) it returns immediately; on finer grids and with slabs it returns once the frame's last re-bin has been
 * enqueued (each re-bin reads a few counts back: the tile count sizes the next launches), i.e. with the last <= 16
 * steps still running. For DataStructure::CompactArray scenes: kernel_compact.cuh:78-92, exactly steps_per_frame
 * all-pairs steps. */
int psim_run_frame_async(PsimStepper* s);

/* Finer-grained control (used by tests and by the driver when it needs it):
 * one fused force + kick + drift step (= bucket_step, kernel_bucket.cuh:40-94) ... */
int psim_step_async(PsimStepper* s, uint32_t steps);
/* ... one re-binning pass (= bucket_move, kernel_bucket.cuh:5-39, without its 16-per-cell and
 * one-cell-per-move losses) ... */
int psim_rebin_async(PsimStepper* s);
/* ... and packing the current state into the snapshot buffer. */
int psim_snapshot_async(PsimStepper* s);
/* Decimated snapshots: from now on a snapshot holds every stride-th particle of the cell-sorted state (a spatially
 * uniform sample for display; at 10 M particles a full frame is 200 MB per snapshot, which is what the reference's
 * Kernel::read (kernel.cuh:117-129) would move). 1 = every particle (default). The state itself is untouched. */
int psim_set_snapshot_stride(PsimStepper* s, uint32_t stride);

/* = Kernel::sync (kernel.cuh:88-94). */
int psim_sync(PsimStepper* s);

/* = Kernel::read (kernel.cuh:117-129) + frame_compact (frontend.hpp:50-56): copies the last packed
 * snapshot to `dst` as a compact frame (header + live particles in cell-major order, the order the
 * reference's compacted slot array has). dst->particle_count must hold dst's capacity on entry.
 * Waits only for the snapshot, not for frames enqueued after it. */
int psim_download_frame(PsimStepper* s, FrameHeader* dst);
/* The snapshot packed `age` snapshots ago: 0 = the latest (= psim_download_frame), 1 = the one before it
 * (needs PsimConfig.snapshot_buffers = 2). */
int psim_download_frame_ex(PsimStepper* s, uint32_t age, FrameHeader* dst);

/* Pipelined frames (streaming scenes at full PCIe duplex): the copies run on their own streams while frames compute.
 *   psim_stage_frame_async   starts the host-to-device copy of `frame` into a second ingest buffer and returns; the
 *                            frame's memory (page-locked for a real overlap) must stay untouched until
 *                            psim_upload_staged has returned;
 *   psim_upload_staged       = psim_upload_frame of the staged frame, minus the copy: bins it as soon as it has
 *                            arrived and the work enqueued so far has finished;
 *   psim_download_frame_begin starts the device-to-host copy of a snapshot (age as psim_download_frame_ex) and returns;
 *                            the header of `dst` is valid at once, the records after psim_download_frame_end.
 * With PsimConfig.snapshot_buffers = 2 the loop  upload_staged(k); stage(k+1); run_frame_async(k); download_end(k-1);
 * download_begin(k)  moves frame k+1 in and frame k-1 out while frame k is stepped (bench.py's e2e leg). */
int psim_stage_frame_async(PsimStepper* s, const FrameHeader* frame);
int psim_upload_staged(PsimStepper* s);
int psim_download_frame_begin(PsimStepper* s, uint32_t age, FrameHeader* dst);
int psim_download_frame_end(PsimStepper* s);

/* Page-locked host memory for frames (uploads and downloads from pageable memory are staged by the driver
 * and run at a fraction of the link rate). NULL on failure. */
void* psim_host_alloc(size_t bytes);
void psim_host_free(void* p);

/* Introspection for parity checks and benchmarks. */
uint32_t psim_particle_count(const PsimStepper* s);          /* live particles on the device      */
uint64_t psim_steps_executed(const PsimStepper* s);          /* leapfrog steps enqueued so far    */
uint64_t psim_rebins_executed(const PsimStepper* s);         /* re-binning passes enqueued so far */
uint64_t psim_kernel_launches(const PsimStepper* s);         /* kernels of this library launched  */
uint64_t psim_migrants_sent(const PsimStepper* s);           /* particles handed to neighbour slabs by re-bins */
uint32_t psim_cell_count(const PsimStepper* s);              /* BX * BY                           */
/* cell_start[0..cells] of the current binning (cells+1 entries, exclusive prefix sum of the per-cell
 * particle counts; synchronises). */
int psim_get_cell_start(PsimStepper* s, uint32_t* out);
/* Last step-kernel duration statistics gathered with CUDA events around every step launch when
 * enabled (adds two event records per step). */
int psim_enable_step_timing(PsimStepper* s, int enable);
int psim_get_step_timing(PsimStepper* s, double* total_ms, uint64_t* launches);

/* How the step kernel's tiles look after the last binning: which kernel runs (float_path = 1: step_kernel_c,
 * two cell-mates per thread on exact fp32 offsets, grids of >= 1024 cells per axis; 0: step_kernel, one particle
 * per thread on integer separations) and how many tiles stage their stencil in shared memory (the others read
 * global memory: very sparse or very crowded spots). Synchronises. */
typedef struct PsimTileStats {
    uint32_t float_path;
    uint32_t tiles, tiles_staged;
    uint32_t max_columns, max_row_particles;
    uint32_t _reserved;
    uint64_t threads_live, threads_launched;
} PsimTileStats;
int psim_tile_stats(PsimStepper* s, PsimTileStats* out);

/* Row boundaries that give every slab about the same number of live particles of `scene` (SURVEY.md section 8e:
 * a clustered scene cut into equal numbers of rows is badly balanced). Host code, no GPU involved: a histogram of
 * cell rows (row = y >> (32 - grid_y_log2), kernel.cuh:225), cut at the multiples of count / slab_count, every slab
 * at least 2 rows. bounds[0] = 0 <= ... <= bounds[slab_count] = 1 << grid_y_log2; slab k's PsimConfig.slab_bounds are
 * bounds[k-1 .. k+2] clamped to the array (psim_slab_bounds_of). Re-balancing a running decomposition = download,
 * psim_balance_rows on the snapshot, re-create the steppers, upload. */
int psim_balance_rows(const FrameHeader* scene, uint32_t grid_y_log2, uint32_t slab_count, uint32_t* bounds);
/* The same cut from a histogram: row_counts[r] = live particles in global cell row r (1 << grid_y_log2 entries). With one
 * process per slab every rank histograms its own particles, the histograms are summed over the ranks (one all-reduce)
 * and every rank cuts the same boundaries (particle_simulator_b200/slabs.py: rebalance_across_ranks). */
int psim_balance_rows_hist(const uint64_t* row_counts, uint32_t grid_y_log2, uint32_t slab_count, uint32_t* bounds);
void psim_slab_bounds_of(const uint32_t* bounds, uint32_t slab_rank, uint32_t slab_count, uint32_t out[4]);

/* Where this stepper's slab sits and what it currently holds. */
typedef struct PsimSlabInfo {
    uint32_t slab_rank, slab_count;
    uint32_t first_row, rows;             /* owned cell rows (global numbering)                       */
    uint32_t first_local_row, local_rows; /* rows of psim_get_cell_start: owned + ghost rows          */
    uint32_t particles;                   /* owned                                                    */
    uint32_t ghost_below, ghost_above;    /* particles currently in the two ghost rows                */
    uint32_t ghost_capacity, migrant_capacity;
    uint32_t _reserved[5];
} PsimSlabInfo;
int psim_slab_info(const PsimStepper* s, PsimSlabInfo* out);

/* Device pointers of the live state (cell-sorted structure of arrays), for zero-copy consumers.
 * pos: uint2[n] fixed-point (x, y); vel: float2[n]; ty: int32[n]; cell_start: uint32[cells+1]. */
int psim_device_state(PsimStepper* s, const void** pos, const void** vel, const void** ty,
                      const void** cell_start);

/* ---- slab decomposition, one process per slab (one GPU each): NCCL over NVLink -----------------
 * Every rank creates its stepper with slab_rank = its rank and slab_count = world size, then joins the
 * communicator. After that the ordinary calls above are COLLECTIVE: every rank makes the same sequence
 * of psim_upload_frame / psim_step_async / psim_rebin_async / psim_run_frame_async calls. Per step each
 * slab sends the positions of its two boundary cell rows to the adjacent slabs (halo exchange); at a
 * re-bin, particles whose row left the slab migrate to the adjacent slab. Results are bit-identical to
 * a single-slab run of the same scene. psim_upload_frame may be handed the whole scene (records outside
 * the slab's rows are skipped); psim_download_frame returns the slab's own particles, and the slabs'
 * frames concatenated in rank order are the single-slab frame.
 * NCCL (libnccl.so.2, or $PSIM_NCCL_LIB) is loaded on first use; a single-slab user never needs it. */
int psim_comm_unique_id(void* out128);                         /* rank 0: ncclGetUniqueId, 128 bytes    */
/* All ranks: ncclCommInitRank, then the slabs exchange CUDA IPC handles of their position buffers. When every
 * rank could map its neighbours (same node, peer access over NVLink), the per-step halo exchange needs no
 * collective call at all: the step kernel stores the new positions of its boundary rows straight into the
 * neighbours' ghost rows and publishes an epoch word; only the neighbours' boundary tiles wait for it.
 * Otherwise (or with PSIM_HALO=nccl) every step is followed by an ncclSend/ncclRecv pair per neighbour.
 * Re-binning (migrants, ghost-row cell counts, fresh ghost rows) always uses ncclSend/ncclRecv. */
int psim_comm_init(PsimStepper* s, const void* unique_id128);
/* 0: single slab, 1: halo by send/recv after every step, 2: halo pushed by the step kernel over peer memory. */
int psim_halo_mode(const PsimStepper* s);

/* ---- slab decomposition inside one process on one device ----------------------------------------
 * The same slabs, exchanges done by device-to-device copies on one stream. It validates the
 * decomposition (bit-identical to the single-slab run) on a single GPU. `steppers[r]` must have been
 * created with slab_rank = r, slab_count = count and equal grids, schedules and capacities; while they
 * belong to a group they are driven through the group calls only (download / introspection excepted). */
int psim_group_create(PsimStepper* const* steppers, uint32_t count, PsimGroup** out);
void psim_group_destroy(PsimGroup* g); /* does not destroy the steppers */
const char* psim_group_last_error(const PsimGroup* g);
int psim_group_upload_frame(PsimGroup* g, const FrameHeader* frame);
int psim_group_set_metadata(PsimGroup* g, const FrameMetadata* meta);
int psim_group_run_frame_async(PsimGroup* g);
int psim_group_step_async(PsimGroup* g, uint32_t steps);
int psim_group_rebin_async(PsimGroup* g);
int psim_group_snapshot_async(PsimGroup* g);
int psim_group_sync(PsimGroup* g);
uint32_t psim_group_particle_count(const PsimGroup* g);
/* All slabs' snapshots concatenated in rank order: the frame a single-slab stepper would return. */
int psim_group_download_frame(PsimGroup* g, FrameHeader* dst);

#ifdef __cplusplus
} /* extern "C" */
#endif
