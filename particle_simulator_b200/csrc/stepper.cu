// stepper.cu -- the B200 (sm_100a) particle stepper behind include/psim_b200.h.
//
// What it replaces in the reference (paths under /root/reference/cuda_simulator/src):
//   kernel_prepare_frame  kernel.cuh:200-250      -> psim_upload_frame  (GPU stable counting sort)
//   bucket_move_kernel    kernel_bucket.cuh:5-39  -> rebin()            (same sort, on the live state)
//   bucket_step_kernel    kernel_bucket.cuh:40-94 -> step_kernel        (fused force + kick + drift)
//   bucket_kernel_run_async kernel_bucket.cuh:181-206 -> run_frame()    (same step / re-bin schedule)
//   Kernel::{write,read,sync,write_metadata} kernel.cuh:88-129 -> upload / download / sync / set_metadata
//
// Data layout in HBM (all cell-sorted, structure of arrays, no per-cell capacity, no null slots):
//   pos[2][n]  uint2  fixed-point (x, y), ping-pong: a step reads pos[cur] and writes pos[cur^1]
//   vel[n]     float2 half-step velocities, updated in place by a step
//   ty[n]      int32  species label, untouched by a step
//   cell_start[cells+1] uint32 exclusive prefix sum of per-cell counts (CSR); cell = cx + cy*BX
//   cell_id[n] uint32 cell of every particle as of the last binning (membership, not position)
//   tiles[ceil(n/128)] TileDesc: what the CTA of each tile of 128 consecutive particles stages
// Membership is by the LAST binning, exactly like the reference's slot array: between re-bins a
// particle keeps its index and its cell even if it has drifted out of it (kernel_bucket.cuh:71-91).
//
// One translation unit: the device code lives in device_common.cuh (parameter blocks, pair arithmetic), step_int.cuh
// (halo protocol, epilogue, step_kernel, all-pairs), step_float.cuh (step_kernel_c, couples and tiles) and binning.cuh
// (counting sort, neighbour records, snapshots, migration); this file holds the host side and the C API.
// There is no CPU path in this file and nothing here includes or links oracle/.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "psim_b200.h"

namespace {

#include "device_common.cuh"
#include "step_int.cuh"
#include "step_float.cuh"
#include "binning.cuh"

}  // namespace

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------

// NCCL is bound at run time (dlopen) and only when psim_comm_init is called: a single-GPU user of this
// library needs no NCCL. Minimal declarations of the stable NCCL 2.x C API.
typedef struct ncclComm* ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
constexpr int kNcclUint8 = 1;   // ncclDataType_t::ncclUint8
constexpr int kNcclInt32 = 2;   // ncclDataType_t::ncclInt32
constexpr int kNcclMin = 3;     // ncclRedOp_t::ncclMin

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*CommAbort)(ncclComm_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

// What one rank sends to / receives from its lower ([0]) and upper ([1]) neighbour slab.
struct XferOp {
    const void* send[2] = {nullptr, nullptr};
    void* recv[2] = {nullptr, nullptr};
    size_t send_bytes[2] = {0, 0};
    size_t recv_bytes[2] = {0, 0};
};

struct PsimGroup;

// A host <-> device copy of a pipelined frame (psim_stage_frame_async, psim_download_frame_begin) issued in pieces. A slab's
// frame makes a host round trip at every re-bin (counts over PCIe into mapped memory, a stream synchronisation, the next
// launches); with all the ranks of a node copying 200 MB frames at once those round trips queue behind the copies and a
// 30 ms frame takes 45 ms (170 ms under back-to-back copies: tools/diag_copy_load.py). So a slab issues one piece right
// after each re-bin's round trip -- it is done long before the next one, 17 steps later -- and whatever is left when the
// copy is needed. A single slab's frame makes no round trip: its copies go out whole.
struct PacedCopy {
    char* dst = nullptr;
    const char* src = nullptr;
    size_t total = 0, issued = 0, piece = 0;  // piece == 0: all at once
    cudaMemcpyKind kind = cudaMemcpyDefault;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;  // recorded behind the last piece
    bool active = false;
};

struct FrameGraph {  // one captured frame (run_frame_with_graph)
    cudaGraphExec_t exec = nullptr;
    uint64_t steps = 0, rebins = 0, launches = 0;
    int end_pos = 0, end_vel = 0, end_ty = 0;
    uint32_t n = 0, n_total = 0, tiles_launch = 0;  // what the captured launches were sized for
};

struct PsimStepper {
    PsimConfig cfg{};
    Grid grid{};
    Phys phys{};
    int kernel_kn = 0;  // step-kernel variant: integer part of n/2+1 (0: run-time exponents)
    int kernel_frac = kFracEx2;
    bool kernel_aniso = false;
    FrameMetadata meta{};
    int device = 0;

    // slab decomposition (nranks == 1: the whole grid, no ghost rows, no exchange)
    int rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;   // one process per slab (psim_comm_init)
    PsimGroup* group = nullptr;  // all slabs in this process (psim_group_create)

    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;       // where the step loop runs (own_stream, the caller's, or the group's)
    cudaStream_t copy_stream = nullptr;  // snapshot download
    // Snapshots: one buffer, or two that alternate (PsimConfig.snapshot_buffers = 2) so that the previous
    // frame's snapshot can be copied out while the next frame, pack included, is already enqueued.
    int nsnap = 1, snap_latest = 0;
    uint32_t snapshot_stride = 1;  // snapshots hold every stride-th particle
    uint64_t snaps_taken = 0;
    cudaEvent_t snapshot_ready[2] = {nullptr, nullptr};
    cudaEvent_t snapshot_consumed[2] = {nullptr, nullptr};

    // capacities
    uint32_t ghost_cap = 0;     // particles per ghost row
    uint32_t cap_total = 0;     // max_particles + 2 * ghost_cap
    uint32_t box_capacity = 0;  // migrants per direction per re-bin
    size_t box_bytes = 0;

    // live state (see the layout table at the top of this file)
    uint2* pos[2] = {nullptr, nullptr};
    float2* vel[2] = {nullptr, nullptr};
    int32_t* ty[2] = {nullptr, nullptr};
    int cur_pos = 0, cur_vel = 0, cur_ty = 0;
    float4* nbr[2] = {nullptr, nullptr};  // neighbour records (fine grids), same ping-pong index as pos
    bool nbr_stale = true;                // nbr[cur_pos] does not match pos[cur_pos] / the current scale
    uint32_t* cell_id = nullptr;
    TileDesc* tiles = nullptr;
    bool tiles_stale = false;  // step_kernel's descriptors were not rebuilt by the last binning (step_kernel_c ran)
    uint32_t* cell_start = nullptr;  // cells + 1 (+ padding)
    uint32_t* pad_start = nullptr;   // cells + 1: prefix sum of the cell counts rounded up to even (step_float.cuh)
    uint2* couple_i0 = nullptr;      // per couple: (first particle | has-a-second << 31, cell)
    uint32_t* tile_base = nullptr;   // own_rows + 1: first tile of every owned row
    uint32_t* d_couple_tiles = nullptr;
    TileC* tiles_c = nullptr;
    uint32_t tiles_c_cap = 0, n_tiles_c = 0;  // n_tiles_c: the count the host last read (exact with slabs: read at every binning)
    uint32_t tiles_c_launch = 0;              // CTAs a step launches: n_tiles_c plus a margin on a single slab (step_float.cuh)
    bool float_grid = false;         // the grid is fine enough for step_kernel_c's exact fp32 offsets
    bool float_path = false;         // ... and the metadata's physics has a step_kernel_c variant
    PhysF physf{};
    bool force_int_path = false;     // PSIM_FORCE_INT_PATH=1: step_kernel on every grid (A/B measurements, tests)
    bool pdl = true;                 // step launches allow programmatic dependent launch (PSIM_PDL=0 turns it off)
    bool species_mode = false;       // PsimConfig.species_physics and the metadata's two species differ: step_kernel_species
    SpeciesArgs species{};
    uint32_t* cell_count = nullptr;  // cells
    uint2* block_sum = nullptr;      // per scan block: (particles, particles with every cell rounded up to even)
    uint32_t* rank_in_cell = nullptr;
    uint32_t* perm = nullptr;
    Particle* staging = nullptr;   // ingest buffer (wire-format records)
    // pipelined ingest (psim_stage_frame_async / psim_upload_staged): a second ingest buffer filled over its own
    // stream while the running frame computes
    Particle* staging_async = nullptr;
    cudaStream_t h2d_stream = nullptr;
    cudaEvent_t staged_ready = nullptr;
    FrameMetadata staged_meta{};
    uint32_t staged_count = 0;
    bool has_staged = false;
    // pipelined download (psim_download_frame_begin / _end)
    FrameHeader* pending_dst = nullptr;
    int pending_k = -1;
    PacedCopy paced_up, paced_down;  // the staged upload and the download in flight
    size_t copy_piece_bytes = 0;     // 0: copies go out whole (single slab); slabs: a sixth of the copy, PSIM_COPY_PIECE_MB
    bool copy_pace_steps = false;    // slabs: a small piece behind every step (pump_copies_behind_step); PSIM_COPY_PACE=rebin: the sixths
    cudaEvent_t pace_event = nullptr;
    Particle* snapshot[2] = {nullptr, nullptr};  // packed snapshots of the owned particles (wire-format records)
    uint32_t ingest_cap = 0;       // records the ingest buffer holds
    unsigned char* outbox[2] = {nullptr, nullptr};
    unsigned char* inbox[2] = {nullptr, nullptr};
    uint32_t* mig_counters = nullptr;  // 2
    uint32_t* mig_blk_cnt = nullptr;   // 2 x blocks of 1024 owned particles: migrants per block and direction (zero between re-bins)
    uint32_t* mig_blk_off = nullptr;   // their exclusive scans
    uint32_t* mig_blk_list = nullptr;  // [0] how many blocks have migrants, then which
    // halo push (HaloArgs): this slab's header, the neighbours' buffers and headers as this device sees them
    HaloHeader* hdr = nullptr;
    uint2* peer_pos[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [side][buffer]
    float4* peer_nbr[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    HaloHeader* peer_hdr[2] = {nullptr, nullptr};
    void* ipc_mapped[2][5] = {};  // to close at destroy: pos[0], pos[1], header, nbr[0], nbr[1]
    bool push = false;           // steps push their boundary rows (else: exchange after every step)
    bool ghosts_by_push = false; // the ghost rows the next step reads are being written by the neighbours' last step
    uint32_t halo_epoch = 0;     // epoch the last step published
    uint32_t tiles_lo = 0, tile_hi0 = 0;  // couple tiles of the first owned row: [0, tiles_lo); of the last: [tile_hi0, ..)
    uint32_t* d_flags = nullptr;   // 1
    uint32_t* d_counts = nullptr;  // 16
    uint32_t* h_counts = nullptr;  // pinned and mapped, 16: [0..11] written by slab_counts_kernel, [12..13] by row_tiles_kernel, [15] check_halo_error
    uint32_t* h_counts_dev = nullptr;  // the device's address of h_counts

    uint32_t n = 0;        // particles this stepper owns
    uint32_t n_total = 0;  // with ghost rows
    uint32_t own_lo = 0, own_hi = 0, b_lo_end = 0, b_hi_start = 0;
    Source src{};          // candidates of the binning in flight
    uint64_t migrants_sent = 0;  // particles handed to a neighbour slab by the re-bins so far

    uint32_t snapshot_n[2] = {0, 0};
    FrameMetadata snapshot_meta[2]{};
    bool compact_mode = false;  // the scene was uploaded with DataStructure::CompactArray: all-pairs steps, no grid
    bool has_scene = false;
    bool has_snapshot = false;
    bool fresh_scene = false;  // nothing has been stepped since the ingest: the binning is current
    int native_countdown = 0;  // native schedule: steps until the next re-bin

    uint64_t steps_executed = 0, rebins_executed = 0, launches = 0;
    std::map<uint32_t, FrameGraph> frame_graphs;  // PsimConfig.use_graph: captured frames by starting buffer parity

    bool commit_pending_pump = false;
    bool timing = false;
    cudaEvent_t timing_after_main = nullptr;  // to be recorded right behind the step kernel of the launch in flight
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_events;
    size_t timing_used = 0;
    double timing_total_ms = 0;
    uint64_t timing_launches = 0;

    std::string error;
    bool failed = false;  // a call failed while a communicator exists: peers may be waiting for this slab
};

struct PsimGroup {
    std::vector<PsimStepper*> ranks;
    cudaStream_t stream = nullptr;
    int native_countdown = 0;
    std::string error;
};

namespace {

std::string g_create_error;
NcclApi g_nccl;

int fail(PsimStepper* s, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (s) {
        s->error = buf;
        s->failed = true;
    } else {
        g_create_error = buf;
    }
    return code;
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t err__ = (call);                                                                       \
        if (err__ != cudaSuccess)                                                                         \
            return fail(s, PSIM_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, \
                        __LINE__);                                                                        \
    } while (0)

#define CKN(call)                                                                                      \
    do {                                                                                               \
        int err__ = (call);                                                                            \
        if (err__ != 0)                                                                                \
            return fail(s, PSIM_ENCCL, "%s failed: %s (%s:%d)", #call,                                 \
                        g_nccl.GetErrorString ? g_nccl.GetErrorString(err__) : "?", __FILE__, __LINE__); \
    } while (0)

inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// Issue the next piece of a paced copy (`all`: everything that is left); the last piece is followed by its event.
int pump_copy(PsimStepper* s, PacedCopy& c, bool all) {
    if (!c.active) return PSIM_OK;
    while (c.issued < c.total) {
        const size_t n = (all || c.piece == 0) ? c.total - c.issued : std::min(c.piece, c.total - c.issued);
        CK(cudaMemcpyAsync(c.dst + c.issued, c.src + c.issued, n, c.kind, c.stream));
        c.issued += n;
        if (!all && c.piece) break;
    }
    if (c.issued == c.total) {
        if (c.done) CK(cudaEventRecord(c.done, c.stream));
        c.active = false;
    }
    return PSIM_OK;
}

void start_copy(PsimStepper* s, PacedCopy& c, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind,
                cudaStream_t stream, cudaEvent_t done) {
    c.dst = static_cast<char*>(dst);
    c.src = static_cast<const char*>(src);
    c.total = bytes;
    c.issued = 0;
    c.kind = kind;
    c.stream = stream;
    c.done = done;
    c.active = true;
    // a slab's frame on the reference schedule has six re-bins: a sixth of the copy behind each
    c.piece = s->copy_piece_bytes == 1 ? std::max<size_t>((bytes + 5) / 6, (size_t)4 << 20) : s->copy_piece_bytes;
    // paced by the steps instead: a piece behind every step of the frame that runs meanwhile, done a few steps before its
    // end (a frame of the benchmark has 101 steps: 96 pieces)
    if (s->copy_pace_steps && s->copy_piece_bytes <= 1) {  // PSIM_COPY_PIECE_MB keeps its say
        const size_t steps = s->meta.steps_per_frame;
        const size_t parts = std::min<size_t>(96, steps > 5 ? steps - 5 : 1);
        c.piece = std::max<size_t>(((bytes + parts - 1) / parts + 255) & ~(size_t)255, (size_t)256 << 10);
    }
}

// Slabs: the copies in flight advance by one small piece per step, each piece ordered behind that step on the device (an
// event), so that the traffic is spread evenly over the frame. The first version sent a sixth behind every re-bin
// (PSIM_COPY_PACE=rebin): six bursts that all slabs fire at the same moment, since they re-bin in lockstep -- 35.6 ms per
// end-to-end step on 4 GPUs against 33.1 ms this way (frame alone: 30.2 ms).
int pump_copies_behind_step(PsimStepper* s) {
    if (!s->copy_pace_steps || !(s->paced_up.active || s->paced_down.active)) return PSIM_OK;
    if (!s->pace_event) CK(cudaEventCreateWithFlags(&s->pace_event, cudaEventDisableTiming));
    CK(cudaEventRecord(s->pace_event, s->stream));
    int rc;
    if (s->paced_up.active) {
        CK(cudaStreamWaitEvent(s->paced_up.stream, s->pace_event, 0));
        if ((rc = pump_copy(s, s->paced_up, false))) return rc;
    }
    if (s->paced_down.active) {
        CK(cudaStreamWaitEvent(s->paced_down.stream, s->pace_event, 0));
        if ((rc = pump_copy(s, s->paced_down, false))) return rc;
    }
    return PSIM_OK;
}


// Cubic for 2^z on z in [z_lo, z_hi] (weighted least squares on Chebyshev nodes; the constant term is
// pinned to 1 so that z = 0 is exact). Returns the largest error of (n/m) q^(kn+fn) it causes, i.e. the
// absolute error of g, over the q range that maps onto [z_lo, z_hi].
double fit_exp2_cubic(double fn, double kn, double nm, double q_lo, double q_hi, float c[3]) {
    const int N = 256;
    const double ln2 = 0.69314718055994530942;
    // normal equations for e(z) - 1 = z (c1 + c2 z + c3 z^2), weight w = how much an error costs in g
    double A[3][3] = {{0}}, B[3] = {0};
    auto weight = [&](double q) { return std::min(nm * std::pow(q, kn + fn), 0.15) + 1e-6; };
    for (int k = 0; k < N; ++k) {
        double t = std::cos(3.14159265358979323846 * (k + 0.5) / N);
        double lq = 0.5 * (std::log2(q_hi) + std::log2(q_lo)) + 0.5 * (std::log2(q_hi) - std::log2(q_lo)) * t;
        double q = std::exp2(lq), z = fn * lq, w = weight(q);
        double basis[3] = {z, z * z, z * z * z}, rhs = std::exp(z * ln2) - 1.0;
        for (int r = 0; r < 3; ++r) {
            for (int cc = 0; cc < 3; ++cc) A[r][cc] += w * w * basis[r] * basis[cc];
            B[r] += w * w * basis[r] * rhs;
        }
    }
    // 3x3 solve (Gaussian elimination with partial pivoting)
    int idx[3] = {0, 1, 2};
    double M[3][4];
    for (int r = 0; r < 3; ++r) {
        for (int cc = 0; cc < 3; ++cc) M[r][cc] = A[r][cc];
        M[r][3] = B[r];
    }
    for (int col = 0; col < 3; ++col) {
        int piv = col;
        for (int r = col + 1; r < 3; ++r)
            if (std::fabs(M[r][col]) > std::fabs(M[piv][col])) piv = r;
        std::swap(idx[col], idx[piv]);
        for (int cc = 0; cc < 4; ++cc) std::swap(M[col][cc], M[piv][cc]);
        if (M[col][col] == 0.0) {  // fn == 0 or a degenerate range: Taylor coefficients
            c[0] = (float)ln2;
            c[1] = (float)(ln2 * ln2 / 2);
            c[2] = (float)(ln2 * ln2 * ln2 / 6);
            return fn == 0.0 ? 0.0 : 1.0;
        }
        for (int r = 0; r < 3; ++r) {
            if (r == col) continue;
            double f = M[r][col] / M[col][col];
            for (int cc = col; cc < 4; ++cc) M[r][cc] -= f * M[col][cc];
        }
    }
    for (int r = 0; r < 3; ++r) c[r] = (float)(M[r][3] / M[r][r]);
    double worst = 0;
    for (int k = 0; k <= 4 * N; ++k) {
        double lq = std::log2(q_lo) + (std::log2(q_hi) - std::log2(q_lo)) * k / (4.0 * N);
        double q = std::exp2(lq), z = fn * lq;
        double approx = 1.0 + z * ((double)c[0] + z * ((double)c[1] + z * (double)c[2]));
        double err = std::fabs(approx - std::exp(z * ln2));
        // absolute error of g, and -- where the repulsive term dominates -- relative to the term itself
        double cost = std::min(nm * std::pow(q, kn + fn) * err, err / 2e-7 * 3e-8);
        worst = std::max(worst, cost);
    }
    return worst;
}

Phys make_phys(const FrameMetadata& m, int* kernel_kn, int* kernel_frac, bool* aniso) {
    const MiePotentialParams& p = m.particles[0];  // kernel_bucket.cuh:52
    Phys ph{};
    const float two32 = 4294967296.f;
    ph.kx = m.box_width / two32;
    ph.ky = m.box_height / two32;
    ph.yscale = ph.ky / ph.kx;
    *aniso = ph.yscale != 1.f;
    ph.inv_c2 = (ph.kx / p.sigma) * (ph.kx / p.sigma);
    ph.nm = p.n / p.m;
    // exponents of q = sigma^2 / r^2: m/2 + 1 and n/2 + 1, split into the nearest integer and a rest
    float em = p.m / 2.f + 1.f, en = p.n / 2.f + 1.f;
    ph.km = (int)lrintf(em);
    ph.kn = (int)lrintf(en);
    ph.fm = em - (float)ph.km;
    ph.fn = en - (float)ph.kn;
    ph.m = p.m;
    float C = (p.n / (p.n - p.m)) * powf(p.n / p.m, p.m / (p.n - p.m));  // particle.cuh:53-55
    ph.pair_scale = C * p.epsilon * p.m * ph.kx / (p.sigma * p.sigma);
    ph.pair_scale_y = ph.pair_scale;
    ph.wall_scale = C * p.epsilon * p.m;
    ph.wall_m6 = p.m == 6.f;
    {   // where a wall's term stops mattering (Phys::wall_skip_x): d = sigma 10^(8 / (m + 1)), in fixed-point units
        const double d_skip = (double)p.sigma * std::pow(10.0, 8.0 / ((double)p.m + 1.0));
        auto units = [&](double box) {
            const double u = d_skip / box * 4294967296.0;
            return p.m > 0.f && u < 2147483647.0 ? (uint32_t)u : 0xFFFFFFFFu;
        };
        ph.wall_skip_x = units(m.box_width);
        ph.wall_skip_y = units(m.box_height);
        if (getenv("PSIM_NO_WALL_SKIP")) ph.wall_skip_x = ph.wall_skip_y = 0xFFFFFFFFu;
    }
    ph.sigma = p.sigma;
    ph.inv_mass = 1.f / (float)6.63352599e-26;  // particle.cuh:51
    ph.dt = m.step_dt;
    ph.ux = m.step_dt / m.box_width * two32;
    ph.uy = m.step_dt / m.box_height * two32;
    ph.cursor_x = m.cursor_pos[0];
    ph.cursor_y = m.cursor_pos[1];
    ph.cursor_r2 = m.cursor_size * m.cursor_size / 4;
    {   // particles live in [0, 1)^2 of cursor units: the cursor matters only if its disc meets that square
        const float r = std::sqrt(std::max(ph.cursor_r2, 0.f));
        const bool reach = ph.cursor_x > -r && ph.cursor_x < 1.f + r && ph.cursor_y > -r && ph.cursor_y < 1.f + r;
        ph.cursor_on = reach && ph.cursor_r2 > 0.f ? 1 : 0;
    }

    // Which step-kernel variant evaluates these exponents.
    //   m = 6 and 5 <= kn <= 10: the integer powers are compile-time products (KN = kn);
    //   the rest fn of the repulsive exponent: none / cubic in z / MUFU.EX2, see pair2().
    // The cubic is accepted only if the error it adds to g stays below 3e-8 (g is O(0.1..1); fp32
    // rounding of the terms themselves is ~1e-7) over r from 0.5 sigma to beyond the stencil reach.
    ph.c1 = 0.69314718f;
    ph.c2 = 0.24022651f;
    ph.c3 = 0.05550411f;
    bool fixed_powers = ph.km == 4 && ph.fm == 0.f && ph.kn >= 5 && ph.kn <= 10;
    *kernel_kn = fixed_powers ? ph.kn : 0;
    if (ph.fn == 0.f && (fixed_powers || ph.fm == 0.f)) {
        *kernel_frac = kFracNone;
    } else if (fixed_powers) {
        float c[3];
        double cost = fit_exp2_cubic(ph.fn, ph.kn, ph.nm, 1e-3, 4.0, c);
        if (cost <= 3e-8) {
            ph.c1 = c[0];
            ph.c2 = c[1];
            ph.c3 = c[2];
            *kernel_frac = kFracPoly;
        } else {
            *kernel_frac = kFracEx2;
        }
    } else {
        *kernel_frac = kFracEx2;
    }
    return ph;
}

void drop_frame_graphs(PsimStepper* s);

// Constants of step_kernel_c (step_float.cuh). Returns false when these physics / this grid have no fp32 variant.
bool make_phys_f(const FrameMetadata& m, const Phys& ph, const Grid& g, int kn, int frac, PhysF* out) {
    if (kn == 0 || frac == kFracEx2) return false;  // run-time exponents, MUFU.EX2 sliver: step_kernel only
    int ye;
    if (std::frexp((double)ph.yscale, &ye) != 0.5) return false;  // ky/kx must fold into a power-of-two scale
    const MiePotentialParams& p = m.particles[0];
    PhysF pf{};
    // kx / sigma = f * 2^ex with f in [1, 2): offsets are scaled by 2^ex, f goes into the constants
    int ex;
    const double inv_c = (double)ph.kx / (double)p.sigma;
    const double f = 2.0 * std::frexp(inv_c, &ex);
    ex -= 1;
    pf.sx = (float)std::ldexp(1.0, ex);
    pf.sy = pf.sx * ph.yscale;
    pf.sxbits = g.sx;
    pf.zl = g.sx <= 21 ? 2 : 1;
    const uint32_t stride = 1u << pf.zl;
    pf.half_span = (stride + 2u) << (g.sx - 1);
    pf.zone_shift = (float)std::ldexp((double)stride, (int)g.sx + ex);
    pf.row_shift = (float)(std::ldexp(1.0, (int)g.sy) * (double)pf.sy);
    // scaled r^2 = true (r/sigma)^2 / f^2, so with qs = 1/scaled r^2 = q f^2:
    //   g f^(2 km) = qs^km - (n/m) f^(-2(kn-km)) qs^kn q^fn ,   q^fn = 2^z,  z = fn log2 q = -fn (l + 2 log2 f)
    const double km = ph.km, knd = ph.kn, fn = ph.fn;
    const double Kc = (double)ph.nm * std::pow(f, -2.0 * (knd - km));
    if (frac == kFracNone) {
        pf.d0 = (float)-Kc;
    } else {
        const double c1 = ph.c1, c2 = ph.c2, c3 = ph.c3;  // 2^z ~ 1 + z (c1 + z (c2 + z c3)), fitted by make_phys
        const double a = -fn, b = -2.0 * fn * std::log2(f);
        pf.d0 = (float)(-Kc * (1.0 + b * (c1 + b * (c2 + b * c3))));
        pf.d1 = (float)(-Kc * a * (c1 + b * (2.0 * c2 + 3.0 * c3 * b)));
        pf.d2 = (float)(-Kc * a * a * (c2 + 3.0 * c3 * b));
        pf.d3 = (float)(-Kc * a * a * a * c3);
    }
    pf.pair_scale = (float)((double)ph.pair_scale * std::pow(f, -2.0 * km) / (double)pf.sx);
    *out = pf;
    return true;
}

// Pair tables of step_kernel_species: like pairs use their own parameters, the unlike pair the Lorentz-Berthelot mix.
void make_species_tables(const FrameMetadata& m, const Phys& ph, SpeciesTab tab[3]) {
    const MiePotentialParams& a = m.particles[0];
    const MiePotentialParams& b = m.particles[1];
    MiePotentialParams mix;
    mix.sigma = (a.sigma + b.sigma) * 0.5f;
    mix.epsilon = std::sqrt(a.epsilon * b.epsilon);
    mix.n = (a.n + b.n) * 0.5f;
    mix.m = (a.m + b.m) * 0.5f;
    const MiePotentialParams pairs[3] = {a, mix, b};
    for (int k = 0; k < 3; ++k) {
        const MiePotentialParams& p = pairs[k];
        const float C = (p.n / (p.n - p.m)) * powf(p.n / p.m, p.m / (p.n - p.m));  // particle.cuh:53-55
        SpeciesTab& t = tab[k];
        t.inv_c2 = (ph.kx / p.sigma) * (ph.kx / p.sigma);
        t.em = p.m / 2.f + 1.f;
        t.en = p.n / 2.f + 1.f;
        t.nm = p.n / p.m;
        t.pair_scale = C * p.epsilon * p.m * ph.kx / (p.sigma * p.sigma);
        t.sigma = p.sigma;
        t.wall_scale = C * p.epsilon * p.m;
        t.m = p.m;
    }
}

void apply_metadata(PsimStepper* s, const FrameMetadata& m) {
    // the same metadata again (a scene re-uploaded every frame by a pipelined host): captured frames stay valid
    const bool same = s->has_scene && std::memcmp(&s->meta, &m, sizeof m) == 0;
    s->meta = m;
    s->phys = make_phys(m, &s->kernel_kn, &s->kernel_frac, &s->kernel_aniso);
    s->species_mode = s->cfg.species_physics && std::memcmp(&m.particles[0], &m.particles[1], sizeof(MiePotentialParams)) != 0;
    if (s->species_mode) make_species_tables(m, s->phys, s->species.tab);
    s->float_path = s->float_grid && !s->force_int_path && !s->species_mode &&
                    make_phys_f(m, s->phys, s->grid, s->kernel_kn, s->kernel_frac, &s->physf);
    s->nbr_stale = true;  // the records carry the old scale (or were not kept at all): rebuilt before the next step
    if (!same) drop_frame_graphs(s);  // captured launches carry the old constants
}

template <int KN, int FRAC>
void launch_step_aniso(PsimStepper* s, const StepArgs& a, uint32_t tiles) {
    if (s->kernel_aniso) step_kernel<KN, FRAC, true><<<tiles, kTile, 0, s->stream>>>(a);
    else step_kernel<KN, FRAC, false><<<tiles, kTile, 0, s->stream>>>(a);
}

template <int KN>
void launch_step_frac(PsimStepper* s, const StepArgs& a, uint32_t tiles) {
    switch (s->kernel_frac) {
        case kFracNone: launch_step_aniso<KN, kFracNone>(s, a, tiles); break;
        case kFracPoly: launch_step_aniso<KN, kFracPoly>(s, a, tiles); break;
        default: launch_step_aniso<KN, kFracEx2>(s, a, tiles); break;
    }
}

// Launch with programmatic dependent launch allowed (the kernel starts with griddepcontrol.wait): its CTAs may be placed
// while the previous kernel of the stream drains, so the launch latency hides behind that kernel's last wave.
template <typename... Params, typename... Args>
void launch_pdl(PsimStepper* s, void (*kernel)(Params...), uint32_t grid, uint32_t block, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.stream = s->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = s->pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <int KN>
void launch_step_c(PsimStepper* s, const StepArgs& a) {
    StepArgsC ac;
    ac.couple_i0 = s->couple_i0;
    ac.tiles = s->tiles_c;
    ac.n_tiles = s->d_couple_tiles;
    if (s->kernel_frac == kFracNone) launch_pdl(s, step_kernel_c<KN, kFracNone>, s->tiles_c_launch, kCouples, a, ac);
    else launch_pdl(s, step_kernel_c<KN, kFracPoly>, s->tiles_c_launch, kCouples, a, ac);
    if (s->timing_after_main) {  // step timing brackets the step kernel itself, not its tail launch
        cudaEventRecord(s->timing_after_main, s->stream);
        s->timing_after_main = nullptr;
    }
    if (s->nranks > 1) return;  // slabs launch exactly the tiles there are
    // a single slab sized the launch from the count it last saw: whatever the last re-bin made beyond it
    if (s->kernel_frac == kFracNone)
        launch_pdl(s, step_kernel_c_surplus<KN, kFracNone>, kSurplusCtas, kCouples, a, ac, s->tiles_c_launch);
    else
        launch_pdl(s, step_kernel_c_surplus<KN, kFracPoly>, kSurplusCtas, kCouples, a, ac, s->tiles_c_launch);
    s->launches += 1;
}

template <int KN, int FRAC>
void launch_allpairs_aniso(PsimStepper* s, const StepArgs& a, uint32_t tiles) {
    if (s->kernel_aniso) allpairs_step_kernel<KN, FRAC, true><<<tiles, kTile, 0, s->stream>>>(a);
    else allpairs_step_kernel<KN, FRAC, false><<<tiles, kTile, 0, s->stream>>>(a);
}

// All-pairs mode keeps two variants: the default exponents' fast path and the general one.
void launch_allpairs(PsimStepper* s, const StepArgs& a, uint32_t tiles) {
    if (s->kernel_kn == 8 && s->kernel_frac == kFracPoly) launch_allpairs_aniso<8, kFracPoly>(s, a, tiles);
    else if (s->kernel_kn == 8 && s->kernel_frac == kFracNone) launch_allpairs_aniso<8, kFracNone>(s, a, tiles);
    else launch_allpairs_aniso<0, kFracEx2>(s, a, tiles);
}

void launch_step(PsimStepper* s, const StepArgs& a, uint32_t tiles) {
    if (s->compact_mode) {
        launch_allpairs(s, a, tiles);
        return;
    }
    if (s->species_mode) {
        s->species.ty = s->ty[s->cur_ty];
        if (s->kernel_aniso) step_kernel_species<true><<<tiles, kTile, 0, s->stream>>>(a, s->species);
        else step_kernel_species<false><<<tiles, kTile, 0, s->stream>>>(a, s->species);
        return;
    }
    if (s->float_path) {
        switch (s->kernel_kn) {
            case 5: launch_step_c<5>(s, a); break;
            case 6: launch_step_c<6>(s, a); break;
            case 7: launch_step_c<7>(s, a); break;
            case 8: launch_step_c<8>(s, a); break;
            case 9: launch_step_c<9>(s, a); break;
            default: launch_step_c<10>(s, a); break;
        }
        return;
    }
    switch (s->kernel_kn) {
        case 5: launch_step_frac<5>(s, a, tiles); break;
        case 6: launch_step_frac<6>(s, a, tiles); break;
        case 7: launch_step_frac<7>(s, a, tiles); break;
        case 8: launch_step_frac<8>(s, a, tiles); break;
        case 9: launch_step_frac<9>(s, a, tiles); break;
        case 10: launch_step_frac<10>(s, a, tiles); break;
        default:  // run-time exponents; kFracPoly is never selected for them
            if (s->kernel_frac == kFracNone) launch_step_aniso<0, kFracNone>(s, a, tiles);
            else launch_step_aniso<0, kFracEx2>(s, a, tiles);
            break;
    }
}

// ------------------------------------------------------------------------------------------------
// Slab transports.  A "team" is the set of slabs one host thread drives in lock-step:
//   * a lone stepper (nranks == 1: no exchange at all; nranks > 1: one process per slab, exchanges
//     are NCCL send/recv pairs with the two adjacent ranks over NVLink), or
//   * a PsimGroup: every slab in this process on one device and one stream; exchanges are
//     device-to-device copies. It exists to validate the decomposition bit-for-bit on a single GPU.
// Every exchange moves data only between adjacent slabs: [0] = lower neighbour, [1] = upper.
// ------------------------------------------------------------------------------------------------

struct Team {
    PsimStepper* const* ranks;
    int count;
    PsimGroup* group;
};

bool load_nccl(PsimStepper* s) {
    if (g_nccl.lib) return true;
    const char* env = getenv("PSIM_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* name : names) {
        if (!name || !*name) continue;
        lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) {
        fail(s, PSIM_ENCCL, "cannot load NCCL (libnccl.so.2; set PSIM_NCCL_LIB): %s", dlerror());
        return false;
    }
    NcclApi api;
    api.lib = lib;
    bool ok = true;
    auto sym = [&](const char* name) {
        void* p = dlsym(lib, name);
        if (!p) ok = false;
        return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.CommAbort = reinterpret_cast<decltype(api.CommAbort)>(sym("ncclCommAbort"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    if (!ok) {
        fail(s, PSIM_ENCCL, "the NCCL library lacks a required symbol");
        return false;
    }
    g_nccl = api;
    return true;
}

// One exchange of a lone slab with its neighbours: a grouped send/recv pair per side.
int exchange_nccl(PsimStepper* s, const XferOp& op, cudaStream_t stream) {
    if (!s->comm) return fail(s, PSIM_ESTATE, "slab %d of %d has no communicator: call psim_comm_init", s->rank, s->nranks);
    bool any = false;
    for (int dir = 0; dir < 2; ++dir) any = any || op.send_bytes[dir] || op.recv_bytes[dir];
    if (!any) return PSIM_OK;
    CKN(g_nccl.GroupStart());
    for (int dir = 0; dir < 2; ++dir) {
        int peer = dir == 0 ? s->rank - 1 : s->rank + 1;
        if (peer < 0 || peer >= s->nranks) continue;
        if (op.send_bytes[dir]) CKN(g_nccl.Send(op.send[dir], op.send_bytes[dir], kNcclUint8, peer, s->comm, stream));
        if (op.recv_bytes[dir]) CKN(g_nccl.Recv(op.recv[dir], op.recv_bytes[dir], kNcclUint8, peer, s->comm, stream));
    }
    CKN(g_nccl.GroupEnd());
    return PSIM_OK;
}

int team_exchange(const Team& t, const std::vector<XferOp>& ops) {
    if (!t.group) {
        PsimStepper* s = t.ranks[0];
        if (s->nranks == 1) return PSIM_OK;
        return exchange_nccl(s, ops[0], s->stream);
    }
    for (int r = 0; r < t.count; ++r) {
        PsimStepper* s = t.ranks[r];
        for (int dir = 0; dir < 2; ++dir) {
            int peer = dir == 0 ? r - 1 : r + 1;
            if (peer < 0 || peer >= t.count) continue;
            if (ops[r].send_bytes[dir] != ops[peer].recv_bytes[dir ^ 1])
                return fail(s, PSIM_EINVAL, "internal: slab %d sends %zu bytes, slab %d expects %zu", r,
                            ops[r].send_bytes[dir], peer, ops[peer].recv_bytes[dir ^ 1]);
            if (ops[r].send_bytes[dir])
                CK(cudaMemcpyAsync(ops[peer].recv[dir ^ 1], ops[r].send[dir], ops[r].send_bytes[dir],
                                   cudaMemcpyDeviceToDevice, t.group->stream));
        }
    }
    return PSIM_OK;
}

inline bool has_lower(const PsimStepper* s) { return s->rank > 0; }
inline bool has_upper(const PsimStepper* s) { return s->rank + 1 < s->nranks; }

// The positions of this slab's two boundary rows go to the neighbours' ghost rows of the same buffer.
XferOp ghost_positions_op(PsimStepper* s) {
    XferOp op;
    uint2* pos = s->pos[s->cur_pos];
    if (has_lower(s)) {
        op.send[0] = pos + s->own_lo;
        op.send_bytes[0] = sizeof(uint2) * (size_t)(s->b_lo_end - s->own_lo);
        op.recv[0] = pos;
        op.recv_bytes[0] = sizeof(uint2) * (size_t)s->own_lo;
    }
    if (has_upper(s)) {
        op.send[1] = pos + s->b_hi_start;
        op.send_bytes[1] = sizeof(uint2) * (size_t)(s->own_hi - s->b_hi_start);
        op.recv[1] = pos + s->own_hi;
        op.recv_bytes[1] = sizeof(uint2) * (size_t)(s->n_total - s->own_hi);
    }
    return op;
}

// Neighbour records of [lo, hi) of the current buffers from the positions there (nbr_rebuild_kernel).
int enqueue_nbr_rebuild(PsimStepper* s, uint32_t lo, uint32_t hi) {
    if (hi <= lo) return PSIM_OK;
    nbr_rebuild_kernel<<<div_up(hi - lo, 256), 256, 0, s->stream>>>(s->pos[s->cur_pos], s->cell_start, s->grid, s->physf,
                                                                   lo, hi, s->nbr[s->cur_pos]);
    s->launches += 1;
    CK(cudaGetLastError());
    return PSIM_OK;
}

// The ghost rows have just arrived as bare positions (an exchange): give them their neighbour records.
int refresh_ghost_records(PsimStepper* s) {
    if (!s->float_path || s->nranks == 1) return PSIM_OK;
    int rc = enqueue_nbr_rebuild(s, 0, s->own_lo);
    if (rc) return rc;
    return enqueue_nbr_rebuild(s, s->own_hi, s->n_total);
}

// ------------------------------------------------------------------------------------------------
// Binning, phase by phase (every phase is run for all slabs of the team before the next one starts)
// ------------------------------------------------------------------------------------------------

// cell_start = exclusive scan of cell_count; total -> cell_start[cells]. Fine grids: pad_start in the same pass (the
// counts rounded up to even: couples of step_kernel_c), then the couples and the tiles of every owned row.
int enqueue_scan(PsimStepper* s) {
    uint32_t cells = s->grid.cells;
    uint32_t blocks = div_up(cells, kScanBlock);
    scan_reduce_kernel<<<blocks, kScanThreads, 0, s->stream>>>(s->cell_count, cells, s->block_sum);
    scan_top_kernel<<<1, 1024, 0, s->stream>>>(s->block_sum, blocks, s->cell_start + cells,
                                               s->float_grid ? s->pad_start + cells : nullptr);
    if (s->float_grid)
        scan_apply_kernel<true><<<blocks, kScanThreads, 0, s->stream>>>(s->cell_count, cells, s->block_sum, s->cell_start, s->pad_start);
    else
        scan_apply_kernel<false><<<blocks, kScanThreads, 0, s->stream>>>(s->cell_count, cells, s->block_sum, s->cell_start, nullptr);
    s->launches += 3;
    if (s->float_grid) {
        couple_build_kernel<<<div_up(cells, 256), 256, 0, s->stream>>>(s->cell_start, s->pad_start, cells, s->couple_i0);
        row_cut_kernel<false><<<div_up(s->grid.own_rows, kCutRowsPerCta), 32 * kCutRowsPerCta, 0, s->stream>>>(s->cell_start, s->pad_start, s->grid,
                                                                                s->tile_base, s->tiles_c, s->tiles_c_cap);
        row_tiles_kernel<<<1, 1024, 0, s->stream>>>(s->grid, s->tile_base, s->tiles_c_cap, s->d_couple_tiles,
                                                    s->h_counts_dev + 12);
        s->launches += 3;
    }
    CK(cudaGetLastError());
    return PSIM_OK;
}

// Re-bin, phase 1 (slabs only): particles that left the owned rows go into the two outboxes.
int bin_phase_migrants(PsimStepper* s, XferOp& op) {
    CK(cudaMemsetAsync(s->mig_counters, 0, 2 * sizeof(uint32_t), s->stream));
    const uint32_t nb = div_up(s->n, kMigBlock);  // blocks of consecutive owned particles
    if (nb) {
        migrant_count_kernel<<<nb, kMigBlock, 0, s->stream>>>(s->pos[s->cur_pos], s->own_lo, s->own_hi, s->grid, nb, s->mig_counters,
                                                             s->mig_blk_cnt, s->d_flags);
        s->launches += 1;
    }
    migrant_scan_kernel<<<1, 1024, 0, s->stream>>>(s->mig_blk_cnt, nb, s->box_capacity, s->mig_counters, s->mig_blk_off, s->mig_blk_list,
                                                   s->d_flags);
    const uint32_t workers = std::min(nb, 296u);  // two CTAs per SM of a B200 walk the list
    migrant_pack_kernel<<<workers + div_up(s->box_capacity, kMigBlock), kMigBlock, 0, s->stream>>>(
        s->pos[s->cur_pos], s->vel[s->cur_vel], s->ty[s->cur_ty], s->own_lo, s->own_hi, s->grid, nb, workers, s->mig_counters,
        s->mig_blk_cnt, s->mig_blk_off, s->mig_blk_list, s->box_capacity, s->outbox[0], s->outbox[1]);
    s->launches += 2;
    CK(cudaGetLastError());
    if (has_lower(s)) {
        op.send[0] = s->outbox[0];
        op.recv[0] = s->inbox[0];
        op.send_bytes[0] = op.recv_bytes[0] = s->box_bytes;
    }
    if (has_upper(s)) {
        op.send[1] = s->outbox[1];
        op.recv[1] = s->inbox[1];
        op.send_bytes[1] = op.recv_bytes[1] = s->box_bytes;
    }
    return PSIM_OK;
}

// Phase 2: keys and per-cell counts of the candidates; the counts of the boundary rows are the
// neighbours' ghost-row counts.
int bin_phase_count(PsimStepper* s, const Source& src, XferOp& op) {
    const uint32_t tb = 256;
    s->src = src;
    CK(cudaMemsetAsync(s->cell_count, 0, sizeof(uint32_t) * s->grid.cells, s->stream));
    uint32_t cand = src.n_lo + src.n_soa + src.n_hi;
    if (cand) {
        key_count_kernel<<<div_up(cand, tb), tb, 0, s->stream>>>(src, s->grid, s->cell_count, s->rank_in_cell,
                                                                 s->d_flags);
        s->launches += 1;
        CK(cudaGetLastError());
    }
    const Grid& g = s->grid;
    const size_t row_bytes = sizeof(uint32_t) * g.bx;
    if (has_lower(s)) {
        op.send[0] = s->cell_count + (size_t)g.own_row0 * g.bx;
        op.recv[0] = s->cell_count;
        op.send_bytes[0] = op.recv_bytes[0] = row_bytes;
    }
    if (has_upper(s)) {
        op.send[1] = s->cell_count + (size_t)(g.own_row0 + g.own_rows - 1) * g.bx;
        op.recv[1] = s->cell_count + (size_t)(g.by - 1) * g.bx;
        op.send_bytes[1] = op.recv_bytes[1] = row_bytes;
    }
    return PSIM_OK;
}

// Phase 3: offsets, and the handful of numbers the host needs to size the next launches.
int bin_phase_scan(PsimStepper* s, bool need_counts) {
    int rc = enqueue_scan(s);
    if (rc) return rc;
    if (need_counts) {
        // The counts go straight into page-locked host memory (mapped: the kernel stores over PCIe), not through a
        // device-to-host copy: a copy engine may be busy for milliseconds with the previous frame's snapshot on its way
        // out, and these 48 bytes would queue behind it while the whole stream waits.
        slab_counts_kernel<<<1, 32, 0, s->stream>>>(s->cell_start, s->grid, s->d_flags, s->d_couple_tiles,
                                                    s->float_grid ? s->tile_base : nullptr, s->hdr,
                                                    s->src.strict ? s->mig_counters : nullptr, s->h_counts_dev);
        s->launches += 1;
        CK(cudaGetLastError());
    }
    return PSIM_OK;
}

// The host has read the tile count of step_kernel_c. With slabs it reads it at every binning and launches exactly that
// many CTAs (the halo protocol numbers the boundary tiles). A single slab reads it when a scene is ingested and at
// psim_sync, never inside a frame: the launch keeps a margin over the count last seen (a re-bin changes it by a few
// tiles), the kernel takes the real count from device memory, and the bound moves only when the count has drifted, so
// that captured frames (CUDA graphs) stay valid.
void set_tile_count(PsimStepper* s, uint32_t count) {
    s->n_tiles_c = count;
    if (s->nranks > 1) {
        s->tiles_c_launch = count;
        return;
    }
    uint32_t want = std::min(s->tiles_c_cap, count + count / 64 + 32);
    bool keep = s->tiles_c_launch >= std::min(s->tiles_c_cap, count + count / 256 + 8) && s->tiles_c_launch <= want + want / 16;
    if (getenv("PSIM_TILE_EXACT")) {  // measurements: no margin
        want = count;
        keep = false;
    }
    if (const char* env = getenv("PSIM_TILE_LAUNCH_CAP")) {  // tests: fewer CTAs than tiles, the surplus launch steps the rest
        want = std::max(1u, std::min(want, (uint32_t)std::atoi(env)));
        keep = false;
    }
    if (!keep && s->tiles_c_launch != want) {
        drop_frame_graphs(s);  // the grid size is part of a captured launch
        s->tiles_c_launch = want;
        // CTAs behind the last tile must find empty descriptors (the next binning writes them itself)
        if (want > count) cudaMemsetAsync(s->tiles_c + count, 0, sizeof(TileC) * (size_t)(want - count), s->stream);
    }
}

// Host: wait for the counts (the only host synchronisation of a binning: at an ingest, and at every re-bin with slabs).
int bin_phase_commit(PsimStepper* s) {
    CK(cudaStreamSynchronize(s->stream));
    const uint32_t* h = s->h_counts;
    uint32_t flags = h[5];
    if (flags & kErrMigrantOverflow)
        return fail(s, PSIM_EMIGRATION, "slab %d: more than %u particles left for a neighbour slab in one re-bin "
                    "(PsimConfig.migrant_capacity)", s->rank, s->box_capacity);
    if (flags & kErrHaloTimeout)
        return fail(s, PSIM_ECUDA, "slab %d: a step waited 20 s for a neighbour's halo (did a neighbour rank die?)", s->rank);
    if (flags & (kErrMigrantTooFar | kErrMigrantOutside))
        return fail(s, PSIM_EMIGRATION, "slab %d: a particle moved past the adjacent slab between two re-bins", s->rank);
    uint32_t owned = h[3] - h[0];
    if (owned > s->cfg.max_particles)
        return fail(s, PSIM_ECAPACITY, "%u live particles exceed max_particles = %u", owned, s->cfg.max_particles);
    if (h[0] > s->ghost_cap || h[4] - h[3] > s->ghost_cap)
        return fail(s, PSIM_ECAPACITY, "slab %d: a ghost row holds %u particles, ghost_capacity = %u", s->rank,
                    std::max(h[0], h[4] - h[3]), s->ghost_cap);
    s->own_lo = h[0];
    s->b_lo_end = h[1];
    s->b_hi_start = h[2];
    s->own_hi = h[3];
    s->n_total = h[4];
    s->n = owned;
    s->migrants_sent += std::min(h[10], s->box_capacity) + std::min(h[11], s->box_capacity);
    if (s->float_grid) {
        if (h[9])
            return fail(s, PSIM_ECAPACITY, "internal: more couple tiles than the %u there is room for", s->tiles_c_cap);
        set_tile_count(s, h[6]);
        s->tiles_lo = h[7];
        s->tile_hi0 = h[8];
    }
    s->ghosts_by_push = false;  // phase 4 of this binning delivers the ghost rows itself
    s->commit_pending_pump = true;  // the next piece of the copies in flight goes out behind this binning's launches
    return PSIM_OK;
}

// Phase 4: place the owned particles (stable), flip the buffers; the new boundary rows are the
// neighbours' ghost rows.
int bin_phase_place(PsimStepper* s, bool ingest, XferOp& op) {
    const uint32_t tb = 256;
    const Source& src = s->src;
    uint32_t cand = src.n_lo + src.n_soa + src.n_hi;
    int np = ingest ? 0 : s->cur_pos ^ 1, nv = ingest ? 0 : s->cur_vel ^ 1, nt = ingest ? 0 : s->cur_ty ^ 1;
    if (s->n) {
        scatter_kernel<<<div_up(cand, tb), tb, 0, s->stream>>>(src, s->grid, s->cell_start, s->rank_in_cell, s->perm);
        gather_kernel<<<div_up(s->n, tb), tb, 0, s->stream>>>(src, s->own_lo, s->own_hi, s->grid, s->cell_start,
                                                              s->perm, s->pos[np], s->vel[nv], s->ty[nt], s->cell_id,
                                                              s->float_path ? s->nbr[np] : nullptr, s->physf);
        s->launches += 2;
        CK(cudaGetLastError());
    }
    s->cur_pos = np;
    s->cur_vel = nv;
    s->cur_ty = nt;
    op = ghost_positions_op(s);
    return PSIM_OK;
}

int enqueue_tile_desc(PsimStepper* s) {
    uint32_t tiles = div_up(s->n, kTile);
    if (tiles)
        tile_desc_kernel<<<div_up(tiles, 128), 128, 0, s->stream>>>(s->cell_start, s->grid, s->own_lo, s->own_hi, s->tiles);
    s->launches += 1;
    s->tiles_stale = false;
    CK(cudaGetLastError());
    return PSIM_OK;
}

// Phase 5: staging descriptors of the step kernel's tiles.
int bin_phase_tiles(PsimStepper* s) {
    uint32_t tiles = div_up(s->n, kTile);
    if (tiles == 0) return PSIM_OK;
    // step_kernel's descriptors: built now where it is the kernel that runs, else when a step first needs them
    s->tiles_stale = true;
    if (!s->float_path) {
        int rc = enqueue_tile_desc(s);
        if (rc) return rc;
    }
    if (s->float_grid) {
        row_cut_kernel<true><<<div_up(s->grid.own_rows, kCutRowsPerCta), 32 * kCutRowsPerCta, 0, s->stream>>>(s->cell_start, s->pad_start, s->grid,
                                                                               s->tile_base, s->tiles_c, s->tiles_c_cap);
        s->launches += 1;
        // with slabs the host knows the count of this binning; a single slab covers whatever the count has become
        const uint32_t upto = s->nranks > 1 ? s->n_tiles_c : s->tiles_c_cap;
        tile_build_kernel<<<div_up(std::max(upto, 1u), 128), 128, 0, s->stream>>>(s->cell_start, s->pad_start, s->tile_base,
                                                                                 s->d_couple_tiles, s->tiles_c_launch, s->couple_i0,
                                                                                 s->grid, s->tiles_c);
        s->launches += 1;
    }
    CK(cudaGetLastError());
    return PSIM_OK;
}

// Bin all slabs of the team: an ingested frame (`ingest`: `records`/`count` in device memory), or the
// live state (bucket_move, kernel_bucket.cuh:5-39) including migration between slabs.
int team_bin(const Team& t, bool ingest, const Particle* records, uint32_t count) {
    if (!ingest && t.ranks[0]->compact_mode) return PSIM_OK;  // CompactArray: there is no grid to re-bin into
    const bool slabs = t.ranks[0]->nranks > 1;
    std::vector<XferOp> ops(t.count);
    int rc;
    for (int r = 0; r < t.count; ++r) {
        PsimStepper* s = t.ranks[r];
        CK(cudaMemsetAsync(s->d_flags, 0, sizeof(uint32_t), s->stream));
    }
    if (!ingest && slabs) {
        for (int r = 0; r < t.count; ++r)
            if ((rc = bin_phase_migrants(t.ranks[r], ops[r]))) return rc;
        if ((rc = team_exchange(t, ops))) return rc;
    }
    ops.assign(t.count, XferOp());
    for (int r = 0; r < t.count; ++r) {
        PsimStepper* s = t.ranks[r];
        Source src{};
        if (ingest) {
            src.aos_lo = records;
            src.n_lo = count;
            src.strict = 0;
        } else {
            src.pos = s->pos[s->cur_pos];
            src.vel = s->vel[s->cur_vel];
            src.ty = s->ty[s->cur_ty];
            src.soa_lo = s->own_lo;
            src.n_soa = s->n;
            src.strict = 1;
            if (has_lower(s)) {
                src.aos_lo = reinterpret_cast<const Particle*>(s->inbox[0] + sizeof(MigrantBoxHeader));
                src.n_lo = s->box_capacity;
            }
            if (has_upper(s)) {
                src.aos_hi = reinterpret_cast<const Particle*>(s->inbox[1] + sizeof(MigrantBoxHeader));
                src.n_hi = s->box_capacity;
            }
        }
        if ((rc = bin_phase_count(s, src, ops[r]))) return rc;
    }
    if ((rc = team_exchange(t, ops))) return rc;
    const bool need_counts = ingest || slabs;  // a single slab's re-bin needs nothing back on the host
    for (int r = 0; r < t.count; ++r)
        if ((rc = bin_phase_scan(t.ranks[r], need_counts))) return rc;
    if (need_counts) {
        int first_error = PSIM_OK;
        for (int r = 0; r < t.count; ++r) {
            rc = bin_phase_commit(t.ranks[r]);
            if (rc && !first_error) {
                first_error = rc;
                if (t.group) t.group->error = t.ranks[r]->error;
            }
        }
        if (first_error) return first_error;
    }
    ops.assign(t.count, XferOp());
    for (int r = 0; r < t.count; ++r)
        if ((rc = bin_phase_place(t.ranks[r], ingest, ops[r]))) return rc;
    if ((rc = team_exchange(t, ops))) return rc;
    for (int r = 0; r < t.count; ++r) {
        PsimStepper* s = t.ranks[r];
        if ((rc = refresh_ghost_records(s))) return rc;
        s->nbr_stale = !s->float_path;  // the gather wrote the owned rows' records, the line above the ghosts'
        if ((rc = bin_phase_tiles(t.ranks[r]))) return rc;
        if (!ingest) t.ranks[r]->rebins_executed += 1;
        if (s->commit_pending_pump) {  // the host round trip of this binning is over: the quiet 17 steps begin
            s->commit_pending_pump = false;
            if (!ingest && !s->copy_pace_steps) {
                if ((rc = pump_copy(s, s->paced_down, false))) return rc;
                if ((rc = pump_copy(s, s->paced_up, false))) return rc;
            }
        }
    }
    return PSIM_OK;
}

// ------------------------------------------------------------------------------------------------
// Steps, snapshots, frames
// ------------------------------------------------------------------------------------------------

int enqueue_step(PsimStepper* s) {
    StepArgs a;
    a.pos_in = s->pos[s->cur_pos];
    a.pos_out = s->pos[s->cur_pos ^ 1];
    a.vel = s->vel[s->cur_vel];
    a.cell_id = s->cell_id;
    a.cell_start = s->cell_start;
    a.tiles = s->tiles;
    a.own_lo = s->own_lo;
    a.own_hi = s->own_hi;
    a.g = s->grid;
    a.ph = s->phys;
    a.pf = s->physf;
    a.nbr_in = a.nbr_out = nullptr;
    if (!s->float_path && !s->compact_mode && s->tiles_stale) {  // new metadata took the fine grid to step_kernel
        int rc = enqueue_tile_desc(s);
        if (rc) return rc;
    }
    if (s->float_path && !s->compact_mode) {
        if (s->nbr_stale)  // team_step rebuilds stale records before it enqueues a step
            return fail(s, PSIM_ESTATE, "internal: a step was enqueued on stale neighbour records");
        a.nbr_in = s->nbr[s->cur_pos];
        a.nbr_out = s->nbr[s->cur_pos ^ 1];
    }
    a.push = s->push ? 1u : 0u;
    std::memset(&a.h, 0, sizeof a.h);
    if (s->push) {
        HaloArgs& h = a.h;
        const int out = s->cur_pos ^ 1;
        h.hdr = s->hdr;
        for (int side = 0; side < 2; ++side) {
            h.peer_out[side] = s->peer_pos[side][out];
            h.peer_nbr_out[side] = a.nbr_out ? s->peer_nbr[side][out] : nullptr;
            h.peer_hdr[side] = s->peer_hdr[side];
            h.wait_epoch[side] = s->ghosts_by_push ? s->halo_epoch : 0u;
        }
        h.lo_end = s->b_lo_end;
        h.hi_start = s->b_hi_start;
        if (s->float_path) {
            h.lo_tiles = s->tiles_lo;
            h.hi_tile0 = s->tile_hi0;
        } else {
            h.lo_tiles = div_up(s->b_lo_end - s->own_lo, kTile);
            h.hi_tile0 = (s->b_hi_start - s->own_lo) / kTile;
        }
        s->halo_epoch += 1;
        if (s->halo_epoch == 0) s->halo_epoch = 1;  // 0 means "nothing to wait for"
        h.pub_epoch = s->halo_epoch;
        s->ghosts_by_push = true;
    }
    if (s->n == 0) {  // nothing to step; the neighbours still get this step's epoch
        if (s->push) {
            halo_publish_kernel<<<1, 32, 0, s->stream>>>(a);
            s->launches += 1;
            CK(cudaGetLastError());
        }
        s->cur_pos ^= 1;  // ghost rows arrive in the buffer the next step reads
        s->steps_executed += 1;
        return PSIM_OK;
    }
    uint32_t tiles = div_up(s->n, kTile);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (s->timing) {
        if (s->timing_used == s->timing_events.size()) {
            cudaEvent_t x, y;
            CK(cudaEventCreate(&x));
            CK(cudaEventCreate(&y));
            s->timing_events.push_back({x, y});
        }
        e0 = s->timing_events[s->timing_used].first;
        e1 = s->timing_events[s->timing_used].second;
        s->timing_used += 1;
        CK(cudaEventRecord(e0, s->stream));
    }
    s->timing_after_main = s->timing ? e1 : nullptr;
    launch_step(s, a, tiles);
    if (s->timing_after_main) {
        CK(cudaEventRecord(e1, s->stream));
        s->timing_after_main = nullptr;
    }
    CK(cudaGetLastError());
    s->launches += 1;
    s->cur_pos ^= 1;
    s->steps_executed += 1;
    return PSIM_OK;
}

// New metadata (another scale of the neighbour records, or records that were not kept so far): rebuild the records of
// the positions the next step reads. A slab rebuilds its OWN rows only. Its ghost rows may still be receiving the
// positions and old-scale records its neighbours' last step pushes (that step is not waited for here), so they are
// delivered again by an explicit exchange of the boundary rows -- ordered after the neighbours' last step by the
// transport -- and converted on arrival; the next step then has nothing to wait for, exactly as after a binning.
int team_refresh_stale_records(const Team& t) {
    bool any = false;
    for (int r = 0; r < t.count; ++r) {
        const PsimStepper* s = t.ranks[r];
        any = any || (s->float_path && !s->compact_mode && s->nbr_stale);
    }
    if (!any) return PSIM_OK;
    int rc;
    const bool redeliver = t.ranks[0]->nranks > 1 && t.ranks[0]->push && t.ranks[0]->ghosts_by_push;
    for (int r = 0; r < t.count; ++r) {
        PsimStepper* s = t.ranks[r];
        if (redeliver) rc = enqueue_nbr_rebuild(s, s->own_lo, s->own_hi);
        else rc = enqueue_nbr_rebuild(s, 0, s->n_total);  // ghost rows are in place (a binning or an exchange put them there)
        if (rc) return rc;
    }
    if (redeliver) {
        std::vector<XferOp> ops(t.count);
        for (int r = 0; r < t.count; ++r) ops[r] = ghost_positions_op(t.ranks[r]);
        if ((rc = team_exchange(t, ops))) return rc;
        for (int r = 0; r < t.count; ++r) {
            if ((rc = refresh_ghost_records(t.ranks[r]))) return rc;
            t.ranks[r]->ghosts_by_push = false;
        }
    }
    for (int r = 0; r < t.count; ++r) t.ranks[r]->nbr_stale = false;
    return PSIM_OK;
}

// One leapfrog step of every slab, then the halo exchange: each slab's new boundary-row positions
// become its neighbours' ghost rows for the next step.
int team_step(const Team& t) {
    int rc;
    if ((rc = team_refresh_stale_records(t))) return rc;
    for (int r = 0; r < t.count; ++r) {
        if ((rc = enqueue_step(t.ranks[r]))) return rc;
        t.ranks[r]->fresh_scene = false;
        if ((rc = pump_copies_behind_step(t.ranks[r]))) return rc;
    }
    if (t.ranks[0]->nranks == 1 || t.ranks[0]->push) return PSIM_OK;  // pushed by the step kernel itself
    std::vector<XferOp> ops(t.count);
    for (int r = 0; r < t.count; ++r) ops[r] = ghost_positions_op(t.ranks[r]);
    if ((rc = team_exchange(t, ops))) return rc;
    for (int r = 0; r < t.count; ++r)
        if ((rc = refresh_ghost_records(t.ranks[r]))) return rc;
    return PSIM_OK;
}

int enqueue_snapshot(PsimStepper* s) {
    const int k = s->snaps_taken ? (s->snap_latest + 1) % s->nsnap : 0;
    // the snapshot that lived in this buffer must have left it before it is overwritten: a paced download still reading it
    // sends the rest now
    if (s->paced_down.active && s->paced_down.src == reinterpret_cast<const char*>(s->snapshot[k])) {
        int rc = pump_copy(s, s->paced_down, true);
        if (rc) return rc;
    }
    CK(cudaStreamWaitEvent(s->stream, s->snapshot_consumed[k], 0));
    const uint32_t packed = div_up(s->n, s->snapshot_stride);
    if (packed) {
        pack_kernel<<<div_up(packed, 256), 256, 0, s->stream>>>(s->pos[s->cur_pos], s->vel[s->cur_vel], s->ty[s->cur_ty],
                                                                s->own_lo, s->own_hi, s->snapshot_stride, s->snapshot[k]);
        s->launches += 1;
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(s->snapshot_ready[k], s->stream));
    s->snapshot_n[k] = packed;
    s->snapshot_meta[k] = s->meta;
    s->snap_latest = k;
    s->snaps_taken += 1;
    s->has_snapshot = true;
    return PSIM_OK;
}

int team_snapshot(const Team& t) {
    int rc;
    for (int r = 0; r < t.count; ++r)
        if ((rc = enqueue_snapshot(t.ranks[r]))) return rc;
    return PSIM_OK;
}

int collect_timing(PsimStepper* s) {
    for (size_t k = 0; k < s->timing_used; ++k) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, s->timing_events[k].first, s->timing_events[k].second));
        s->timing_total_ms += ms;
        s->timing_launches += 1;
    }
    s->timing_used = 0;
    return PSIM_OK;
}

// Bin `count` wire-format records that sit in device memory at `records`, on every slab of the team.
int team_ingest(const Team& t, const Particle* records, uint32_t count) {
    for (int r = 0; r < t.count; ++r) t.ranks[r]->compact_mode = false;
    int rc = team_bin(t, true, records, count);
    if (rc) return rc;
    for (int r = 0; r < t.count; ++r) {
        PsimStepper* s = t.ranks[r];
        s->has_scene = true;
        s->native_countdown = 0;
        s->fresh_scene = true;
    }
    // the ingested scene itself can be downloaded (cuda_simulator.cu:28-31)
    if ((rc = team_snapshot(t))) return rc;
    for (int r = 0; r < t.count; ++r) {
        PsimStepper* s = t.ranks[r];
        CK(cudaStreamSynchronize(s->stream));
    }
    return PSIM_OK;
}

// One frame (Kernel::run_async for MatrixBuckets): steps_per_frame steps with re-binning on the
// configured schedule, then a snapshot. All slabs of a team share metadata and schedule state.
int team_frame_steps(const Team& t) {
    PsimStepper* lead = t.ranks[0];
    const uint32_t target = lead->meta.steps_per_frame;
    int rc;
    if (lead->compact_mode) {
        // compact_kernel_run_async, kernel_compact.cuh:78-92: steps_per_frame steps (two when it is 0: the even
        // branch runs its pair of steps unconditionally)
        const uint32_t steps = target == 0 ? 2u : target;
        for (uint32_t k = 0; k < steps; ++k)
            if ((rc = team_step(t))) return rc;
    } else if (lead->cfg.schedule == PSIM_SCHEDULE_REFERENCE) {
        // bucket_kernel_run_async, kernel_bucket.cuh:181-206: the reference always runs one step,
        // then alternates "re-bin + 1 step" with pairs of steps, 16 steps between re-bins counted
        // from the first re-bin, the countdown restarting with every frame. Pairs make it overshoot
        // an odd remainder by one step.
        const int move_every_n = 16;
        int countdown = 0;
        uint32_t steps = 0;
        if ((rc = team_step(t))) return rc;
        steps += 1;
        while (steps < target) {
            if (countdown <= 0) {
                if ((rc = team_bin(t, false, nullptr, 0))) return rc;
                countdown = move_every_n;
                if ((rc = team_step(t))) return rc;
                countdown -= 1;
                steps += 1;
            } else {
                if ((rc = team_step(t))) return rc;
                if ((rc = team_step(t))) return rc;
                countdown -= 2;
                steps += 2;
            }
        }
    } else {
        for (uint32_t k = 0; k < target; ++k) {
            if (lead->native_countdown <= 0) {
                if (!lead->fresh_scene)  // a freshly ingested scene is already binned
                    if ((rc = team_bin(t, false, nullptr, 0))) return rc;
                lead->native_countdown = (int)lead->cfg.rebin_every;
            }
            if ((rc = team_step(t))) return rc;
            lead->native_countdown -= 1;
        }
    }
    for (int r = 0; r < t.count; ++r) t.ranks[r]->native_countdown = lead->native_countdown;
    return PSIM_OK;
}

int team_run_frame(const Team& t) {
    int rc = team_frame_steps(t);
    if (rc) return rc;
    return team_snapshot(t);
}

// ------------------------------------------------------------------------------------------------
// Frames as CUDA graphs (PsimConfig.use_graph). The reference's own scenes are small (<= 65 536 particles): a frame is
// ~100 step launches and a few dozen binning launches of a few microseconds each, bound by launch overhead. On coarse
// grids a frame needs nothing back on the host, so its launches are captured once and replayed with one
// cudaGraphLaunch. A graph is keyed by what the launches depend on besides the metadata: which of the ping-pong
// buffers are current when the frame starts. New metadata or a new scene drops the cache.
// ------------------------------------------------------------------------------------------------
bool frame_graph_usable(const PsimStepper* s) {
    return s->cfg.use_graph && s->nranks == 1 && !s->group && !s->timing &&
           s->cfg.schedule == PSIM_SCHEDULE_REFERENCE && s->n > 0;
}

void drop_frame_graphs(PsimStepper* s) {
    for (auto& kv : s->frame_graphs) cudaGraphExecDestroy(kv.second.exec);
    s->frame_graphs.clear();
}

int run_frame_with_graph(PsimStepper* s) {
    {   // records made stale by new metadata are rebuilt once, outside the captured frame
        PsimStepper* self = s;
        int rc = team_refresh_stale_records(Team{&self, 1, nullptr});
        if (rc) return rc;
    }
    const uint32_t key = (uint32_t)s->cur_pos | (uint32_t)s->cur_vel << 1 | (uint32_t)s->cur_ty << 2;
    auto it = s->frame_graphs.find(key);
    if (it != s->frame_graphs.end() && (it->second.n != s->n || it->second.n_total != s->n_total ||
                                        it->second.tiles_launch != s->tiles_c_launch)) {
        cudaGraphExecDestroy(it->second.exec);  // another scene since: the launches were sized for other counts
        s->frame_graphs.erase(it);
        it = s->frame_graphs.end();
    }
    if (it == s->frame_graphs.end()) {
        const uint64_t steps0 = s->steps_executed, rebins0 = s->rebins_executed, launches0 = s->launches;
        CK(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        PsimStepper* self = s;
        int rc = team_frame_steps(Team{&self, 1, nullptr});  // enqueues into the capture; host-side state advances as usual
        cudaGraph_t graph = nullptr;
        cudaError_t end = cudaStreamEndCapture(s->stream, &graph);
        if (rc) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        if (end != cudaSuccess) return fail(s, PSIM_ECUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(end));
        FrameGraph fg;
        cudaError_t inst = cudaGraphInstantiate(&fg.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (inst != cudaSuccess) return fail(s, PSIM_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(inst));
        fg.steps = s->steps_executed - steps0;
        fg.rebins = s->rebins_executed - rebins0;
        fg.launches = s->launches - launches0;
        fg.end_pos = s->cur_pos;
        fg.end_vel = s->cur_vel;
        fg.end_ty = s->cur_ty;
        fg.n = s->n;
        fg.n_total = s->n_total;
        fg.tiles_launch = s->tiles_c_launch;
        s->frame_graphs[key] = fg;
        CK(cudaGraphLaunch(fg.exec, s->stream));  // the capture executed nothing
        return PSIM_OK;
    }
    const FrameGraph& fg = it->second;
    CK(cudaGraphLaunch(fg.exec, s->stream));
    s->steps_executed += fg.steps;
    s->rebins_executed += fg.rebins;
    s->launches += fg.launches;
    s->cur_pos = fg.end_pos;
    s->cur_vel = fg.end_vel;
    s->cur_ty = fg.end_ty;
    return PSIM_OK;
}

Team lone(PsimStepper* const* s) { return Team{s, 1, nullptr}; }

// A single slab on a fine grid learns the tile count of its last re-bin when the host next waits for it anyway (the
// stream is idle here).
int refresh_tile_count(PsimStepper* s) {
    if (s->nranks > 1 || !s->float_grid || !s->has_scene || s->compact_mode) return PSIM_OK;
    const uint32_t c[2] = {s->h_counts[12], s->h_counts[13]};  // row_tiles_kernel's mirror; the caller has waited for the stream
    if (c[1]) return fail(s, PSIM_ECAPACITY, "internal: more couple tiles than the %u there is room for", s->tiles_c_cap);
    set_tile_count(s, c[0]);
    return PSIM_OK;
}

int check_lone(PsimStepper* s, const char* what) {
    if (s->group) return fail(s, PSIM_ESTATE, "%s: this stepper is a slab of a group; use the psim_group_* call", what);
    return PSIM_OK;
}

void write_header(FrameHeader* dst, const FrameMetadata& meta, uint32_t count) {
    // FrameHeader::new (particle.rs:214-223)
    static const uint8_t sig0[4] = {0x36, 0xbc, 0xe9, 0xbd}, sig1[4] = {0xac, 0xc4, 0x12, 0xec};
    std::memcpy(dst->signature_start, sig0, 4);
    std::memcpy(dst->signature_end, sig1, 4);
    dst->_padding = 0;
    dst->metadata = meta;
    dst->particle_count = count;
}

// Which buffer holds the snapshot taken `age` snapshots ago (0: the latest), or -1.
int snapshot_index(const PsimStepper* s, uint32_t age) {
    if (!s->has_snapshot || age >= (uint32_t)s->nsnap || age >= s->snaps_taken) return -1;
    return (s->snap_latest - (int)age + s->nsnap) % s->nsnap;
}

// A step that gave up waiting for a neighbour's halo ran on stale ghost rows: snapshots taken since are not valid
// frames. The flag is sticky in the slab's HaloHeader; `h_counts[15]` is the host's copy, fetched over the copy stream.
int check_halo_error(PsimStepper* s) {
    if (!s->push || !s->hdr) return PSIM_OK;
    if (s->h_counts[15]) {
        s->failed = true;
        return fail(s, PSIM_ECUDA, "slab %d: a step waited 20 s for a neighbour's halo (did a neighbour rank die?); "
                    "the frames since are invalid", s->rank);
    }
    return PSIM_OK;
}

// Copy one of this slab's packed snapshots to `out` (host); waits only for that snapshot.
int download_records(PsimStepper* s, int k, Particle* out) {
    CK(cudaStreamWaitEvent(s->copy_stream, s->snapshot_ready[k], 0));
    if (s->snapshot_n[k])
        CK(cudaMemcpyAsync(out, s->snapshot[k], sizeof(Particle) * (size_t)s->snapshot_n[k], cudaMemcpyDeviceToHost,
                           s->copy_stream));
    if (s->push && s->hdr)
        CK(cudaMemcpyAsync(&s->h_counts[15], &s->hdr->error, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->copy_stream));
    CK(cudaEventRecord(s->snapshot_consumed[k], s->copy_stream));
    CK(cudaStreamSynchronize(s->copy_stream));
    return check_halo_error(s);
}

}  // namespace

extern "C" {

PsimConfig psim_default_config(void) {
    PsimConfig c;
    std::memset(&c, 0, sizeof c);
    c.grid_x_log2 = 6;  // kernel.cuh:15-16
    c.grid_y_log2 = 6;
    c.max_particles = 65536;  // kernel.cuh:20
    c.schedule = PSIM_SCHEDULE_REFERENCE;
    c.rebin_every = 0;
    c.device = -1;
    c.use_graph = 0;
    c.slab_rank = 0;
    c.slab_count = 1;
    return c;
}

const char* psim_last_error(const PsimStepper* s) { return s ? s->error.c_str() : g_create_error.c_str(); }

void psim_destroy(PsimStepper* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    // after a failure the exchanges still queued on the stream would wait for peers for ever: abort them
    if (s->comm && s->failed && g_nccl.CommAbort) {
        g_nccl.CommAbort(s->comm);
        s->comm = nullptr;
    }
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
    drop_frame_graphs(s);
    if (s->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s->comm);
    for (auto& side : s->ipc_mapped)
        for (void*& m : side)
            if (m) cudaIpcCloseMemHandle(m);
    cudaFree(s->hdr);
    for (auto& ev : s->timing_events) {
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    for (int k = 0; k < 2; ++k) {
        cudaFree(s->pos[k]);
        cudaFree(s->nbr[k]);
        cudaFree(s->vel[k]);
        cudaFree(s->ty[k]);
        cudaFree(s->outbox[k]);
        cudaFree(s->inbox[k]);
    }
    cudaFree(s->cell_start);
    cudaFree(s->pad_start);
    cudaFree(s->couple_i0);
    cudaFree(s->tile_base);
    cudaFree(s->d_couple_tiles);
    cudaFree(s->tiles_c);
    cudaFree(s->cell_count);
    cudaFree(s->block_sum);
    cudaFree(s->rank_in_cell);
    cudaFree(s->perm);
    cudaFree(s->cell_id);
    cudaFree(s->tiles);
    cudaFree(s->staging);
    cudaFree(s->staging_async);
    if (s->h2d_stream) cudaStreamDestroy(s->h2d_stream);
    if (s->staged_ready) cudaEventDestroy(s->staged_ready);
    if (s->pace_event) cudaEventDestroy(s->pace_event);
    cudaFree(s->snapshot[0]);
    cudaFree(s->snapshot[1]);
    cudaFree(s->mig_counters);
    cudaFree(s->mig_blk_cnt);
    cudaFree(s->mig_blk_off);
    cudaFree(s->mig_blk_list);
    cudaFree(s->d_flags);
    cudaFree(s->d_counts);
    if (s->h_counts) cudaFreeHost(s->h_counts);
    for (int k = 0; k < 2; ++k) {
        if (s->snapshot_ready[k]) cudaEventDestroy(s->snapshot_ready[k]);
        if (s->snapshot_consumed[k]) cudaEventDestroy(s->snapshot_consumed[k]);
    }
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    delete s;
}

int psim_create(const PsimConfig* config, PsimStepper** out) {
    PsimStepper* s = nullptr;  // for CK / fail before the object exists
    if (!config || !out) return fail(s, PSIM_EINVAL, "psim_create: null argument");
    *out = nullptr;
    if (config->grid_x_log2 > 15 || config->grid_y_log2 > 15 || config->grid_x_log2 + config->grid_y_log2 > 28)
        return fail(s, PSIM_EINVAL, "psim_create: grid 2^%u x 2^%u is out of range", config->grid_x_log2,
                    config->grid_y_log2);
    // separations inside the 3x3 stencil must fit a signed 32-bit fixed-point difference
    if (config->grid_x_log2 < 3 || config->grid_y_log2 < 3)
        return fail(s, PSIM_EINVAL, "psim_create: the grid needs at least 8 cells per axis");
    if (config->max_particles == 0 || config->max_particles > 0x7FFFFF00u)
        return fail(s, PSIM_EINVAL, "psim_create: max_particles out of range");
    if (config->schedule > PSIM_SCHEDULE_NATIVE) return fail(s, PSIM_EINVAL, "psim_create: unknown schedule");
    if (config->species_physics && config->slab_count > 1)
        return fail(s, PSIM_EINVAL, "psim_create: species_physics is not available with slabs (ghost rows carry positions only)");
    const uint32_t nranks = config->slab_count ? config->slab_count : 1;
    const uint32_t rows_global = 1u << config->grid_y_log2;
    if (config->slab_rank >= nranks) return fail(s, PSIM_EINVAL, "psim_create: slab_rank %u of %u", config->slab_rank, nranks);
    // Rows of this slab and of the two adjacent ones: PsimConfig.slab_bounds, or equal shares of the rows.
    uint32_t bounds[4];
    std::memcpy(bounds, config->slab_bounds, sizeof bounds);
    if (bounds[2] == 0) {
        if (rows_global % nranks != 0 || rows_global / nranks < 2)
            return fail(s, PSIM_EINVAL, "psim_create: %u cell rows do not split into %u slabs of at least 2 rows "
                        "(or give PsimConfig.slab_bounds)", rows_global, nranks);
        const uint32_t per = rows_global / nranks, r = config->slab_rank;
        bounds[0] = r > 0 ? (r - 1) * per : 0;
        bounds[1] = r * per;
        bounds[2] = (r + 1) * per;
        bounds[3] = r + 1 < nranks ? (r + 2) * per : rows_global;
    }
    {
        const bool first = config->slab_rank == 0, last = config->slab_rank + 1 == nranks;
        const bool ok = bounds[0] <= bounds[1] && bounds[1] + 2 <= bounds[2] && bounds[2] <= bounds[3] && bounds[3] <= rows_global &&
                        (first ? bounds[0] == 0 && bounds[1] == 0 : bounds[0] + 2 <= bounds[1]) &&
                        (last ? bounds[2] == rows_global && bounds[3] == rows_global : bounds[2] + 2 <= bounds[3]);
        if (!ok)
            return fail(s, PSIM_EINVAL, "psim_create: slab_bounds {%u, %u, %u, %u} of slab %u of %u: every slab owns at least 2 "
                        "cell rows, the first starts at row 0, the last ends at row %u", bounds[0], bounds[1], bounds[2], bounds[3],
                        config->slab_rank, nranks, rows_global);
    }

    int device = config->device;
    int ndev = 0;
    cudaError_t err = cudaGetDeviceCount(&ndev);
    if (err != cudaSuccess || ndev == 0)
        return fail(s, PSIM_ECUDA, "psim_create: no CUDA device (%s); this library has no CPU path",
                    err == cudaSuccess ? "device count is 0" : cudaGetErrorString(err));
    if (device < 0) CK(cudaGetDevice(&device));
    if (device >= ndev) return fail(s, PSIM_EINVAL, "psim_create: device %d of %d", device, ndev);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(s, PSIM_ECUDA, "psim_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);

    PsimStepper* st = new PsimStepper;
    st->cfg = *config;
    st->cfg.slab_count = nranks;
    if (st->cfg.rebin_every == 0) st->cfg.rebin_every = 17;
    st->device = device;
    if (const char* env = getenv("PSIM_FORCE_INT_PATH")) st->force_int_path = env[0] == '1';
    if (const char* env = getenv("PSIM_PDL")) st->pdl = env[0] == '1';
    st->copy_piece_bytes = nranks > 1 ? 1 : 0;  // 1: a sixth of each copy (PacedCopy)
    if (const char* env = getenv("PSIM_COPY_PIECE_MB")) st->copy_piece_bytes = (size_t)std::atoi(env) << 20;
    st->copy_pace_steps = nranks > 1;  // PSIM_COPY_PACE=rebin: a sixth behind every re-bin instead
    if (const char* env = getenv("PSIM_COPY_PACE")) st->copy_pace_steps = nranks > 1 && std::strcmp(env, "rebin") != 0;
    st->rank = (int)config->slab_rank;
    st->nranks = (int)nranks;
    Grid& g = st->grid;
    g.lx = config->grid_x_log2;
    g.bx = 1u << g.lx;
    g.sx = 32 - g.lx;
    g.sy = 32 - config->grid_y_log2;
    std::memcpy(st->cfg.slab_bounds, bounds, sizeof bounds);
    g.own_rows = bounds[2] - bounds[1];
    g.rows_below = bounds[1] - bounds[0];
    g.rows_above = bounds[3] - bounds[2];
    g.own_row0 = st->rank > 0 ? 1 : 0;
    g.by = g.own_rows + g.own_row0 + (st->rank + 1 < st->nranks ? 1 : 0);
    g.row_offset = (int32_t)bounds[1] - (int32_t)g.own_row0;
    g.cells = g.bx * g.by;

    const size_t cap = config->max_particles;
    if (nranks > 1) {
        // A ghost row holds one cell row of a neighbour: by default 4 times what a row holds on average when the
        // thinnest of the three slabs involved is full, and at least 4096. The migrant boxes are the same size on
        // every slab (they are exchanged): 4 times the mean row of a slab of average height.
        auto rows_of = [&](uint32_t rows) { return (uint32_t)std::min<size_t>(cap, std::max<size_t>(4096, 4 * (cap / rows + 1))); };
        uint32_t thinnest = g.own_rows;
        if (g.rows_below) thinnest = std::min(thinnest, g.rows_below);
        if (g.rows_above) thinnest = std::min(thinnest, g.rows_above);
        st->ghost_cap = config->ghost_capacity ? config->ghost_capacity : rows_of(thinnest);
        st->box_capacity = config->migrant_capacity ? config->migrant_capacity : rows_of(std::max(rows_global / nranks, 1u));
    }
    st->cap_total = (uint32_t)std::min<size_t>(cap + 2 * (size_t)st->ghost_cap, 0x7FFFFF00u);
    st->box_bytes = sizeof(MigrantBoxHeader) + sizeof(Particle) * (size_t)st->box_capacity;
    st->ingest_cap = std::max<uint32_t>(config->ingest_capacity, config->max_particles);
    const size_t cap_total = st->cap_total;
    const size_t cand_cap = std::max<size_t>(st->ingest_cap, cap_total + 2 * (size_t)st->box_capacity);

    s = st;  // from here on failures are recorded on the object (and it is destroyed before returning)
#define CKC(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t err__ = (call);                                                                      \
        if (err__ != cudaSuccess) {                                                                      \
            fail(nullptr, PSIM_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, \
                 __LINE__);                                                                              \
            psim_destroy(st);                                                                            \
            return PSIM_ECUDA;                                                                           \
        }                                                                                                \
    } while (0)
    CKC(cudaStreamCreateWithFlags(&st->own_stream, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&st->copy_stream, cudaStreamNonBlocking));
    st->stream = st->own_stream;
    st->nsnap = config->snapshot_buffers >= 2 ? 2 : 1;
    for (int k = 0; k < st->nsnap; ++k) {
        CKC(cudaEventCreateWithFlags(&st->snapshot_ready[k], cudaEventDisableTiming));
        CKC(cudaEventCreateWithFlags(&st->snapshot_consumed[k], cudaEventDisableTiming));
    }
    for (int k = 0; k < 2; ++k) {
        CKC(cudaMalloc(&st->pos[k], sizeof(uint2) * (cap_total + kPadParticles)));
        CKC(cudaMemset(st->pos[k], 0, sizeof(uint2) * (cap_total + kPadParticles)));
        CKC(cudaMalloc(&st->vel[k], sizeof(float2) * cap_total));
        CKC(cudaMalloc(&st->ty[k], sizeof(int32_t) * cap_total));
        if (nranks > 1) {
            CKC(cudaMalloc(&st->outbox[k], st->box_bytes));
            CKC(cudaMalloc(&st->inbox[k], st->box_bytes));
        }
    }
    if (nranks > 1) {
        const size_t blocks = 2 * ((size_t)div_up((uint32_t)cap, kMigBlock) + 1);
        CKC(cudaMalloc(&st->mig_blk_cnt, sizeof(uint32_t) * blocks));
        CKC(cudaMemset(st->mig_blk_cnt, 0, sizeof(uint32_t) * blocks));
        CKC(cudaMalloc(&st->mig_blk_off, sizeof(uint32_t) * blocks));
        CKC(cudaMalloc(&st->mig_blk_list, sizeof(uint32_t) * (blocks / 2 + 1)));
    }
    CKC(cudaMalloc(&st->cell_start, sizeof(uint32_t) * ((size_t)g.cells + 1 + kPadCells)));
    // step_kernel_c needs cells of at most 2^22 fixed-point units (exact fp32 offsets within a zone / a tile's rows)
    st->float_grid = g.sx <= 22 && g.sy <= 22;
    if (st->float_grid) {
        // couples: every particle, plus one half-empty couple per cell at most
        const size_t couples = (cap_total + std::min<size_t>(cap_total, g.cells)) / 2 + 1;
        // per run of blocks of a row (row_cut_kernel): at most ceil(couples / 128) + its blocks; at most one run per block
        st->tiles_c_cap = (uint32_t)((cap + std::min<size_t>(cap, (size_t)g.own_rows * g.bx)) / 2 / kCouples + 1 +
                                     (size_t)g.own_rows * (2 * (g.bx / kTileCols) + 1));
        CKC(cudaMalloc(&st->pad_start, sizeof(uint32_t) * ((size_t)g.cells + 1 + kPadCells)));
        CKC(cudaMemset(st->pad_start, 0, sizeof(uint32_t) * ((size_t)g.cells + 1 + kPadCells)));
        CKC(cudaMalloc(&st->couple_i0, sizeof(uint2) * couples));
        CKC(cudaMalloc(&st->tile_base, sizeof(uint32_t) * ((size_t)g.own_rows + 1)));
        CKC(cudaMalloc(&st->d_couple_tiles, 2 * sizeof(uint32_t)));  // [0] tiles of the last binning, [1] sticky: they did not fit
        CKC(cudaMemset(st->d_couple_tiles, 0, 2 * sizeof(uint32_t)));
        CKC(cudaMalloc(&st->tiles_c, sizeof(TileC) * (size_t)st->tiles_c_cap));
        for (int k = 0; k < 2; ++k) {
            CKC(cudaMalloc(&st->nbr[k], sizeof(float4) * (cap_total + kPadParticles)));
            CKC(cudaMemset(st->nbr[k], 0, sizeof(float4) * (cap_total + kPadParticles)));
        }
    }
    CKC(cudaMalloc(&st->cell_count, sizeof(uint32_t) * (size_t)g.cells));
    CKC(cudaMalloc(&st->block_sum, sizeof(uint2) * (size_t)div_up(g.cells, kScanBlock)));
    CKC(cudaMalloc(&st->rank_in_cell, sizeof(uint32_t) * cand_cap));
    CKC(cudaMalloc(&st->perm, sizeof(uint32_t) * cap_total));
    CKC(cudaMalloc(&st->cell_id, sizeof(uint32_t) * cap_total));
    CKC(cudaMalloc(&st->tiles, sizeof(TileDesc) * ((size_t)div_up((uint32_t)cap, kTile) + 1)));
    CKC(cudaMalloc(&st->staging, sizeof(Particle) * (size_t)st->ingest_cap));
    for (int k = 0; k < st->nsnap; ++k) CKC(cudaMalloc(&st->snapshot[k], sizeof(Particle) * cap));
    if (nranks > 1) {
        CKC(cudaMalloc(&st->hdr, sizeof(HaloHeader)));
        CKC(cudaMemset(st->hdr, 0, sizeof(HaloHeader)));
    }
    CKC(cudaMalloc(&st->mig_counters, 2 * sizeof(uint32_t)));
    CKC(cudaMalloc(&st->d_flags, sizeof(uint32_t)));
    CKC(cudaMalloc(&st->d_counts, 16 * sizeof(uint32_t)));
    CKC(cudaHostAlloc(&st->h_counts, 16 * sizeof(uint32_t), cudaHostAllocMapped));
    std::memset(st->h_counts, 0, 16 * sizeof(uint32_t));
    CKC(cudaHostGetDevicePointer(reinterpret_cast<void**>(&st->h_counts_dev), st->h_counts, 0));
    CKC(cudaMemset(st->cell_start, 0, sizeof(uint32_t) * ((size_t)g.cells + 1 + kPadCells)));
    CKC(cudaMemset(st->d_flags, 0, sizeof(uint32_t)));
#undef CKC
    *out = st;
    return PSIM_OK;
}

int psim_set_stream(PsimStepper* s, void* cuda_stream) {
    if (!s) return PSIM_EINVAL;
    int rc = check_lone(s, "psim_set_stream");
    if (rc) return rc;
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    drop_frame_graphs(s);
    s->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : s->own_stream;
    return PSIM_OK;
}

int psim_comm_unique_id(void* out128) {
    PsimStepper* s = nullptr;
    if (!out128) return fail(s, PSIM_EINVAL, "psim_comm_unique_id: null argument");
    if (!load_nccl(nullptr)) return PSIM_ENCCL;
    ncclUniqueId id;
    CKN(g_nccl.GetUniqueId(&id));
    std::memcpy(out128, &id, sizeof id);
    return PSIM_OK;
}

// What a slab hands its neighbours so that their step kernels can write its ghost rows: CUDA IPC handles of
// its two position buffers and of its HaloHeader.
struct HaloExport {
    cudaIpcMemHandle_t pos[2];
    cudaIpcMemHandle_t hdr;
    cudaIpcMemHandle_t nbr[2];  // fine grids only
    uint32_t valid;
    uint32_t has_nbr;
    uint32_t row_begin, row_end;  // the exporter's own rows: adjacent slabs must meet
};

// Map the neighbours' buffers (one process per slab). Collective over the communicator. Any failure on any
// rank leaves every rank on the send/recv halo exchange: the decision is an all-reduce.
int connect_peers(PsimStepper* s) {
    HaloExport mine;
    std::memset(&mine, 0, sizeof mine);
    const char* mode = getenv("PSIM_HALO");
    int ok = !(mode && std::strcmp(mode, "nccl") == 0);
    if (ok) {
        ok = cudaIpcGetMemHandle(&mine.pos[0], s->pos[0]) == cudaSuccess &&
             cudaIpcGetMemHandle(&mine.pos[1], s->pos[1]) == cudaSuccess &&
             cudaIpcGetMemHandle(&mine.hdr, s->hdr) == cudaSuccess;
        if (ok && s->nbr[0]) {
            ok = cudaIpcGetMemHandle(&mine.nbr[0], s->nbr[0]) == cudaSuccess &&
                 cudaIpcGetMemHandle(&mine.nbr[1], s->nbr[1]) == cudaSuccess;
            mine.has_nbr = 1;
        }
        cudaGetLastError();
    }
    mine.valid = ok ? 1u : 0u;
    mine.row_begin = s->cfg.slab_bounds[1];
    mine.row_end = s->cfg.slab_bounds[2];
    HaloExport* d_io = nullptr;  // [0] mine, [1] from the lower, [2] from the upper neighbour
    int* d_ok = nullptr;
    CK(cudaMalloc(&d_io, 3 * sizeof(HaloExport)));
    CK(cudaMalloc(&d_ok, sizeof(int)));
    CK(cudaMemsetAsync(d_io, 0, 3 * sizeof(HaloExport), s->stream));
    CK(cudaMemcpyAsync(d_io, &mine, sizeof mine, cudaMemcpyHostToDevice, s->stream));
    XferOp op;
    for (int side = 0; side < 2; ++side) {
        if (side == 0 ? !(s->rank > 0) : !(s->rank + 1 < s->nranks)) continue;
        op.send[side] = d_io;
        op.recv[side] = d_io + 1 + side;
        op.send_bytes[side] = op.recv_bytes[side] = sizeof(HaloExport);
    }
    int rc = exchange_nccl(s, op, s->stream);
    if (rc) return rc;
    HaloExport got[3];
    CK(cudaMemcpyAsync(got, d_io, sizeof got, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    bool rows_meet = true;
    for (int side = 0; side < 2; ++side) {
        if (side == 0 ? !(s->rank > 0) : !(s->rank + 1 < s->nranks)) continue;
        const HaloExport& e = got[1 + side];
        rows_meet = rows_meet && (side == 0 ? e.row_end == s->cfg.slab_bounds[1] && e.row_begin == s->cfg.slab_bounds[0]
                                            : e.row_begin == s->cfg.slab_bounds[2] && e.row_end == s->cfg.slab_bounds[3]);
    }
    for (int side = 0; side < 2 && ok; ++side) {
        if (side == 0 ? !(s->rank > 0) : !(s->rank + 1 < s->nranks)) continue;
        const HaloExport& e = got[1 + side];
        if (!e.valid) {
            ok = 0;
            break;
        }
        const cudaIpcMemHandle_t* handles[5] = {&e.pos[0], &e.pos[1], &e.hdr, &e.nbr[0], &e.nbr[1]};
        for (int k = 0; k < (e.has_nbr ? 5 : 3) && ok; ++k) {
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, *handles[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
            } else {
                s->ipc_mapped[side][k] = p;
            }
        }
    }
    // everybody pushes or nobody does (-1: somebody's rows do not meet its neighbour's, nobody may run)
    if (!rows_meet) ok = -1;
    CK(cudaMemcpyAsync(d_ok, &ok, sizeof ok, cudaMemcpyHostToDevice, s->stream));
    CKN(g_nccl.AllReduce(d_ok, d_ok, 1, kNcclInt32, kNcclMin, s->comm, s->stream));
    CK(cudaMemcpyAsync(&ok, d_ok, sizeof ok, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    cudaFree(d_io);
    cudaFree(d_ok);
    if (ok < 0)
        return fail(s, PSIM_EINVAL, "psim_comm_init: the slab_bounds of adjacent ranks do not meet (slab %d owns rows [%u, %u))",
                    s->rank, s->cfg.slab_bounds[1], s->cfg.slab_bounds[2]);
    s->push = ok != 0;
    if (s->push) {
        for (int side = 0; side < 2; ++side) {
            s->peer_pos[side][0] = static_cast<uint2*>(s->ipc_mapped[side][0]);
            s->peer_pos[side][1] = static_cast<uint2*>(s->ipc_mapped[side][1]);
            s->peer_hdr[side] = static_cast<HaloHeader*>(s->ipc_mapped[side][2]);
            s->peer_nbr[side][0] = static_cast<float4*>(s->ipc_mapped[side][3]);
            s->peer_nbr[side][1] = static_cast<float4*>(s->ipc_mapped[side][4]);
        }
    }
    return PSIM_OK;
}

int psim_comm_init(PsimStepper* s, const void* unique_id128) {
    if (!s || !unique_id128) return PSIM_EINVAL;
    int rc = check_lone(s, "psim_comm_init");
    if (rc) return rc;
    if (s->nranks == 1) return PSIM_OK;  // a single slab never communicates
    if (s->comm) return fail(s, PSIM_ESTATE, "psim_comm_init: already initialised");
    if (!load_nccl(s)) return PSIM_ENCCL;
    CK(cudaSetDevice(s->device));
    ncclUniqueId id;
    std::memcpy(&id, unique_id128, sizeof id);
    CKN(g_nccl.CommInitRank(&s->comm, s->nranks, id, s->rank));
    return connect_peers(s);
}

int psim_halo_mode(const PsimStepper* s) { return s ? (s->nranks == 1 ? 0 : (s->push ? 2 : 1)) : 0; }

}  // extern "C"

namespace {
// Scene ingest for DataStructure::CompactArray = frame_compact_into (kernel.cuh:207-209): the live records in input
// order, no binning. The null records are dropped on the host, as the reference does.
int upload_compact(PsimStepper* s, const FrameHeader* frame) {
    if (s->nranks > 1 || s->group)
        return fail(s, PSIM_EINVAL, "DataStructure::CompactArray (all pairs) cannot be decomposed into slabs");
    std::vector<Particle> live;
    const Particle* rec = frame->particles;
    uint32_t n = frame->particle_count;
    bool has_null = false;
    for (uint32_t i = 0; i < n && !has_null; ++i) has_null = rec[i].ty < 0;
    if (has_null) {
        live.reserve(n);
        for (uint32_t i = 0; i < n; ++i)
            if (rec[i].ty >= 0) live.push_back(rec[i]);
        rec = live.data();
        n = (uint32_t)live.size();
    }
    if (n > s->cfg.max_particles)
        return fail(s, PSIM_ECAPACITY, "%u live particles exceed max_particles = %u", n, s->cfg.max_particles);
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaStreamSynchronize(s->copy_stream));
    apply_metadata(s, frame->metadata);
    s->compact_mode = true;
    s->cur_pos = s->cur_vel = s->cur_ty = 0;
    if (n) {
        CK(cudaMemcpyAsync(s->staging, rec, sizeof(Particle) * (size_t)n, cudaMemcpyHostToDevice, s->stream));
        unpack_kernel<<<div_up(n, 256), 256, 0, s->stream>>>(s->staging, n, s->pos[0], s->vel[0], s->ty[0], s->cell_id);
        s->launches += 1;
        CK(cudaGetLastError());
    }
    s->n = s->n_total = s->own_hi = s->b_hi_start = n;
    s->own_lo = s->b_lo_end = 0;
    s->has_scene = true;
    s->fresh_scene = true;
    s->native_countdown = 0;
    int rc = enqueue_snapshot(s);
    if (rc) return rc;
    CK(cudaStreamSynchronize(s->stream));  // `live` and the caller's frame are free again
    return PSIM_OK;
}
}  // namespace

extern "C" {

int psim_upload_frame(PsimStepper* s, const FrameHeader* frame) {
    if (!s || !frame) return PSIM_EINVAL;
    int rc = check_lone(s, "psim_upload_frame");
    if (rc) return rc;
    CK(cudaSetDevice(s->device));
    if (frame->metadata.data_structure == 0 /* DataStructure::CompactArray */) {
        if (frame->particle_count > s->ingest_cap)
            return fail(s, PSIM_ECAPACITY, "frame holds %u particles, the ingest buffer %u", frame->particle_count, s->ingest_cap);
        return upload_compact(s, frame);
    }
    if (frame->particle_count > s->ingest_cap)
        return fail(s, PSIM_ECAPACITY, "frame holds %u particles, the ingest buffer %u (max_particles / ingest_capacity)",
                    frame->particle_count, s->ingest_cap);
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaStreamSynchronize(s->copy_stream));
    apply_metadata(s, frame->metadata);
    if (frame->particle_count)
        CK(cudaMemcpyAsync(s->staging, frame->particles, sizeof(Particle) * (size_t)frame->particle_count,
                           cudaMemcpyHostToDevice, s->stream));
    return team_ingest(lone(&s), s->staging, frame->particle_count);
}

int psim_upload_device(PsimStepper* s, const FrameMetadata* meta, const void* d_particles, uint32_t count) {
    if (!s || !meta || (count && !d_particles)) return PSIM_EINVAL;
    int rc = check_lone(s, "psim_upload_device");
    if (rc) return rc;
    CK(cudaSetDevice(s->device));
    if (meta->data_structure == 0)
        return fail(s, PSIM_EINVAL, "psim_upload_device: DataStructure::CompactArray scenes are uploaded from the host");
    if (count > s->ingest_cap)
        return fail(s, PSIM_ECAPACITY, "%u particles, the ingest buffer holds %u", count, s->ingest_cap);
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaStreamSynchronize(s->copy_stream));
    apply_metadata(s, *meta);
    if (count)
        CK(cudaMemcpyAsync(s->staging, d_particles, sizeof(Particle) * (size_t)count, cudaMemcpyDeviceToDevice,
                           s->stream));
    return team_ingest(lone(&s), s->staging, count);
}

int psim_stage_frame_async(PsimStepper* s, const FrameHeader* frame) {
    if (!s || !frame) return PSIM_EINVAL;
    int rc = check_lone(s, "psim_stage_frame_async");
    if (rc) return rc;
    CK(cudaSetDevice(s->device));
    if (frame->metadata.data_structure == 0)
        return fail(s, PSIM_EINVAL, "psim_stage_frame_async: DataStructure::CompactArray scenes use psim_upload_frame");
    if (frame->particle_count > s->ingest_cap)
        return fail(s, PSIM_ECAPACITY, "frame holds %u particles, the ingest buffer %u (max_particles / ingest_capacity)",
                    frame->particle_count, s->ingest_cap);
    if (!s->staging_async) {
        CK(cudaMalloc(&s->staging_async, sizeof(Particle) * (size_t)s->ingest_cap));
        CK(cudaStreamCreateWithFlags(&s->h2d_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&s->staged_ready, cudaEventDisableTiming));
    }
    // the buffer is free: the ingest that read it last returned only after its kernels had finished
    if (s->paced_up.active) return fail(s, PSIM_ESTATE, "psim_stage_frame_async: a staged frame is still on its way");
    start_copy(s, s->paced_up, s->staging_async, frame->particles, sizeof(Particle) * (size_t)frame->particle_count,
               cudaMemcpyHostToDevice, s->h2d_stream, s->staged_ready);
    // whole at once for a single slab; a slab sends a piece behind each re-bin of the frame that runs meanwhile
    if (s->copy_piece_bytes == 0 && (rc = pump_copy(s, s->paced_up, true))) return rc;
    s->staged_meta = frame->metadata;
    s->staged_count = frame->particle_count;
    s->has_staged = true;
    return PSIM_OK;
}

int psim_upload_staged(PsimStepper* s) {
    if (!s) return PSIM_EINVAL;
    int rc = check_lone(s, "psim_upload_staged");
    if (rc) return rc;
    if (!s->has_staged) return fail(s, PSIM_ESTATE, "psim_upload_staged: no frame has been staged");
    CK(cudaSetDevice(s->device));
    if ((rc = pump_copy(s, s->paced_up, true))) return rc;  // whatever the re-bins have not sent yet
    CK(cudaStreamWaitEvent(s->stream, s->staged_ready, 0));
    s->has_staged = false;
    std::swap(s->staging, s->staging_async);  // the next psim_stage_frame_async fills the other buffer
    apply_metadata(s, s->staged_meta);
    return team_ingest(lone(&s), s->staging, s->staged_count);
}

int psim_set_metadata(PsimStepper* s, const FrameMetadata* meta) {
    if (!s || !meta) return PSIM_EINVAL;
    // The reference switches between its two layouts on every frame's metadata (kernel.cuh:143-150); here the layout is
    // the one the scene was uploaded in (the all-pairs mode keeps the input order, the grid mode is cell-sorted), so a
    // header-only update that flips data_structure is refused rather than silently ignored: upload the scene again.
    if (s->has_scene && (meta->data_structure == 0) != s->compact_mode)
        return fail(s, PSIM_EINVAL, "psim_set_metadata: data_structure %u does not match the layout the scene was uploaded in (%s); "
                    "upload the scene again to switch", meta->data_structure, s->compact_mode ? "CompactArray" : "MatrixBuckets");
    apply_metadata(s, *meta);  // captured by value at the next enqueue, like kernel_bucket.cuh:121
    return PSIM_OK;
}

int psim_get_metadata(const PsimStepper* s, FrameMetadata* out) {
    if (!s || !out) return PSIM_EINVAL;
    *out = s->meta;
    return PSIM_OK;
}

int psim_step_async(PsimStepper* s, uint32_t steps) {
    if (!s) return PSIM_EINVAL;
    int rc = check_lone(s, "psim_step_async");
    if (rc) return rc;
    if (!s->has_scene) return fail(s, PSIM_ESTATE, "psim_step_async: no scene uploaded");
    CK(cudaSetDevice(s->device));
    for (uint32_t k = 0; k < steps; ++k) {
        if ((rc = team_step(lone(&s)))) return rc;
        s->fresh_scene = false;
    }
    return PSIM_OK;
}

int psim_rebin_async(PsimStepper* s) {
    if (!s) return PSIM_EINVAL;
    int rc = check_lone(s, "psim_rebin_async");
    if (rc) return rc;
    if (!s->has_scene) return fail(s, PSIM_ESTATE, "psim_rebin_async: no scene uploaded");
    CK(cudaSetDevice(s->device));
    s->fresh_scene = false;
    return team_bin(lone(&s), false, nullptr, 0);
}

int psim_set_snapshot_stride(PsimStepper* s, uint32_t stride) {
    if (!s) return PSIM_EINVAL;
    if (stride == 0) return fail(s, PSIM_EINVAL, "psim_set_snapshot_stride: stride must be >= 1");
    s->snapshot_stride = stride;
    return PSIM_OK;
}

int psim_snapshot_async(PsimStepper* s) {
    if (!s) return PSIM_EINVAL;
    if (!s->has_scene) return fail(s, PSIM_ESTATE, "psim_snapshot_async: no scene uploaded");
    CK(cudaSetDevice(s->device));
    return enqueue_snapshot(s);
}

int psim_run_frame_async(PsimStepper* s) {
    if (!s) return PSIM_EINVAL;
    int rc = check_lone(s, "psim_run_frame_async");
    if (rc) return rc;
    if (!s->has_scene) return fail(s, PSIM_ESTATE, "psim_run_frame_async: no scene uploaded");
    CK(cudaSetDevice(s->device));
    if (frame_graph_usable(s)) {
        rc = run_frame_with_graph(s);
        if (!rc) rc = enqueue_snapshot(s);
    } else {
        rc = team_run_frame(lone(&s));
    }
    s->fresh_scene = false;
    return rc;
}

int psim_sync(PsimStepper* s) {
    if (!s) return PSIM_EINVAL;
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    if (s->push && s->hdr) {  // a step that gave up waiting for a neighbour's halo says so here
        uint32_t halo_error = 0;
        CK(cudaMemcpy(&halo_error, &s->hdr->error, sizeof halo_error, cudaMemcpyDeviceToHost));
        if (halo_error)
            return fail(s, PSIM_ECUDA, "slab %d: a step waited 20 s for a neighbour's halo (did a neighbour rank die?)", s->rank);
    }
    int rc = refresh_tile_count(s);
    if (rc) return rc;
    if (s->timing) return collect_timing(s);
    return PSIM_OK;
}

int psim_download_frame_ex(PsimStepper* s, uint32_t age, FrameHeader* dst) {
    if (!s || !dst) return PSIM_EINVAL;
    const int k = snapshot_index(s, age);
    if (k < 0)
        return fail(s, PSIM_ESTATE, "psim_download_frame: no snapshot of age %u (%llu packed so far, %d buffer%s)", age,
                    (unsigned long long)s->snaps_taken, s->nsnap, s->nsnap > 1 ? "s" : "");
    CK(cudaSetDevice(s->device));
    if (dst->particle_count < s->snapshot_n[k])
        return fail(s, PSIM_ECAPACITY, "psim_download_frame: destination holds %u particles, snapshot has %u",
                    dst->particle_count, s->snapshot_n[k]);
    int rc = download_records(s, k, dst->particles);
    if (rc) return rc;
    write_header(dst, s->snapshot_meta[k], s->snapshot_n[k]);
    return PSIM_OK;
}

int psim_download_frame(PsimStepper* s, FrameHeader* dst) { return psim_download_frame_ex(s, 0, dst); }

int psim_download_frame_begin(PsimStepper* s, uint32_t age, FrameHeader* dst) {
    if (!s || !dst) return PSIM_EINVAL;
    if (s->pending_dst) return fail(s, PSIM_ESTATE, "psim_download_frame_begin: a download is already in flight");
    const int k = snapshot_index(s, age);
    if (k < 0)
        return fail(s, PSIM_ESTATE, "psim_download_frame_begin: no snapshot of age %u (%llu packed so far, %d buffer%s)",
                    age, (unsigned long long)s->snaps_taken, s->nsnap, s->nsnap > 1 ? "s" : "");
    CK(cudaSetDevice(s->device));
    if (dst->particle_count < s->snapshot_n[k])
        return fail(s, PSIM_ECAPACITY, "psim_download_frame_begin: destination holds %u particles, snapshot has %u",
                    dst->particle_count, s->snapshot_n[k]);
    CK(cudaStreamWaitEvent(s->copy_stream, s->snapshot_ready[k], 0));
    start_copy(s, s->paced_down, dst->particles, s->snapshot[k], sizeof(Particle) * (size_t)s->snapshot_n[k],
               cudaMemcpyDeviceToHost, s->copy_stream, s->snapshot_consumed[k]);
    // whole at once for a single slab; a slab sends nothing yet (the ingest of the next frame is about to make its round
    // trip) and a piece behind each re-bin of the frame that follows
    if (s->copy_piece_bytes == 0) {
        int rc = pump_copy(s, s->paced_down, true);
        if (rc) return rc;
    }
    write_header(dst, s->snapshot_meta[k], s->snapshot_n[k]);  // the header is host data: valid at once
    s->pending_dst = dst;
    s->pending_k = k;
    return PSIM_OK;
}

int psim_download_frame_end(PsimStepper* s) {
    if (!s) return PSIM_EINVAL;
    if (!s->pending_dst) return fail(s, PSIM_ESTATE, "psim_download_frame_end: no download in flight");
    CK(cudaSetDevice(s->device));
    s->pending_dst = nullptr;
    int rc = pump_copy(s, s->paced_down, true);
    if (rc) return rc;
    if (s->push && s->hdr)
        CK(cudaMemcpyAsync(&s->h_counts[15], &s->hdr->error, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->copy_stream));
    CK(cudaStreamSynchronize(s->copy_stream));
    return check_halo_error(s);
}

void* psim_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void psim_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

uint32_t psim_particle_count(const PsimStepper* s) { return s ? s->n : 0; }
uint64_t psim_steps_executed(const PsimStepper* s) { return s ? s->steps_executed : 0; }
uint64_t psim_rebins_executed(const PsimStepper* s) { return s ? s->rebins_executed : 0; }
uint64_t psim_kernel_launches(const PsimStepper* s) { return s ? s->launches : 0; }
uint32_t psim_cell_count(const PsimStepper* s) { return s ? s->grid.cells : 0; }
uint64_t psim_migrants_sent(const PsimStepper* s) { return s ? s->migrants_sent : 0; }

int psim_get_cell_start(PsimStepper* s, uint32_t* out) {
    if (!s || !out) return PSIM_EINVAL;
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaMemcpy(out, s->cell_start, sizeof(uint32_t) * ((size_t)s->grid.cells + 1), cudaMemcpyDeviceToHost));
    return PSIM_OK;
}

int psim_enable_step_timing(PsimStepper* s, int enable) {
    if (!s) return PSIM_EINVAL;
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    int rc = collect_timing(s);
    if (rc) return rc;
    s->timing = enable != 0;
    s->timing_total_ms = 0;
    s->timing_launches = 0;
    return PSIM_OK;
}

int psim_get_step_timing(PsimStepper* s, double* total_ms, uint64_t* launches) {
    if (!s) return PSIM_EINVAL;
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    int rc = collect_timing(s);
    if (rc) return rc;
    if (total_ms) *total_ms = s->timing_total_ms;
    if (launches) *launches = s->timing_launches;
    return PSIM_OK;
}

int psim_tile_stats(PsimStepper* s, PsimTileStats* out) {
    if (!s || !out) return PSIM_EINVAL;
    std::memset(out, 0, sizeof *out);
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    if (s->compact_mode) {  // all pairs: plain tiles of 128 particles, nothing staged per tile
        out->tiles = div_up(s->n, kTile);
        out->threads_live = s->n;
        out->threads_launched = (uint64_t)out->tiles * kTile;
        return PSIM_OK;
    }
    out->float_path = s->float_path ? 1u : 0u;
    int rc = refresh_tile_count(s);
    if (rc) return rc;
    if (s->float_path) {
        std::vector<TileC> t(s->n_tiles_c);
        if (!t.empty()) CK(cudaMemcpy(t.data(), s->tiles_c, sizeof(TileC) * t.size(), cudaMemcpyDeviceToHost));
        out->tiles = s->n_tiles_c;
        for (const TileC& x : t) {
            out->tiles_staged += x.fits ? 1u : 0u;
            out->threads_live += x.nk;
            out->max_columns = std::max(out->max_columns, x.cs_cnt[1]);
            for (int d = 0; d < 3; ++d) out->max_row_particles = std::max(out->max_row_particles, x.p_cnt[d]);
        }
        out->threads_launched = (uint64_t)s->n_tiles_c * kCouples;
    } else {
        std::vector<TileDesc> t(div_up(s->n, kTile));
        if (!t.empty()) CK(cudaMemcpy(t.data(), s->tiles, sizeof(TileDesc) * t.size(), cudaMemcpyDeviceToHost));
        out->tiles = (uint32_t)t.size();
        for (const TileDesc& x : t) {
            out->tiles_staged += x.fits ? 1u : 0u;
            out->max_columns = std::max(out->max_columns, x.last - x.first + 3);
            for (int d = 0; d < 3; ++d) out->max_row_particles = std::max(out->max_row_particles, x.p_cnt[d]);
        }
        out->threads_live = s->n;
        out->threads_launched = (uint64_t)t.size() * kTile;
    }
    return PSIM_OK;
}

int psim_device_state(PsimStepper* s, const void** pos, const void** vel, const void** ty, const void** cell_start) {
    if (!s) return PSIM_EINVAL;
    if (pos) *pos = s->pos[s->cur_pos] + s->own_lo;
    if (vel) *vel = s->vel[s->cur_vel] + s->own_lo;
    if (ty) *ty = s->ty[s->cur_ty] + s->own_lo;
    if (cell_start) *cell_start = s->cell_start;
    return PSIM_OK;
}

int psim_balance_rows_hist(const uint64_t* row_counts, uint32_t grid_y_log2, uint32_t slab_count, uint32_t* bounds) {
    PsimStepper* s = nullptr;
    if (!row_counts || !bounds || slab_count == 0 || grid_y_log2 < 3 || grid_y_log2 > 15 || (2ull * slab_count) > (1ull << grid_y_log2))
        return fail(s, PSIM_EINVAL, "psim_balance_rows: bad argument (every slab needs 2 of the %llu cell rows)",
                    1ull << (grid_y_log2 & 31));
    const uint32_t rows = 1u << grid_y_log2;
    std::vector<uint64_t> upto(rows + 1, 0);  // live particles in rows < r
    for (uint32_t r = 0; r < rows; ++r) upto[r + 1] = upto[r] + row_counts[r];
    const uint64_t live = upto[rows];
    bounds[0] = 0;
    bounds[slab_count] = rows;
    for (uint32_t k = 1; k < slab_count; ++k) {
        // the boundary whose share of the particles is nearest to k / slab_count
        const uint64_t want = live * k / slab_count;
        uint32_t r = (uint32_t)(std::lower_bound(upto.begin(), upto.end(), want) - upto.begin());
        if (r > 0 && want - upto[r - 1] < upto[std::min(r, rows)] - want) r -= 1;
        bounds[k] = std::min(r, rows);
    }
    for (uint32_t k = 1; k < slab_count; ++k) bounds[k] = std::max(bounds[k], bounds[k - 1] + 2);
    for (uint32_t k = slab_count - 1; k >= 1; --k) bounds[k] = std::min(bounds[k], bounds[k + 1] - 2);
    return PSIM_OK;
}

int psim_balance_rows(const FrameHeader* scene, uint32_t grid_y_log2, uint32_t slab_count, uint32_t* bounds) {
    PsimStepper* s = nullptr;
    if (!scene || grid_y_log2 < 3 || grid_y_log2 > 15)
        return fail(s, PSIM_EINVAL, "psim_balance_rows: bad argument");
    std::vector<uint64_t> hist((size_t)1 << grid_y_log2, 0);
    const Particle* p = scene->particles;
    for (uint32_t i = 0; i < scene->particle_count; ++i)
        if (p[i].ty >= 0) hist[p[i].y >> (32 - grid_y_log2)] += 1;
    return psim_balance_rows_hist(hist.data(), grid_y_log2, slab_count, bounds);
}

void psim_slab_bounds_of(const uint32_t* bounds, uint32_t slab_rank, uint32_t slab_count, uint32_t out[4]) {
    out[0] = bounds[slab_rank > 0 ? slab_rank - 1 : 0];
    out[1] = bounds[slab_rank];
    out[2] = bounds[slab_rank + 1];
    out[3] = bounds[std::min(slab_rank + 2, slab_count)];
}

int psim_slab_info(const PsimStepper* s, PsimSlabInfo* out) {
    if (!s || !out) return PSIM_EINVAL;
    std::memset(out, 0, sizeof *out);
    out->slab_rank = (uint32_t)s->rank;
    out->slab_count = (uint32_t)s->nranks;
    out->first_row = s->cfg.slab_bounds[1];
    out->rows = s->grid.own_rows;
    out->local_rows = s->grid.by;
    out->first_local_row = (uint32_t)((int32_t)out->first_row - (int32_t)s->grid.own_row0);
    out->particles = s->n;
    out->ghost_below = s->own_lo;
    out->ghost_above = s->n_total - s->own_hi;
    out->ghost_capacity = s->ghost_cap;
    out->migrant_capacity = s->box_capacity;
    return PSIM_OK;
}

// ---- slabs of one process on one device ---------------------------------------------------------

const char* psim_group_last_error(const PsimGroup* g) { return g ? g->error.c_str() : g_create_error.c_str(); }

int psim_group_create(PsimStepper* const* steppers, uint32_t count, PsimGroup** out) {
    PsimStepper* s = nullptr;
    if (!steppers || !out || count == 0) return fail(s, PSIM_EINVAL, "psim_group_create: bad argument");
    *out = nullptr;
    for (uint32_t r = 0; r < count; ++r) {
        PsimStepper* m = steppers[r];
        if (!m || m->group || m->comm || m->nranks != (int)count || m->rank != (int)r ||
            m->device != steppers[0]->device || m->cfg.grid_x_log2 != steppers[0]->cfg.grid_x_log2 ||
            m->cfg.grid_y_log2 != steppers[0]->cfg.grid_y_log2 || m->cfg.schedule != steppers[0]->cfg.schedule ||
            m->cfg.rebin_every != steppers[0]->cfg.rebin_every || m->box_capacity != steppers[0]->box_capacity ||
            m->ingest_cap != steppers[0]->ingest_cap ||  // every slab scans the whole ingested frame
            (r > 0 && (m->cfg.slab_bounds[1] != steppers[r - 1]->cfg.slab_bounds[2] ||
                       m->cfg.slab_bounds[0] != steppers[r - 1]->cfg.slab_bounds[1] ||
                       m->cfg.slab_bounds[2] != steppers[r - 1]->cfg.slab_bounds[3])))  // adjacent slabs meet
            return fail(s, PSIM_EINVAL, "psim_group_create: stepper %u must be slab %u of %u on the group's device "
                        "with the group's grid, schedule and capacities (ingest_capacity included), its rows starting where slab %u's end",
                        r, r, count, r ? r - 1 : 0);
    }
    CK(cudaSetDevice(steppers[0]->device));
    PsimGroup* g = new PsimGroup;
    g->ranks.assign(steppers, steppers + count);
    g->stream = steppers[0]->own_stream;
    const char* mode = getenv("PSIM_HALO");
    const bool push = !(mode && std::strcmp(mode, "nccl") == 0);
    for (uint32_t r = 0; r < count; ++r) {
        PsimStepper* m = g->ranks[r];
        cudaStreamSynchronize(m->stream);
        m->group = g;
        m->stream = g->stream;
        m->push = push && count > 1;  // same device: the neighbours' buffers are plain pointers
        for (int side = 0; side < 2; ++side) {
            const int peer = side == 0 ? (int)r - 1 : (int)r + 1;
            const bool has = m->push && peer >= 0 && peer < (int)count;
            m->peer_pos[side][0] = has ? g->ranks[peer]->pos[0] : nullptr;
            m->peer_pos[side][1] = has ? g->ranks[peer]->pos[1] : nullptr;
            m->peer_nbr[side][0] = has ? g->ranks[peer]->nbr[0] : nullptr;
            m->peer_nbr[side][1] = has ? g->ranks[peer]->nbr[1] : nullptr;
            m->peer_hdr[side] = has ? g->ranks[peer]->hdr : nullptr;
        }
    }
    *out = g;
    return PSIM_OK;
}

void psim_group_destroy(PsimGroup* g) {
    if (!g) return;
    for (PsimStepper* m : g->ranks) {
        cudaStreamSynchronize(g->stream);
        m->group = nullptr;
        m->stream = m->own_stream;
        m->push = false;
        for (int side = 0; side < 2; ++side) {
            m->peer_pos[side][0] = m->peer_pos[side][1] = nullptr;
            m->peer_nbr[side][0] = m->peer_nbr[side][1] = nullptr;
            m->peer_hdr[side] = nullptr;
        }
    }
    delete g;
}

}  // extern "C"

namespace {
Team team_of(PsimGroup* g) { return Team{g->ranks.data(), (int)g->ranks.size(), g}; }

int group_fail(PsimGroup* g, int rc) {
    if (rc && g->error.empty())
        for (PsimStepper* m : g->ranks)
            if (!m->error.empty()) g->error = m->error;
    return rc;
}
}  // namespace

extern "C" {

int psim_group_upload_frame(PsimGroup* g, const FrameHeader* frame) {
    if (!g || !frame) return PSIM_EINVAL;
    g->error.clear();
    PsimStepper* s = g->ranks[0];
    if (frame->metadata.data_structure == 0)
        return group_fail(g, fail(s, PSIM_EINVAL, "DataStructure::CompactArray (all pairs) cannot be decomposed into slabs"));
    for (PsimStepper* m : g->ranks) m->error.clear();
    CK(cudaSetDevice(s->device));
    if (frame->particle_count > s->ingest_cap)
        return group_fail(g, fail(s, PSIM_ECAPACITY, "frame holds %u particles, the ingest buffer of slab 0 holds %u",
                                  frame->particle_count, s->ingest_cap));
    CK(cudaStreamSynchronize(g->stream));
    for (PsimStepper* m : g->ranks) {
        CK(cudaStreamSynchronize(m->copy_stream));
        apply_metadata(m, frame->metadata);
    }
    if (frame->particle_count)
        CK(cudaMemcpyAsync(s->staging, frame->particles, sizeof(Particle) * (size_t)frame->particle_count,
                           cudaMemcpyHostToDevice, g->stream));
    return group_fail(g, team_ingest(team_of(g), s->staging, frame->particle_count));
}

int psim_group_set_metadata(PsimGroup* g, const FrameMetadata* meta) {
    if (!g || !meta) return PSIM_EINVAL;
    for (PsimStepper* m : g->ranks) apply_metadata(m, *meta);
    return PSIM_OK;
}

int psim_group_step_async(PsimGroup* g, uint32_t steps) {
    if (!g) return PSIM_EINVAL;
    g->error.clear();
    PsimStepper* s = g->ranks[0];
    if (!s->has_scene) return group_fail(g, fail(s, PSIM_ESTATE, "psim_group_step_async: no scene uploaded"));
    CK(cudaSetDevice(s->device));
    for (uint32_t k = 0; k < steps; ++k) {
        int rc = team_step(team_of(g));
        if (rc) return group_fail(g, rc);
        for (PsimStepper* m : g->ranks) m->fresh_scene = false;
    }
    return PSIM_OK;
}

int psim_group_rebin_async(PsimGroup* g) {
    if (!g) return PSIM_EINVAL;
    g->error.clear();
    PsimStepper* s = g->ranks[0];
    if (!s->has_scene) return group_fail(g, fail(s, PSIM_ESTATE, "psim_group_rebin_async: no scene uploaded"));
    CK(cudaSetDevice(s->device));
    for (PsimStepper* m : g->ranks) m->fresh_scene = false;
    return group_fail(g, team_bin(team_of(g), false, nullptr, 0));
}

int psim_group_snapshot_async(PsimGroup* g) {
    if (!g) return PSIM_EINVAL;
    g->error.clear();
    PsimStepper* s = g->ranks[0];
    if (!s->has_scene) return group_fail(g, fail(s, PSIM_ESTATE, "psim_group_snapshot_async: no scene uploaded"));
    CK(cudaSetDevice(s->device));
    return group_fail(g, team_snapshot(team_of(g)));
}

int psim_group_run_frame_async(PsimGroup* g) {
    if (!g) return PSIM_EINVAL;
    g->error.clear();
    PsimStepper* s = g->ranks[0];
    if (!s->has_scene) return group_fail(g, fail(s, PSIM_ESTATE, "psim_group_run_frame_async: no scene uploaded"));
    CK(cudaSetDevice(s->device));
    int rc = team_run_frame(team_of(g));
    for (PsimStepper* m : g->ranks) m->fresh_scene = false;
    return group_fail(g, rc);
}

int psim_group_sync(PsimGroup* g) {
    if (!g) return PSIM_EINVAL;
    PsimStepper* s = g->ranks[0];
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(g->stream));
    for (PsimStepper* m : g->ranks)
        if (m->timing) {
            int rc = collect_timing(m);
            if (rc) return rc;
        }
    return PSIM_OK;
}

uint32_t psim_group_particle_count(const PsimGroup* g) {
    uint32_t n = 0;
    if (g)
        for (const PsimStepper* m : g->ranks) n += m->n;
    return n;
}

int psim_group_download_frame(PsimGroup* g, FrameHeader* dst) {
    if (!g || !dst) return PSIM_EINVAL;
    g->error.clear();
    PsimStepper* s = g->ranks[0];
    uint64_t total = 0;
    for (PsimStepper* m : g->ranks) {
        const int k = snapshot_index(m, 0);
        if (k < 0) return group_fail(g, fail(s, PSIM_ESTATE, "psim_group_download_frame: no snapshot has been packed"));
        total += m->snapshot_n[k];
    }
    if (dst->particle_count < total)
        return group_fail(g, fail(s, PSIM_ECAPACITY, "psim_group_download_frame: destination holds %u particles, "
                                  "snapshot has %llu", dst->particle_count, (unsigned long long)total));
    CK(cudaSetDevice(s->device));
    // slabs in rank order = ascending cell rows = the cell-major order of a single-slab snapshot
    Particle* out = dst->particles;
    for (PsimStepper* m : g->ranks) {
        const int k = snapshot_index(m, 0);
        int rc = download_records(m, k, out);
        if (rc) return group_fail(g, rc);
        out += m->snapshot_n[k];
    }
    write_header(dst, s->snapshot_meta[snapshot_index(s, 0)], (uint32_t)total);
    return PSIM_OK;
}

}  // extern "C"
