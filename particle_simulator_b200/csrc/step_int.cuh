// step_int.cuh -- the halo protocol inside a step, the per-particle epilogue (wall, cursor, kick, drift), step_kernel (integer separations: coarse grids and exponents without an fp32 variant) and the all-pairs kernel of CompactArray.
// Included by stepper.cu inside its anonymous namespace (one translation unit: the kernels, their parameter blocks and
// the host code that launches them are compiled together). Not a stand-alone header.

// ------------------------------------------------------------------------------------------------
// The step kernel: force over the 3x3 cell stencil + kick + drift, one HBM round trip of the state.
//
// CTA b owns particles [b*kTile, (b+1)*kTile). Because the arrays are cell-sorted and cells are
// row-major, everything those particles interact with lies in three contiguous index ranges, one
// per stencil row: cells [first-1, last+1] shifted by -BX, 0, +BX. One thread reads the tile's
// descriptor and issues up to six 1-D bulk copies (TMA, cp.async.bulk -> mbarrier) that stage the
// cell_start entries and the positions of those ranges in shared memory; meanwhile every thread
// loads its own particle.  Then every thread walks its own three windows (cells cx-1..cx+1 of rows
// cy-1..cy+1, clipped at the grid edge exactly like kernel_bucket.cuh:74-77) in ascending index
// order, the (row, column, slot) order of the reference's loop.
// Tiles whose stencil does not fit the staging buffers (very sparse or very clustered spots) run
// the same code with the pointers aimed at global memory instead.
// ------------------------------------------------------------------------------------------------

// ------------------------------------------------------------------------------------------------
// Halo exchange inside the step kernel (slab decomposition).
//
// Each slab keeps one ghost row of each neighbour. Instead of a send/recv after every step, the threads
// that step a particle of a boundary row store its new position twice: in the slab's own array and, through
// peer-mapped memory (NVLink P2P: CUDA IPC between processes, plain pointers inside one process), in the
// neighbour's ghost row of the buffer the neighbour's NEXT step reads. Synchronisation is one epoch word
// per direction in the receiver's HaloHeader:
//   * the thread that pushes the last particle of a boundary row publishes this step's epoch
//     (__threadfence_system + store to the neighbour's header);
//   * only the CTAs that read a ghost row (the tiles of the first / last owned row -- the same CTAs that
//     produce the outgoing halo) wait, before staging, until the neighbour has published the previous
//     step's epoch. Interior tiles never wait, so the transfer overlaps the interior's pair loops.
// That wait also covers the write-after-read hazard: a neighbour publishes epoch k only after ITS
// boundary tiles of step k have finished, i.e. have finished reading the ghost rows step k+1 overwrites.
// Boundary tiles come first / last in the grid, so their halo is on the wire while the interior computes.
// ------------------------------------------------------------------------------------------------
// Constants of the fp32-offset step kernel (step_float.cuh) and of the neighbour records it reads.
struct PhysF {
    float sx, sy;          // fixed-point units -> scaled units (sx a power of two; sy = sx * ky/kx, also one)
    float zone_shift;      // zone stride * cell width * sx: distance between the even-zone and odd-zone origins
    float row_shift;       // cell height * sy: distance between the centres of two adjacent cell rows
    float d0, d1, d2, d3;  // -(n/m) f^(-2(kn-km)) q^fn as a cubic in l = log2(scaled r^2)  (d0 alone if fn == 0)
    float pair_scale;      // scaled pair sum -> newtons
    uint32_t zl;           // log2 of the zone stride in cell columns
    uint32_t half_span;    // (2^zl + 2) cells / 2 in fixed-point units: centre of a zone's used span
    uint32_t sxbits;       // 32 - LX
};

struct HaloHeader {     // one per slab, in device memory its two neighbours can reach
    uint32_t flags[2];  // [0]: last epoch published by the lower neighbour, [1]: by the upper one
    uint32_t done[2];   // boundary particles pushed so far in the running step, per side
    uint32_t own_hi;    // where this slab's upper ghost row starts (written at every binning)
    uint32_t error;     // sticky: a wait timed out (the neighbour died); waits stop blocking
    uint32_t _pad[2];
};

struct HaloArgs {
    uint2* peer_out[2];       // the neighbour's position buffer this step writes ([0] lower, [1] upper); null: none
    float4* peer_nbr_out[2];  // ... and its neighbour-record buffer (fine grids), or null
    HaloHeader* peer_hdr[2];
    HaloHeader* hdr;
    uint32_t lo_end, hi_start;    // [own_lo, lo_end) goes to the lower neighbour, [hi_start, own_hi) to the upper one
    uint32_t lo_tiles, hi_tile0;  // the tiles that hold (and read the ghost row next to) them: [0, lo_tiles), [hi_tile0, ..)
    uint32_t wait_epoch[2];       // 0: the ghost row is already in place (a binning delivered it)
    uint32_t pub_epoch;
};

constexpr unsigned long long kHaloTimeoutNs = 20ull * 1000 * 1000 * 1000;

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct StepArgs {
    const uint2* __restrict__ pos_in;
    uint2* __restrict__ pos_out;
    float2* __restrict__ vel;
    const uint32_t* __restrict__ cell_id;
    const uint32_t* __restrict__ cell_start;
    const TileDesc* __restrict__ tiles;
    uint32_t own_lo, own_hi;  // the particles this stepper steps: [own_lo, own_hi) (ghost rows lie outside)
    Grid g;
    Phys ph;
    uint32_t push;  // 1: boundary rows are pushed into the neighbours' ghost rows by this kernel (HaloArgs)
    HaloArgs h;
    // fine grids: every particle also has a NEIGHBOUR RECORD (x_even, y, x_odd, y): its position as exact scaled fp32
    // offsets from the centres of its membership cell's even / odd zone and of its membership row (step_float.cuh).
    // A step reads nbr_in (staged by TMA, no conversion) and writes nbr_out for the next one. Null on coarse grids.
    const float4* __restrict__ nbr_in;
    float4* __restrict__ nbr_out;
    PhysF pf;
};

// Where the offsets of a neighbour record are measured from: the centre of the membership cell's even zone span (x),
// the centre of its row (y), and the signed distance to the odd zone's origin. The same for all particles of a cell.
struct RecOrigin {
    uint32_t xo0, yc;
    float odd_shift;
};

__device__ __forceinline__ RecOrigin rec_origin(uint32_t cell, const Grid& g, const PhysF& pf) {
    const uint32_t cx = cell & (g.bx - 1);
    const long long row = (long long)(cell >> g.lx) + g.row_offset;              // global cell row
    const uint32_t zq = cx >> pf.zl;
    RecOrigin o;
    o.yc = (uint32_t)((2ll * row + 1) << (g.sy - 1));                            // centre of that row
    o.xo0 = (((zq & ~1u) << pf.zl) << pf.sxbits) + pf.half_span;                 // centre of the even zone's span
    o.odd_shift = (zq & 1u) ? -pf.zone_shift : pf.zone_shift;
    return o;
}

__device__ __forceinline__ float4 nbr_record(uint2 p, const RecOrigin& o, const PhysF& pf) {
    const float xe = __int2float_rn((int)(p.x - o.xo0)) * pf.sx;
    const float y = __int2float_rn((int)(p.y - o.yc)) * pf.sy;
    return make_float4(xe, y, xe + o.odd_shift, y);
}

// The neighbour record of a particle at `p` whose membership cell is `cell` (local numbering).
__device__ __forceinline__ float4 nbr_record(uint2 p, uint32_t cell, const Grid& g, const PhysF& pf) {
    return nbr_record(p, rec_origin(cell, g, pf), pf);
}

// Before a tile that reads a ghost row stages anything: wait for the neighbour's previous step (one thread).
__device__ __forceinline__ void halo_wait(const StepArgs& a, uint32_t tile) {
    const HaloArgs& h = a.h;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const bool reads_ghost = side == 0 ? tile < h.lo_tiles : tile >= h.hi_tile0;
        if (!h.peer_out[side] || !h.wait_epoch[side] || !reads_ghost) continue;
        if (ld_acquire_sys(&h.hdr->error)) continue;
        const unsigned long long t0 = global_timer_ns();
        while ((int32_t)(ld_acquire_sys(&h.hdr->flags[side]) - h.wait_epoch[side]) < 0) {
            if (global_timer_ns() - t0 > kHaloTimeoutNs) {
                st_release_sys(&h.hdr->error, 1u);
                break;
            }
        }
    }
}

__device__ __forceinline__ void halo_publish(const HaloArgs& h, int side) {
    h.hdr->done[side] = 0;  // for the next step (its launch is ordered after this kernel)
    __threadfence_system();
    st_release_sys(&h.peer_hdr[side]->flags[side ^ 1], h.pub_epoch);
}

// A boundary row without particles has nobody to publish its epoch: the first thread of the grid does.
__device__ __forceinline__ void halo_publish_empty(const StepArgs& a) {
    const HaloArgs& h = a.h;
    if (h.peer_out[0] && h.lo_end == a.own_lo) halo_publish(h, 0);
    if (h.peer_out[1] && h.hi_start == a.own_hi) halo_publish(h, 1);
}

// The new position of boundary-row particle i also goes into the neighbour's ghost row.
__device__ __forceinline__ void halo_push(const StepArgs& a, uint32_t i, uint2 po, float4 nb) {
    const HaloArgs& h = a.h;
    if (h.peer_out[0] && i < h.lo_end) {
        const uint32_t base = ld_acquire_sys(&h.peer_hdr[0]->own_hi);  // the lower slab's upper ghost row
        h.peer_out[0][base + (i - a.own_lo)] = po;
        if (h.peer_nbr_out[0]) h.peer_nbr_out[0][base + (i - a.own_lo)] = nb;
        __threadfence_system();
        if (atomicAdd(&h.hdr->done[0], 1u) + 1u == h.lo_end - a.own_lo) halo_publish(h, 0);
    }
    if (h.peer_out[1] && i >= h.hi_start) {
        h.peer_out[1][i - h.hi_start] = po;  // the upper slab's lower ghost row starts at 0
        if (h.peer_nbr_out[1]) h.peer_nbr_out[1][i - h.hi_start] = nb;
        __threadfence_system();
        if (atomicAdd(&h.hdr->done[1], 1u) + 1u == a.own_hi - h.hi_start) halo_publish(h, 1);
    }
}

// Which tile the b-th CTA of the grid steps. With a pushed halo the tiles of BOTH boundary rows come first (the first
// owned row's, then the last owned row's, then the interior in order): the halo is on the wire, fenced and published
// while the interior computes, and no neighbour ever finds a flag late because its producer ran at the grid's tail.
__device__ __forceinline__ uint32_t halo_tile_order(const StepArgs& a, uint32_t b, uint32_t tiles) {
    if (!a.push) return b;
    const uint32_t lo = a.h.lo_tiles, hi0 = max(a.h.hi_tile0, lo), hi_count = tiles - min(hi0, tiles);
    if (b < lo) return b;
    if (b < lo + hi_count) return hi0 + (b - lo);
    return b - hi_count;
}

// A slab without particles still owes its neighbours the epoch of every step.
__global__ void halo_publish_kernel(StepArgs a) {
    if (threadIdx.x == 0 && blockIdx.x == 0) halo_publish_empty(a);
}

// Cursor + wall force, pair sum, kick + drift and the stores of one particle (shared by all step kernels).
// `origin`: where the particle's neighbour record is measured from (null: derived from `cell` when records are kept).
template <bool M6 = false>
__device__ __forceinline__ void finish_particle(uint32_t i, uint2 pi, float2 vi, uint32_t cell, float sum_x, float sum_y,
                                                float scale_x, float scale_y, const StepArgs& a,
                                                const RecOrigin* origin = nullptr) {
    float2 f = field_force<M6>(pi, a.ph);
    f.x = fmaf(scale_x, sum_x, f.x);
    f.y = fmaf(scale_y, sum_y, f.y);
    uint2 po;
    float2 vo;
    integrate(pi, vi, f, a.ph, po, vo);
    a.pos_out[i] = po;
    a.vel[i] = vo;
    float4 nb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (origin) {
        nb = nbr_record(po, *origin, a.pf);
        a.nbr_out[i] = nb;
    } else if (a.nbr_out) {
        nb = nbr_record(po, cell, a.g, a.pf);
        a.nbr_out[i] = nb;
    }
    if (a.push) halo_push(a, i, po, nb);
}

template <int KN, int FRAC, bool ANISO, bool CG>
__device__ __forceinline__ void step_particle(uint32_t i, uint2 pi, float2 vi, uint32_t cell,
                                              const uint32_t* const cs[3], const uint32_t cs_lo[3],
                                              const uint2* const pp[3], const uint32_t pp_lo[3], const StepArgs& a) {
    const Grid& g = a.g;
    uint32_t cx = cell & (g.bx - 1), cy = cell >> g.lx;
    uint32_t x0 = cx == 0 ? 0 : cx - 1, x1 = cx == g.bx - 1 ? cx : cx + 1;
    float2 gx = splat(0.f), gy = splat(0.f);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        int row = (int)cy + d - 1;
        if (row < 0 || row >= (int)g.by) continue;
        uint32_t c0 = ((uint32_t)row << g.lx) + x0, c1 = ((uint32_t)row << g.lx) + x1;
        uint32_t s = cs[d][c0 - cs_lo[d]], e = cs[d][c1 + 1 - cs_lo[d]];
        const uint2* win = pp[d] + (s - pp_lo[d]);  // window [s, e) of this row
        if (d == 1) window_accumulate<KN, FRAC, ANISO, true, CG>(win, (int)(e - s), pi, a.ph, gx, gy);  // contains i
        else window_accumulate<KN, FRAC, ANISO, false, CG>(win, (int)(e - s), pi, a.ph, gx, gy);
    }
    finish_particle<(KN > 0)>(i, pi, vi, cell, gx.x + gx.y, gy.x + gy.y, a.ph.pair_scale, a.ph.pair_scale_y, a);
}

template <int KN, int FRAC, bool ANISO>
__global__ void __launch_bounds__(kTile, 8) step_kernel(const StepArgs a) {
    __shared__ __align__(16) uint32_t s_cs[3][kCsCap];
    __shared__ __align__(16) uint2 s_pos[3][kPosCap];
    __shared__ __align__(8) uint64_t s_bar;

    const uint32_t b = halo_tile_order(a, blockIdx.x, gridDim.x);
    const uint32_t i = a.own_lo + b * kTile + threadIdx.x;
    const TileDesc t = a.tiles[b];

    if (a.push) {  // uniform over the grid
        if (threadIdx.x == 0) {
            if (blockIdx.x == 0) halo_publish_empty(a);
            halo_wait(a, b);
        }
        if (!t.fits) __syncthreads();  // the global-memory path reads the ghost rows directly
    }
    if (t.fits) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar, 1);
            uint32_t bytes = 0;
#pragma unroll
            for (int d = 0; d < 3; ++d) bytes += t.cs_cnt[d] * 4u + t.p_cnt[d] * 8u;
            mbar_arrive_expect_tx(&s_bar, bytes);
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                if (t.cs_cnt[d]) bulk_copy_g2s(s_cs[d], a.cell_start + t.cs_lo[d], t.cs_cnt[d] * 4u, &s_bar);
                if (t.p_cnt[d]) bulk_copy_g2s(s_pos[d], a.pos_in + t.p_lo[d], t.p_cnt[d] * 8u, &s_bar);
            }
        }
        __syncthreads();  // the barrier is initialised before anyone polls it
    }
    const bool live = i < a.own_hi;
    uint2 pi = make_uint2(0, 0);
    float2 vi = make_float2(0.f, 0.f);
    uint32_t cell = 0;
    if (live) {
        pi = a.pos_in[i];
        vi = a.vel[i];
        cell = a.cell_id[i];
    }
    if (t.fits) {
        mbar_wait(&s_bar, 0);
        if (!live) return;
        const uint32_t* cs[3] = {s_cs[0], s_cs[1], s_cs[2]};
        const uint2* pp[3] = {s_pos[0], s_pos[1], s_pos[2]};
        step_particle<KN, FRAC, ANISO, false>(i, pi, vi, cell, cs, t.cs_lo, pp, t.p_lo, a);
    } else {
        if (!live) return;
        const uint32_t* cs[3] = {a.cell_start, a.cell_start, a.cell_start};
        const uint2* pp[3] = {a.pos_in, a.pos_in, a.pos_in};
        const uint32_t zero[3] = {0, 0, 0};
        step_particle<KN, FRAC, ANISO, true>(i, pi, vi, cell, cs, zero, pp, zero, a);
    }
}

// ------------------------------------------------------------------------------------------------
// Per-species Mie parameters (PsimConfig.species_physics). An EXTENSION: the metadata carries two
// MiePotentialParams (particle.rs:114) but the reference steps every particle with species 0's (kernel_bucket.cuh:52).
// Here a pair (i, j) uses the parameters of its species pair -- like pairs their own, unlike pairs the Lorentz-
// Berthelot mix sigma = (s0 + s1) / 2, epsilon = sqrt(e0 e1), exponents (n0 + n1) / 2, (m0 + m1) / 2 -- and the wall
// term the particle's own. species = min(ty, 1). One particle per thread on step_kernel's tiles, general exponents
// (2^(-e log2 r^2): MUFU.LG2 + two MUFU.EX2 per pair); the neighbours' labels are read from the sorted `ty` array.
// Checked against oracle_step_species (oracle/psim_oracle.c), the same extension of the CPU restatement.
// ------------------------------------------------------------------------------------------------
struct SpeciesTab {
    float inv_c2;      // kx^2 / sigma^2
    float em, en;      // m / 2 + 1, n / 2 + 1
    float nm;          // n / m
    float pair_scale;  // C eps m kx / sigma^2
    float sigma, wall_scale, m;  // the wall term of a particle of this (like-pair) species
};

struct SpeciesArgs {
    SpeciesTab tab[3];  // species pairs 00, 01, 11 (index = species_i + species_j)
    const int32_t* __restrict__ ty;
};

__device__ __forceinline__ int species_of(int32_t ty) { return ty > 0 ? 1 : 0; }

__device__ __forceinline__ float species_wall_term(float d, const SpeciesTab& t) {
    const float inv_d = fast_rcp(d);
    return t.wall_scale * fast_ex2(t.m * fast_lg2(t.sigma * inv_d)) * inv_d;
}

template <bool ANISO>
__global__ void __launch_bounds__(kTile) step_kernel_species(const StepArgs a, const SpeciesArgs sp) {
    __shared__ __align__(16) uint32_t s_cs[3][kCsCap];
    __shared__ __align__(16) uint2 s_pos[3][kPosCap];
    __shared__ __align__(8) uint64_t s_bar;

    const uint32_t b = blockIdx.x;
    const uint32_t i = a.own_lo + b * kTile + threadIdx.x;
    const TileDesc t = a.tiles[b];
    if (t.fits) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar, 1);
            uint32_t bytes = 0;
#pragma unroll
            for (int d = 0; d < 3; ++d) bytes += t.cs_cnt[d] * 4u + t.p_cnt[d] * 8u;
            mbar_arrive_expect_tx(&s_bar, bytes);
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                if (t.cs_cnt[d]) bulk_copy_g2s(s_cs[d], a.cell_start + t.cs_lo[d], t.cs_cnt[d] * 4u, &s_bar);
                if (t.p_cnt[d]) bulk_copy_g2s(s_pos[d], a.pos_in + t.p_lo[d], t.p_cnt[d] * 8u, &s_bar);
            }
        }
        __syncthreads();
    }
    const bool live = i < a.own_hi;
    uint2 pi = make_uint2(0, 0);
    float2 vi = make_float2(0.f, 0.f);
    uint32_t cell = 0;
    int si = 0;
    if (live) {
        pi = a.pos_in[i];
        vi = a.vel[i];
        cell = a.cell_id[i];
        si = species_of(sp.ty[i]);
    }
    if (t.fits) mbar_wait(&s_bar, 0);
    if (!live) return;
    const Grid& g = a.g;
    const uint32_t cx = cell & (g.bx - 1), cy = cell >> g.lx;
    const uint32_t x0 = cx == 0 ? 0 : cx - 1, x1 = cx == g.bx - 1 ? cx : cx + 1;
    float fx = 0.f, fy = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int row = (int)cy + d - 1;
        if (row < 0 || row >= (int)g.by) continue;
        const uint32_t c0 = ((uint32_t)row << g.lx) + x0, c1 = ((uint32_t)row << g.lx) + x1;
        const uint32_t* cs = t.fits ? s_cs[d] - t.cs_lo[d] : a.cell_start;
        const uint32_t ws = cs[c0], we = cs[c1 + 1];
        const uint2* pp = t.fits ? s_pos[d] - t.p_lo[d] : a.pos_in;
        for (uint32_t j = ws; j < we; ++j) {
            const uint2 pj = t.fits ? pp[j] : __ldcg(pp + j);
            const SpeciesTab& tab = sp.tab[si + species_of(__ldg(sp.ty + j))];
            const float x = __int2float_rn((int)(pj.x - pi.x));
            float y = __int2float_rn((int)(pj.y - pi.y));
            if (ANISO) y *= a.ph.yscale;
            const float r2 = fmaxf((x * x + y * y) * tab.inv_c2, 1e-2f);  // j == i: x = y = 0 contributes an exact 0
            const float l = fast_lg2(r2);
            const float gq = tab.pair_scale * (fast_ex2(-tab.em * l) - tab.nm * fast_ex2(-tab.en * l));
            fx = fmaf(gq, x, fx);
            fy = fmaf(gq, y, fy);
        }
    }
    // cursor (kernel_bucket.cuh:54-67) and the walls with the particle's own species (particle.cuh:125-144)
    const Phys& ph = a.ph;
    const SpeciesTab& own = sp.tab[2 * si];
    float2 f = make_float2(fx, fy);
    if (ph.cursor_on) {
        const float inv32 = 1.f / 4294967296.f;
        const float dx = ph.cursor_x - __uint2float_rn(pi.x) * inv32, dy = ph.cursor_y - __uint2float_rn(pi.y) * inv32;
        const float sq = dx * dx + dy * dy;
        if (sq < ph.cursor_r2) {
            const float c = 8e-12f * fast_rcp(sq + 1.f);
            f.x += dx > 0 ? -c : c;
            f.y += dy > 0 ? -c : c;
        }
    }
    const bool left = pi.x < 0xFFFFFFFFu / 2, low = pi.y < 0xFFFFFFFFu / 2;
    const float wx = species_wall_term(__uint2float_rn(left ? pi.x : 0xFFFFFFFFu - pi.x) * ph.kx, own);
    const float wy = species_wall_term(__uint2float_rn(low ? pi.y : 0xFFFFFFFFu - pi.y) * ph.ky, own);
    f.x += left ? wx : -wx;
    f.y += low ? wy : -wy;
    uint2 po;
    float2 vo;
    integrate(pi, vi, f, ph, po, vo);
    a.pos_out[i] = po;
    a.vel[i] = vo;
}

// ------------------------------------------------------------------------------------------------
// DataStructure::CompactArray (kernel_compact.cuh:4-34): every particle interacts with every other one, particles
// keep their input order, there is no grid. O(N^2): the reference's teaching baseline, offered so that the
// metadata's data_structure switch (kernel.cuh:143-150) means here what it means there. One thread per particle;
// the array streams through shared memory 128 positions at a time; same pair arithmetic as step_kernel, with
// f_dist's unsigned separation (two particles can be more than half the box apart).
// ------------------------------------------------------------------------------------------------
template <int KN, int FRAC, bool ANISO>
__global__ void __launch_bounds__(kTile) allpairs_step_kernel(const StepArgs a) {
    __shared__ uint2 s_pos[kTile];
    const uint32_t n = a.own_hi;
    const uint32_t i = blockIdx.x * kTile + threadIdx.x;
    const bool live = i < n;
    const uint2 pi = live ? a.pos_in[i] : make_uint2(0, 0);
    float2 gx = splat(0.f), gy = splat(0.f);
    for (uint32_t base = 0; base < n; base += kTile) {
        __syncthreads();
        if (base + threadIdx.x < n) s_pos[threadIdx.x] = a.pos_in[base + threadIdx.x];
        __syncthreads();
        const int count = (int)min((uint32_t)kTile, n - base);
        int k = 0;
        for (; k + 1 < count; k += 2)  // j == i contributes an exact 0 (the clamp), like the reference's `continue`
            pair2<KN, FRAC, ANISO, false, true, true>(pi, s_pos[k], s_pos[k + 1], a.ph, gx, gy);
        if (k < count) pair2<KN, FRAC, ANISO, true, true, true>(pi, s_pos[k], pi, a.ph, gx, gy);
    }
    if (live) finish_particle(i, pi, a.vel[i], 0u, gx.x + gx.y, gy.x + gy.y, a.ph.pair_scale, a.ph.pair_scale_y, a);
}

// CompactArray ingest: wire-format records (already free of nulls) -> the structure of arrays, input order kept.
__global__ void unpack_kernel(const Particle* __restrict__ rec, uint32_t n, uint2* __restrict__ pos,
                              float2* __restrict__ vel, int32_t* __restrict__ ty, uint32_t* __restrict__ cell_id) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Particle q = rec[i];
    pos[i] = make_uint2(q.x, q.y);
    vel[i] = make_float2(q.vx, q.vy);
    ty[i] = q.ty;
    cell_id[i] = 0;
}
