// binning.cuh -- the stable counting sort by cell (ingest and re-bin), the tile tables, neighbour records, snapshots, and migration between slabs.
// Included by stepper.cu inside its anonymous namespace (one translation unit: the kernels, their parameter blocks and
// the host code that launches them are compiled together). Not a stand-alone header.

// ------------------------------------------------------------------------------------------------
// Binning: stable counting sort by cell (count -> scan -> scatter -> order fix-up + gather).
//
//   key_count : key = cell(pos); rank = atomicAdd(count[key], 1)          (arbitrary rank in cell)
//   scan      : cell_start = exclusive prefix sum of count                 (3 small kernels)
//   scatter   : perm[cell_start[key] + rank] = candidate index
//   gather    : slot p holds candidate c = perm[p]; its final place inside its cell is the number of
//               cell-mates with a smaller candidate index, which makes the result the STABLE sort
//               whatever order the atomics were served in -- the order the reference's serial
//               append (kernel.cuh:219-229) and its (row, column, slot) pull (kernel_bucket.cuh:17-33)
//               both produce.
// The candidates are, in this order: wire-format records ahead of the live state (an ingested frame,
// or the particles that migrated in from the lower slab), the live particles this stepper owns, and
// records behind them (migrants from the upper slab) -- the order those particles have in the global
// cell-sorted array, so that a slab-decomposed run sorts exactly like a single-slab one.
// ------------------------------------------------------------------------------------------------

struct Source {
    const Particle* aos_lo;  // records that sort ahead of the live state (may contain nulls, ty < 0)
    uint32_t n_lo;
    const uint2* pos;  // the live state, particles [soa_lo, soa_lo + n_soa)
    const float2* vel;
    const int32_t* ty;
    uint32_t soa_lo, n_soa;
    const Particle* aos_hi;  // records that sort behind it
    uint32_t n_hi;
    uint32_t strict;  // 1: a record outside the owned rows is an error (migrants), 0: it is skipped (ingest)
};

__device__ __forceinline__ uint32_t source_count(const Source& s) { return s.n_lo + s.n_soa + s.n_hi; }

// Position / velocity / label of candidate c; returns false for a null record (kernel.cuh:222).
__device__ __forceinline__ bool source_fetch(const Source& s, uint32_t c, uint2& pos, float2& vel, int32_t& ty,
                                             bool& is_record) {
    const Particle* rec = nullptr;
    if (c < s.n_lo) rec = s.aos_lo + c;
    else if (c >= s.n_lo + s.n_soa) rec = s.aos_hi + (c - s.n_lo - s.n_soa);
    is_record = rec != nullptr;
    if (rec) {
        Particle q = *rec;
        pos = make_uint2(q.x, q.y);
        vel = make_float2(q.vx, q.vy);
        ty = q.ty;
        return q.ty >= 0;
    }
    uint32_t i = s.soa_lo + (c - s.n_lo);
    pos = s.pos[i];
    vel = s.vel[i];
    ty = s.ty[i];
    return true;
}

// device-side error bits (PsimStepper::d_flags[0])
constexpr uint32_t kErrMigrantOutside = 1u;   // a migrant record does not belong to this slab
constexpr uint32_t kErrMigrantOverflow = 2u;  // more migrants than the exchange boxes hold
constexpr uint32_t kErrMigrantTooFar = 4u;    // a particle left for a slab that is not adjacent

__global__ void key_count_kernel(Source src, Grid g, uint32_t* __restrict__ cell_count, uint32_t* __restrict__ rank,
                                 uint32_t* __restrict__ flags) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t key = kKeyUp;  // nothing to count: past the end, a null record, or outside the owned rows
    if (c < source_count(src)) {
        uint2 pos;
        float2 vel;
        int32_t ty;
        bool is_record;
        if (source_fetch(src, c, pos, vel, ty, is_record)) {
            key = cell_of(pos, g);
            // outside the owned rows: filtered (ingest) or already extracted (live state)
            if (key >= kKeyUp && is_record && src.strict) atomicOr(flags, kErrMigrantOutside);
        }
    }
    // one atomic per particle: lanes of a warp that hit the same cell are combined by the hardware at L2; a
    // __match_any_sync aggregation in front of it was measured slower (78 vs 55 us at 10M, sorted input)
    if (key >= kKeyUp) return;
    rank[c] = atomicAdd(&cell_count[key], 1u);
}

// The scans produce cell_start (exclusive prefix sum of the counts) and, on fine grids, pad_start (the same over the
// counts rounded up to even: couples of step_kernel_c) in one pass over cell_count: .x sums the counts, .y the padded ones.
__device__ __forceinline__ uint2 scan_item(uint32_t v) { return make_uint2(v, (v + 1u) & ~1u); }
__device__ __forceinline__ uint2 add2(uint2 a, uint2 b) { return make_uint2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ uint2 shfl_up2(uint2 v, int o) {
    return make_uint2(__shfl_up_sync(0xFFFFFFFFu, v.x, o), __shfl_up_sync(0xFFFFFFFFu, v.y, o));
}

// kScanItems = 8 consecutive counts of one thread as two 16-byte loads (the arrays are 16-byte aligned and padded past
// `count`); entries past the end read as 0.
__device__ __forceinline__ void load_items(const uint32_t* __restrict__ in, uint32_t base, uint32_t count, uint32_t item[kScanItems]) {
    static_assert(kScanItems == 8, "two uint4 per thread");
    if (base + kScanItems <= count) {
        const uint4 a = *reinterpret_cast<const uint4*>(in + base), b = *reinterpret_cast<const uint4*>(in + base + 4);
        item[0] = a.x, item[1] = a.y, item[2] = a.z, item[3] = a.w, item[4] = b.x, item[5] = b.y, item[6] = b.z, item[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) item[k] = base + k < count ? in[base + k] : 0u;
    }
}

__device__ __forceinline__ void store_items(uint32_t* __restrict__ out, uint32_t base, uint32_t count, const uint32_t item[kScanItems]) {
    if (base + kScanItems <= count) {
        *reinterpret_cast<uint4*>(out + base) = make_uint4(item[0], item[1], item[2], item[3]);
        *reinterpret_cast<uint4*>(out + base + 4) = make_uint4(item[4], item[5], item[6], item[7]);
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k)
            if (base + k < count) out[base + k] = item[k];
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const uint32_t* __restrict__ in, uint32_t count,
                                                                   uint2* __restrict__ block_sum) {
    __shared__ uint2 warp_sum[kScanThreads / 32];
    uint32_t base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
    uint32_t item[kScanItems];
    load_items(in, base, count, item);
    uint2 v = make_uint2(0, 0);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) v = add2(v, scan_item(item[k]));
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_down_sync(0xFFFFFFFFu, v.x, o);
        v.y += __shfl_down_sync(0xFFFFFFFFu, v.y, o);
    }
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint2 t = make_uint2(0, 0);
        for (int w = 0; w < kScanThreads / 32; ++w) t = add2(t, warp_sum[w]);
        block_sum[blockIdx.x] = t;
    }
}

// single block: exclusive scan of block_sum in place, totals -> *total_out (.x) and *pad_total_out (.y, if not null)
__global__ void __launch_bounds__(1024) scan_top_kernel(uint2* __restrict__ block_sum, uint32_t blocks,
                                                        uint32_t* __restrict__ total_out, uint32_t* __restrict__ pad_total_out) {
    __shared__ uint2 warp_sum[32];
    __shared__ uint2 carry;
    if (threadIdx.x == 0) carry = make_uint2(0, 0);
    __syncthreads();
    for (uint32_t base = 0; base < blocks; base += 1024) {
        uint32_t idx = base + threadIdx.x;
        uint2 v = idx < blocks ? block_sum[idx] : make_uint2(0, 0);
        uint2 incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint2 t = shfl_up2(incl, o);
            if ((threadIdx.x & 31) >= o) incl = add2(incl, t);
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint2 w = warp_sum[threadIdx.x], wi = w;
            for (int o = 1; o < 32; o <<= 1) {
                uint2 t = shfl_up2(wi, o);
                if (threadIdx.x >= o) wi = add2(wi, t);
            }
            warp_sum[threadIdx.x] = make_uint2(wi.x - w.x, wi.y - w.y);  // exclusive
        }
        __syncthreads();
        uint2 ws = warp_sum[threadIdx.x >> 5];
        uint2 excl = make_uint2(carry.x + ws.x + incl.x - v.x, carry.y + ws.y + incl.y - v.y);
        if (idx < blocks) block_sum[idx] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = add2(excl, v);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *total_out = carry.x;
        if (pad_total_out) *pad_total_out = carry.y;
    }
}

// PAD: also write pad_start.
template <bool PAD>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const uint32_t* __restrict__ in, uint32_t count,
                                                                  const uint2* __restrict__ block_offset,
                                                                  uint32_t* __restrict__ out, uint32_t* __restrict__ pad_out) {
    __shared__ uint2 warp_sum[kScanThreads / 32];
    uint32_t base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    load_items(in, base, count, v);
    uint2 t = make_uint2(0, 0);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) t = add2(t, scan_item(v[k]));
    uint2 incl = t;
    for (int o = 1; o < 32; o <<= 1) {
        uint2 u = shfl_up2(incl, o);
        if ((threadIdx.x & 31) >= o) incl = add2(incl, u);
    }
    if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint2 woff = make_uint2(0, 0);
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff = add2(woff, warp_sum[w]);
    const uint2 off = block_offset[blockIdx.x];
    uint2 run = make_uint2(off.x + woff.x + incl.x - t.x, off.y + woff.y + incl.y - t.y);
    uint32_t o0[kScanItems], o1[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        o0[k] = run.x;
        o1[k] = run.y;
        run = add2(run, scan_item(v[k]));
    }
    store_items(out, base, count, o0);
    if (PAD) store_items(pad_out, base, count, o1);
}

__global__ void scatter_kernel(Source src, Grid g, const uint32_t* __restrict__ cell_start,
                               const uint32_t* __restrict__ rank, uint32_t* __restrict__ perm) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= source_count(src)) return;
    uint2 pos;
    float2 vel;
    int32_t ty;
    bool is_record;
    if (!source_fetch(src, c, pos, vel, ty, is_record)) return;
    uint32_t key = cell_of(pos, g);
    if (key >= kKeyUp) return;
    perm[cell_start[key] + rank[c]] = c;
}

// Slots [p_lo, p_hi) of the sorted arrays are the owned rows; ghost rows are filled by the exchange.
__global__ void gather_kernel(Source src, uint32_t p_lo, uint32_t p_hi, Grid g,
                              const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ perm,
                              uint2* __restrict__ pos_out, float2* __restrict__ vel_out,
                              int32_t* __restrict__ ty_out, uint32_t* __restrict__ cell_id_out,
                              float4* __restrict__ nbr_out, PhysF pf) {
    uint32_t p = p_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= p_hi) return;
    uint32_t c = perm[p];
    uint2 pos;
    float2 vel;
    int32_t ty;
    bool is_record;
    source_fetch(src, c, pos, vel, ty, is_record);
    uint32_t key = cell_of(pos, g);
    uint32_t s = cell_start[key], e = cell_start[key + 1];
    uint32_t r = 0;
    for (uint32_t k = s; k < e; ++k) r += perm[k] < c ? 1u : 0u;
    uint32_t dst = s + r;
    pos_out[dst] = pos;
    vel_out[dst] = vel;
    ty_out[dst] = ty;
    cell_id_out[dst] = key;
    if (nbr_out) nbr_out[dst] = nbr_record(pos, key, g, pf);
}

// Neighbour records of particles [lo, hi) from their positions and membership cells (found in cell_start): ghost
// rows that arrived as bare positions, or everything after the metadata changed the scale.
__global__ void nbr_rebuild_kernel(const uint2* __restrict__ pos, const uint32_t* __restrict__ cell_start, Grid g,
                                   PhysF pf, uint32_t lo, uint32_t hi, float4* __restrict__ nbr) {
    const uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    nbr[i] = nbr_record(pos[i], (uint32_t)last_le(cell_start, (int)g.cells, i), g, pf);
}

// The few numbers the host needs after a binning: where the owned rows and their two boundary rows
// start and end in the sorted arrays. out[0] = own_lo, [1] = end of the first owned row,
// [2] = start of the last owned row, [3] = own_hi, [4] = total (with ghost rows), [5] = error flags,
// [6] = tiles of step_kernel_c, [7] = its tiles of the first owned row, [8] = its first tile of the last owned row,
// [9] = the tile table overflowed, [10], [11] = migrants of this re-bin to the lower / upper slab.
// The start of the upper ghost row is also published in the slab's HaloHeader for the lower neighbour's pushes.
constexpr uint32_t kErrHaloTimeout = 8u;  // a step waited 20 s for a neighbour's halo
__global__ void slab_counts_kernel(const uint32_t* __restrict__ cell_start, Grid g, const uint32_t* __restrict__ flags,
                                   const uint32_t* __restrict__ couple_tiles, const uint32_t* __restrict__ tile_base,
                                   HaloHeader* __restrict__ hdr, const uint32_t* __restrict__ mig_counters,
                                   uint32_t* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    out[10] = mig_counters ? mig_counters[0] : 0u;  // particles this re-bin found below / above the owned rows
    out[11] = mig_counters ? mig_counters[1] : 0u;
    out[6] = couple_tiles ? couple_tiles[0] : 0u;  // tiles of step_kernel_c (step_float.cuh)
    out[9] = couple_tiles ? couple_tiles[1] : 0u;  // ... did not fit the tile table
    out[7] = tile_base ? tile_base[1] : 0u;
    out[8] = tile_base ? tile_base[g.own_rows - 1] : 0u;
    if (hdr) {
        hdr->own_hi = cell_start[(g.own_row0 + g.own_rows) * g.bx];
        out[5] = flags[0] | (hdr->error ? kErrHaloTimeout : 0u);
        out[0] = cell_start[g.own_row0 * g.bx];
        out[1] = cell_start[(g.own_row0 + 1) * g.bx];
        out[2] = cell_start[(g.own_row0 + g.own_rows - 1) * g.bx];
        out[3] = hdr->own_hi;
        out[4] = cell_start[g.cells];
        return;
    }
    out[0] = cell_start[g.own_row0 * g.bx];
    out[1] = cell_start[(g.own_row0 + 1) * g.bx];
    out[2] = cell_start[(g.own_row0 + g.own_rows - 1) * g.bx];
    out[3] = cell_start[(g.own_row0 + g.own_rows) * g.bx];
    out[4] = cell_start[g.cells];
    out[5] = flags[0];
}

// Cell ids of a ghost row: its particles arrive as bare positions; the ids are needed by nobody (ghosts
// are never stepped) -- but the row's cell_start entries are, and those come from the scan.

// One descriptor per tile of kTile consecutive OWNED particles: the cell range of the tile, and for each
// of the three stencil rows the (16-byte aligned) slices of cell_start and of the position array that its
// CTA stages in shared memory (see step_kernel).
__global__ void tile_desc_kernel(const uint32_t* __restrict__ cell_start, Grid g, uint32_t own_lo, uint32_t own_hi,
                                 TileDesc* __restrict__ tiles) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t ntiles = (own_hi - own_lo + kTile - 1) / kTile;
    if (b >= ntiles) return;
    uint32_t i0 = own_lo + b * kTile, i1 = min(own_hi, i0 + kTile) - 1;
    TileDesc t;
    t.first = (uint32_t)last_le(cell_start, (int)g.cells, i0);
    t.last = (uint32_t)last_le(cell_start, (int)g.cells, i1);
    t._pad = 0;
    bool fits = true;
    for (int d = 0; d < 3; ++d) {
        long long shift = (long long)(d - 1) * (long long)g.bx;
        long long lo = max((long long)t.first - 1 + shift, 0ll);
        long long hi = min((long long)t.last + 1 + shift, (long long)g.cells - 1);
        if (hi < lo) {  // the whole row lies outside the grid
            t.cs_lo[d] = t.cs_cnt[d] = t.p_lo[d] = t.p_cnt[d] = 0;
            continue;
        }
        uint32_t cs_lo = (uint32_t)lo & ~3u;
        uint32_t cs_cnt = (((uint32_t)hi + 2 - cs_lo) + 3u) & ~3u;  // entries lo .. hi+1
        uint32_t p_lo = cell_start[lo] & ~1u;
        uint32_t p_cnt = ((cell_start[hi + 1] - p_lo) + 1u) & ~1u;
        t.cs_lo[d] = cs_lo;
        t.cs_cnt[d] = cs_cnt;
        t.p_lo[d] = p_lo;
        t.p_cnt[d] = p_cnt;
        fits = fits && cs_cnt <= (uint32_t)kCsCap && p_cnt <= (uint32_t)kPosCap;
    }
    t.fits = fits ? 1u : 0u;
    tiles[b] = t;
}

// snapshot: pack the owned particles back into wire-format records (particle.rs:10-18)
// With stride > 1 only every stride-th particle of the (cell-sorted, hence spatially coherent) state is packed: a
// decimated snapshot for display, 1/stride of the bytes to copy out and send.
__global__ void pack_kernel(const uint2* __restrict__ pos, const float2* __restrict__ vel,
                            const int32_t* __restrict__ ty, uint32_t lo, uint32_t hi, uint32_t stride,
                            Particle* __restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t i64 = (uint64_t)lo + (uint64_t)k * stride;
    if (i64 >= hi) return;
    const uint32_t i = (uint32_t)i64;
    uint2 p = pos[i];
    float2 v = vel[i];
    Particle q;
    q.x = p.x;
    q.y = p.y;
    q.vx = v.x;
    q.vy = v.y;
    q.ty = ty[i];
    out[k] = q;
}

// ------------------------------------------------------------------------------------------------
// Migration between slabs at re-bin time: owned particles whose position left the owned rows are
// collected (atomics, arbitrary order), then written to a fixed-size box in ascending index order
// (= the order they have in the global sorted array) for the neighbour to merge.
// ------------------------------------------------------------------------------------------------

struct MigrantBoxHeader {
    uint32_t count;
    uint32_t _pad[3];
};

// Migrants leave in ascending index order (the order they have in the global sorted array: what keeps slabs bit-identical
// to a single slab). Three passes, O(n / 1024 + migrants) whatever the migration rate: count the migrants of every block of
// 1024 consecutive particles per direction, scan the block counts, and let the blocks that have any rank theirs with ballots
// and write them to their place in the outbox. (Round 1 appended indices with an atomic and ranked each against all others:
// O(count^2), 2 ms for a full box.)
constexpr uint32_t kMigBlock = 1024;

__device__ __forceinline__ int migrant_dir(uint2 p, const Grid& g) {  // -1: stays, 0: down, 1: up
    const uint32_t key = cell_of(p, g);
    return key < kKeyUp ? -1 : (key == kKeyDown ? 0 : 1);
}

// blk_cnt[dir * nb + b]: migrants of block b (zero on entry: the pack kernel clears what it has used)
__global__ void __launch_bounds__(kMigBlock) migrant_count_kernel(const uint2* __restrict__ pos, uint32_t own_lo, uint32_t own_hi,
                                                                  Grid g, uint32_t nb, uint32_t* __restrict__ counters,
                                                                  uint32_t* __restrict__ blk_cnt, uint32_t* __restrict__ flags) {
    const uint32_t i = own_lo + blockIdx.x * kMigBlock + threadIdx.x;
    if (i >= own_hi) return;
    const uint2 p = pos[i];
    const int dir = migrant_dir(p, g);
    if (dir < 0) return;
    const int32_t row = (int32_t)(p.y >> g.sy) - g.row_offset;
    // anything beyond the adjacent slabs cannot be delivered
    if (row < (int32_t)g.own_row0 - (int32_t)g.rows_below || row >= (int32_t)(g.own_row0 + g.own_rows + g.rows_above))
        atomicOr(flags, kErrMigrantTooFar);
    atomicAdd(&counters[dir], 1u);
    atomicAdd(&blk_cnt[(uint32_t)dir * nb + blockIdx.x], 1u);
}

// blk_off = exclusive scan of blk_cnt, per direction, and blk_list = the blocks that have migrants at all (blk_list[0] =
// how many, then their indices in no particular order). One CTA; nb is a few thousand.
__global__ void __launch_bounds__(1024) migrant_scan_kernel(const uint32_t* __restrict__ blk_cnt, uint32_t nb, uint32_t box_capacity,
                                                            const uint32_t* __restrict__ counters, uint32_t* __restrict__ blk_off,
                                                            uint32_t* __restrict__ blk_list, uint32_t* __restrict__ flags) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_listed;
    if (threadIdx.x == 0) {
        s_listed = 0;
        if (counters[0] > box_capacity || counters[1] > box_capacity) atomicOr(flags, kErrMigrantOverflow);
    }
    __syncthreads();
    const uint32_t per = (nb + 1023u) / 1024u;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t b = min(threadIdx.x * per, nb); b < min(threadIdx.x * per + per, nb); ++b)
        if (blk_cnt[b] | blk_cnt[nb + b]) blk_list[1 + atomicAdd(&s_listed, 1u)] = b;
    __syncthreads();
    if (threadIdx.x == 0) blk_list[0] = s_listed;
    for (uint32_t dir = 0; dir < 2; ++dir) {
        const uint32_t* c = blk_cnt + dir * nb;
        uint32_t* o = blk_off + dir * nb;
        const uint32_t b0 = min(threadIdx.x * per, nb), b1 = min(b0 + per, nb);
        uint32_t sum = 0;
        for (uint32_t b = b0; b < b1; ++b) sum += c[b];
        uint32_t incl = sum;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if ((int)lane >= d) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane];
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, w, d);
                if ((int)lane >= d) w += v;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        uint32_t run = incl - sum + (warp ? s_warp[warp - 1] : 0u);
        for (uint32_t b = b0; b < b1; ++b) {
            o[b] = run;
            run += c[b];
        }
        __syncthreads();
    }
}

// CTAs [0, workers): the blocks of blk_list in turn, their migrants to their places in the two outboxes. CTAs [workers,
// workers + ceil(capacity / 1024)): the headers and the null records (ty < 0) behind the last migrant -- the receiver scans
// the whole box.
__global__ void __launch_bounds__(kMigBlock) migrant_pack_kernel(const uint2* __restrict__ pos, const float2* __restrict__ vel,
                                                                 const int32_t* __restrict__ ty, uint32_t own_lo, uint32_t own_hi,
                                                                 Grid g, uint32_t nb, uint32_t workers,
                                                                 const uint32_t* __restrict__ counters, uint32_t* __restrict__ blk_cnt,
                                                                 const uint32_t* __restrict__ blk_off,
                                                                 const uint32_t* __restrict__ blk_list, uint32_t box_capacity,
                                                                 unsigned char* __restrict__ box_down, unsigned char* __restrict__ box_up) {
    __shared__ uint32_t s_warp[2][32];
    if (blockIdx.x >= workers) {
        const uint32_t e = (blockIdx.x - workers) * kMigBlock + threadIdx.x;
        for (int dir = 0; dir < 2; ++dir) {
            unsigned char* box = dir ? box_up : box_down;
            const uint32_t count = min(counters[dir], box_capacity);
            if (e == 0) {
                MigrantBoxHeader* header = reinterpret_cast<MigrantBoxHeader*>(box);
                header->count = count;
                header->_pad[0] = header->_pad[1] = header->_pad[2] = 0;
            }
            if (e >= count && e < box_capacity) {
                Particle q;
                q.x = q.y = 0;
                q.vx = q.vy = 0.f;
                q.ty = -1;
                reinterpret_cast<Particle*>(box + sizeof(MigrantBoxHeader))[e] = q;
            }
        }
        return;
    }
    const uint32_t listed = blk_list[0];
    for (uint32_t k = blockIdx.x; k < listed; k += workers) {
        const uint32_t b = blk_list[1 + k];
        const uint32_t i = own_lo + b * kMigBlock + threadIdx.x;
        uint2 p = make_uint2(0, 0);
        int dir = -1;
        if (i < own_hi) {
            p = pos[i];
            dir = migrant_dir(p, g);
        }
        const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
        const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, dir == 0), m1 = __ballot_sync(0xFFFFFFFFu, dir == 1);
        if (lane == 0) {
            s_warp[0][warp] = __popc(m0);
            s_warp[1][warp] = __popc(m1);
        }
        __syncthreads();
        if (warp == 0) {
            for (int d = 0; d < 2; ++d) {
                const uint32_t own = s_warp[d][lane];
                uint32_t w = own;
                for (int k = 1; k < 32; k <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, w, k);
                    if ((int)lane >= k) w += v;
                }
                s_warp[d][lane] = w - own;  // exclusive
            }
        }
        __syncthreads();
        if (dir >= 0) {
            const uint32_t below = (1u << lane) - 1u;
            const uint32_t slot = blk_off[(uint32_t)dir * nb + b] + s_warp[dir][warp] + __popc((dir ? m1 : m0) & below);
            if (slot < box_capacity) {
                const float2 v = vel[i];
                Particle q;
                q.x = p.x;
                q.y = p.y;
                q.vx = v.x;
                q.vy = v.y;
                q.ty = ty[i];
                reinterpret_cast<Particle*>((dir ? box_up : box_down) + sizeof(MigrantBoxHeader))[slot] = q;
            }
        }
        if (threadIdx.x == 0) blk_cnt[b] = blk_cnt[nb + b] = 0;  // ready for the next re-bin
        __syncthreads();  // s_warp is reused by the next block of the list
    }
}
