// step_float.cuh -- the step kernel fine grids (>= 1024 cells per axis) run.  Included by stepper.cu inside its
// anonymous namespace.
//
// Three ideas on top of step_kernel (stepper.cu):
//
// 1. EXACT fp32 OFFSETS, KEPT IN HBM.  f_dist (particle.cuh:41-47) subtracts the u32 coordinates exactly and rounds
//    the difference to fp32 once.  An integer of magnitude < 2^24 is exact in fp32 and so is the difference of two
//    such numbers, so  float(xj - O) - float(xi - O) == float(xj - xi)  whenever both offsets from a common origin O
//    stay below 2^24 fixed-point units: with neighbours available as fp32 offsets the pair loop needs no integer
//    subtract and no int -> float conversion.  A cell is 2^(32-LX) units wide, so an origin serves only a few cells:
//      y: a particle's offset is taken from the centre of its membership cell row; a thread adds the (exact)
//         distance between row centres once per stencil row;
//      x: cell columns are grouped into ZONES of 2^zl columns (zl = 2 for LX >= 11, 1 for LX = 10).  A thread
//         whose cell is column cx works in zone z = max(cx-1, 0) >> zl; its three stencil columns lie within the
//         first 2^zl + 2 columns of that zone.  Every column belongs to exactly one even and one odd zone (z and
//         z-1 of its own zone z), so a particle carries two x offsets, from the centre of the used span of its even
//         zone and of its odd zone, and a thread reads the one with the parity of its own zone.
//    Offsets are pre-multiplied by a power of two (exact) that brings (r/sigma)^2 to within [1/4, 1) of its true
//    value; the remaining factor f in [1, 2) of kx/sigma is folded into the constants of the force law (PhysF).
//    The offsets depend only on a particle's position and membership cell, so they are produced by whoever
//    produces the position -- the previous step's epilogue, the re-bin's gather, the neighbour slab's halo push --
//    as a 16-byte NEIGHBOUR RECORD (x_even, y, x_odd, y) next to the position (nbr_record, stepper.cu).  Staging a
//    tile's stencil is then nothing but bulk copies (TMA); there is no conversion pass in this kernel.
//
// 2. TWO CELL-MATES PER THREAD.  A thread steps a COUPLE: two consecutive particles of one cell (a cell with an
//    odd count ends in a half-empty couple).  Both see the same stencil, so one neighbour is loaded once and the
//    packed fp32x2 pipe (FADD2 / FMUL2 / FFMA2) evaluates it against both: per neighbour 1 LDS.64, 16 packed
//    instructions and 4 MUFU for two pair forces, and all per-thread set-up is shared by two particles.
//
// 3. ROW-ALIGNED TILES.  A tile is up to 128 consecutive couples of ONE cell row (TileC, built at re-bin time); its
//    stencil is three contiguous ranges of the sorted arrays (the rows below, of, and above it, its columns +- 1).
//    The accumulation order is the reference's (row, column, slot) order.

constexpr int kColCap = 128;   // cell columns a tile's stencil rows may span (its own columns +- 1)
constexpr int kCsRow = 136;    // cell_start entries staged per row: kColCap + 1, + 3 alignment slack, multiple of 4
#ifndef PSIM_ROW_CAP
#define PSIM_ROW_CAP 448
#endif
constexpr int kRowCap = PSIM_ROW_CAP;   // neighbour records staged per stencil row (a crystal at r0 has rows of 6 per cell)
constexpr int kCouples = 128;  // couples (= threads) per tile
#ifndef PSIM_MIN_CTAS
#define PSIM_MIN_CTAS 9  // CTAs per SM the register allocation aims at (56 registers per thread)
#endif
#ifndef PSIM_EARLY_COUPLE
#define PSIM_EARLY_COUPLE 1  // the couple is loaded before warp 0 turns to the bulk copies (0: after, the round-1 order)
#endif
#ifndef PSIM_LATE_LOADS
#define PSIM_LATE_LOADS 2  // own position / velocity: 0 loaded up front, 1 after the pair loops, 2 before the last loop
#endif
#ifndef PSIM_OPAQUE_CONSTS
#define PSIM_OPAQUE_CONSTS 1
#endif
#ifndef PSIM_PAIR_UNROLL
#define PSIM_PAIR_UNROLL 1
#endif
constexpr int kPairUnroll = PSIM_PAIR_UNROLL;  // neighbours per trip of the pair loops

// One tile: couples [k0, k0 + nk) of local cell row `row`.
struct __align__(16) TileC {
    uint32_t k0, nk;
    uint32_t row;
    uint32_t fits;        // bytes staged in shared memory; 0: a stencil row exceeds the staging buffers -> global-memory path
    uint32_t cs_lo[3];    // first cell_start entry staged per row (multiple of 4)
    uint32_t cs_cnt[3];   // entries staged per row (multiple of 4; 0: row outside the grid)
    uint32_t p_lo[3];     // first particle staged per row (even)
    uint32_t p_cnt[3];    // particles staged per row (even)
};
static_assert(sizeof(TileC) == 64, "TileC is read as four 16-byte words");

struct StepArgsC {
    const uint2* __restrict__ couple_i0;  // per couple: (index of its first particle | (has a second one) << 31, its cell)
    const TileC* __restrict__ tiles;
    const uint32_t* __restrict__ n_tiles;  // tiles of the last binning, in device memory (read by the surplus launch only)
};

// One staged neighbour (its x offset of the thread's zone parity, its y) against the thread's two particles.
// nx, ny: minus their own offsets.  g f^(2 km) = qs^4 - (n/m) f^(-2(kn-4)) qs^KN q^fn  with qs = 1 / (scaled r^2);
// see make_phys_f().
struct PolyC {  // the cubic's coefficients, held in registers across the pair loops (step_tile_c)
    float d0, d1, d2, d3;
};

template <int KN, int FRAC, bool CLAMP>
__device__ __forceinline__ void pairc(float xj, float yj, float2 nx, float2 ny, const PolyC& pf, float2& gx, float2& gy) {
    float2 x = __fadd2_rn(nx, splat(xj));
    float2 y = __fadd2_rn(ny, splat(yj));
    float2 r2 = __ffma2_rn(y, y, __fmul2_rn(x, x));
    if (CLAMP) {  // the own row contains the particle itself: an exact 0 instead of 0 * inf
        r2.x = fmaxf(r2.x, 1e-3f);
        r2.y = fmaxf(r2.y, 1e-3f);
    }
#ifdef PSIM_SHARED_RCP  // one MUFU.RCP for both lanes: 1 / (a b), times b and a (A/B switch)
    const float rr = fast_rcp(r2.x * r2.y);
    float2 q = make_float2(r2.y * rr, r2.x * rr);
#else
    float2 q = make_float2(fast_rcp(r2.x), fast_rcp(r2.y));
#endif
    float2 q2 = __fmul2_rn(q, q);
    float2 q4 = __fmul2_rn(q2, q2);
    float2 pn;
    if (KN == 5) pn = __fmul2_rn(q4, q);
    else if (KN == 6) pn = __fmul2_rn(q4, q2);
    else if (KN == 7) pn = __fmul2_rn(__fmul2_rn(q4, q2), q);
    else if (KN == 8) pn = __fmul2_rn(q4, q4);
    else if (KN == 9) pn = __fmul2_rn(__fmul2_rn(q4, q4), q);
    else pn = __fmul2_rn(__fmul2_rn(q4, q4), q2);
    float2 e;
    if (FRAC == kFracPoly) {
        float2 l = make_float2(fast_lg2(r2.x), fast_lg2(r2.y));
        e = __ffma2_rn(l, splat(pf.d3), splat(pf.d2));
        e = __ffma2_rn(l, e, splat(pf.d1));
        e = __ffma2_rn(l, e, splat(pf.d0));
    } else {
        e = splat(pf.d0);
    }
    float2 g = __ffma2_rn(pn, e, q4);
    gx = __ffma2_rn(g, x, gx);
    gy = __ffma2_rn(g, y, gy);
}

struct SmemC {
    float4 nb[3][kRowCap];     // neighbour records of the three stencil rows
    uint32_t cs[3][kCsRow];    // their cell_start slices
    uint64_t bar;
};

// One tile: stage its stencil, step its couples. MAIN: called by step_kernel_c itself (whose first CTA publishes the
// epoch of a boundary row without particles), not for a surplus tile.
template <int KN, int FRAC, bool MAIN>
__device__ __forceinline__ void step_tile_c(const StepArgs& a, const StepArgsC& ac, uint32_t tile, SmemC& sm) {
    float4 (&s_nb)[3][kRowCap] = sm.nb;
    uint32_t (&s_cs)[3][kCsRow] = sm.cs;
    uint64_t& s_bar = sm.bar;

    const Grid& g = a.g;
    const PhysF& pf = a.pf;
    const TileC t = ac.tiles[tile];
    const bool live = threadIdx.x < t.nk;

    if (a.push) {  // uniform over the grid
        if (threadIdx.x == 0) {
            if (MAIN && blockIdx.x == 0) halo_publish_empty(a);
            halo_wait(a, tile);  // tiles next to a ghost row: the neighbour's last step has landed
        }
        if (!t.fits) __syncthreads();  // the global-memory path reads the ghost rows directly
    }
    if (!t.fits) {  // very sparse or very crowded spot: same physics straight from global memory, one particle at a time
        if (!live) return;
        const uint32_t w = ac.couple_i0[t.k0 + threadIdx.x].x;
        const uint32_t i0 = w & 0x7FFFFFFFu;
        const uint32_t* cs[3] = {a.cell_start, a.cell_start, a.cell_start};
        const uint2* pp[3] = {a.pos_in, a.pos_in, a.pos_in};
        const uint32_t zero[3] = {0, 0, 0};
        for (uint32_t i = i0; i <= i0 + (w >> 31); ++i)
            step_particle<KN, FRAC, true, true>(i, a.pos_in[i], a.vel[i], a.cell_id[i], cs, zero, pp, zero, a);
        return;
    }

#if PSIM_EARLY_COUPLE
    // this thread's couple: the load is issued BEFORE warp 0 turns to the bulk copies, so that warp 0's own couple does not
    // arrive a whole L2 latency after everybody else's (the other three warps were waiting for it at the barrier below:
    // 7 % of the warp samples of the round's ncu capture)
    uint2 cw = make_uint2(a.own_lo, t.row << g.lx);
    if (live) cw = ac.couple_i0[t.k0 + threadIdx.x];
#endif

    // TMA: the cell_start slices and the neighbour records of the three stencil rows (through L2: a ghost row is
    // written by the neighbour slab while this kernel runs). Six lanes of warp 0 issue one 1-D bulk copy each.
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar, 1);
            mbar_arrive_expect_tx(&s_bar, t.fits);
        }
        __syncwarp();
        if (threadIdx.x < 6) {
            // ghost rows were written by the neighbour GPU through the generic proxy; the bulk copies read through the
            // async proxy: order them after thread 0's acquire in halo_wait
            if (a.push) asm volatile("fence.proxy.async.global;" ::: "memory");
            const uint32_t d = threadIdx.x >> 1;
            const uint32_t* tw = reinterpret_cast<const uint32_t*>(ac.tiles + tile);  // cs_lo @4, cs_cnt @7, p_lo @10, p_cnt @13
            if (threadIdx.x & 1) {
                const uint32_t lo = tw[10 + d], cnt = tw[13 + d];
                if (cnt) bulk_copy_g2s(s_nb[d], a.nbr_in + lo, cnt * 16u, &s_bar);
            } else {
                const uint32_t lo = tw[4 + d], cnt = tw[7 + d];
                if (cnt) bulk_copy_g2s(s_cs[d], a.cell_start + lo, cnt * 4u, &s_bar);
            }
        }
    }

    // this thread's couple
#if PSIM_EARLY_COUPLE
    const uint32_t i0 = cw.x & 0x7FFFFFFFu, has1 = cw.x >> 31, cell = cw.y;  // own_lo has no bit 31
#else
    uint32_t i0 = a.own_lo, has1 = 0, cell = t.row << g.lx;
    if (live) {
        const uint2 w = ac.couple_i0[t.k0 + threadIdx.x];
        i0 = w.x & 0x7FFFFFFFu;
        has1 = w.x >> 31;
        cell = w.y;
    }
#endif
    const uint32_t i1 = i0 + has1;  // a half-empty couple computes its only particle twice
#if PSIM_LATE_LOADS
    // positions and velocities are needed by the epilogue only: fetched into L1 now, read after the pair loops (eight
    // registers that would otherwise stay live across the loops and push the cubic's constants out of the register file)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.pos_in + i0));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.vel + i0));
#else
    // positions and velocities are needed by the epilogue only: these loads fly during the pair loops
    const uint2 p0 = a.pos_in[i0], p1 = a.pos_in[i1];
    const float2 v0 = a.vel[i0], v1 = a.vel[i1];
#endif
    const uint32_t cx = cell & (g.bx - 1);
    const uint32_t x0c = cx == 0 ? 0 : cx - 1, x1c = cx == g.bx - 1 ? cx : cx + 1;
    const uint32_t par = (x0c >> pf.zl) & 1u;  // parity of the thread's zone: which x offset of a record it reads

    __syncthreads();  // the barrier is initialised before anyone polls it
    mbar_wait(&s_bar, 0);
    if (!live) return;

    // own offsets: the thread's two particles are records of the own row too (x of the same parity: their cell is
    // one of the thread's stencil columns)
    const float2* own = reinterpret_cast<const float2*>(s_nb[1]) + par;
    const float2 o0 = own[2 * (i0 - t.p_lo[1])], o1 = own[2 * (i1 - t.p_lo[1])];
    const float2 nx = make_float2(-o0.x, -o1.x), ny0 = make_float2(-o0.y, -o1.y);

    // The cubic's four coefficients are read ONCE into registers through an opaque move: left to itself the compiler
    // re-reads them from the constant bank in every trip of two of the three loops (two LDC.64 of 24 instructions).
    PolyC pc;
#if PSIM_OPAQUE_CONSTS
    const float zero = __int_as_float(t.row & 0x80000000u);  // +0.0 (a row index has no bit 31), but only at run time
    pc.d0 = pf.d0 + zero;
    pc.d1 = pf.d1 + zero;
    pc.d2 = pf.d2 + zero;
    pc.d3 = pf.d3 + zero;
#else
    pc.d0 = pf.d0, pc.d1 = pf.d1, pc.d2 = pf.d2, pc.d3 = pf.d3;
#endif
    float2 gx = splat(0.f), gy = splat(0.f);
#if PSIM_LATE_LOADS == 2
    uint2 p0, p1;
    float2 v0, v1;
#endif
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#if PSIM_LATE_LOADS == 2
        if (d == 2) {  // the L1 prefetch above rarely survives two loops of nine CTAs (4 % of the warp samples sat at the
                       // loads behind the loops): the real loads fly during the last loop
            p0 = a.pos_in[i0], p1 = a.pos_in[i1];
            v0 = a.vel[i0], v1 = a.vel[i1];
        }
#endif
        if (t.cs_cnt[d] == 0) continue;  // row outside the grid (uniform)
        const uint32_t rowbase = ((t.row + d - 1) << g.lx) - t.cs_lo[d];
        const uint32_t ws = s_cs[d][rowbase + x0c] - t.p_lo[d], we = s_cs[d][rowbase + x1c + 1] - t.p_lo[d];
        // a neighbour in the row below / above sits one row distance further down / up than its own-row offset says
        const float2 ny = __fadd2_rn(ny0, splat((float)(d - 1) * pf.row_shift));
        // (x_even, y) or (x_odd, y): the 8 bytes at offset 0 or 8 of a record. The loop runs on the 32-bit shared-memory
        // address itself (one add, one compare per trip; a generic pointer gets a second counter next to it)
        uint32_t pa = smem_u32(s_nb[d] + ws) + par * 8u;
        const uint32_t pa_end = smem_u32(s_nb[d] + we);
#pragma unroll kPairUnroll
        for (; pa < pa_end; pa += 16u) {
            float2 j;
            asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(j.x), "=f"(j.y) : "r"(pa));
            if (d == 1) pairc<KN, FRAC, true>(j.x, j.y, nx, ny, pc, gx, gy);
            else pairc<KN, FRAC, false>(j.x, j.y, nx, ny, pc, gx, gy);
        }
    }
#if PSIM_LATE_LOADS == 1
    const uint2 p0 = a.pos_in[i0], p1 = a.pos_in[i1];
    const float2 v0 = a.vel[i0], v1 = a.vel[i1];
#endif
    const RecOrigin org = rec_origin(cell, g, pf);  // both particles are members of the same cell
    finish_particle<true>(i0, p0, v0, cell, gx.x, gy.x, pf.pair_scale, pf.pair_scale, a, &org);  // KN > 0 implies m == 6
    if (has1) finish_particle<true>(i1, p1, v1, cell, gx.y, gy.y, pf.pair_scale, pf.pair_scale, a, &org);
}

// The grid is the tile count the host last saw plus a margin (exactly the count when it has just read it: slabs). The
// binning writes EMPTY descriptors (nk = 0) behind the last tile up to the size of the launch, so a surplus CTA leaves as
// soon as it has read its descriptor and nobody waits for the count. The kernel steps exactly one tile and ends with the
// tile body: any code BEHIND it -- a loop over more tiles, a call, even one never taken -- was measured at +2 to +4 %
// (threads can no longer exit early, branches lose their reconvergence hints: tools/ab_step.py, profiles/r02_ab_step.txt).
template <int KN, int FRAC>
__global__ void __launch_bounds__(kCouples, PSIM_MIN_CTAS) step_kernel_c(const StepArgs a, const StepArgsC ac) {
    __shared__ __align__(16) SmemC sm;
    asm volatile("griddepcontrol.wait;" ::: "memory");  // a no-op unless the launch allowed programmatic dependent launch
    step_tile_c<KN, FRAC, true>(a, ac, halo_tile_order(a, blockIdx.x, gridDim.x), sm);
}

// Surplus TILES: the count grew past the launch, which the host's margin makes rare. A single slab follows every step
// launch with this one (a few CTAs that read the count from device memory and, almost always, leave at once: 2.6 us):
// tiles [first, count) in turn.
constexpr uint32_t kSurplusCtas = 128;

template <int KN, int FRAC>
__global__ void __launch_bounds__(kCouples) step_kernel_c_surplus(const StepArgs a, const StepArgsC ac, uint32_t first) {
    __shared__ __align__(16) SmemC sm;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t count = *ac.n_tiles;
    for (uint32_t tile = first + blockIdx.x; tile < count; tile += gridDim.x) {
        step_tile_c<KN, FRAC, false>(a, ac, tile, sm);
        __syncthreads();  // everybody has left the staging buffers and the barrier before the next tile re-arms them
        if (threadIdx.x == 0) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&sm.bar)) : "memory");
        __syncthreads();
    }
}

// ---- re-bin side of the couples ------------------------------------------------------------------------

// couple_i0[k] for the couples of every cell (the cell is the one of this binning, i.e. the membership cell of the
// particles until the next one); couples are numbered by pad_start / 2.
__global__ void couple_build_kernel(const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ pad_start,
                                    uint32_t cells, uint2* __restrict__ couple_i0) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    const uint32_t s = cell_start[c], n = cell_start[c + 1] - s;
    uint32_t k = pad_start[c] >> 1;
    for (uint32_t m = 0; m < n; m += 2, ++k) couple_i0[k] = make_uint2((s + m) | (m + 1 < n ? 0x80000000u : 0u), c);
}

// Cutting a row's couples into tiles (one thread per row, twice: to count the tiles, then to write their spans).
//
// The row is looked at in aligned blocks of kTileCols columns. A block is DENSE when it holds at least kSparseCouples
// couples, or holds anything at all next to such a block (the rim of a droplet belongs to the droplet). A maximal run of
// dense blocks is cut like this: enough tiles for 128 couples per tile, for at most kTileColsMax occupied columns per
// tile, AND for each of the three stencil rows over the run's columns to fit the staging buffer (a thin row -- one couple
// per cell -- next to a thick one would otherwise make tiles whose neighbour rows overflow it); the run's couples are
// split evenly over them. A gas at 2.4 per cell fills its 128 threads this way (89 columns per tile), a crystal is cut by
// the couples (58 columns). A run of SPARSE blocks (thin gas in an otherwise empty box) is not worth staging --
// column-capped tiles would be CTAs with a handful of live threads, and their number would follow the area of the box
// rather than the particles -- so it gets full tiles as wide as they come; tile_build_kernel finds that those do not fit
// the staging buffers and they run from global memory.
// A row that is dense from its first particle to its last (a crystal, a liquid) is one run.
constexpr int kTileCols = 64;             // block width of the dense / sparse classification
#ifndef PSIM_TILE_COLS_MAX
#define PSIM_TILE_COLS_MAX (kColCap - 2)
#endif
constexpr int kTileColsMax = PSIM_TILE_COLS_MAX; // occupied columns a staged tile may span (its stencil rows: +- 1 column)
#ifndef PSIM_ROW_FILL
#define PSIM_ROW_FILL (PSIM_ROW_CAP - 64)
#endif
constexpr uint32_t kRowFill = PSIM_ROW_FILL;  // staged particles per stencil row a cut aims at (kRowCap with slack for uneven rows)
constexpr uint32_t kSparseCouples = 32;

template <bool EMIT>
__device__ __forceinline__ uint32_t cut_run(const uint32_t* __restrict__ cell_start, const Grid& g, uint32_t row, uint32_t b0,
                                            uint32_t b1, uint32_t k0, uint32_t k1, bool dense, uint32_t first_tile,
                                            TileC* __restrict__ tiles, uint32_t tiles_cap) {
    const uint32_t m = k1 - k0;
    if (m == 0) return 0;
    uint32_t n = (m + kCouples - 1) / kCouples;
    if (dense) {
        const uint32_t c0 = row << g.lx;
        const uint32_t* cs = cell_start + c0 + b0 * kTileCols;
        const int ncols = (int)((b1 - b0) * kTileCols);
        // first occupied column: the last cell whose start is still the run's; last occupied: the last cell that starts
        // below the run's end
        const uint32_t first = b0 * kTileCols + (uint32_t)last_le(cs, ncols, cs[0]);
        const uint32_t last = b0 * kTileCols + (uint32_t)last_le(cs, ncols, cs[ncols] - 1);
        n = max(n, (last - first + kTileColsMax) / kTileColsMax);
        // the three stencil rows over the run's columns +- 1: enough tiles for each to fit the staging buffer when the run's
        // couples are split evenly (a thin row next to a thick one must not make tiles whose neighbour rows overflow)
        const uint32_t lo = first == 0 ? 0 : first - 1, hi = min(last + 1, g.bx - 1);
#pragma unroll
        for (int d = -1; d <= 1; ++d) {
            const int rd = (int)row + d;
            if (rd < 0 || rd >= (int)g.by) continue;
            const uint32_t* rs = cell_start + ((uint32_t)rd << g.lx);
            n = max(n, (rs[hi + 1] - rs[lo] + kRowFill - 1) / kRowFill);
        }
    }
    if (EMIT) {
        const uint32_t per = (m + n - 1) / n;  // <= kCouples
        for (uint32_t t = 0; t < n && first_tile + t < tiles_cap; ++t) {
            const uint32_t k = k0 + t * per;
            tiles[first_tile + t].k0 = k;
            tiles[first_tile + t].nk = k < k1 ? min(k1 - k, per) : 0u;  // the even split can leave the last tile empty
        }
    }
    return n;
}

// EMIT = false: tile_base[r] = number of tiles of owned row r (row_tiles_kernel turns the counts into bases in place).
// EMIT = true: tiles[tile_base[r] ..] get their k0 / nk.
// One WARP per row: the lanes fetch the couple index at every block boundary at once (a row of 2048 columns has 33 of them;
// read one after the other by a single thread they were 33 dependent trips to L2), then lane 0 walks the blocks.
constexpr int kCutRowsPerCta = 4;
constexpr int kCutMaxBlocks = 32768 / kTileCols;  // grid_x_log2 <= 15

template <bool EMIT>
__global__ void __launch_bounds__(32 * kCutRowsPerCta) row_cut_kernel(const uint32_t* __restrict__ cell_start,
                                                                      const uint32_t* __restrict__ pad_start, Grid g,
                                                                      uint32_t* __restrict__ tile_base, TileC* __restrict__ tiles,
                                                                      uint32_t tiles_cap) {
    __shared__ uint32_t s_kb[kCutRowsPerCta][kCutMaxBlocks + 1];  // first couple of every block of the row, and the row's end
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t r = blockIdx.x * kCutRowsPerCta + warp;
    if (r >= g.own_rows) return;
    const uint32_t c0 = (g.own_row0 + r) << g.lx;
    const uint32_t nb = g.bx / kTileCols;
    uint32_t* kb = s_kb[warp];
    for (uint32_t b = lane; b <= nb; b += 32) kb[b] = pad_start[c0 + b * kTileCols] >> 1;
    __syncwarp();
    if (lane != 0) return;
    const uint32_t base = EMIT ? tile_base[r] : 0u;
    uint32_t cnt_prev = 0, cnt_cur = kb[1] - kb[0];
    uint32_t run_b0 = 0, n = 0;
    bool run_dense = false;
    for (uint32_t b = 0; b < nb; ++b) {
        const uint32_t cnt_next = b + 1 < nb ? kb[b + 2] - kb[b + 1] : 0u;
        const bool dense = cnt_cur >= kSparseCouples || (cnt_cur > 0 && (cnt_prev >= kSparseCouples || cnt_next >= kSparseCouples));
        if (b == 0) {
            run_dense = dense;
        } else if (dense != run_dense) {
            n += cut_run<EMIT>(cell_start, g, g.own_row0 + r, run_b0, b, kb[run_b0], kb[b], run_dense, base + n, tiles, tiles_cap);
            run_b0 = b;
            run_dense = dense;
        }
        cnt_prev = cnt_cur;
        cnt_cur = cnt_next;
    }
    n += cut_run<EMIT>(cell_start, g, g.own_row0 + r, run_b0, nb, kb[run_b0], kb[nb], run_dense, base + n, tiles, tiles_cap);
    if (!EMIT) tile_base[r] = n;
}

// Tiles of every owned row (row_cut_kernel<false> left the counts in tile_base); single block: exclusive scan in place
// into tile_base[0..own_rows]; total (at most the room in the tile table) -> total_out[0], total_out[1] |= it was more.
// `host_mirror`: the same two words in mapped host memory (the host reads them at its next wait, without a copy).
__global__ void __launch_bounds__(1024) row_tiles_kernel(Grid g, uint32_t* __restrict__ tile_base, uint32_t tiles_cap,
                                                         uint32_t* __restrict__ total_out, uint32_t* __restrict__ host_mirror) {
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < g.own_rows; base += 1024) {
        const uint32_t r = base + threadIdx.x;
        uint32_t v = 0;
        if (r < g.own_rows) v = tile_base[r];
        uint32_t incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += u;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warp_sum[threadIdx.x], wi = w;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t u = __shfl_up_sync(0xFFFFFFFFu, wi, o);
                if (threadIdx.x >= o) wi += u;
            }
            warp_sum[threadIdx.x] = wi - w;
        }
        __syncthreads();
        const uint32_t excl = carry + warp_sum[threadIdx.x >> 5] + incl - v;
        if (r < g.own_rows) tile_base[r] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        tile_base[g.own_rows] = carry;
        total_out[0] = min(carry, tiles_cap);
        if (carry > tiles_cap) total_out[1] = 1u;
        host_mirror[0] = min(carry, tiles_cap);
        host_mirror[1] = total_out[1];
    }
}

// One TileC per tile index b in [0, n_tiles[0]); row_cut_kernel<true> has written its k0 / nk.
__global__ void tile_build_kernel(const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ pad_start,
                                  const uint32_t* __restrict__ tile_base, const uint32_t* __restrict__ n_tiles,
                                  uint32_t launch, const uint2* __restrict__ couple_i0, Grid g,
                                  TileC* __restrict__ tiles) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t count = n_tiles[0];
    if (b >= count) {  // behind the last tile, as far as a step launches CTAs: empty descriptors
        if (b < launch) {
            TileC t;
            t.k0 = t.nk = t.row = t.fits = 0;
            for (int d = 0; d < 3; ++d) t.cs_lo[d] = t.cs_cnt[d] = t.p_lo[d] = t.p_cnt[d] = 0;
            tiles[b] = t;
        }
        return;
    }
    // the last row whose base is <= b (rows without tiles share their successor's base and are skipped over)
    const uint32_t r = (uint32_t)last_le(tile_base, (int)g.own_rows, b);
    const uint32_t row = g.own_row0 + r;
    const uint32_t nk = tiles[b].nk;  // as row_cut_kernel<true> left it: <= kCouples
    TileC t;
    t.k0 = tiles[b].k0;
    t.nk = nk;
    t.row = row;
    if (nk == 0) {  // nothing to step: no staging either
        t.fits = 0;
        for (int d = 0; d < 3; ++d) t.cs_lo[d] = t.cs_cnt[d] = t.p_lo[d] = t.p_cnt[d] = 0;
        tiles[b] = t;
        return;
    }
    const uint32_t c_first = couple_i0[t.k0].y & (g.bx - 1);
    const uint32_t c_last = couple_i0[t.k0 + nk - 1].y & (g.bx - 1);
    const uint32_t col_lo = c_first == 0 ? 0 : c_first - 1;
    const uint32_t col_hi = c_last == g.bx - 1 ? c_last : c_last + 1;
    bool fits = col_hi - col_lo + 1 <= (uint32_t)kColCap;
    uint32_t bytes = 0;
    for (int d = 0; d < 3; ++d) {
        const long long rd = (long long)row + d - 1;
        if (rd < 0 || rd >= (long long)g.by) {
            t.cs_lo[d] = t.cs_cnt[d] = t.p_lo[d] = t.p_cnt[d] = 0;
            continue;
        }
        const uint32_t lo = ((uint32_t)rd << g.lx) + col_lo, hi = ((uint32_t)rd << g.lx) + col_hi;
        t.cs_lo[d] = lo & ~3u;
        t.cs_cnt[d] = ((hi + 2 - t.cs_lo[d]) + 3u) & ~3u;  // entries lo .. hi+1
        t.p_lo[d] = cell_start[lo] & ~1u;
        t.p_cnt[d] = ((cell_start[hi + 1] - t.p_lo[d]) + 1u) & ~1u;
        fits = fits && t.cs_cnt[d] <= (uint32_t)kCsRow && t.p_cnt[d] <= (uint32_t)kRowCap;
        bytes += t.cs_cnt[d] * 4u + t.p_cnt[d] * 16u;
    }
    t.fits = fits ? bytes : 0u;
    tiles[b] = t;
}

