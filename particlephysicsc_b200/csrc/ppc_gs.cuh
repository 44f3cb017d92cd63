// ppc_gs.cuh -- K7 of the "resolve"=3 path: narrow phase, position pushes and a bounded-pass, deterministic Gauss-Seidel
// velocity resolve, tile by tile in shared memory.  No per-particle contact lists in global memory, no cooperative launch,
// no grid-wide dependency rounds: one launch of k_gs_pairs and four launches (one per tile colour) of k_gs_resolve per substep.
//
// Relation to the reference (src/particles_solver.c):
//  * candidate pairs: the reference's (3x3 cell neighbourhood :331-340, AABB reject and inclusive circle test :352-362,
//    same float operations) -- as in k_collide_tile2 (ppc_collide.cuh), whose fine x-bins k_gs_pairs shares;
//  * contact geometry frozen before anything moves (pass 1, :84-127); position pushes (:145-156) summed per particle in
//    the order the reference's pair list touches it, so POSITIONS ARE BIT-IDENTICAL to the reference;
//  * velocity impulses (:158-190): every pair is applied exactly once, from the CURRENT velocities of its two particles,
//    by the reference's formulas -- i.e. the reference's sequential pass 2 with the pair list in a different,
//    deterministic order.  The reference's own results move by as much when its list is reversed or shuffled (SURVEY.md
//    section 8c); oracle/ppc_oracle.c:ppc_oracle_tilegs_order is the CPU statement of exactly this order, and
//    tests/test_gpu_gs.py checks the device against it bit for bit.
//
// The order (see also the oracle):
//  1. Tiles of GS_R cell rows x wc cell columns, anchored at the whole field's cell (0, 0), coloured
//     (tx & 1) + 2 (ty & 1).  Tiles of one colour touch disjoint particles (a tile touches its cells and a one-cell halo;
//     tiles are >= 2 cells wide).  A tile whose staged particles would overflow shared memory is cut into column strips
//     (greedily from the left, as wide as fits), which its CTA runs one after the other.
//  2. A pair belongs to the sub-tile (tile, strip) holding both particles, else to the EARLIER of the two sub-tiles
//     (lower colour; inside a tile the lower strip).  The owner applies the impulse to both particles; what it did to a
//     halo particle is in global memory before that particle's own sub-tile runs.
//  3. Inside a sub-tile the pairs are EDGE-COLOURED greedily: they are visited in a fixed order -- nine phases of disjoint cell
//     pairs (pairs inside one cell; pairs with the east neighbour cell, base column even / odd; with the south, south-east,
//     south-west neighbour, base row even / odd); inside a cell pair by listing particle (ascending index), then by position in
//     its contact list -- and each takes the lowest pair-colour (0..31) that no earlier pair has used at either of its particles.
//  4. The sub-tile's pairs fire colour by colour.  Pairs of one colour share no particle, so their mutual order is immaterial.
//     (Round 2 first fired them by "level" = 1 + the level of the previous pair of either particle in the visiting order: the
//     same passes, but ~14 dependent steps per sub-tile of a dense pile where the colouring needs ~8.)
//
// Execution.  Positions do not depend on velocities, so the work splits into two kernels:
//   k_gs_pairs   (all tiles at once)  stages a tile and its halo, finds the contacts of every centre particle, adds its
//                position pushes in the reference's order, and writes the pairs the sub-tile owns as 16-byte records
//                {particle numbers, unit normal, Baumgarte term} into one contiguous chunk per sub-tile, sorted by pair-colour.
//   k_gs_resolve (one launch per tile colour)  loads the current velocities of a sub-tile and its halo, then fires the records
//                colour by colour, one thread per pair, and writes back what changed: the sequential pass 2 of the reference
//                over the pair list in the order (tile colour, sub-tile, pair-colour).
//
// Data movement: the canonical gather (ppc_grid.cuh) leaves, besides the ordinary sorted state, one 16-byte record
// {x, y, radius, API index} and one 16-bit word {mass byte, cell column byte} per particle in sorted order, so the GS_R + 2
// slot runs of a tile and its halo are contiguous byte ranges, staged by the TMA (cp.async.bulk + mbarrier, ppc_tma.cuh):
// no load instructions and no address arithmetic per staged particle.
#pragma once
#include "ppc_collide.cuh"
#include "ppc_tma.cuh"

namespace ppc {

#ifndef PPC_GS_THREADS
#define PPC_GS_THREADS 512
#endif
#ifndef PPC_GS_MINBLOCKS
#define PPC_GS_MINBLOCKS 3
#endif
#ifndef PPC_GS_CAP
#define PPC_GS_CAP 1024    // staged particles (centre + halo) per sub-tile
#endif
#ifndef PPC_GS_POOL
#define PPC_GS_POOL 6144   // contact entries per sub-tile (a centre-centre pair takes two)
#endif
#ifndef PPC_GS_TARGET
#define PPC_GS_TARGET 960  // staged particles (tile + halo) per tile of average density the host aims for
#endif
#ifndef PPC_GS_FLAT
#define PPC_GS_FLAT 0      // 1: the three staged rows a particle scans are one loop (measured 10-20 % SLOWER: the state machine costs more than the lanes it keeps together); 0: one loop per row
#endif
#ifndef PPC_GS_TMA
#define PPC_GS_TMA 1       // 0: stage with ordinary loads (debugging aid)
#endif
#ifndef PPC_GSB_THREADS
#define PPC_GSB_THREADS 128
#endif
#ifndef PPC_GSB_NBUF
#define PPC_GSB_NBUF 2     // slices of the record chunk in flight in k_gs_resolve.  The slice length is chosen per launch (GsLaunch::slice,
#endif                     // dynamic shared memory): measured on the 16M pile 0.81 / 0.74 ms with 256 / 512 records per slice (a short slice
                           // lasts about two levels, less than the latency of the copy that refills its buffer), but 0.24 / 0.32 ms on the
                           // dilute uniform field, whose tiles own a few dozen pairs and want many resident CTAs instead; deeper rings
                           // (3, 4, 6, 8 buffers) cost more residency than the latency they hide on both
constexpr int      GS_THREADS  = PPC_GS_THREADS;
constexpr int      GSB_THREADS = PPC_GSB_THREADS;
constexpr int      GSB_NBUF    = PPC_GSB_NBUF;
static_assert( GSB_NBUF >= 2 && GSB_NBUF <= 8, "ring of slice buffers" );
constexpr int      GS_R        = TILE_R;              // centre rows per tile
constexpr int      GS_CAP      = PPC_GS_CAP;
constexpr int      GS_POOL     = PPC_GS_POOL;
constexpr int      GS_STAGE_TARGET = PPC_GS_TARGET;
constexpr int      GS_F        = 4;                   // fine bins per cell
constexpr int      GS_MAXC     = 64;                  // staged columns (centre + halo)
constexpr int      GS_NB       = GS_MAXC * GS_F;      // fine bins per staged row
constexpr int      GS_BINS     = ( GS_R + 2 ) * GS_NB;
constexpr int      GS_STRIP_LIMIT = GS_CAP;           // staged particles (strip + halo) a strip may hold
constexpr int      GS_BUCKETS  = 64;                  // buckets of the record chunk: colour + 1 of a pair (32 colours; bucket 0 is unused)
constexpr int      GS_MAXSTRIPS = 64;                 // descriptor slots per tile (a strip is at least one column, a tile <= 62)
constexpr int      GS_MC_PAD   = 8;                   // mass/column words are copied from an 8-element (16-byte) boundary
constexpr uint32_t GS_NONE     = 0xffffu;
constexpr int      GS_FEW      = 256;                 // contact entries up to which a sub-tile takes the few-contacts path to its levels
constexpr uint32_t GS_LAST_STRIP = 0x80000000u;
constexpr uint32_t GS_HDR_UNITS = ( 2 * ( GS_R + 2 ) + 1 + 3 ) / 4;   // 21 words
constexpr uint32_t GS_ERR_CAP = 1u, GS_ERR_POOL = 2u, GS_ERR_DEPTH = 4u, GS_ERR_CHUNK = 8u;
static_assert( GS_CAP <= 1024, "T.ord holds a staged index in 10 bits" );
static_assert( GS_BINS * 4 <= GS_POOL * 4, "the bin counters alias the contact pool" );
static_assert( GS_CAP + ( GS_R + 2 ) * 2 * GS_MC_PAD <= GS_POOL, "the mass/column words land in pool_aux" );
static_assert( GS_BINS % GS_THREADS == 0, "bins per thread" );

struct GsSmem {
    float4   a[ GS_CAP ];           // x, y, radius, API index (bits), in storage order
    uint16_t pool_val[ GS_POOL ];   // contact entry: partner [10:0], scan-cell code 3 (row offset + 1) + column offset + 1 [14:11], anchored [15]
    uint16_t pool_next[ GS_POOL ];  // next entry of the same particle; the two arrays double as 32-bit bin counters during binning
    uint16_t pool_aux[ GS_POOL ];   // level of an entry that stands for a pair this sub-tile owns
    uint16_t fend[ GS_BINS ];       // end (exclusive) of fine bin f = row * GS_NB + bin in perm; its begin is fend[f - 1]
    uint16_t perm[ GS_CAP ];        // staged indices sorted by (staged row, fine bin)
    uint8_t  colb[ GS_CAP ];        // cell column byte of every staged particle
    uint8_t  massb[ GS_CAP ];       // mass byte (:141-142)
    uint16_t head[ GS_CAP ];        // first contact entry of a centre particle
    uint16_t ord[ GS_CAP ];         // centre particle t: staged index [9:0] | staged row << 10
    uint16_t perm2[ GS_CAP ];       // centre particle t: rank inside its class of contact count
    uint8_t  kk[ GS_CAP ];          // centre particle t: number of contacts (saturating)
    uint32_t kcnt[ 4 ];             // centre particles with contacts, per class of contact count
    uint16_t cstart[ ( GS_R + 2 ) * ( GS_MAXC + 1 ) ];   // first staged index of staged cell (row, column); [row][ncol] = end of the row
    float    inv_lut[ 256 ];        // 1 / (float)k, the inverse masses of :146-147
    uint32_t run_g0[ GS_R + 2 ], run_s0[ GS_R + 3 ], cen_g0[ GS_R + 1 ], cen_s0[ GS_R + 1 ];
    uint32_t mc_s0[ GS_R + 2 ];     // mc index of the first staged particle of each row
    uint32_t wsum[ 32 ], wsum2[ 32 ];
    uint32_t colp[ GS_MAXC + 2 ];   // column sums of cell_start over the staged rows (strip planning)
    uint32_t bcount[ GS_BUCKETS ];  // pairs per level, then running begin / end of each level in the record chunk
    uint32_t rmax_bits, pool_n, err, chunk_base, n_buckets, phase_mask, n_owned;
    unsigned long long mbar;
};

static_assert( ( sizeof( GsSmem ) + 1024 ) * PPC_GS_MINBLOCKS <= 232448, "PPC_GS_MINBLOCKS CTAs of k_gs_pairs must fit the 227 KB of an SM" );

struct GsLaunch {
    int wc;          // tile width in cells
    int tiles_x;     // tiles per tile row
    int ty_first;    // global tile row of the grid's first row (slab windows; 0 for a whole field)
    int colour;      // k_gs_resolve: 0..3 = (tx & 1) + 2 (ty & 1) of the tiles this launch runs
    int ntx;         // k_gs_resolve: tiles of this colour per tile row
    int ty_start;    // k_gs_resolve: first local tile row of this colour
    int strip_limit; // staged particles (strip + halo) above which a tile is cut into strips, and which a strip may hold (<= GS_CAP)
    int slice_log2;  // k_gs_resolve: log2 of the pair records per slice of a chunk (dynamic shared memory: GSB_NBUF * slice * 16 bytes)
    // k_gs_resolve, all four colours in ONE launch (merged != 0): CTAs take tickets; ticket t runs tile number t of the colour-major
    // list (col_begin[c] <= t < col_begin[c + 1]: colour c) once its neighbouring tiles of lower colour have reported done
    int      merged, tiles_y;
    int      col_ntx[ 4 ], col_ty_start[ 4 ];
    uint32_t col_begin[ 5 ];
    uint32_t epoch;   // value of a done[] word that means "done in THIS substep" (the words are never cleared)
};

// Pair records of one substep: one chunk per sub-tile, allocated with one atomic add, in units of 16 bytes:
//   GS_HDR_UNITS      run table of the sub-tile: first storage slot of each of its GS_R + 2 staged rows, then the running
//                     staged count at the begin of each row and in total (k_gs_resolve maps staged numbers to slots with it)
//   then              bucket table: end (exclusive) of every bucket in use, 16-bit, in bucket order, padded to whole units
//   then              one record per pair, sorted by bucket: {lo [10:0] | hi [21:11] (staged indices; lo = the lower API
//                     index = the reference's i), n_x, n_y, Baumgarte velocity term}
struct GsPairs {
    uint4    *chunks;
    uint32_t *cursor;   // units handed out so far
    uint32_t  cap;      // units
    uint2    *desc;     // [tile * GS_MAXSTRIPS + strip] = {first unit of that sub-tile's chunk + 1 (0: it owns no pair),
                        //  pairs | buckets in use << 16 | GS_LAST_STRIP for the tile's last strip}
};

// optional dump of the pairs k_gs_pairs hands to the resolve (parity tests: the hot kernel's own pair set)
struct GsDump {
    uint32_t *pairs;   // 2 words per pair: API index of i, of j
    uint32_t *count;
    uint32_t  cap;
};

__device__ __forceinline__ int gs_fine_bin( const float x, const float x0, const float inv_w, const int nb ) {
    const float q = __fmul_rn( __fsub_rn( x, x0 ), inv_w );
    int         b = ( q >= 0.0f ) ? ( q < 16777216.0f ? __float2int_rz( q ) : nb ) : 0;   // NaN -> 0
    return b < nb ? b : nb - 1;
}

// Column sums of the tile and its halo: cell_start is a running sum along a grid row, so with P[c] = sum over the staged
// rows of cell_start[row][c0 + c] the staged particle count of the columns [a, b) is P[b] - P[a].  Ends with a barrier.
// (k_gs_resolve runs one-warp CTAs, where the barrier costs no more than a __syncwarp.)
__device__ __forceinline__ void gs_column_sums( uint32_t *colp, const uint32_t *__restrict__ cell_start, const GridDims &G, const int c0, const int c1,
                                                const int cy0, const int nthreads ) {
    for ( int c = threadIdx.x; c <= c1 - c0; c += nthreads ) {
        uint32_t sum = 0;
        for ( int j = 0; j < GS_R + 2; j++ ) {
            const int gy = cy0 - 1 + j;
            if ( gy >= 0 && gy < G.rows )
                sum += cell_start[ (size_t)gy * G.cols + c0 + c ];
        }
        colp[ c ] = sum;
    }
    __syncthreads();
}

// End of the strip that starts at column sx: a tile that would overflow shared memory is cut greedily from the left, each strip
// taking as many columns as keep its staged particles (strip + one halo column either side) within the limit, at least one.
__device__ __forceinline__ int gs_strip_end( const uint32_t *colp, const uint32_t total, const int sx, const int cx1, const int c0, const int c1,
                                             const uint32_t limit ) {
    if ( total <= limit )
        return cx1;
    // the first end that cannot grow any further; the test is monotone in the end, so a warp tries 32 ends at once (every warp of the
    // CTA, or the one warp of k_gs_resolve, computes the same value; all their lanes are here)
    const uint32_t below = colp[ max( sx - 1, c0 ) - c0 ];
    for ( int ex = sx + 1;; ex += 32 ) {
        const int      e  = ex + (int)( threadIdx.x & 31u );
        const bool     ok = e < cx1 && colp[ min( e + 2, c1 ) - c0 ] - below <= limit;
        const unsigned m  = __ballot_sync( 0xffffffffu, ok );
        if ( m != 0xffffffffu )
            return ex + __ffs( (int)~m ) - 1;
    }
}

// Run table of a sub-tile: staged row j = 0 .. GS_R+1 is grid row cy0 - 1 + j, columns [c0, c1); centre = columns [cx0, cx1) of rows
// 1 .. GS_R.  One thread then turns the lengths into running sums.  Needs a barrier before, between and after (the caller's).
__device__ __forceinline__ void gs_run_lengths( uint32_t *run_g0, uint32_t *run_s0, uint32_t *cen_g0, uint32_t *cen_s0,
                                                const uint32_t *__restrict__ cell_start, const GridDims &G, const int c0, const int c1, const int cx0,
                                                const int cx1, const int cy0 ) {
    if ( threadIdx.x < GS_R + 2 ) {
        const int j = threadIdx.x, gy = cy0 - 1 + j;
        uint32_t  g0 = 0, g1 = 0, m0 = 0, m1 = 0;
        if ( gy >= 0 && gy < G.rows ) {
            const size_t row = (size_t)gy * G.cols;
            g0 = cell_start[ row + c0 ], g1 = cell_start[ row + c1 ];
            if ( j >= 1 && j <= GS_R )
                m0 = cell_start[ row + cx0 ], m1 = cell_start[ row + cx1 ];
        }
        run_g0[ j ] = g0;
        run_s0[ j ] = g1 - g0;   // length for now
        if ( j >= 1 && j <= GS_R ) {
            cen_g0[ j - 1 ] = m0;
            cen_s0[ j - 1 ] = m1 - m0;   // length for now
        }
    }
}
__device__ __forceinline__ void gs_run_sums( uint32_t *run_s0, uint32_t *cen_s0 ) {   // one thread
    uint32_t run = 0;
    for ( int j = 0; j < GS_R + 2; j++ ) {
        const uint32_t len = run_s0[ j ];
        run_s0[ j ]        = run;
        run += len;
    }
    run_s0[ GS_R + 2 ] = run;
    uint32_t cen = 0;
    for ( int j = 0; j < GS_R; j++ ) {
        const uint32_t len = cen_s0[ j ];
        cen_s0[ j ]        = cen;
        cen += len;
    }
    cen_s0[ GS_R ] = cen;
}

// ====================================================================================================================
// k_gs_pairs: one sub-tile = centre columns [cx0, cx1) of the tile [tcx0, tcx1), centre rows [cy0, cy0 + GS_R) (clipped).
// ====================================================================================================================
// Returns true when the sub-tile holds more than stage_limit staged particles (then nothing has been done).
__device__ __forceinline__ bool gs_pairs_subtile( GsSmem &T, const float4 *__restrict__ rec_a, const uint16_t *__restrict__ mcol, float4 *__restrict__ pv_out,
                                                  const uint32_t *__restrict__ cell_start, const GridDims &G, const float dt, const int cx0, const int cx1,
                                                  const int tcx0, const int tcx1, const int cy0, const int tx, const int tyg, const int my_colour,
                                                  uint32_t &phase_parity, uint2 *__restrict__ desc_slot, const uint32_t last_flag, const GsPairs PR,
                                                  uint32_t *__restrict__ err_flags, const GsDump D, const uint32_t stage_limit ) {
    const int c0 = max( cx0 - 1, 0 ), c1 = min( cx1 + 1, G.cols );      // staged columns [c0, c1)
    const int ncol = c1 - c0, nb = ncol * GS_F;
    uint32_t *cnt32 = reinterpret_cast<uint32_t *>( T.pool_val );       // bin counters / cursors while binning (pool unused then)
    uint16_t *mcland = T.pool_aux;   // where the bulk copies put the mass/column words, row by row from 16-byte boundaries (pool_aux is not used
                                     // before the levels are assigned); the binning pass below spreads them into colb / massb
    __syncthreads();                                                    // shared memory may still be in use by the previous strip

    // Run table and staging by warp 0, one lane per staged row: row j = 0 .. GS_R+1 is grid row cy0 - 1 + j, columns [c0, c1); centre =
    // columns [cx0, cx1) of rows 1 .. GS_R.  Lengths from the cell-start table, running sums by warp scans, and every lane posts the
    // two bulk copies of its own row (records; 16-bit mass/column words from the 16-byte boundary at or below the row's first slot).
    if ( threadIdx.x < 32 ) {
        const int      j  = (int)threadIdx.x, gy = cy0 - 1 + j;
        const bool     row = j < GS_R + 2 && gy >= 0 && gy < G.rows, mid = row && j >= 1 && j <= GS_R;
        uint32_t       g0 = 0, g1 = 0, m0 = 0, m1 = 0;
        if ( row ) {
            const uint32_t *cs = cell_start + (size_t)gy * G.cols;
            g0 = cs[ c0 ], g1 = cs[ c1 ];
            if ( mid )
                m0 = cs[ cx0 ], m1 = cs[ cx1 ];
        }
        const uint32_t len = g1 - g0, clen = m1 - m0, lead = g0 & ( GS_MC_PAD - 1 );
        const uint32_t mclen = len ? ( ( lead + len + GS_MC_PAD - 1 ) & ~(uint32_t)( GS_MC_PAD - 1 ) ) : 0u;
        uint32_t s_inc = len, c_inc = clen, m_inc = mclen;
#pragma unroll
        for ( int d = 1; d < 16; d <<= 1 ) {
            const uint32_t a = __shfl_up_sync( 0xffffffffu, s_inc, d ), b = __shfl_up_sync( 0xffffffffu, c_inc, d ), c = __shfl_up_sync( 0xffffffffu, m_inc, d );
            if ( j >= d )
                s_inc += a, c_inc += b, m_inc += c;
        }
        const uint32_t run = __shfl_sync( 0xffffffffu, s_inc, GS_R + 1 ), cen = __shfl_sync( 0xffffffffu, c_inc, GS_R + 1 );
        const uint32_t mcb = __shfl_sync( 0xffffffffu, m_inc, GS_R + 1 );
        if ( j < GS_R + 2 ) {
            T.run_g0[ j ] = g0;
            T.run_s0[ j ] = s_inc - len;
            T.mc_s0[ j ]  = m_inc - mclen + lead;
            if ( j >= 1 && j <= GS_R ) {
                T.cen_g0[ j - 1 ] = m0;
                T.cen_s0[ j - 1 ] = c_inc - clen;
            }
        }
        if ( j == 0 ) {
            T.run_s0[ GS_R + 2 ] = run;
            T.cen_s0[ GS_R ]     = cen;
            T.rmax_bits = 0u;
            T.err       = 0u;
            T.pool_n    = 0u;
            T.phase_mask = 0u;
            T.n_owned    = 0u;
        }
        if ( j < 4 )
            T.kcnt[ j ] = 0u;
#if PPC_GS_TMA
        if ( run != 0 && cen != 0 && run <= stage_limit ) {   // stage the runs: one barrier phase
            fence_proxy_async();   // the previous strip read this shared memory with ordinary loads
            if ( j == 0 )
                mbar_arrive_expect_tx( reinterpret_cast<uint64_t *>( &T.mbar ), run * 16u + mcb * 2u );
            __syncwarp();
            if ( len ) {
                tma_load_1d( &T.a[ s_inc - len ], rec_a + g0, len * 16u, reinterpret_cast<uint64_t *>( &T.mbar ) );
                tma_load_1d( &mcland[ m_inc - mclen ], mcol + ( g0 - lead ), mclen * 2u, reinterpret_cast<uint64_t *>( &T.mbar ) );
            }
        }
#endif
    }
    for ( int f = threadIdx.x; f < GS_BINS; f += GS_THREADS )
        cnt32[ f ] = 0;
    __syncthreads();
    const uint32_t staged = T.run_s0[ GS_R + 2 ], centre = T.cen_s0[ GS_R ];
    if ( staged == 0 || centre == 0 )
        return false;   // (the kernel has already marked this sub-tile as owning no pair)
    if ( staged > stage_limit )
        return true;    // nothing was staged: the caller cuts the tile into strips (or reports a strip that cannot be cut further)

    // ---- while the copies are in flight: staged cell table, inverse-mass table, list heads ----
    for ( int e = threadIdx.x; e < ( GS_R + 2 ) * ( ncol + 1 ); e += GS_THREADS ) {
        const int j = e / ( ncol + 1 ), i = e - j * ( ncol + 1 ), gy = cy0 - 1 + j;
        uint32_t  v = T.run_s0[ j ];
        if ( gy >= 0 && gy < G.rows )
            v += cell_start[ (size_t)gy * G.cols + c0 + i ] - T.run_g0[ j ];
        T.cstart[ j * ( GS_MAXC + 1 ) + i ] = (uint16_t)v;
    }
    if ( threadIdx.x < 256 )
        T.inv_lut[ threadIdx.x ] = __fdiv_rn( 1.0f, (float)threadIdx.x );
    for ( uint32_t t = threadIdx.x; t < staged; t += GS_THREADS ) {
        T.head[ t ] = (uint16_t)GS_NONE;
    }
    for ( int b = threadIdx.x; b < GS_BUCKETS; b += GS_THREADS )
        T.bcount[ b ] = 0u;
#if PPC_GS_TMA
    mbar_wait( reinterpret_cast<uint64_t *>( &T.mbar ), phase_parity );
    phase_parity ^= 1u;
#else
    for ( uint32_t t = threadIdx.x; t < staged; t += GS_THREADS ) {
        int j = 0;
        while ( j < GS_R + 1 && t >= T.run_s0[ j + 1 ] )
            j++;
        const uint32_t g = T.run_g0[ j ] + ( t - T.run_s0[ j ] );
        T.a[ t ] = rec_a[ g ];
        mcland[ T.mc_s0[ j ] + ( t - T.run_s0[ j ] ) ] = mcol[ g ];
    }
    __syncthreads();
#endif
    // ---- histogram of (row, fine bin); largest radius ----
    const float x0 = __fmul_rn( (float)c0, G.cell_size ), inv_w = __fdiv_rn( (float)GS_F, G.cell_size );
    float       rmax = 0.0f;
    for ( uint32_t t = threadIdx.x; t < staged; t += GS_THREADS ) {
        int j = 0;
        while ( j < GS_R + 1 && t >= T.run_s0[ j + 1 ] )
            j++;
        const float4 p = T.a[ t ];
        rmax           = fmaxf( rmax, p.z );
        atomicAdd( &cnt32[ j * GS_NB + gs_fine_bin( p.x, x0, inv_w, nb ) ], 1u );
        const uint32_t w = mcland[ T.mc_s0[ j ] + ( t - T.run_s0[ j ] ) ];
        T.colb[ t ]  = (uint8_t)( w >> 8 );
        T.massb[ t ] = (uint8_t)w;
    }
#pragma unroll
    for ( int d = 16; d > 0; d >>= 1 )
        rmax = fmaxf( rmax, __shfl_xor_sync( 0xffffffffu, rmax, d ) );
    if ( ( threadIdx.x & 31u ) == 0 && rmax > 0.0f )
        atomicMax( &T.rmax_bits, __float_as_uint( rmax ) );   // positive floats order like their bit patterns
    __syncthreads();

    // ---- exclusive scan of the bin counts (each thread a contiguous chunk; warp scan; block scan) ----
    {
        constexpr int PER = GS_BINS / GS_THREADS;
        const int     f0  = threadIdx.x * PER;
        uint32_t      loc[ PER ], sum = 0;
#pragma unroll
        for ( int i = 0; i < PER; i++ ) {
            loc[ i ] = cnt32[ f0 + i ];
            sum += loc[ i ];
        }
        uint32_t inc = sum;
#pragma unroll
        for ( int d = 1; d < 32; d <<= 1 ) {
            const uint32_t o = __shfl_up_sync( 0xffffffffu, inc, d );
            if ( ( threadIdx.x & 31u ) >= (unsigned)d )
                inc += o;
        }
        if ( ( threadIdx.x & 31u ) == 31u )
            T.wsum[ threadIdx.x >> 5 ] = inc;
        __syncthreads();
        uint32_t ws = ( threadIdx.x & 31u ) < (unsigned)( GS_THREADS / 32 ) ? T.wsum[ threadIdx.x & 31u ] : 0u;   // every warp scans the warp sums
#pragma unroll
        for ( int d = 1; d < 32; d <<= 1 ) {
            const uint32_t o = __shfl_up_sync( 0xffffffffu, ws, d );
            if ( ( threadIdx.x & 31u ) >= (unsigned)d )
                ws += o;
        }
        const unsigned wid  = threadIdx.x >> 5;
        const uint32_t base = wid ? __shfl_sync( 0xffffffffu, ws, wid - 1 ) : 0u;
        uint32_t       run  = base + inc - sum;
#pragma unroll
        for ( int i = 0; i < PER; i++ ) {
            cnt32[ f0 + i ] = run;   // begin of the bin; the scatter below advances it to the end
            run += loc[ i ];
        }
    }
    __syncthreads();
    for ( uint32_t t = threadIdx.x; t < staged; t += GS_THREADS ) {
        int j = 0;
        while ( j < GS_R + 1 && t >= T.run_s0[ j + 1 ] )
            j++;
        T.perm[ atomicAdd( &cnt32[ j * GS_NB + gs_fine_bin( T.a[ t ].x, x0, inv_w, nb ) ], 1u ) ] = (uint16_t)t;
    }
    __syncthreads();
    for ( int f = threadIdx.x; f < GS_BINS; f += GS_THREADS )
        T.fend[ f ] = (uint16_t)cnt32[ f ];
    __syncthreads();   // the counters are dead from here on: their memory becomes the contact pool
    const float R = __uint_as_float( T.rmax_bits );

    // ---- one thread per centre particle: candidates -> contact list (unordered) ----
    for ( uint32_t t = threadIdx.x; t < centre; t += GS_THREADS ) {
        int j = 1;
        while ( j < GS_R && t >= T.cen_s0[ j ] )
            j++;
        const uint32_t gk   = T.cen_g0[ j - 1 ] + ( t - T.cen_s0[ j - 1 ] );   // storage slot
        const uint32_t self = T.run_s0[ j ] + ( gk - T.run_g0[ j ] );          // staged index
        const float4   me4  = T.a[ self ];
        const float    x = me4.x, y = me4.y, r = me4.z;
        const int      mycol = (int)T.colb[ self ];
        const int      gy    = cy0 - 1 + j;
        const float    reach = __fadd_rn( __fadd_rn( r, R ), 0.02f );
        const int      b_lo  = gs_fine_bin( __fsub_rn( x, reach ), x0, inv_w, nb ), b_hi = gs_fine_bin( __fadd_rn( x, reach ), x0, inv_w, nb );
        const float    top   = __fmul_rn( (float)( gy + G.row0 ), G.cell_size ), bot = __fmul_rn( (float)( gy + G.row0 + 1 ), G.cell_size );

        uint32_t head = GS_NONE, tail = GS_NONE, k = 0;
#if PPC_GS_FLAT
        // The three staged rows a particle scans are ONE loop (a lane that has finished a row moves on to its next one while the
        // others are still scanning theirs: the lanes of a warp stay together for the sum of their trip counts, not row by row).
        int      d = -1, jr = 0;
        uint32_t q = 0, qe = 0;
#pragma unroll 1
        for ( ;; ) {
            if ( q >= qe ) {   // next row with candidates
                bool more = false;
                while ( ++d < 3 ) {
                    const int ry = gy - 1 + d;
                    if ( ry < 0 || ry >= G.rows )
                        continue;
                    if ( d == 0 && __fsub_rn( y, reach ) >= top )   // nothing in the row above can be within reach
                        continue;
                    if ( d == 2 && __fadd_rn( y, reach ) < bot )
                        continue;
                    jr = j - 1 + d;   // staged row scanned
                    const int f_lo = jr * GS_NB + b_lo, f_hi = jr * GS_NB + b_hi;
                    q  = f_lo > 0 ? T.fend[ f_lo - 1 ] : 0u;
                    qe = T.fend[ f_hi ];
                    if ( q < qe ) {
                        more = true;
                        break;
                    }
                }
                if ( !more )
                    break;
            }
            const uint32_t s  = T.perm[ q++ ];
            const float4   c  = T.a[ s ];
            const float    rs = __fadd_rn( r, c.z );
            const float    dx = __fsub_rn( x, c.x );
            if ( fabsf( dx ) > rs )   // :353
                continue;
            const float dy = __fsub_rn( y, c.y );
            if ( fabsf( dy ) > rs )   // :356
                continue;
            const float d2 = __fadd_rn( __fmul_rn( dx, dx ), __fmul_rn( dy, dy ) );
            if ( d2 <= __fmul_rn( rs, rs ) && s != self ) {   // :360-362
                const int dc = (int)(int8_t)( (uint32_t)T.colb[ s ] - (uint32_t)mycol );
                if ( dc < -1 || dc > 1 )   // two columns apart: outside the reference's 3x3 scan (:331-340)
                    continue;
                const uint32_t at = atomicAdd( &T.pool_n, 1u );
                if ( at < (uint32_t)GS_POOL ) {
                    T.pool_val[ at ]  = (uint16_t)( s | ( (uint32_t)( d * 3 + dc + 1 ) << 11 ) );
                    T.pool_next[ at ] = (uint16_t)GS_NONE;
                    if ( tail == GS_NONE )
                        head = at;
                    else
                        T.pool_next[ tail ] = (uint16_t)at;
                    tail = at;
                    k++;
                } else
                    T.err = GS_ERR_POOL;
            }
        }
#else
#pragma unroll 1
        for ( int d = 0; d < 3; d++ ) {
            const int ry = gy - 1 + d;
            if ( ry < 0 || ry >= G.rows )
                continue;
            if ( d == 0 && __fsub_rn( y, reach ) >= top )   // nothing in the row above can be within reach
                continue;
            if ( d == 2 && __fadd_rn( y, reach ) < bot )
                continue;
            const int      jr = j - 1 + d;   // staged row scanned
            const int      f_lo = jr * GS_NB + b_lo, f_hi = jr * GS_NB + b_hi;
            const uint32_t rb = f_lo > 0 ? T.fend[ f_lo - 1 ] : 0u, re = T.fend[ f_hi ];
#pragma unroll 2
            for ( uint32_t q = rb; q < re; q++ ) {
                const uint32_t s  = T.perm[ q ];
                const float4   c  = T.a[ s ];
                const float    rs = __fadd_rn( r, c.z );
                const float    dx = __fsub_rn( x, c.x );
                if ( fabsf( dx ) > rs )   // :353
                    continue;
                const float dy = __fsub_rn( y, c.y );
                if ( fabsf( dy ) > rs )   // :356
                    continue;
                const float d2 = __fadd_rn( __fmul_rn( dx, dx ), __fmul_rn( dy, dy ) );
                if ( d2 <= __fmul_rn( rs, rs ) && s != self ) {   // :360-362
                    const int dc = (int)(int8_t)( (uint32_t)T.colb[ s ] - (uint32_t)mycol );
                    if ( dc < -1 || dc > 1 )   // two columns apart: outside the reference's 3x3 scan (:331-340)
                        continue;
                    const uint32_t at = atomicAdd( &T.pool_n, 1u );
                    if ( at < (uint32_t)GS_POOL ) {
                        T.pool_val[ at ]  = (uint16_t)( s | ( (uint32_t)( d * 3 + dc + 1 ) << 11 ) );
                        T.pool_next[ at ] = (uint16_t)GS_NONE;
                        if ( tail == GS_NONE )
                            head = at;
                        else
                            T.pool_next[ tail ] = (uint16_t)at;
                        tail = at;
                        k++;
                    } else
                        T.err = GS_ERR_POOL;
                }
            }
        }
#endif
        // A particle without contacts is done: its position stands as the gather wrote it.  The others are handed out again below,
        // grouped by their number of contacts, so that the lanes of a warp walk lists of similar length.
        T.ord[ t ] = (uint16_t)( self | ( (uint32_t)j << 10 ) );
        T.kk[ t ]  = (uint8_t)( k < 255u ? k : 255u );
        if ( k ) {
            T.head[ self ] = (uint16_t)head;
            const uint32_t cls   = k <= 2u ? 3u : ( k <= 4u ? 2u : ( k <= 7u ? 1u : 0u ) );   // long lists first
            const unsigned peers = __match_any_sync( __activemask(), cls );
            const unsigned lane  = threadIdx.x & 31u, leader = (unsigned)( __ffs( peers ) - 1 );
            uint32_t       base  = 0;
            if ( lane == leader )
                base = atomicAdd( &T.kcnt[ cls ], (uint32_t)__popc( peers ) );
            base = __shfl_sync( peers, base, leader );
            T.perm2[ t ] = (uint16_t)( base + (uint32_t)__popc( peers & ( ( 1u << lane ) - 1u ) ) );   // rank inside its class
        }
    }
    __syncthreads();   // every list is complete; fend / perm are dead
    uint16_t *order = T.fend;   // centre particles with contacts, grouped by class
    const uint32_t n0 = T.kcnt[ 0 ], n1 = n0 + T.kcnt[ 1 ], n2 = n1 + T.kcnt[ 2 ], with_contacts = n2 + T.kcnt[ 3 ];
    for ( uint32_t t = threadIdx.x; t < centre; t += GS_THREADS ) {
        const uint32_t k = T.kk[ t ];
        if ( k )
            order[ ( k <= 2u ? n2 : ( k <= 4u ? n1 : ( k <= 7u ? n0 : 0u ) ) ) + T.perm2[ t ] ] = T.ord[ t ];
    }
    __syncthreads();
    // colours in use at every staged particle (bit c: a pair of colour c of this sub-tile touches it); lives where ord / perm2 were
    uint32_t *used = reinterpret_cast<uint32_t *>( T.ord );
    static_assert( sizeof( T.ord ) + sizeof( T.perm2 ) >= GS_CAP * sizeof( uint32_t ) && offsetof( GsSmem, perm2 ) == offsetof( GsSmem, ord ) + sizeof( T.ord ) &&
                       offsetof( GsSmem, ord ) % 4 == 0, "the colour masks alias ord + perm2" );
    for ( uint32_t t = threadIdx.x; t < staged; t += GS_THREADS )
        used[ t ] = 0u;

    // ---- one thread per centre particle with contacts: list into the reference's order, position pushes, who owns each pair ----
    for ( uint32_t u = threadIdx.x; u < with_contacts; u += GS_THREADS ) {
        const uint32_t ov   = order[ u ];
        const uint32_t self = ov & 0x3ffu;
        const int      j    = (int)( ov >> 10 );
        const uint32_t gk   = T.run_g0[ j ] + ( self - T.run_s0[ j ] );
        const float4   me4  = T.a[ self ];
        const float    x = me4.x, y = me4.y, r = me4.z;
        const uint32_t gid   = __float_as_uint( me4.w );
        const int      mycol = (int)T.colb[ self ];
        const uint32_t head  = T.head[ self ];
        float X = x, Y = y;
        {
            // The list is put into the order the reference's pair list touches this particle -- pairs (i, me), i < me, by
            // ascending i; then pairs (me, j) by scan cell (dy outer, dx inner), descending index inside a cell (:311-383) --
            // by selection, swapping entry values along the links; each entry is consumed as soon as it is in place.
            // Keys: a pair (i, me) sorts by i; a pair (me, j) by 2^31 | scan cell << 11 | (2047 - staged number of j), the
            // staged numbers of one cell ascending with the index.
            auto sort_key = [ & ]( const uint32_t v ) -> uint32_t {
                const uint32_t s = v & 0x7ffu, g = __float_as_uint( T.a[ s ].w );
                return g < gid ? g : 0x80000000u | ( v & 0x7800u ) | ( 2047u - s );   // the scan-cell code sits in the entry where the key wants it
            };
            const float wme      = T.inv_lut[ T.massb[ self ] ];
            const bool  row_last = ( j == GS_R ), row_first = ( j == 1 );
            const int   my_gc    = c0 + ( ( mycol - c0 ) & 0xff );   // this particle's grid column
#pragma unroll 1
            for ( uint32_t e = head; e != GS_NONE; e = T.pool_next[ e ] ) {
                uint32_t best = e, v = T.pool_val[ e ];
                uint32_t bkey = sort_key( v );
#pragma unroll 1
                for ( uint32_t f = T.pool_next[ e ]; f != GS_NONE; f = T.pool_next[ f ] ) {
                    const uint32_t vf = T.pool_val[ f ];
                    const uint32_t kf = sort_key( vf );
                    if ( kf < bkey ) {
                        bkey = kf;
                        best = f;
                        v    = vf;
                    }
                }
                if ( best != e )
                    T.pool_val[ best ] = T.pool_val[ e ];
                const uint32_t s     = v & 0x7ffu;
                const uint32_t code  = ( v >> 11 ) & 15u, code3 = ( code * 11u ) >> 5;                  // code / 3 for 0..8
                const int      d     = (int)code3 - 1, dc = (int)( code - 3u * code3 ) - 1;                  // partner's cell relative to mine
                const float4   c     = T.a[ s ];
                const uint32_t cg    = __float_as_uint( c.w );
                const bool     me_lo = gid < cg;
                const float    wq    = T.inv_lut[ T.massb[ s ] ];
                {   // position push, operands selected by role (lo = the pair's i), :145-156
                    const PushTerms P = push_terms( contact_geom( me_lo ? x : c.x, me_lo ? y : c.y, me_lo ? r : c.z, me_lo ? c.x : x, me_lo ? c.y : y,
                                                                  me_lo ? c.z : r, dt ), me_lo ? wme : wq, me_lo ? wq : wme );
                    if ( P.push ) {
                        X = me_lo ? __fadd_rn( X, P.dpx_lo ) : __fsub_rn( X, P.dpx_hi );
                        Y = me_lo ? __fadd_rn( Y, P.dpy_lo ) : __fsub_rn( Y, P.dpy_hi );
                    }
                }
                // who resolves this pair's impulse, and which of its two particles lists it (see the header)
                const int  pc       = my_gc + dc;   // partner's grid column
                const bool p_row_in = !( ( row_first && d < 0 ) || ( row_last && d > 0 ) );
                const bool p_col_in = pc >= cx0 && pc < cx1;
                bool       anchored;
                if ( p_row_in && p_col_in )   // both in this sub-tile: the particle of the base cell (smaller (row, column)); same cell: lower index
                    anchored = d > 0 || ( d == 0 && ( dc > 0 || ( dc == 0 && me_lo ) ) );
                else {                        // partner in the halo: mine iff its sub-tile runs later
                    const int dty = p_row_in ? 0 : d, dtx = pc < tcx0 ? -1 : ( pc >= tcx1 ? 1 : 0 );
                    if ( dty == 0 && dtx == 0 )
                        anchored = pc >= cx1;   // a later strip of this tile
                    else
                        anchored = ( ( ( tx + dtx ) & 1 ) | ( ( ( tyg + dty ) & 1 ) << 1 ) ) > my_colour;
                }
                T.pool_val[ e ] = (uint16_t)( v | ( anchored ? 0x8000u : 0u ) );
            }
        }
        *reinterpret_cast<float2 *>( pv_out + gk ) = make_float2( X, Y );   // the velocity half stays as the gather left it
    }
    __syncthreads();
    if ( T.err ) {
        if ( threadIdx.x == 0 ) {
            atomicOr( err_flags, T.err );
            *desc_slot = make_uint2( 0u, last_flag );
        }
        return false;
    }

    if ( T.pool_n == 0u ) {   // no contact at all in this sub-tile
        if ( threadIdx.x == 0 )
            *desc_slot = make_uint2( 0u, last_flag );
        return false;
    }

    // ---- the pairs this sub-tile owns, as records sorted by LEVEL ----
    // ("level" = pair-colour + 1 from here on.)  The pairs are VISITED in the nine phases of disjoint cell pairs described in the
    // header and each takes the lowest colour free at both its particles, a cheap integer pass done here; the pairs of one colour
    // share no particle and fire together in k_gs_resolve.  A dense packing has ~70 (phase, position) steps per sub-tile, ~14
    // levels under list scheduling of that order, and ~8 colours.
    // One thread per staged cell (the host bounds the tile width accordingly).  All pairs of a cell pair are listed by particles
    // of ONE cell with ONE direction code (the base cell when it is inside the sub-tile, else the other one): ascending particle,
    // list order -- the order inside the cell pair.
    const int  ncell  = ( GS_R + 2 ) * ncol;
    const int  row_g0 = G.row0 + cy0 - 1;      // whole-field row of staged row 0
    const int  xj = (int)threadIdx.x / ncol, xc = (int)threadIdx.x - xj * ncol;                       // this thread's cell
    const bool x_in = (int)threadIdx.x < ncell && xj >= 1 && xj <= GS_R && c0 + xc >= cx0 && c0 + xc < cx1;   // anchors are centre particles
    const uint32_t x_pb = x_in ? T.cstart[ xj * ( GS_MAXC + 1 ) + xc ] : 0u, x_pe = x_in ? T.cstart[ xj * ( GS_MAXC + 1 ) + xc + 1 ] : 0u;
    uint16_t *seg = T.fend;   // fend + perm are dead: entry numbers grouped by (cell, direction code)
    constexpr uint32_t SEG_CAP = ( sizeof( T.fend ) + sizeof( T.perm ) ) / sizeof( uint16_t );
    // phase of the pairs a cell lists under direction code d * 3 + dc: class by the direction as seen from the base cell, parity of
    // the base cell's whole-field column (east) or row (the southward classes)
    auto phase_at = [ & ]( const uint32_t code, const int xj, const int xc ) -> uint32_t {   // (xj, xc): staged cell of the listing particle
        const int  dj = (int)( code / 3u ) - 1, dcl = (int)( code % 3u ) - 1;
        const bool fwd = dj > 0 || ( dj == 0 && dcl >= 0 );                    // this cell is the base
        const int  bj = fwd ? xj : xj + dj, bc = fwd ? xc : xc + dcl;          // base cell
        const int  uj = fwd ? dj : -dj, uc = fwd ? dcl : -dcl;                 // direction from the base
        const int  cls = uj == 0 ? ( uc == 0 ? 0 : 1 ) : ( uc == 0 ? 2 : ( uc > 0 ? 3 : 4 ) );
        const int  par = cls == 1 ? ( ( c0 + bc ) & 1 ) : ( ( row_g0 + bj ) & 1 );
        return cls == 0 ? 0u : (uint32_t)( 2 * cls - 1 + par );
    };
    auto phase_of = [ & ]( const uint32_t code ) -> uint32_t { return phase_at( code, xj, xc ); };
    const bool few = T.pool_n <= (uint32_t)GS_FEW;
    if ( few ) {
        // ---- a sub-tile with a handful of contacts (every sub-tile of a dilute field): the same order and levels without the per-cell machinery ----
        // Every owned pair gets the key {phase, listing particle, position in its list}: pairs of one phase that share a particle belong
        // to one cell pair, whose order is exactly (particle, position); pairs that share none may come in any order.  Rank by
        // counting, then one thread assigns the levels in rank order.
        uint32_t *skey = reinterpret_cast<uint32_t *>( T.fend + GS_CAP );       // behind `order` (GS_CAP 16-bit words), inside fend
        uint16_t *sent = T.fend + GS_CAP + 2 * GS_FEW;
        static_assert( ( GS_CAP + 3 * GS_FEW ) * 2 <= (int)sizeof( T.fend ), "keys and entry numbers of the few-contacts path live in fend" );
        for ( uint32_t u = threadIdx.x; u < with_contacts; u += GS_THREADS ) {
            const uint32_t ov = order[ u ], self = ov & 0x3ffu;
            const int      pj = (int)( ov >> 10 ), pc = ( (int)T.colb[ self ] - c0 ) & 0xff;
            uint32_t       pos = 0;
            for ( uint32_t e = T.head[ self ]; e != GS_NONE; pos++ ) {
                const uint32_t nxt = T.pool_next[ e ], v = T.pool_val[ e ];
                if ( v & 0x8000u ) {
                    const uint32_t at = atomicAdd( &T.n_owned, 1u );
                    skey[ at ] = ( phase_at( ( v >> 11 ) & 15u, pj, pc ) << 20 ) | ( self << 10 ) | ( pos < 1023u ? pos : 1023u );
                    sent[ at ] = (uint16_t)e;
                    T.pool_next[ e ] = (uint16_t)self;   // the entry remembers its anchor where its link was
                }
                e = nxt;
            }
        }
        __syncthreads();
        const uint32_t P = T.n_owned;
        for ( uint32_t i = threadIdx.x; i < P; i += GS_THREADS ) {
            const uint32_t k = skey[ i ];
            uint32_t       rank = 0;
            for ( uint32_t o = 0; o < P; o++ )
                rank += skey[ o ] < k ? 1u : 0u;
            seg[ rank ] = sent[ i ];   // (`order` is dead: every list has been walked)
        }
        __syncthreads();
        if ( threadIdx.x == 0 )
            for ( uint32_t r = 0; r < P; r++ ) {
                const uint32_t e = seg[ r ], a = T.pool_next[ e ], q = T.pool_val[ e ] & 0x7ffu;
                const uint32_t ua = used[ a ], uq = used[ q ], taken = ua | uq;
                uint32_t       lv = (uint32_t)__ffs( (int)~taken );   // 1 + the lowest colour neither particle has used yet
                if ( lv == 0u ) {   // 32 colours in use around this pair
                    T.err = GS_ERR_DEPTH;
                    lv    = 32u;
                }
                used[ a ]       = ua | ( 1u << ( lv - 1u ) );
                used[ q ]       = uq | ( 1u << ( lv - 1u ) );
                T.pool_aux[ e ] = (uint16_t)lv;
                T.bcount[ lv ]++;
            }
        __syncthreads();
    } else {
    // (a) this cell's pairs per direction code: nine 7-bit counters in one word
    unsigned long long cnt = 0ull;
    uint32_t           mine = 0;
#pragma unroll 1
    for ( uint32_t a = x_pb; a < x_pe; a++ )
#pragma unroll 1
        for ( uint32_t e = T.head[ a ]; e != GS_NONE; e = T.pool_next[ e ] ) {
            const uint32_t v = T.pool_val[ e ];
            if ( v & 0x8000u ) {
                cnt += 1ull << ( 7u * ( ( v >> 11 ) & 15u ) );
                mine++;
            }
        }
    unsigned long long pre = 0ull;   // running sums of the nine counters, same packing (all below 128 when mine is)
    uint32_t           phases = 0;   // phases in which this cell has pairs
    {
        uint32_t run = 0;
#pragma unroll
        for ( uint32_t c = 0; c < 9; c++ ) {
            pre |= (unsigned long long)( run & 127u ) << ( 7u * c );
            const uint32_t k = (uint32_t)( cnt >> ( 7u * c ) ) & 127u;
            run += k;
            if ( k && x_in )
                phases |= 1u << phase_of( c );
        }
    }
    if ( phases )
        atomicOr( &T.phase_mask, phases );
    uint32_t seg_base;   // exclusive scan of `mine` over the CTA
    {
        uint32_t inc = mine;
#pragma unroll
        for ( int d = 1; d < 32; d <<= 1 ) {
            const uint32_t o = __shfl_up_sync( 0xffffffffu, inc, d );
            if ( ( threadIdx.x & 31u ) >= (unsigned)d )
                inc += o;
        }
        if ( ( threadIdx.x & 31u ) == 31u )
            T.wsum[ threadIdx.x >> 5 ] = inc;
        __syncthreads();
        uint32_t ws = ( threadIdx.x & 31u ) < (unsigned)( GS_THREADS / 32 ) ? T.wsum[ threadIdx.x & 31u ] : 0u;
#pragma unroll
        for ( int d = 1; d < 32; d <<= 1 ) {
            const uint32_t o = __shfl_up_sync( 0xffffffffu, ws, d );
            if ( ( threadIdx.x & 31u ) >= (unsigned)d )
                ws += o;
        }
        const unsigned wid = threadIdx.x >> 5;
        seg_base           = ( wid ? __shfl_sync( 0xffffffffu, ws, wid - 1 ) : 0u ) + inc - mine;
        if ( __shfl_sync( 0xffffffffu, ws, 31 ) > SEG_CAP )
            T.err = GS_ERR_POOL;   // more pairs than the entry list takes (same value in every thread)
    }
    if ( mine > 127u )   // a 7-bit counter may have wrapped (dozens of particles on one spot): not a field for this mode
        T.err = GS_ERR_DEPTH;
    __syncthreads();
    if ( T.err ) {
        if ( threadIdx.x == 0 ) {
            atomicOr( err_flags, T.err );
            *desc_slot = make_uint2( 0u, last_flag );
        }
        return false;
    }
    // (b) entry numbers of this cell, grouped by direction code; the entry remembers its anchor where its link was
    {
        unsigned long long cur = 0ull;
#pragma unroll 1
        for ( uint32_t a = x_pb; a < x_pe; a++ )
#pragma unroll 1
            for ( uint32_t e = T.head[ a ]; e != GS_NONE; ) {
                const uint32_t nxt = T.pool_next[ e ], v = T.pool_val[ e ];
                if ( v & 0x8000u ) {
                    const uint32_t code = ( v >> 11 ) & 15u;
                    seg[ seg_base + ( (uint32_t)( pre >> ( 7u * code ) ) & 127u ) + ( (uint32_t)( cur >> ( 7u * code ) ) & 127u ) ] = (uint16_t)e;
                    cur += 1ull << ( 7u * code );
                    T.pool_next[ e ] = (uint16_t)a;
                }
                e = nxt;
            }
    }
    __syncthreads();
    // (c) levels, phase by phase: in a phase a cell's thread walks the pairs of ONE direction code (towards the phase's other
    //     cell when this cell is the base, i.e. has the phase's parity; backwards to a halo base cell otherwise)
    const uint32_t phase_mask = T.phase_mask;
#pragma unroll 1
    for ( int phase = 0; phase < 9; phase++ ) {
        if ( !( ( phase_mask >> phase ) & 1u ) )
            continue;   // no pair of this sub-tile in this phase
        if ( ( phases >> phase ) & 1u ) {
            const int      cls = phase == 0 ? 0 : ( phase + 1 ) >> 1;   // 0 same cell, 1 east, 2 south, 3 south-east, 4 south-west
            const int      par = phase == 0 ? 0 : ( phase + 1 ) & 1;
            const bool     base = cls == 0 || ( ( cls == 1 ? c0 + xc : row_g0 + xj ) & 1 ) == par;
            const int      dj = cls >= 2 ? 1 : 0, dcl = cls == 1 || cls == 3 ? 1 : ( cls == 4 ? -1 : 0 );
            const uint32_t code = (uint32_t)( ( ( base ? dj : -dj ) + 1 ) * 3 + ( base ? dcl : -dcl ) + 1 );
            const uint32_t kb = seg_base + ( (uint32_t)( pre >> ( 7u * code ) ) & 127u ), ke = kb + ( (uint32_t)( cnt >> ( 7u * code ) ) & 127u );
#pragma unroll 1
            for ( uint32_t k = kb; k < ke; k++ ) {
                const uint32_t e = seg[ k ];
                const uint32_t a = T.pool_next[ e ], q = T.pool_val[ e ] & 0x7ffu;
                const uint32_t ua = used[ a ], uq = used[ q ], taken = ua | uq;
                uint32_t       lv = (uint32_t)__ffs( (int)~taken );   // 1 + the lowest colour neither particle has used yet
                if ( lv == 0u ) {   // 32 colours in use around this pair
                    T.err = GS_ERR_DEPTH;
                    lv    = 32u;
                }
                used[ a ]       = ua | ( 1u << ( lv - 1u ) );
                used[ q ]       = uq | ( 1u << ( lv - 1u ) );
                T.pool_aux[ e ] = (uint16_t)lv;
                atomicAdd( &T.bcount[ lv ], 1u );
            }
        }
        __syncthreads();
    }
    }   // (the per-cell path)
    if ( T.err ) {
        if ( threadIdx.x == 0 ) {
            atomicOr( err_flags, T.err );
            *desc_slot = make_uint2( 0u, last_flag );
        }
        return false;
    }
    // exclusive scan of the level counts and of the "level in use" flags over the CTA; the chunk; the table of level ends
    {
        constexpr int PER = ( GS_BUCKETS + GS_THREADS - 1 ) / GS_THREADS;
        const int     f0  = threadIdx.x * PER;
        uint32_t      loc[ PER ], sum = 0, used = 0;
#pragma unroll
        for ( int i = 0; i < PER; i++ ) {
            loc[ i ] = f0 + i < GS_BUCKETS ? T.bcount[ f0 + i ] : 0u;
            sum += loc[ i ];
            used += loc[ i ] ? 1u : 0u;
        }
        uint32_t inc = sum, uinc = used;
#pragma unroll
        for ( int d = 1; d < 32; d <<= 1 ) {
            const uint32_t o = __shfl_up_sync( 0xffffffffu, inc, d ), u = __shfl_up_sync( 0xffffffffu, uinc, d );
            if ( ( threadIdx.x & 31u ) >= (unsigned)d ) {
                inc += o;
                uinc += u;
            }
        }
        if ( ( threadIdx.x & 31u ) == 31u ) {
            T.wsum[ threadIdx.x >> 5 ]  = inc;
            T.wsum2[ threadIdx.x >> 5 ] = uinc;
        }
        __syncthreads();
        const bool in = ( threadIdx.x & 31u ) < (unsigned)( GS_THREADS / 32 );
        uint32_t   ws = in ? T.wsum[ threadIdx.x & 31u ] : 0u, us = in ? T.wsum2[ threadIdx.x & 31u ] : 0u;   // every warp scans the warp sums
#pragma unroll
        for ( int d = 1; d < 32; d <<= 1 ) {
            const uint32_t o = __shfl_up_sync( 0xffffffffu, ws, d ), u = __shfl_up_sync( 0xffffffffu, us, d );
            if ( ( threadIdx.x & 31u ) >= (unsigned)d ) {
                ws += o;
                us += u;
            }
        }
        const unsigned wid     = threadIdx.x >> 5;
        const uint32_t n_pairs = __shfl_sync( 0xffffffffu, ws, 31 ), n_used = __shfl_sync( 0xffffffffu, us, 31 );
        uint32_t       run     = ( wid ? __shfl_sync( 0xffffffffu, ws, wid - 1 ) : 0u ) + inc - sum;
        uint32_t       urun    = ( wid ? __shfl_sync( 0xffffffffu, us, wid - 1 ) : 0u ) + uinc - used;
        if ( threadIdx.x == 0 ) {
            const uint32_t units = GS_HDR_UNITS + ( n_used * 2u + 15u ) / 16u + n_pairs;
            uint32_t       base  = 0xffffffffu;
            if ( n_pairs ) {
                base = atomicAdd( PR.cursor, units );
                if ( base + units > PR.cap ) {   // pair buffer full
                    atomicOr( err_flags, GS_ERR_CHUNK );
                    base = 0xffffffffu;
                }
            }
            T.chunk_base = base;
            T.n_buckets  = n_used;
            T.n_owned    = n_pairs;
            *desc_slot   = base != 0xffffffffu ? make_uint2( base + 1u, n_pairs | ( n_used << 16 ) | last_flag ) : make_uint2( 0u, last_flag );
        }
        __syncthreads();
        uint16_t *table = T.chunk_base != 0xffffffffu ? reinterpret_cast<uint16_t *>( PR.chunks + T.chunk_base + GS_HDR_UNITS ) : nullptr;
        if ( table && threadIdx.x < 2 * ( GS_R + 2 ) + 1 )   // the run table
            reinterpret_cast<uint32_t *>( PR.chunks + T.chunk_base )[ threadIdx.x ] =
                threadIdx.x < GS_R + 2 ? T.run_g0[ threadIdx.x ] : T.run_s0[ threadIdx.x - ( GS_R + 2 ) ];
#pragma unroll
        for ( int i = 0; i < PER; i++ )
            if ( f0 + i < GS_BUCKETS ) {
                T.bcount[ f0 + i ] = run;   // begin of the bucket; the scatter below advances it to the end
                run += loc[ i ];
                if ( loc[ i ] && table )
                    table[ urun++ ] = (uint16_t)run;   // end of the bucket, among the buckets in use
            }
    }
    __syncthreads();
    if ( T.chunk_base == 0xffffffffu )
        return false;   // no pair, or no room
    {
        uint4         *rec = PR.chunks + T.chunk_base + GS_HDR_UNITS + ( T.n_buckets * 2u + 15u ) / 16u;
        // one thread per owned pair: seg lists exactly the entries that stand for one (every lane has work, unlike a sweep over the pool)
        const uint32_t n_owned = T.n_owned;
        for ( uint32_t k = threadIdx.x; k < n_owned; k += GS_THREADS ) {
            const uint32_t e = seg[ k ], v = T.pool_val[ e ];
            const uint32_t a = T.pool_next[ e ], q = v & 0x7ffu;
            const float4   A = T.a[ a ], Q = T.a[ q ];
            const bool     a_lo = __float_as_uint( A.w ) < __float_as_uint( Q.w );   // lo = the lower API index = the reference's i
            const float4   L = a_lo ? A : Q, H = a_lo ? Q : A;
            const ContactGeom C = contact_geom( L.x, L.y, L.z, H.x, H.y, H.z, dt );
            const uint32_t at = atomicAdd( &T.bcount[ T.pool_aux[ e ] ], 1u );   // the order inside a bucket is immaterial
            rec[ at ] = make_uint4( ( a_lo ? a : q ) | ( ( a_lo ? q : a ) << 11 ), __float_as_uint( C.nx ), __float_as_uint( C.ny ), __float_as_uint( C.vbias ) );
            if ( D.pairs ) {
                const uint32_t k = atomicAdd( D.count, 1u );
                if ( k < D.cap ) {
                    D.pairs[ 2 * k ]     = __float_as_uint( L.w );
                    D.pairs[ 2 * k + 1 ] = __float_as_uint( H.w );
                }
            }
        }
    }
    return false;
}

__global__ void __launch_bounds__( GS_THREADS, PPC_GS_MINBLOCKS ) k_gs_pairs( const float4 *__restrict__ rec_a, const uint16_t *__restrict__ mcol,
                                                             float4 *__restrict__ pv_out, const uint32_t *__restrict__ cell_start, const GridDims G,
                                                             const float dt, const GsLaunch L, const GsPairs PR, uint32_t *__restrict__ err_flags,
                                                             const GsDump D ) {
    extern __shared__ __align__( 16 ) unsigned char smem_raw[];
    GsSmem &T = *reinterpret_cast<GsSmem *>( smem_raw );
    const int tx = (int)( blockIdx.x % (unsigned)L.tiles_x ), ty = (int)( blockIdx.x / (unsigned)L.tiles_x );   // ty: tile row inside this grid
    const int tyg = L.ty_first + ty;                                       // tile row of the whole field
    const int cx0 = tx * L.wc, cx1 = min( cx0 + L.wc, G.cols );
    const int cy0 = tyg * GS_R - G.row0;                                   // may be negative for a slab window's first tile row
    if ( threadIdx.x == 0 )
        mbar_init( reinterpret_cast<uint64_t *>( &T.mbar ), 1u );
    uint2         *desc  = PR.desc + (size_t)blockIdx.x * GS_MAXSTRIPS;
    const int      my_colour = ( tx & 1 ) | ( ( tyg & 1 ) << 1 );
    uint32_t       parity = 0;
    // Most tiles fit as they are: one sub-tile, and its run table tells so without looking at the columns.
    if ( threadIdx.x == 0 )
        desc[ 0 ] = make_uint2( 0u, GS_LAST_STRIP );   // until the sub-tile turns out to own pairs (same thread, program order)
    if ( !gs_pairs_subtile( T, rec_a, mcol, pv_out, cell_start, G, dt, cx0, cx1, cx0, cx1, cy0, tx, tyg, my_colour, parity, desc, GS_LAST_STRIP, PR, err_flags,
                            D, (uint32_t)L.strip_limit ) )
        return;
    // A tile whose staged particles exceed the limit is cut into column strips, which this CTA runs one after the other.
    const int c0 = max( cx0 - 1, 0 ), c1 = min( cx1 + 1, G.cols );
    __syncthreads();
    gs_column_sums( T.colp, cell_start, G, c0, c1, cy0, GS_THREADS );
    const uint32_t total = T.colp[ c1 - c0 ] - T.colp[ 0 ];
    int            strip = 0;
    for ( int sx = cx0; sx < cx1; strip++ ) {
        const int      ex   = gs_strip_end( T.colp, total, sx, cx1, c0, c1, (uint32_t)L.strip_limit );
        const uint32_t last = ex >= cx1 ? GS_LAST_STRIP : 0u;
        if ( threadIdx.x == 0 )
            desc[ strip ] = make_uint2( 0u, last );
        if ( gs_pairs_subtile( T, rec_a, mcol, pv_out, cell_start, G, dt, sx, ex, cx0, cx1, cy0, tx, tyg, my_colour, parity, desc + strip, last, PR, err_flags, D,
                               (uint32_t)GS_CAP ) &&
             threadIdx.x == 0 )
            atomicOr( err_flags, GS_ERR_CAP );   // a one-column strip that still does not fit: this mode has no generic path
        sx = ex;
    }
}

// ====================================================================================================================
// k_gs_resolve: the velocity impulses of the pairs a sub-tile owns, bucket by bucket (one launch per tile colour).
// ====================================================================================================================
// A sub-tile's records are one contiguous chunk, read front to back: the TMA streams it through shared memory in slices of
// GSB_SLICE records, GSB_NBUF slices in flight, so a record is never waited for at HBM latency.  CTAs are small (the pairs of a bucket
// are a few dozen) and hold little shared memory, so many tiles are resident per SM and their dependent steps interleave.
struct GsbSmem {
    float2   v[ GS_CAP ];                    // current velocity of every staged particle
    uint8_t  m[ GS_CAP ];                    // mass byte (:141-142)
    uint8_t  dirty[ GS_CAP ];                // an impulse changed this particle's velocity
    uint16_t bend[ GS_BUCKETS ];             // end of every bucket in use
    float    inv_lut[ 256 ];                 // 1 / (float)k, the inverse masses of :146-147
    uint32_t run[ 2 * ( GS_R + 2 ) + 1 + 3 ];   // run table of the sub-tile, as k_gs_pairs left it: run_g0[GS_R + 2], run_s0[GS_R + 3]
    uint2    desc;
    unsigned long long mbar[ GSB_NBUF ];
};

__global__ void __launch_bounds__( GSB_THREADS ) k_gs_resolve( const uint16_t *__restrict__ mcol, float4 *pv, const GsLaunch L, const GsPairs PR,
                                                                const uint32_t *__restrict__ err_flags, uint32_t *__restrict__ ticket,
                                                                uint32_t *done ) {
    __shared__ __align__( 16 ) GsbSmem T;
    __shared__ uint32_t s_ticket;
    extern __shared__ __align__( 16 ) uint4 gsb_recs[];   // ring of GSB_NBUF slices of the record chunk, L.slice records each
    const uint32_t SL = (uint32_t)L.slice_log2, GSB_SLICE = 1u << SL;
    if ( *err_flags )   // some sub-tile did not fit: the host runs the exact resolve for this substep instead
        return;
    int tx, ty;   // ty: tile row inside this grid
    if ( L.merged ) {
        // Tickets are handed out in colour-major order and a CTA only ever waits for tiles of LOWER colour, i.e. for lower tickets,
        // whose CTAs are already running: no deadlock whatever order the hardware starts CTAs in.
        if ( threadIdx.x == 0 )
            s_ticket = atomicAdd( ticket, 1u );
        __syncthreads();
        const uint32_t t = s_ticket;
        // (selected with constant indices: a dynamically indexed kernel parameter would be copied to local memory)
        const int      c = t >= L.col_begin[ 3 ] ? 3 : ( t >= L.col_begin[ 2 ] ? 2 : ( t >= L.col_begin[ 1 ] ? 1 : 0 ) );
        const uint32_t i = t - ( c == 3 ? L.col_begin[ 3 ] : ( c == 2 ? L.col_begin[ 2 ] : ( c == 1 ? L.col_begin[ 1 ] : 0u ) ) );
        const unsigned ntx = (unsigned)( c == 3 ? L.col_ntx[ 3 ] : ( c == 2 ? L.col_ntx[ 2 ] : ( c == 1 ? L.col_ntx[ 1 ] : L.col_ntx[ 0 ] ) ) );
        const int      ty0 = c == 3 ? L.col_ty_start[ 3 ] : ( c == 2 ? L.col_ty_start[ 2 ] : ( c == 1 ? L.col_ty_start[ 1 ] : L.col_ty_start[ 0 ] ) );
        const int      bx = (int)( i % ntx ), by = (int)( i / ntx );
        tx = 2 * bx + ( c & 1 ), ty = 2 * by + ty0;
        if ( threadIdx.x < 8 ) {   // one lane per neighbouring tile
            const int k = (int)threadIdx.x, ddx = k < 3 ? k - 1 : ( k == 3 ? -1 : ( k == 4 ? 1 : k - 6 ) ), ddy = k < 3 ? -1 : ( k < 5 ? 0 : 1 );
            const int nx = tx + ddx, ny = ty + ddy;
            if ( nx >= 0 && nx < L.tiles_x && ny >= 0 && ny < L.tiles_y && ( ( nx & 1 ) | ( ( ( L.ty_first + ny ) & 1 ) << 1 ) ) < c ) {
                const volatile uint32_t *flag = done + (size_t)ny * L.tiles_x + nx;
                while ( *flag != L.epoch )
                    __nanosleep( 100 );
            }
        }
        __threadfence();
        __syncthreads();
    } else {
        const int bx = (int)( blockIdx.x % (unsigned)L.ntx ), by = (int)( blockIdx.x / (unsigned)L.ntx );
        tx = 2 * bx + ( L.colour & 1 ), ty = 2 * by + L.ty_start;
    }
    const uint2 *desc = PR.desc + ( (size_t)ty * L.tiles_x + tx ) * GS_MAXSTRIPS;
    uint32_t     parity = 0u;   // bit i: phase of the barrier of ring buffer i
    bool         ready = false;
    for ( int strip = 0;; strip++ ) {
        __syncthreads();   // shared memory may still be in use by the previous strip
        if ( threadIdx.x == 0 )
            T.desc = desc[ strip ];
        __syncthreads();
        const uint2 d = T.desc;
        const uint32_t n_pairs = d.y & 0xffffu, n_used = ( d.y >> 16 ) & 0x7fffu;
        if ( d.x != 0 && n_pairs <= (uint32_t)GSB_THREADS ) {
            // A sub-tile with a handful of pairs (every sub-tile of a dilute field): one thread per pair, everything straight from
            // global memory -- its record, the two velocities when its level comes up -- instead of staging the velocities of a
            // thousand particles for a few dozen impulses.  Same pairs, same levels, same arithmetic.
            const uint4 *chunk = PR.chunks + ( d.x - 1u );
            const uint4 *rec   = chunk + GS_HDR_UNITS + ( n_used * 2u + 15u ) / 16u;
            if ( threadIdx.x < 2 * ( GS_R + 2 ) + 1 )
                T.run[ threadIdx.x ] = reinterpret_cast<const uint32_t *>( chunk )[ threadIdx.x ];
            for ( uint32_t b = threadIdx.x; b < n_used; b += GSB_THREADS )
                T.bend[ b ] = reinterpret_cast<const uint16_t *>( chunk + GS_HDR_UNITS )[ b ];
            const bool mine = threadIdx.x < n_pairs;
            uint4      r    = make_uint4( 0u, 0u, 0u, 0u );
            if ( mine )
                r = rec[ threadIdx.x ];
            __syncthreads();
            const uint32_t *run_g0 = T.run, *run_s0 = T.run + ( GS_R + 2 );
            uint32_t my_level = 0xffffffffu, slot_lo = 0, slot_hi = 0;
            float    wlo = 0.0f, whi = 0.0f;
            if ( mine ) {
                my_level = 0;
                while ( my_level < n_used && threadIdx.x >= T.bend[ my_level ] )
                    my_level++;
                const uint32_t lo = r.x & 0x7ffu, hi = ( r.x >> 11 ) & 0x7ffu;
                int jl = 0, jh = 0;
                while ( jl < GS_R + 1 && lo >= run_s0[ jl + 1 ] )
                    jl++;
                while ( jh < GS_R + 1 && hi >= run_s0[ jh + 1 ] )
                    jh++;
                slot_lo = run_g0[ jl ] + ( lo - run_s0[ jl ] );
                slot_hi = run_g0[ jh ] + ( hi - run_s0[ jh ] );
                wlo     = __fdiv_rn( 1.0f, (float)( mcol[ slot_lo ] & 0xffu ) );   // :141-142, :146-147
                whi     = __fdiv_rn( 1.0f, (float)( mcol[ slot_hi ] & 0xffu ) );
            }
            for ( uint32_t b = 0; b < n_used; b++ ) {
                if ( my_level == b ) {
                    float2 *pl = reinterpret_cast<float2 *>( pv + slot_lo ) + 1, *ph = reinterpret_cast<float2 *>( pv + slot_hi ) + 1;
                    float2  ul = __ldcg( pl ), uh = __ldcg( ph );
                    ContactGeom C;
                    C.nx = __uint_as_float( r.y ), C.ny = __uint_as_float( r.z ), C.vbias = __uint_as_float( r.w ), C.raw = 0.0f, C.corrected = 0.0f;
                    const ImpulseTerms I = impulse_terms( C, wlo, whi, ul.x, ul.y, uh.x, uh.y );   // :158-190
                    if ( I.impulse ) {
                        ul.x = __fadd_rn( ul.x, I.dvx_lo );
                        ul.y = __fadd_rn( ul.y, I.dvy_lo );
                        uh.x = __fsub_rn( uh.x, I.dvx_hi );
                        uh.y = __fsub_rn( uh.y, I.dvy_hi );
                        __stcg( pl, ul );
                        __stcg( ph, uh );
                    }
                }
                if ( b + 1 < n_used )
                    __syncthreads();   // the next level reads what this one wrote (same CTA: the barrier orders the global accesses)
            }
        } else if ( d.x != 0 ) {
            if ( !ready ) {   // first sub-tile with many pairs
                if ( threadIdx.x < GSB_NBUF )
                    mbar_init( reinterpret_cast<uint64_t *>( &T.mbar[ threadIdx.x ] ), 1u );
                for ( int i = threadIdx.x; i < 256; i += GSB_THREADS )
                    T.inv_lut[ i ] = __fdiv_rn( 1.0f, (float)i );
                ready = true;
            }
            const uint4   *chunk   = PR.chunks + ( d.x - 1u );
            const uint4   *rec     = chunk + GS_HDR_UNITS + ( n_used * 2u + 15u ) / 16u;
            const uint32_t n_slices = ( n_pairs + GSB_SLICE - 1 ) >> SL;
            if ( threadIdx.x < 2 * ( GS_R + 2 ) + 1 )
                T.run[ threadIdx.x ] = reinterpret_cast<const uint32_t *>( chunk )[ threadIdx.x ];
            for ( uint32_t b = threadIdx.x; b < n_used; b += GSB_THREADS )
                T.bend[ b ] = reinterpret_cast<const uint16_t *>( chunk + GS_HDR_UNITS )[ b ];
            __syncthreads();   // run table; the barriers initialised
            auto load_slice = [ & ]( const uint32_t i ) {   // one thread
                const uint32_t count = min( GSB_SLICE, n_pairs - ( i << SL ) );
                fence_proxy_async();
                mbar_arrive_expect_tx( reinterpret_cast<uint64_t *>( &T.mbar[ i % GSB_NBUF ] ), count * 16u );
                tma_load_1d( gsb_recs + ( ( i % GSB_NBUF ) << SL ), rec + ( (size_t)i << SL ), count * 16u, reinterpret_cast<uint64_t *>( &T.mbar[ i % GSB_NBUF ] ) );
            };
            if ( threadIdx.x == 0 )
                for ( uint32_t i = 0; i < (uint32_t)GSB_NBUF && i < n_slices; i++ )
                    load_slice( i );
            const uint32_t *run_g0 = T.run, *run_s0 = T.run + ( GS_R + 2 );
            // current velocities (earlier sub-tiles may have changed them) and mass bytes of the staged particles, row by row so
            // that the loads of a row are independent of each other and many are in flight
            for ( int j = 0; j < GS_R + 2; j++ ) {
                const uint32_t s0 = run_s0[ j ], len = run_s0[ j + 1 ] - s0, g0 = run_g0[ j ];
#pragma unroll 8
                for ( uint32_t t = threadIdx.x; t < len; t += GSB_THREADS ) {
                    T.v[ s0 + t ]     = __ldcg( reinterpret_cast<const float2 *>( pv + g0 + t ) + 1 );
                    T.m[ s0 + t ]     = (uint8_t)( mcol[ g0 + t ] & 0xffu );
                    T.dirty[ s0 + t ] = 0;
                }
            }
            __syncthreads();
            // Fire, bucket by bucket (the pairs of a bucket share no particle), in segments that end where a bucket or a slice ends.
            uint32_t k = 0, b = 0;
#pragma unroll 1
            while ( k < n_pairs ) {
                const uint32_t slice = k >> SL, s_end = min( ( slice + 1u ) << SL, n_pairs );
                const uint32_t b_end = T.bend[ b ], seg_end = min( b_end, s_end );
                const uint32_t ring = slice % GSB_NBUF;
                if ( k == slice << SL ) {   // first record of a slice: it must have landed
                    mbar_wait( reinterpret_cast<uint64_t *>( &T.mbar[ ring ] ), ( parity >> ring ) & 1u );
                    parity ^= 1u << ring;
                }
                const uint4 *buf = gsb_recs + ( ring << SL ) - ( (size_t)slice << SL );
#pragma unroll 1
                for ( uint32_t q = k + threadIdx.x; q < seg_end; q += GSB_THREADS ) {
                    const uint4    r  = buf[ q ];
                    const uint32_t lo = r.x & 0x7ffu, hi = ( r.x >> 11 ) & 0x7ffu;
                    ContactGeom    C;
                    C.nx = __uint_as_float( r.y ), C.ny = __uint_as_float( r.z ), C.vbias = __uint_as_float( r.w ), C.raw = 0.0f, C.corrected = 0.0f;
                    float2             ul = T.v[ lo ], uh = T.v[ hi ];
                    const ImpulseTerms I  = impulse_terms( C, T.inv_lut[ T.m[ lo ] ], T.inv_lut[ T.m[ hi ] ], ul.x, ul.y, uh.x, uh.y );   // :158-190
                    if ( I.impulse ) {
                        ul.x = __fadd_rn( ul.x, I.dvx_lo );
                        ul.y = __fadd_rn( ul.y, I.dvy_lo );
                        uh.x = __fsub_rn( uh.x, I.dvx_hi );
                        uh.y = __fsub_rn( uh.y, I.dvy_hi );
                        T.v[ lo ]     = ul;
                        T.v[ hi ]     = uh;
                        T.dirty[ lo ] = 1;
                        T.dirty[ hi ] = 1;
                    }
                }
                __syncthreads();
                if ( seg_end == s_end && slice + GSB_NBUF < n_slices && threadIdx.x == 0 )
                    load_slice( slice + GSB_NBUF );   // the slice just finished frees its buffer
                k = seg_end;
                if ( k == b_end )
                    b++;
            }
            // what changed goes back: this sub-tile's own particles are final, halo particles are picked up by their own sub-tile later
            for ( int j = 0; j < GS_R + 2; j++ ) {
                const uint32_t s0 = run_s0[ j ], len = run_s0[ j + 1 ] - s0, g0 = run_g0[ j ];
                for ( uint32_t t = threadIdx.x; t < len; t += GSB_THREADS )
                    if ( T.dirty[ s0 + t ] )
                        *( reinterpret_cast<float2 *>( pv + g0 + t ) + 1 ) = T.v[ s0 + t ];
            }
        }
        if ( d.y & GS_LAST_STRIP )
            break;
    }
    if ( L.merged ) {   // everything this tile wrote is visible before its flag is
        __threadfence();
        __syncthreads();
        if ( threadIdx.x == 0 )
            *( (volatile uint32_t *)( done + (size_t)ty * L.tiles_x + tx ) ) = L.epoch;
    }
}

}   // namespace ppc
