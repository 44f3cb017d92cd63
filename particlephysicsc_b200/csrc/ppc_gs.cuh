// ppc_gs.cuh -- K7 of the "resolve"=3 path: narrow phase, position pushes and a bounded-pass, deterministic Gauss-Seidel
// velocity resolve, all inside the shared memory of one tile.  No contact lists in global memory, no cooperative launch,
// no grid-wide rounds: four ordinary launches (one per tile colour) per substep.
//
// Relation to the reference (src/particles_solver.c):
//  * candidate pairs: the reference's (3x3 cell neighbourhood :331-340, AABB reject and inclusive circle test :352-362,
//    same float operations) -- as in k_collide_tile2 (ppc_collide.cuh), whose fine x-bins this kernel shares;
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
//     (tx & 1) + 2 (ty & 1).  One launch per colour; tiles of one colour touch disjoint particles (a tile touches its
//     cells and a one-cell halo; tiles are >= 2 cells wide).  A tile whose staged particles would overflow shared memory
//     is cut into column strips (greedily from the left, as wide as fits), which its CTA runs one after the other.
//  2. A pair belongs to the sub-tile (tile, strip) holding both particles, else to the EARLIER of the two sub-tiles
//     (lower colour; inside a tile the lower strip).  The owner applies the impulse to both particles; what it did to a
//     halo particle is written back to global memory, where that particle's own sub-tile picks it up later.  After a
//     sub-tile has run, all pairs of its own particles are done, so it writes their final state.
//  3. Inside a sub-tile, nine phases separated by CTA barriers: pairs inside one cell; pairs with the east neighbour cell,
//     base column even / odd; with the south, south-east, south-west neighbour, base row even / odd.  The cell pairs of
//     one phase are disjoint, so one thread per cell pair runs its pairs one after the other with no conflicts.
//  4. Inside a cell pair: by anchor particle (ascending index), then by position in the anchor's contact list.
//
// Data movement: the canonical gather (ppc_grid.cuh) leaves two 16-byte records per particle in sorted order,
//   rec_a = {x, y, radius, API index}    rec_b = {v_x, v_y, mass byte | cell column byte << 8, list head = none}
// so the GS_R + 2 slot runs of a tile and its halo are contiguous byte ranges, staged by the TMA (cp.async.bulk +
// mbarrier, ppc_tma.cuh): no load instructions, no address arithmetic per particle.
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
#define PPC_GS_CAP 1152    // staged particles (centre + halo) per sub-tile
#endif
#ifndef PPC_GS_POOL
#define PPC_GS_POOL 6144   // contact entries per sub-tile (a centre-centre pair takes two)
#endif
#ifndef PPC_GS_TARGET
#define PPC_GS_TARGET 430  // centre particles per tile the host aims for
#endif
#ifndef PPC_GS_TMA
#define PPC_GS_TMA 1       // 0: stage with ordinary loads (debugging aid)
#endif
constexpr int      GS_THREADS = PPC_GS_THREADS;
constexpr int      GS_R       = TILE_R;              // centre rows per tile
constexpr int      GS_CAP     = PPC_GS_CAP;
constexpr int      GS_POOL    = PPC_GS_POOL;
constexpr int      GS_TARGET  = PPC_GS_TARGET;
constexpr int      GS_F       = 4;                   // fine bins per cell
constexpr int      GS_MAXC    = 64;                  // staged columns (centre + halo)
constexpr int      GS_NB      = GS_MAXC * GS_F;      // fine bins per staged row
constexpr int      GS_BINS    = ( GS_R + 2 ) * GS_NB;
constexpr int      GS_STRIP_LIMIT = GS_CAP;          // staged particles (strip + halo) a strip may hold
constexpr uint32_t GS_NONE    = 0xffffu;
constexpr uint32_t GS_TOUCHED = 0x10000u;            // rec_b.aux: a sub-tile that owns a pair with this halo particle changed its velocity
constexpr uint32_t GS_ERR_CAP = 1u, GS_ERR_POOL = 2u, GS_ERR_DEPTH = 4u;
static_assert( GS_CAP <= 2048, "contact entries hold the partner in 11 bits" );
static_assert( GS_BINS * 4 <= GS_POOL * 4, "the bin counters alias the contact pool" );
static_assert( GS_BINS % GS_THREADS == 0, "bins per thread" );

struct GsSmem {
    float4   a[ GS_CAP ];           // x, y, radius, API index (bits), in storage order
    float4   b[ GS_CAP ];           // v_x, v_y (current), mass byte | column byte << 8, list head | flags
    uint16_t pool_val[ GS_POOL ];   // contact entry: partner [10:0], row offset + 1 [12:11], column offset + 1 [14:13], anchored [15]
    uint16_t pool_next[ GS_POOL ];  // next entry of the same particle; the two arrays double as 32-bit bin counters during binning
    uint16_t fend[ GS_BINS ];       // end (exclusive) of fine bin f = row * GS_NB + bin in perm; its begin is fend[f - 1]
    uint16_t perm[ GS_CAP ];        // staged indices sorted by (staged row, fine bin)
    uint16_t cstart[ ( GS_R + 2 ) * ( GS_MAXC + 1 ) ];   // first staged index of staged cell (row, column); [row][ncol] = end of the row
    float    inv_lut[ 256 ];        // 1 / (float)k, the inverse masses of :146-147
    uint32_t run_g0[ GS_R + 2 ], run_s0[ GS_R + 3 ], cen_g0[ GS_R + 1 ], cen_s0[ GS_R + 1 ];
    uint32_t wsum[ 32 ];
    uint32_t colp[ GS_MAXC + 2 ];   // column sums of cell_start over the staged rows (strip planning)
    uint32_t lcount[ 256 ];         // pairs per level, then running begin / end of each level in the level-sorted entry list
    uint32_t rmax_bits, pool_n, err;
    unsigned long long mbar;
};

struct GsLaunch {
    int wc;          // tile width in cells
    int tiles_x;     // tiles per tile row
    int ty_first;    // global tile row of the grid's first row (slab windows; 0 for a whole field)
    int colour;      // 0..3: (tx & 1) + 2 (ty & 1) of the tiles this launch runs
    int ntx;         // tiles of this colour per tile row
    int ty_start;    // first local tile row of this colour
};

// optional dump of the pairs this kernel resolves (parity tests: the hot kernel's own pair set)
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

// One Gauss-Seidel step: the impulse of :158-190 for the pair (a, p) from the current velocities in shared memory.
__device__ __forceinline__ void gs_fire( GsSmem &T, const uint32_t a, const uint32_t p, const float dt ) {
    const float4   A = T.a[ a ], P = T.a[ p ];
    const bool     a_lo = __float_as_uint( A.w ) < __float_as_uint( P.w );   // lo = the lower API index = the reference's i
    const uint32_t lo = a_lo ? a : p, hi = a_lo ? p : a;
    const float4   L = a_lo ? A : P, H = a_lo ? P : A;
    const ContactGeom C = contact_geom( L.x, L.y, L.z, H.x, H.y, H.z, dt );
    float2      *vl = reinterpret_cast<float2 *>( &T.b[ lo ] ), *vh = reinterpret_cast<float2 *>( &T.b[ hi ] );
    float2       ul = *vl, uh = *vh;
    const float  wl = T.inv_lut[ __float_as_uint( T.b[ lo ].z ) & 0xffu ], wh = T.inv_lut[ __float_as_uint( T.b[ hi ].z ) & 0xffu ];
    const ImpulseTerms I = impulse_terms( C, wl, wh, ul.x, ul.y, uh.x, uh.y );
    if ( I.impulse ) {
        ul.x = __fadd_rn( ul.x, I.dvx_lo );
        ul.y = __fadd_rn( ul.y, I.dvy_lo );
        uh.x = __fsub_rn( uh.x, I.dvx_hi );
        uh.y = __fsub_rn( uh.y, I.dvy_hi );
        *vl  = ul;
        *vh  = uh;
    }
}

// One sub-tile: centre columns [cx0, cx1) of the tile [tcx0, tcx1), centre rows [cy0, cy0 + GS_R) (clipped to the grid).
__device__ __forceinline__ void gs_subtile( GsSmem &T, const float4 *__restrict__ rec_a, float4 *rec_b, float4 *__restrict__ pv_out,
                                            const uint32_t *__restrict__ cell_start, const GridDims &G, const float dt, const int cx0, const int cx1,
                                            const int tcx0, const int tcx1, const int cy0, const int tx, const int tyg, const int my_colour,
                                            uint32_t &phase_parity, const bool later_strip, uint32_t *__restrict__ err_flags, const GsDump D ) {
    const int c0 = max( cx0 - 1, 0 ), c1 = min( cx1 + 1, G.cols );      // staged columns [c0, c1)
    const int ncol = c1 - c0, nb = ncol * GS_F;
    uint32_t *cnt32 = reinterpret_cast<uint32_t *>( T.pool_val );       // bin counters / cursors while binning (pool unused then)
    __syncthreads();                                                    // shared memory may still be in use by the previous strip

    // ---- run table: staged row j = 0 .. GS_R+1 is grid row cy0 - 1 + j ----
    if ( threadIdx.x < GS_R + 2 ) {
        const int j = threadIdx.x, gy = cy0 - 1 + j;
        uint32_t  g0 = 0, g1 = 0, m0 = 0, m1 = 0;
        if ( gy >= 0 && gy < G.rows ) {
            const size_t row = (size_t)gy * G.cols;
            g0 = cell_start[ row + c0 ], g1 = cell_start[ row + c1 ];
            if ( j >= 1 && j <= GS_R )
                m0 = cell_start[ row + cx0 ], m1 = cell_start[ row + cx1 ];
        }
        T.run_g0[ j ] = g0;
        T.run_s0[ j ] = g1 - g0;   // length for now
        if ( j >= 1 && j <= GS_R ) {
            T.cen_g0[ j - 1 ] = m0;
            T.cen_s0[ j - 1 ] = m1 - m0;   // length for now
        }
    }
    for ( int f = threadIdx.x; f < GS_BINS; f += GS_THREADS )
        cnt32[ f ] = 0;
    __syncthreads();
    if ( threadIdx.x == 0 ) {
        uint32_t run = 0;
        for ( int j = 0; j < GS_R + 2; j++ ) {
            const uint32_t len = T.run_s0[ j ];
            T.run_s0[ j ]      = run;
            run += len;
        }
        T.run_s0[ GS_R + 2 ] = run;
        uint32_t cen = 0;
        for ( int j = 0; j < GS_R; j++ ) {
            const uint32_t len = T.cen_s0[ j ];
            T.cen_s0[ j ]      = cen;
            cen += len;
        }
        T.cen_s0[ GS_R ] = cen;
        T.rmax_bits      = 0u;
        T.err            = 0u;
#if PPC_GS_TMA
        if ( run != 0 && cen != 0 && run <= (uint32_t)GS_CAP ) {   // stage the runs: two bulk copies per row, one barrier phase
            if ( later_strip )
                __threadfence();   // velocities the previous strip wrote back are read by these copies
            fence_proxy_async();
            mbar_arrive_expect_tx( reinterpret_cast<uint64_t *>( &T.mbar ), run * 32u );
            uint32_t at = 0;
            for ( int j = 0; j < GS_R + 2; j++ ) {
                const uint32_t len = ( j < GS_R + 1 ? T.run_s0[ j + 1 ] : run ) - at;
                if ( len ) {
                    tma_load_1d( &T.a[ at ], rec_a + T.run_g0[ j ], len * 16u, reinterpret_cast<uint64_t *>( &T.mbar ) );
                    tma_load_1d( &T.b[ at ], rec_b + T.run_g0[ j ], len * 16u, reinterpret_cast<uint64_t *>( &T.mbar ) );
                }
                at += len;
            }
        }
#endif
    }
    __syncthreads();
    const uint32_t staged = T.run_s0[ GS_R + 2 ], centre = T.cen_s0[ GS_R ];
    if ( staged == 0 || centre == 0 )
        return;
    if ( staged > (uint32_t)GS_CAP ) {   // a one-column strip that still does not fit: this mode has no generic path
        if ( threadIdx.x == 0 )
            atomicOr( err_flags, GS_ERR_CAP );
        return;
    }

    // ---- while the copies are in flight: staged cell table, inverse-mass table ----
    for ( int e = threadIdx.x; e < ( GS_R + 2 ) * ( ncol + 1 ); e += GS_THREADS ) {
        const int j = e / ( ncol + 1 ), i = e - j * ( ncol + 1 ), gy = cy0 - 1 + j;
        uint32_t  v = T.run_s0[ j ];
        if ( gy >= 0 && gy < G.rows )
            v += cell_start[ (size_t)gy * G.cols + c0 + i ] - T.run_g0[ j ];
        T.cstart[ j * ( GS_MAXC + 1 ) + i ] = (uint16_t)v;
    }
    if ( threadIdx.x < 256 )
        T.inv_lut[ threadIdx.x ] = __fdiv_rn( 1.0f, (float)threadIdx.x );
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
        T.b[ t ] = __ldcg( rec_b + g );
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
    if ( threadIdx.x == 0 )
        T.pool_n = 0u;
    __syncthreads();
    const float R = __uint_as_float( T.rmax_bits );

    // ---- one thread per centre particle: candidates, contact list, position pushes ----
    for ( uint32_t t = threadIdx.x; t < centre; t += GS_THREADS ) {
        int j = 1;
        while ( j < GS_R && t >= T.cen_s0[ j ] )
            j++;
        const uint32_t gk   = T.cen_g0[ j - 1 ] + ( t - T.cen_s0[ j - 1 ] );   // storage slot
        const uint32_t self = T.run_s0[ j ] + ( gk - T.run_g0[ j ] );          // staged index
        const float4   me4  = T.a[ self ];
        const float    x = me4.x, y = me4.y, r = me4.z;
        const uint32_t gid   = __float_as_uint( me4.w );
        const uint32_t mcol  = __float_as_uint( T.b[ self ].z );
        const int      mycol = (int)( ( mcol >> 8 ) & 0xffu );
        const int      gy    = cy0 - 1 + j;
        const float    reach = __fadd_rn( __fadd_rn( r, R ), 0.02f );
        const int      b_lo  = gs_fine_bin( __fsub_rn( x, reach ), x0, inv_w, nb ), b_hi = gs_fine_bin( __fadd_rn( x, reach ), x0, inv_w, nb );
        const float    top   = __fmul_rn( (float)( gy + G.row0 ), G.cell_size ), bot = __fmul_rn( (float)( gy + G.row0 + 1 ), G.cell_size );

        uint32_t head = GS_NONE, tail = GS_NONE;
#pragma unroll 1
        for ( int d = 0; d < 3; d++ ) {
            const int ry = gy - 1 + d;
            if ( ry < 0 || ry >= G.rows )
                continue;
            if ( d == 0 && __fsub_rn( y, reach ) >= top )   // nothing in the row above can be within reach
                continue;
            if ( d == 2 && __fadd_rn( y, reach ) < bot )
                continue;
            const int      f_lo = ( j - 1 + d ) * GS_NB + b_lo, f_hi = ( j - 1 + d ) * GS_NB + b_hi;
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
                    const int dc = (int)(int8_t)( ( ( __float_as_uint( T.b[ s ].z ) >> 8 ) & 0xffu ) - (uint32_t)mycol );
                    if ( dc < -1 || dc > 1 )   // two columns apart: outside the reference's 3x3 scan (:331-340)
                        continue;
                    const uint32_t at = atomicAdd( &T.pool_n, 1u );
                    if ( at < (uint32_t)GS_POOL ) {
                        T.pool_val[ at ]  = (uint16_t)( s | ( (uint32_t)d << 11 ) | ( (uint32_t)( dc + 1 ) << 13 ) );
                        T.pool_next[ at ] = (uint16_t)GS_NONE;
                        if ( tail == GS_NONE )
                            head = at;
                        else
                            T.pool_next[ tail ] = (uint16_t)at;
                        tail = at;
                    } else
                        T.err = GS_ERR_POOL;
                }
            }
        }
        float X = x, Y = y;
        if ( head != GS_NONE ) {
            // The list is put into the order the reference's pair list touches this particle -- pairs (i, me), i < me, by
            // ascending i; then pairs (me, j) by scan cell (dy outer, dx inner), descending index inside a cell (:311-383) --
            // by selection, swapping entry values along the links; each entry is consumed as soon as it is in place.
            auto sort_key = [ & ]( const uint32_t v ) -> unsigned long long {
                const uint32_t g = __float_as_uint( T.a[ v & 0x7ffu ].w );
                return g < gid ? (unsigned long long)g
                               : ( 1ull << 48 ) | ( (unsigned long long)( ( ( v >> 11 ) & 3u ) * 3u + ( ( v >> 13 ) & 3u ) ) << 32 ) | ( 0xffffffffu - g );
            };
            const float wme      = T.inv_lut[ mcol & 0xffu ];
            const bool  row_last = ( j == GS_R ), row_first = ( j == 1 );
#pragma unroll 1
            for ( uint32_t e = head; e != GS_NONE; e = T.pool_next[ e ] ) {
                uint32_t           best = e, v = T.pool_val[ e ];
                unsigned long long bkey = sort_key( v );
#pragma unroll 1
                for ( uint32_t f = T.pool_next[ e ]; f != GS_NONE; f = T.pool_next[ f ] ) {
                    const uint32_t           vf = T.pool_val[ f ];
                    const unsigned long long kf = sort_key( vf );
                    if ( kf < bkey ) {
                        bkey = kf;
                        best = f;
                        v    = vf;
                    }
                }
                if ( best != e )
                    T.pool_val[ best ] = T.pool_val[ e ];
                const uint32_t s     = v & 0x7ffu;
                const int      d     = (int)( ( v >> 11 ) & 3u ) - 1, dc = (int)( ( v >> 13 ) & 3u ) - 1;   // partner's cell relative to mine
                const float4   c     = T.a[ s ];
                const uint32_t cg    = __float_as_uint( c.w );
                const bool     me_lo = gid < cg;
                const float    wq    = T.inv_lut[ __float_as_uint( T.b[ s ].z ) & 0xffu ];
                {   // position push, operands selected by role (lo = the pair's i), :145-156
                    const PushTerms P = push_terms( contact_geom( me_lo ? x : c.x, me_lo ? y : c.y, me_lo ? r : c.z, me_lo ? c.x : x, me_lo ? c.y : y,
                                                                  me_lo ? c.z : r, dt ), me_lo ? wme : wq, me_lo ? wq : wme );
                    if ( P.push ) {
                        X = me_lo ? __fadd_rn( X, P.dpx_lo ) : __fsub_rn( X, P.dpx_hi );
                        Y = me_lo ? __fadd_rn( Y, P.dpy_lo ) : __fsub_rn( Y, P.dpy_hi );
                    }
                }
                // who resolves this pair's impulse, and which of its two particles lists it (see the header)
                const int  pc       = c0 + ( mycol - ( c0 & 0xff ) & 0xff ) + dc;          // partner's grid column
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
                    if ( anchored )
                        atomicOr( reinterpret_cast<uint32_t *>( &T.b[ s ].w ), GS_TOUCHED );
                }
                T.pool_val[ e ] = (uint16_t)( v | ( anchored ? 0x8000u : 0u ) );
                if ( anchored && D.pairs ) {
                    const uint32_t at = atomicAdd( D.count, 1u );
                    if ( at < D.cap ) {
                        D.pairs[ 2 * at ]     = me_lo ? gid : cg;
                        D.pairs[ 2 * at + 1 ] = me_lo ? cg : gid;
                    }
                }
            }
            T.b[ self ].w = __uint_as_float( ( __float_as_uint( T.b[ self ].w ) & ~0xffffu ) | head );
        }
        *reinterpret_cast<float2 *>( pv_out + gk ) = make_float2( X, Y );
    }
    __syncthreads();
    if ( T.err ) {
        if ( threadIdx.x == 0 )
            atomicOr( err_flags, T.err );
        return;
    }

    // ---- Gauss-Seidel over the pairs of this sub-tile ----
    // The ORDER is the nine phases of disjoint cell pairs described in the header; the EXECUTION is levelled: every pair gets
    // level = 1 + max(level of the previous pair of either particle in that order), a cheap integer pass; pairs of one level
    // share no particle and fire together, one thread per pair, all lanes busy.  Each particle still receives its impulses
    // in exactly the order of the phases, so the result is the sequential one.
    const int ncell  = ( GS_R + 2 ) * ncol;   // <= GS_THREADS (the host bounds the tile width): one thread per staged cell
    const int row_g0 = G.row0 + cy0 - 1;      // whole-field row of staged row 0
    uint16_t *seg    = T.fend;                // fend + perm are dead: entry numbers, first by (cell, direction), later by level
    static_assert( sizeof( T.fend ) + sizeof( T.perm ) >= GS_POOL / 2 * sizeof( uint16_t ), "an anchored entry per pair" );
    const int  xj = (int)threadIdx.x / ncol, xc = (int)threadIdx.x - xj * ncol;                       // this thread's cell
    const bool x_in = (int)threadIdx.x < ncell && xj >= 1 && xj <= GS_R && c0 + xc >= cx0 && c0 + xc < cx1;   // anchors are centre particles
    const uint32_t x_pb = x_in ? T.cstart[ xj * ( GS_MAXC + 1 ) + xc ] : 0u, x_pe = x_in ? T.cstart[ xj * ( GS_MAXC + 1 ) + xc + 1 ] : 0u;
    // (a) anchored entries of this cell per direction code d * 3 + dc (nine 7-bit counters in one word)
    unsigned long long cnt = 0ull;
    uint32_t           mine = 0;
#pragma unroll 1
    for ( uint32_t a = x_pb; a < x_pe; a++ )
#pragma unroll 1
        for ( uint32_t e = __float_as_uint( T.b[ a ].w ) & 0xffffu; e != GS_NONE; e = T.pool_next[ e ] ) {
            const uint32_t v = T.pool_val[ e ];
            if ( v & 0x8000u ) {
                cnt += 1ull << ( 7u * ( ( ( v >> 11 ) & 3u ) * 3u + ( ( v >> 13 ) & 3u ) ) );
                mine++;
            }
        }
    if ( threadIdx.x < 256 )
        T.lcount[ threadIdx.x ] = 0u;
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
        if ( __shfl_sync( 0xffffffffu, ws, GS_THREADS / 32 - 1 ) > (uint32_t)( ( sizeof( T.fend ) + sizeof( T.perm ) ) / sizeof( uint16_t ) ) )
            T.err = GS_ERR_POOL;   // more pairs than the entry list takes (same value in every thread)
    }
    if ( mine > 127u )   // a 7-bit counter may have wrapped (dozens of particles on one spot): not a field for this mode
        T.err = GS_ERR_POOL;
    __syncthreads();
    if ( T.err ) {
        if ( threadIdx.x == 0 )
            atomicOr( err_flags, T.err );
        return;
    }
    // (b) entry numbers of this cell, grouped by direction code; the entry remembers its anchor where its link was
    {
        unsigned long long cur = 0ull;
#pragma unroll 1
        for ( uint32_t a = x_pb; a < x_pe; a++ )
#pragma unroll 1
            for ( uint32_t e = __float_as_uint( T.b[ a ].w ) & 0xffffu; e != GS_NONE; ) {
                const uint32_t nxt = T.pool_next[ e ], v = T.pool_val[ e ];
                if ( v & 0x8000u ) {
                    const uint32_t code = ( ( v >> 11 ) & 3u ) * 3u + ( ( v >> 13 ) & 3u );
                    uint32_t       pre  = 0;
#pragma unroll
                    for ( uint32_t k = 0; k < 8; k++ )
                        pre += k < code ? (uint32_t)( cnt >> ( 7u * k ) ) & 127u : 0u;
                    seg[ seg_base + pre + ( (uint32_t)( cur >> ( 7u * code ) ) & 127u ) ] = (uint16_t)e;
                    cur += 1ull << ( 7u * code );
                    T.pool_next[ e ] = (uint16_t)a;   // bits [10:0] anchor; [15:11] and pool_val [14:11] will hold the level
                    T.pool_val[ e ]  = (uint16_t)( v & 0x87ffu );
                }
                e = nxt;
            }
    }
    __syncthreads();
    // (c) levels, phase by phase: in a phase a cell's thread walks the pairs of ONE direction code (towards the phase's other
    //     cell when this cell is the base, i.e. has the phase's parity; backwards to a halo base cell otherwise)
#pragma unroll 1
    for ( int phase = 0; phase < 9; phase++ ) {
        const int cls = phase == 0 ? 0 : ( phase + 1 ) >> 1;   // 0 same cell, 1 east, 2 south, 3 south-east, 4 south-west
        const int par = phase == 0 ? 0 : ( phase + 1 ) & 1;
        if ( x_in && mine ) {
            const bool     base = cls == 0 || ( ( cls == 1 ? c0 + xc : row_g0 + xj ) & 1 ) == par;
            const int      dj = cls >= 2 ? 1 : 0, dcl = cls == 1 || cls == 3 ? 1 : ( cls == 4 ? -1 : 0 );
            const uint32_t code = (uint32_t)( ( ( base ? dj : -dj ) + 1 ) * 3 + ( base ? dcl : -dcl ) + 1 );
            uint32_t       pre  = 0;
#pragma unroll
            for ( uint32_t k = 0; k < 8; k++ )
                pre += k < code ? (uint32_t)( cnt >> ( 7u * k ) ) & 127u : 0u;
            const uint32_t kb = seg_base + pre, ke = kb + ( (uint32_t)( cnt >> ( 7u * code ) ) & 127u );
#pragma unroll 1
            for ( uint32_t k = kb; k < ke; k++ ) {
                const uint32_t e = seg[ k ];
                const uint32_t a = T.pool_next[ e ] & 0x7ffu, q = T.pool_val[ e ] & 0x7ffu;
                uint32_t      *wa = reinterpret_cast<uint32_t *>( &T.b[ a ].w ), *wq = reinterpret_cast<uint32_t *>( &T.b[ q ].w );
                const uint32_t ua = *wa, uq = *wq;
                uint32_t       lv = max( ( ua >> 17 ) & 0x3ffu, ( uq >> 17 ) & 0x3ffu ) + 1u;
                if ( lv > 255u ) {
                    T.err = GS_ERR_DEPTH;
                    lv    = 255u;
                }
                *wa = ( ua & 0x1ffffu ) | ( lv << 17 );
                *wq = ( uq & 0x1ffffu ) | ( lv << 17 );
                T.pool_val[ e ]  = (uint16_t)( 0x8000u | q | ( ( lv & 15u ) << 11 ) );
                T.pool_next[ e ] = (uint16_t)( a | ( ( lv >> 4 ) << 11 ) );
                atomicAdd( &T.lcount[ lv ], 1u );
            }
        }
        __syncthreads();
    }
    if ( T.err ) {
        if ( threadIdx.x == 0 )
            atomicOr( err_flags, T.err );
        return;
    }
    // (d) entries by level: scan of the level counts, scatter (the order inside a level is immaterial)
    const uint32_t n_entries = min( T.pool_n, (uint32_t)GS_POOL );
    if ( threadIdx.x < 32 ) {
        uint32_t loc[ 8 ], sum = 0;
#pragma unroll
        for ( int i = 0; i < 8; i++ ) {
            loc[ i ] = T.lcount[ threadIdx.x * 8 + i ];
            sum += loc[ i ];
        }
        uint32_t inc = sum;
#pragma unroll
        for ( int d = 1; d < 32; d <<= 1 ) {
            const uint32_t o = __shfl_up_sync( 0xffffffffu, inc, d );
            if ( threadIdx.x >= (unsigned)d )
                inc += o;
        }
        uint32_t run = inc - sum;
#pragma unroll
        for ( int i = 0; i < 8; i++ ) {
            T.lcount[ threadIdx.x * 8 + i ] = run;   // begin of the level; the scatter advances it to the end
            run += loc[ i ];
        }
        if ( threadIdx.x == 31 )
            T.wsum[ 0 ] = run;   // pairs of this sub-tile
    }
    __syncthreads();
    const uint32_t n_pairs = T.wsum[ 0 ];
    for ( uint32_t e = threadIdx.x; e < n_entries; e += GS_THREADS ) {
        const uint32_t v = T.pool_val[ e ];
        if ( v & 0x8000u )
            seg[ atomicAdd( &T.lcount[ ( ( v >> 11 ) & 15u ) | ( ( (uint32_t)T.pool_next[ e ] >> 11 ) << 4 ) ], 1u ) ] = (uint16_t)e;
    }
    __syncthreads();
    // (e) fire, level by level
    {
        uint32_t kb = 0;
#pragma unroll 1
        for ( uint32_t lv = 1; kb < n_pairs; lv++ ) {
            const uint32_t ke = T.lcount[ lv ];
            for ( uint32_t k = kb + threadIdx.x; k < ke; k += GS_THREADS ) {
                const uint32_t e = seg[ k ];
                gs_fire( T, T.pool_next[ e ] & 0x7ffu, T.pool_val[ e ] & 0x7ffu, dt );
            }
            kb = ke;
            __syncthreads();
        }
    }

    // ---- final velocities of this sub-tile's particles; velocities of halo particles it changed go back in place ----
    for ( uint32_t t = threadIdx.x; t < centre; t += GS_THREADS ) {
        int j = 1;
        while ( j < GS_R && t >= T.cen_s0[ j ] )
            j++;
        const uint32_t gk   = T.cen_g0[ j - 1 ] + ( t - T.cen_s0[ j - 1 ] );
        const uint32_t self = T.run_s0[ j ] + ( gk - T.run_g0[ j ] );
        *( reinterpret_cast<float2 *>( pv_out + gk ) + 1 ) = *reinterpret_cast<const float2 *>( &T.b[ self ] );
    }
    for ( uint32_t t = threadIdx.x; t < staged; t += GS_THREADS )
        if ( __float_as_uint( T.b[ t ].w ) & GS_TOUCHED ) {
            int j = 0;
            while ( j < GS_R + 1 && t >= T.run_s0[ j + 1 ] )
                j++;
            *reinterpret_cast<float2 *>( rec_b + ( T.run_g0[ j ] + ( t - T.run_s0[ j ] ) ) ) = *reinterpret_cast<const float2 *>( &T.b[ t ] );
        }
}

__global__ void __launch_bounds__( GS_THREADS, PPC_GS_MINBLOCKS ) k_collide_gs( const float4 *__restrict__ rec_a, float4 *rec_b, float4 *__restrict__ pv_out,
                                                               const uint32_t *__restrict__ cell_start, const GridDims G, const float dt,
                                                               const GsLaunch L, uint32_t *__restrict__ err_flags, const GsDump D ) {
    extern __shared__ __align__( 16 ) unsigned char smem_raw[];
    GsSmem &T = *reinterpret_cast<GsSmem *>( smem_raw );
    const int bx = (int)( blockIdx.x % (unsigned)L.ntx ), by = (int)( blockIdx.x / (unsigned)L.ntx );
    const int tx = 2 * bx + ( L.colour & 1 ), ty = 2 * by + L.ty_start;   // ty: tile row inside this grid
    const int tyg = L.ty_first + ty;                                       // tile row of the whole field
    const int cx0 = tx * L.wc, cx1 = min( cx0 + L.wc, G.cols );
    const int cy0 = tyg * GS_R - G.row0;                                   // may be negative for a slab window's first tile row
    if ( threadIdx.x == 0 )
        mbar_init( reinterpret_cast<uint64_t *>( &T.mbar ), 1u );
    // Column sums of the tile and its halo: cell_start is a running sum along a grid row, so P[c] = sum over the staged rows of
    // cell_start[row][c0 + c] makes the staged particle count of the columns [a, b) equal to P[b] - P[a].
    const int c0 = max( cx0 - 1, 0 ), c1 = min( cx1 + 1, G.cols );
    for ( int c = threadIdx.x; c <= c1 - c0; c += GS_THREADS ) {
        uint32_t sum = 0;
        for ( int j = 0; j < GS_R + 2; j++ ) {
            const int gy = cy0 - 1 + j;
            if ( gy >= 0 && gy < G.rows )
                sum += cell_start[ (size_t)gy * G.cols + c0 + c ];
        }
        T.colp[ c ] = sum;
    }
    __syncthreads();
    const uint32_t total = T.colp[ c1 - c0 ] - T.colp[ 0 ];
    if ( total == 0 )
        return;
    // A tile that would overflow shared memory is cut into column strips, greedily from the left: each strip takes as many
    // columns as keep its staged particles (strip + one halo column either side) within GS_STRIP_LIMIT, and at least one.
    uint32_t parity = 0;
    for ( int sx = cx0; sx < cx1; ) {
        int ex = cx1;
        if ( total > (uint32_t)GS_STRIP_LIMIT ) {
            ex = sx + 1;
            while ( ex < cx1 && T.colp[ min( ex + 2, c1 ) - c0 ] - T.colp[ max( sx - 1, c0 ) - c0 ] <= (uint32_t)GS_STRIP_LIMIT )
                ex++;
        }
        gs_subtile( T, rec_a, rec_b, pv_out, cell_start, G, dt, sx, ex, cx0, cx1, cy0, tx, tyg, L.colour, parity, sx != cx0, err_flags, D );
        sx = ex;
    }
}

}   // namespace ppc
