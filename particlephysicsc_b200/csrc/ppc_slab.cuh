// ppc_slab.cuh -- kernels of the multi-GPU slab decomposition (no reference counterpart: the
// reference is one thread on one address space).
//
// The cell index is row-major (reference src/particles_solver.c:291), so a slab of whole cell rows is a contiguous key
// range and -- storage being sorted -- a contiguous slot range.  Rank k OWNS the particles whose cell row lies in
// [own_begin, own_end) and works on a WINDOW [row_lo, row_hi) = owned rows plus `halo` ghost rows on either side.
// Per substep, after the integrate + wall pass of its owned particles, a rank sends each neighbour every owned particle
// within `halo` rows of (or beyond) the common border -- that one message carries both the migrants and the ghosts -- and
// keeps its own emigrants as ghosts.  The receiver sorts everything it now holds by local cell key; what falls outside the
// window lands in a drop bin at the end of the sorted storage.  Narrow phase and resolve then run over the window exactly
// as on one GPU; afterwards the owned particles are the slot range of the owned rows, ghosts are forgotten.
//
// Exactness.  Keys, sorted order, cell-start table and candidate pairs of the owned rows need a halo of one row
// (cellSize = 2 maxRadius, :241-244).  Position pushes are sums of terms frozen before anything moves (:84-156): one row.
// Velocities follow the reference's sequential pair order; an error made at the outer edge of the halo (whose particles
// miss their outward contacts) can travel one contact, i.e. at most one cell row, per level of the pair dependency graph.
// The resolve therefore marks those particles, and every particle that resolves a pair with a marked one (ST_TAINT in
// ppc_collide.cuh); while no owned particle gets marked, owned velocities are bit-identical to the single-GPU run.  The
// host checks that every substep, and sizes the halo from how far the mark was seen to travel (1-5 rows in practice).
#pragma once
#include "ppc_collide.cuh"
#include "ppc_common.cuh"
#include "ppc_grid.cuh"

namespace ppc {

// one particle on the wire (32 B): everything a rank needs to own it or to use it as a ghost
struct __align__( 16 ) SlabRec {
    float4   pv;      // p_x, p_y, v_x, v_y (after integrate + walls)
    float    radius, mass;
    uint32_t gid;     // global API index: fixes the pair order (j > i) on every rank alike
    uchar4   col;
};

struct SlabRows {
    int own_begin, own_end;   // owned cell rows
    int halo;                 // ghost rows on either side
    int has_lo, has_hi;       // a neighbour below (smaller rows) / above
    int cols;
};

// which neighbours particle with cell row cy goes to (bit 0: lower, bit 1: upper), and whether it left the window
__device__ __forceinline__ unsigned slab_targets( const SlabRows &R, const int cy, bool &lost ) {
    unsigned t = 0;
    if ( R.has_lo && cy < R.own_begin + R.halo )
        t |= 1u;
    if ( R.has_hi && cy >= R.own_end - R.halo )
        t |= 2u;
    lost = ( R.has_lo && cy < R.own_begin - R.halo ) || ( R.has_hi && cy >= R.own_end + R.halo );
    return t;
}

// counts[0..1] = records for the lower / upper neighbour, counts[2] = particles that jumped past the halo (fatal)
__global__ void __launch_bounds__( STREAM_THREADS ) k_slab_count( const uint32_t *__restrict__ key, const size_t n, const SlabRows R,
                                                                  uint32_t *__restrict__ counts ) {
    const size_t k = (size_t)blockIdx.x * STREAM_THREADS + threadIdx.x;
    unsigned     t = 0;
    bool         lost = false;
    if ( k < n )
        t = slab_targets( R, (int)( key[ k ] / (uint32_t)R.cols ), lost );
    const unsigned lo = __ballot_sync( 0xffffffffu, t & 1u ), hi = __ballot_sync( 0xffffffffu, t & 2u ), ls = __ballot_sync( 0xffffffffu, lost );
    if ( ( threadIdx.x & 31u ) == 0 ) {
        if ( lo )
            atomicAdd( &counts[ 0 ], (uint32_t)__popc( lo ) );
        if ( hi )
            atomicAdd( &counts[ 1 ], (uint32_t)__popc( hi ) );
        if ( ls )
            atomicAdd( &counts[ 2 ], (uint32_t)__popc( ls ) );
    }
}

// append the records to the two send buffers (order inside a buffer is arbitrary: the receiver sorts by (key, gid))
__global__ void __launch_bounds__( STREAM_THREADS ) k_slab_pack( const ParticleSet S, const size_t n, const SlabRows R, SlabRec *__restrict__ send_lo,
                                                                 SlabRec *__restrict__ send_hi, uint32_t *__restrict__ fill ) {
    const size_t   k    = (size_t)blockIdx.x * STREAM_THREADS + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    unsigned       t    = 0;
    bool           lost = false;
    if ( k < n )
        t = slab_targets( R, (int)( S.key[ k ] / (uint32_t)R.cols ), lost );
    const unsigned lo = __ballot_sync( 0xffffffffu, t & 1u ), hi = __ballot_sync( 0xffffffffu, t & 2u );
    if ( !( lo | hi ) )
        return;
    uint32_t base_lo = 0, base_hi = 0;
    if ( lane == 0 ) {
        if ( lo )
            base_lo = atomicAdd( &fill[ 0 ], (uint32_t)__popc( lo ) );
        if ( hi )
            base_hi = atomicAdd( &fill[ 1 ], (uint32_t)__popc( hi ) );
    }
    base_lo = __shfl_sync( 0xffffffffu, base_lo, 0 );
    base_hi = __shfl_sync( 0xffffffffu, base_hi, 0 );
    if ( t ) {
        SlabRec      r;
        const float2 a = S.rm[ k ];
        r.pv = S.pv[ k ], r.radius = a.x, r.mass = a.y, r.gid = S.gid[ k ], r.col = S.col[ k ];
        const unsigned below = ( 1u << lane ) - 1u;
        if ( t & 1u )
            send_lo[ base_lo + (uint32_t)__popc( lo & below ) ] = r;
        if ( t & 2u )
            send_hi[ base_hi + (uint32_t)__popc( hi & below ) ] = r;
    }
}

// ---- the fused first pass of a slab substep ---------------------------------------------------------------------------------
// K1 of the single-GPU path (integrate, colour, walls, cell key: same float operations, ppc_math.cuh) over the owned particles,
// and in the same pass everything that used to be three more sweeps over them: which neighbour each particle goes to (counts +
// the two lists of slots to pack, instead of k_slab_count and a k_slab_pack over all owned particles), its key LOCAL to the window
// and the window's histogram (instead of k_slab_localize over the owned particles), and its rank inside its cell (so that the
// counting sort places it without atomics, k_cell_place).
// The pass runs over the slots [first, first + count) of the owned particles.  `interior`: these slots lay more than 2 * halo rows
// from both borders at the last grid build, so none of them can reach a send zone without moving more than `halo` rows in one
// substep -- which is the fatal "lost" condition anyway: the border slots are done first and their records are on the wire while
// this (much larger) part is still being integrated.
__global__ void __launch_bounds__( STREAM_THREADS ) k_slab_integrate( ParticleSet S, const size_t first, const size_t count, const StepParams P,
                                                                      const SlabRows R, const int row_lo, const int row_hi, const uint32_t drop,
                                                                      uint32_t *__restrict__ cell_count, uint32_t *__restrict__ cell_rank,
                                                                      uint32_t *__restrict__ counts, uint32_t *__restrict__ list_lo,
                                                                      uint32_t *__restrict__ list_hi, const int interior ) {
    const size_t   t_in  = (size_t)blockIdx.x * STREAM_THREADS + threadIdx.x;
    const size_t   k     = first + t_in;
    const bool     valid = t_in < count;
    const unsigned lane  = threadIdx.x & 31u;
    uint32_t       lk    = 0xffffffffu;
    unsigned       t     = 0;
    bool           lost  = false;
    if ( valid ) {
        float4       s = S.pv[ k ];
        const float2 a = S.rm[ k ];
        integrate( s.x, s.y, s.z, s.w, P.dt_update );
        if ( P.do_colour )
            S.col[ k ] = energy_colour( s.z, s.w, a.y );
        walls( s.x, s.y, s.z, s.w, a.x, P.wall_w, P.wall_h );
        const int cx = cell_coord( s.x, P.grid.cell_size, P.grid.cols );   // P.grid: the WHOLE field's grid
        const int cy = cell_coord( s.y, P.grid.cell_size, P.grid.rows );
        t            = slab_targets( R, cy, lost );
        if ( interior && t ) {   // came from deep inside the slab into a send zone: more than `halo` rows in one substep
            lost = true;
            t    = 0;
        }
        lk           = ( cy >= row_lo && cy < row_hi ) ? (uint32_t)( ( cy - row_lo ) * P.grid.cols + cx ) : drop;
        S.key[ k ]   = lk;
        S.pv[ k ]    = s;
    }
    const unsigned lo = __ballot_sync( 0xffffffffu, t & 1u ), hi = __ballot_sync( 0xffffffffu, t & 2u ), ls = __ballot_sync( 0xffffffffu, lost );
    if ( lo | hi | ls ) {   // rare: only warps near a border
        uint32_t base_lo = 0, base_hi = 0;
        if ( lane == 0 ) {
            if ( lo )
                base_lo = atomicAdd( &counts[ 0 ], (uint32_t)__popc( lo ) );
            if ( hi )
                base_hi = atomicAdd( &counts[ 1 ], (uint32_t)__popc( hi ) );
            if ( ls )
                atomicAdd( &counts[ 2 ], (uint32_t)__popc( ls ) );
        }
        base_lo = __shfl_sync( 0xffffffffu, base_lo, 0 );
        base_hi = __shfl_sync( 0xffffffffu, base_hi, 0 );
        const unsigned below = ( 1u << lane ) - 1u;
        if ( t & 1u )
            list_lo[ base_lo + (uint32_t)__popc( lo & below ) ] = (uint32_t)k;
        if ( t & 2u )
            list_hi[ base_hi + (uint32_t)__popc( hi & below ) ] = (uint32_t)k;
    }
    const unsigned peers  = __match_any_sync( 0xffffffffu, lk );
    const unsigned leader = (unsigned)( __ffs( peers ) - 1 );
    uint32_t       base   = 0;
    if ( valid && lane == leader )
        base = atomicAdd( &cell_count[ lk ], (uint32_t)__popc( peers ) );
    base = __shfl_sync( 0xffffffffu, base, leader );
    if ( valid )
        cell_rank[ k ] = base + (uint32_t)__popc( peers & ( ( 1u << lane ) - 1u ) );
}

// the records of the listed slots -> one send buffer (order inside a buffer is arbitrary: the receiver sorts by (key, gid))
__global__ void __launch_bounds__( STREAM_THREADS ) k_slab_pack_list( const ParticleSet S, const uint32_t *__restrict__ list, const size_t count,
                                                                      SlabRec *__restrict__ send ) {
    const size_t i = (size_t)blockIdx.x * STREAM_THREADS + threadIdx.x;
    if ( i >= count )
        return;
    const uint32_t k = list[ i ];
    SlabRec        r;
    const float2   a = S.rm[ k ];
    r.pv = S.pv[ k ], r.radius = a.x, r.mass = a.y, r.gid = S.gid[ k ], r.col = S.col[ k ];
    send[ i ] = r;
}

// received records -> storage slots [first, first + capacity), with their global cell key (:280-291, same float operations
// as K1).  With a header (fixed-size messages: record 0 holds the number of records that follow) the slots past that number
// get a key beyond every window, which the sort puts into the drop bin.
__global__ void __launch_bounds__( STREAM_THREADS ) k_slab_append( const ParticleSet S, const size_t first, const SlabRec *__restrict__ recv,
                                                                   const size_t capacity, const GridDims G, const int with_header ) {
    const size_t t = (size_t)blockIdx.x * STREAM_THREADS + threadIdx.x;
    if ( t >= capacity )
        return;
    const size_t k = first + t;
    if ( with_header ) {
        const uint32_t count = *reinterpret_cast<const uint32_t *>( recv );
        if ( t >= count ) {
            S.key[ k ] = 0xffffffffu;
            S.gid[ k ] = 0xffffffffu;
            return;
        }
        recv += 1;
    }
    const SlabRec r = recv[ t ];
    S.pv[ k ]  = r.pv;
    S.rm[ k ]  = make_float2( r.radius, r.mass );
    S.gid[ k ] = r.gid;
    S.col[ k ] = r.col;
    const int cx = cell_coord( r.pv.x, G.cell_size, G.cols ), cy = cell_coord( r.pv.y, G.cell_size, G.rows );
    S.key[ k ]   = (uint32_t)( cy * G.cols + cx );
}

// global key -> key local to the window [row_lo, row_hi) (or the drop bin `drop`), and the histogram of those (K3)
// (with cell_rank: also the rank of every particle inside its cell, as the fused first pass leaves it for the owned particles)
__global__ void __launch_bounds__( STREAM_THREADS ) k_slab_localize( uint32_t *__restrict__ key, const size_t n, const int cols, const int row_lo,
                                                                     const int row_hi, const uint32_t drop, uint32_t *__restrict__ cell_count,
                                                                     uint32_t *__restrict__ cell_rank ) {
    const size_t   k     = (size_t)blockIdx.x * STREAM_THREADS + threadIdx.x;
    const bool     valid = k < n;
    const unsigned lane  = threadIdx.x & 31u;
    uint32_t       lk    = 0xffffffffu;
    if ( valid ) {
        const uint32_t g  = key[ k ];
        const int      cy = (int)( g / (uint32_t)cols );
        lk                = ( cy >= row_lo && cy < row_hi ) ? g - (uint32_t)row_lo * (uint32_t)cols : drop;
        key[ k ]          = lk;
    }
    const unsigned peers  = __match_any_sync( 0xffffffffu, lk );
    const unsigned leader = (unsigned)( __ffs( peers ) - 1 );
    uint32_t       base   = 0;
    if ( valid && lane == leader )
        base = atomicAdd( &cell_count[ lk ], (uint32_t)__popc( peers ) );
    if ( cell_rank ) {
        base = __shfl_sync( 0xffffffffu, base, leader );
        if ( valid )
            cell_rank[ k ] = base + (uint32_t)__popc( peers & ( ( 1u << lane ) - 1u ) );
    }
}

// After the resolve: did the inexactness mark (ST_TAINT) reach an owned particle, and how many rows did it travel?  A contact
// spans at most one cell row, so a mark inside the owned rows implies a marked particle in the first or the last owned
// row: only the two ghost bands plus those two rows are looked at (slots [0, lo_end) and [hi_begin, n)).
// `state` is the array the resolve engine kept its state words in (stride in words: 1 for WaveBuffers::state, 8 for WaveRec).
__global__ void __launch_bounds__( STREAM_THREADS ) k_slab_taint_scan( const uint32_t *__restrict__ state, const uint32_t stride, const uint16_t *__restrict__ deg,
                                                                       const uint32_t *__restrict__ key, const GridDims G, const uint32_t lo_end,
                                                                       const uint32_t hi_begin, const uint32_t n, const uint32_t own_first,
                                                                       const uint32_t own_end, uint32_t *__restrict__ counters ) {
    const uint32_t t = blockIdx.x * STREAM_THREADS + threadIdx.x;
    const uint32_t k = t < lo_end ? t : hi_begin + ( t - lo_end );
    if ( k >= n || ( t >= lo_end && k < lo_end ) )
        return;
    if ( deg[ k ] == 0 || !( __ldcg( &state[ (size_t)k * stride ] ) & ST_TAINT ) )
        return;
    if ( k >= own_first && k < own_end )
        atomicOr( &counters[ 10 ], 1u );
    const int row = (int)( key[ k ] / (uint32_t)G.cols );
    int       depth = 0x7fffffff;
    if ( G.taint_lo >= 0 )
        depth = row - G.taint_lo;
    if ( G.taint_hi >= 0 )
        depth = min( depth, G.taint_hi - row );
    atomicMax( &counters[ 11 ], (uint32_t)max( depth, 0 ) );
}

// host -> device with explicit global indices (slab ranks hold an arbitrary subset of the particles)
__global__ void __launch_bounds__( STREAM_THREADS ) k_slab_upload( const ParticleSet S, const float *__restrict__ px, const float *__restrict__ py,
                                                                   const float *__restrict__ vx, const float *__restrict__ vy,
                                                                   const float *__restrict__ mass, const float *__restrict__ radius,
                                                                   const uchar4 *__restrict__ col, const uint32_t *__restrict__ gid, const size_t n ) {
    const size_t k = (size_t)blockIdx.x * STREAM_THREADS + threadIdx.x;
    if ( k >= n )
        return;
    S.pv[ k ]  = make_float4( px[ k ], py[ k ], vx[ k ], vy[ k ] );
    S.rm[ k ]  = make_float2( radius[ k ], mass[ k ] );
    S.gid[ k ] = gid[ k ];
    S.key[ k ] = 0u;
    S.col[ k ] = col[ k ];
}

// device -> host: the owned slot range, in storage order (ascending (cell key, global index)), with its global indices
__global__ void __launch_bounds__( STREAM_THREADS ) k_slab_download( const ParticleSet S, const size_t first, const size_t n, float *__restrict__ px,
                                                                     float *__restrict__ py, float *__restrict__ vx, float *__restrict__ vy,
                                                                     float *__restrict__ mass, float *__restrict__ radius, uchar4 *__restrict__ col,
                                                                     uint32_t *__restrict__ gid ) {
    const size_t t = (size_t)blockIdx.x * STREAM_THREADS + threadIdx.x;
    if ( t >= n )
        return;
    const float4 s = S.pv[ first + t ];
    const float2 a = S.rm[ first + t ];
    px[ t ] = s.x, py[ t ] = s.y, vx[ t ] = s.z, vy[ t ] = s.w;
    radius[ t ] = a.x, mass[ t ] = a.y;
    col[ t ] = S.col[ first + t ];
    gid[ t ] = S.gid[ first + t ];
}

// candidate pairs (i, j), i < j by global index, whose i is an OWNED particle: every pair of the whole field is listed by
// exactly one rank.  Two passes (count per slot, then emit at the scanned offset); order inside a rank is by slot.
template <bool EMIT>
__global__ void __launch_bounds__( COLLIDE_THREADS ) k_slab_pairs( const ParticleSet S, const uint32_t *__restrict__ cell_start, const size_t first,
                                                                   const size_t n_own, const GridDims G, uint32_t *__restrict__ count_or_offset,
                                                                   int32_t *__restrict__ pairs, const size_t max_pairs ) {
    const size_t t = (size_t)blockIdx.x * COLLIDE_THREADS + threadIdx.x;
    if ( t >= n_own )
        return;
    const size_t         k  = first + t;
    const Self           me = load_self( S, k );
    const GlobalAccessor acc{ S.pv, S.rm, S.gid, cell_start };
    const uint32_t       key = S.key[ k ];
    const int            cy = (int)( key / (uint32_t)G.cols ), cx = (int)( key - (uint32_t)cy * (uint32_t)G.cols );
    uint32_t             cnt = 0;
    size_t               at  = EMIT ? count_or_offset[ t ] : 0;
    for_each_candidate( acc, G, cx, cy, [ & ]( const uint32_t q ) {
        const Cand c = acc.geom( q );
        if ( c.gid > me.gid && in_contact( me, c ) ) {
            if ( EMIT && at < max_pairs ) {
                pairs[ 2 * at ]     = (int32_t)me.gid;
                pairs[ 2 * at + 1 ] = (int32_t)c.gid;
            }
            at++;
            cnt++;
        }
    } );
    if ( !EMIT )
        count_or_offset[ t ] = cnt;
}

}   // namespace ppc
