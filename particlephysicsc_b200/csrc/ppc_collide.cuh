// ppc_collide.cuh -- narrow phase + contact resolution (K7), and the candidate-pair exporter used by parity tests.
//
// Input is the cell-sorted particle set of this substep (storage order = ascending (cell key, API index)) plus the
// cell-start table.  Each thread owns ONE particle and gathers the contributions of all its contacts, so there are
// no atomics on particle data and no write conflicts: the result is deterministic, run to run and across GPU counts.
//
// Relation to the reference (src/particles_solver.c):
//  * The candidate set is the reference's: 3x3 neighbourhood of the particle's cell (:331-340), the AABB reject
//    and inclusive circle test of :352-362 evaluated with the same float operations.  The test is symmetric in the
//    two particles, so both ends of a pair find each other.
//  * Contact geometry is frozen before anything moves (pass 1, :84-127), exactly as in the reference.
//  * Position pushes (:145-156) depend only on frozen values, so they are pure sums.  Each thread adds its terms in
//    the order the reference's pair list would touch this particle -- first the pairs in which it is j (ascending
//    i), then the pairs in which it is i (neighbour cells dy-major / dx-minor, members in descending index,
//    :311-383) -- which makes positions bit-identical to the reference's sequential loop.
//  * Velocity impulses (:158-190) are order dependent in the reference: every pair reads the velocities left by the
//    pairs before it in the list (Gauss-Seidel).  Two resolve modes:
//      - SEQUENTIAL-EQUIVALENT (default): k_collide_* records, per particle, its contacts in list order; then
//        k_resolve_wavefront fires, round by round, every pair that is at the head of BOTH its particles' lists.
//        Two pairs that share a particle never fire in the same round, and a pair fires only after all earlier pairs
//        of both particles have: the schedule is a topological order of the reference's dependency graph, so every
//        pair sees exactly the velocities it sees in the reference.  Velocities are bit-identical to the reference.
//      - JACOBI: every pair reads the velocities the resolve started with; one pass, no rounds.  Matches the
//        reference statistically while particles have at most a few simultaneous contacts, diverges in dense packings.
//        oracle/ppc_oracle.c:ppc_oracle_resolve_jacobi is the CPU model of exactly this.
#pragma once
#include "ppc_common.cuh"
#include "ppc_math.cuh"

#include <cooperative_groups.h>

namespace ppc {

constexpr int      COLLIDE_THREADS = 128;
constexpr int      COLLIDE_STASH   = 24;    // contacts kept per thread before falling back to the streaming path
constexpr int      INC_MAXDEG      = 16;    // contacts per particle the TILE kernels handle; the stored lists hold WaveBuffers::maxdeg >= 16
constexpr uint32_t INC_LO_FLAG     = 0x80000000u;   // this particle is the pair's i (lower API index)
constexpr uint32_t PC_BITS         = 12;    // wavefront state word: (round << 12) | contacts already resolved
constexpr uint32_t PC_MASK         = ( 1u << PC_BITS ) - 1u;
// Bit 31 of the state word: TAINT.  Slab windows only (GridDims::taint_lo / taint_hi): the particles of the outermost ghost
// row miss their outward contacts, so what the resolve computes for them -- and for every particle that resolves a pair
// with a tainted one -- may differ from the whole-field run.  The bit travels with the pairs; the resolve reports whether
// bit travels with the pairs; afterwards k_slab_taint_scan (ppc_slab.cuh) looks whether it reached an owned particle
// (counters[10]) and how many rows it got into the window (counters[11]).
constexpr uint32_t ST_TAINT        = 0x80000000u;
constexpr uint32_t ST_ROUND_LIMIT  = ( 1u << ( 31 - PC_BITS ) ) - 1u;
__device__ __forceinline__ uint32_t st_round( const uint32_t w ) { return ( w & ~ST_TAINT ) >> PC_BITS; }
__device__ __forceinline__ uint32_t initial_state( const GridDims &G, const uint32_t key ) {
    const int row = (int)( key / (uint32_t)G.cols );
    return ( row == G.taint_lo || row == G.taint_hi ) ? ST_TAINT : 0u;
}
constexpr int      WAVE_THREADS    = 256;

struct Cand {
    float    x, y, r;
    uint32_t gid;
};

// Candidate access straight from global memory (through L1/L2): sorted storage keeps the 3 cell rows of a
// neighbourhood contiguous, so neighbouring threads hit the same lines.
struct GlobalAccessor {
    const float4   *__restrict__ pv;
    const float2   *__restrict__ rm;
    const uint32_t *__restrict__ gid;
    const uint32_t *__restrict__ cell_start;
    __device__ __forceinline__ Cand geom( const uint32_t q ) const {
        const float4 s = pv[ q ];
        Cand         c;
        c.x   = s.x;
        c.y   = s.y;
        c.r   = rm[ q ].x;
        c.gid = gid[ q ];
        return c;
    }
    __device__ __forceinline__ void dyn( const uint32_t q, float &vx, float &vy, float &m ) const {
        const float4 s = pv[ q ];
        vx             = s.z;
        vy             = s.w;
        m              = rm[ q ].y;
    }
    __device__ __forceinline__ float mass( const uint32_t q ) const { return rm[ q ].y; }
    __device__ __forceinline__ void  cell_range( const uint32_t c, uint32_t &b, uint32_t &e ) const {
        b = cell_start[ c ];
        e = cell_start[ c + 1 ];
    }
};

// Visit the candidates of cell (cx, cy) in the reference's scan order (:331-342): dy outer, dx inner, members of a
// cell in DESCENDING API index (= descending sorted slot, storage being ascending).
template <class Acc, class F>
__device__ __forceinline__ void for_each_candidate( const Acc &acc, const GridDims &G, const int cx, const int cy, F &&f ) {
    for ( int oy = -1; oy <= 1; oy++ ) {
        const int ny = cy + oy;
        if ( ny < 0 || ny >= G.rows )
            continue;
        for ( int ox = -1; ox <= 1; ox++ ) {
            const int nx = cx + ox;
            if ( nx < 0 || nx >= G.cols )
                continue;
            uint32_t b, e;
            acc.cell_range( (uint32_t)( ny * G.cols + nx ), b, e );
            for ( uint32_t q = e; q-- > b; )
                f( q );
        }
    }
}

struct Self {
    float    x, y, vx, vy, r, m;
    uint32_t gid;
};

// d = p_lo - p_hi with lo the lower API index (the reference's i), then the test of :352-362.
__device__ __forceinline__ bool in_contact( const Self &me, const Cand &c ) {
    const bool  me_lo = me.gid < c.gid;
    const float dx    = me_lo ? __fsub_rn( me.x, c.x ) : __fsub_rn( c.x, me.x );
    const float dy    = me_lo ? __fsub_rn( me.y, c.y ) : __fsub_rn( c.y, me.y );
    return overlaps( dx, dy, __fadd_rn( me.r, c.r ) );
}

// Call sink(q) for every contact of `me`, in the order the reference's pair list touches this particle:
// pairs (i, me) with i < me by ascending i, then pairs (me, j) in scan order.
template <class Acc, int STASH, class Sink>
__device__ __forceinline__ void for_each_contact_in_list_order( const Acc &acc, const GridDims &G, const Self &me, const uint32_t key,
                                                                Sink &&sink ) {
    const int cy = (int)( key / (uint32_t)G.cols ), cx = (int)( key - (uint32_t)cy * (uint32_t)G.cols );

    // One sweep in scan order.  Contacts in which this particle is i go to the front of the stash in the order met
    // (already the list order); contacts in which it is j go to the back and are sorted by index afterwards.
    uint32_t stash_q[ STASH > 0 ? STASH : 1 ], stash_g[ STASH > 0 ? STASH : 1 ];
    int      n_i = 0, n_j = 0;
    bool     overflow = false;
    for_each_candidate( acc, G, cx, cy, [ & ]( const uint32_t q ) {
        const Cand c = acc.geom( q );
        if ( c.gid == me.gid || !in_contact( me, c ) )
            return;
        if ( n_i + n_j >= STASH ) {
            overflow = true;
            return;
        }
        if ( me.gid < c.gid ) {
            stash_q[ n_i ] = q;
            stash_g[ n_i ] = c.gid;
            n_i++;
        } else {
            n_j++;
            stash_q[ STASH - n_j ] = q;
            stash_g[ STASH - n_j ] = c.gid;
        }
    } );

    if ( !overflow ) {
        for ( int a = 0; a < n_j; a++ ) {   // selection sort of the j-role contacts by ascending index
            const int at   = STASH - n_j + a;
            int       best = at;
            for ( int b = at + 1; b < STASH; b++ )
                if ( stash_g[ b ] < stash_g[ best ] )
                    best = b;
            const uint32_t tq = stash_q[ best ], tg = stash_g[ best ];
            stash_q[ best ] = stash_q[ at ];
            stash_g[ best ] = stash_g[ at ];
            stash_q[ at ]   = tq;
            stash_g[ at ]   = tg;
            sink( tq );
        }
        for ( int a = 0; a < n_i; a++ )
            sink( stash_q[ a ] );
        return;
    }

    // Streaming path for particles with more contacts than the stash holds (e.g. the coincident 10-particle
    // clusters that particlesAddParticle creates, reference src/particles_arrays.c:265-277): select the j-role
    // contacts one at a time in ascending index, then sweep once more for the i-role contacts.
    bool     have_last = false;
    uint32_t last      = 0;
    for ( ;; ) {
        bool     found  = false;
        uint32_t best_q = 0, best_g = 0;
        for_each_candidate( acc, G, cx, cy, [ & ]( const uint32_t q ) {
            const Cand c = acc.geom( q );
            if ( c.gid >= me.gid || ( have_last && c.gid <= last ) || ( found && c.gid >= best_g ) )
                return;
            if ( !in_contact( me, c ) )
                return;
            found  = true;
            best_q = q;
            best_g = c.gid;
        } );
        if ( !found )
            break;
        sink( best_q );
        have_last = true;
        last      = best_g;
    }
    for_each_candidate( acc, G, cx, cy, [ & ]( const uint32_t q ) {
        const Cand c = acc.geom( q );
        if ( c.gid <= me.gid || !in_contact( me, c ) )
            return;
        sink( q );
    } );
}

// The c-th entry (0-based) of a particle's contact list, recomputed from the neighbourhood.  Used by the wavefront
// only for particles with more than INC_MAXDEG contacts, whose list is not stored.  Returns slot | INC_LO_FLAG if
// the particle is the pair's i.  Cost O((c + 2) x candidates).
template <class Acc>
__device__ uint32_t nth_contact( const Acc &acc, const GridDims &G, const Self &me, const uint32_t key, const uint32_t c ) {
    const int cy = (int)( key / (uint32_t)G.cols ), cx = (int)( key - (uint32_t)cy * (uint32_t)G.cols );
    uint32_t  n_j = 0;
    for_each_candidate( acc, G, cx, cy, [ & ]( const uint32_t q ) {
        const Cand k = acc.geom( q );
        if ( k.gid < me.gid && in_contact( me, k ) )
            n_j++;
    } );
    if ( c < n_j ) {
        bool     have_last = false;
        uint32_t last = 0, best_q = 0;
        for ( uint32_t t = 0; t <= c; t++ ) {
            bool     found  = false;
            uint32_t best_g = 0;
            for_each_candidate( acc, G, cx, cy, [ & ]( const uint32_t q ) {
                const Cand k = acc.geom( q );
                if ( k.gid >= me.gid || ( have_last && k.gid <= last ) || ( found && k.gid >= best_g ) )
                    return;
                if ( !in_contact( me, k ) )
                    return;
                found  = true;
                best_q = q;
                best_g = k.gid;
            } );
            have_last = true;
            last      = best_g;
        }
        return best_q;
    }
    uint32_t seen = 0, hit = 0;
    for_each_candidate( acc, G, cx, cy, [ & ]( const uint32_t q ) {
        const Cand k = acc.geom( q );
        if ( k.gid <= me.gid || !in_contact( me, k ) )
            return;
        if ( seen == c - n_j )
            hit = q;
        seen++;
    } );
    return hit | INC_LO_FLAG;
}

// position push of one contact onto `me` (:145-156)
template <class Acc>
__device__ __forceinline__ void apply_push( const Acc &acc, const Self &me, const uint32_t q, const float dt, float4 &out ) {
    const Cand  c  = acc.geom( q );
    const float qm = acc.mass( q );
    if ( me.gid < c.gid ) {   // this particle is the pair's i
        const PushTerms T = push_terms( contact_geom( me.x, me.y, me.r, c.x, c.y, c.r, dt ), inv_mass_u8( me.m ), inv_mass_u8( qm ) );
        if ( T.push ) {
            out.x = __fadd_rn( out.x, T.dpx_lo );
            out.y = __fadd_rn( out.y, T.dpy_lo );
        }
    } else {   // this particle is the pair's j
        const PushTerms T = push_terms( contact_geom( c.x, c.y, c.r, me.x, me.y, me.r, dt ), inv_mass_u8( qm ), inv_mass_u8( me.m ) );
        if ( T.push ) {
            out.x = __fsub_rn( out.x, T.dpx_hi );
            out.y = __fsub_rn( out.y, T.dpy_hi );
        }
    }
}

// Jacobi velocity impulse of one contact onto `me`, from the velocities the resolve started with
template <class Acc>
__device__ __forceinline__ void apply_impulse_frozen( const Acc &acc, const Self &me, const uint32_t q, const float dt, float4 &out ) {
    const Cand c = acc.geom( q );
    float      qvx, qvy, qm;
    acc.dyn( q, qvx, qvy, qm );
    if ( me.gid < c.gid ) {
        const ImpulseTerms T = impulse_terms( contact_geom( me.x, me.y, me.r, c.x, c.y, c.r, dt ), inv_mass_u8( me.m ), inv_mass_u8( qm ),
                                              me.vx, me.vy, qvx, qvy );
        if ( T.impulse ) {
            out.z = __fadd_rn( out.z, T.dvx_lo );
            out.w = __fadd_rn( out.w, T.dvy_lo );
        }
    } else {
        const ImpulseTerms T = impulse_terms( contact_geom( c.x, c.y, c.r, me.x, me.y, me.r, dt ), inv_mass_u8( qm ), inv_mass_u8( me.m ),
                                              qvx, qvy, me.vx, me.vy );
        if ( T.impulse ) {
            out.z = __fsub_rn( out.z, T.dvx_hi );
            out.w = __fsub_rn( out.w, T.dvy_hi );
        }
    }
}

__device__ __forceinline__ Self load_self( const ParticleSet &S, const size_t k ) {
    const float4 s = S.pv[ k ];
    const float2 a = S.rm[ k ];
    Self         me;
    me.x   = s.x;
    me.y   = s.y;
    me.vx  = s.z;
    me.vy  = s.w;
    me.r   = a.x;
    me.m   = a.y;
    me.gid = S.gid[ k ];
    return me;
}

// Per-particle bookkeeping of the sequential-equivalent resolve.
struct WaveBuffers {
    uint16_t *deg;       // contacts of each particle (sorted slot)
    uint32_t *state;     // (round of last advance << PC_BITS) | contacts already resolved
    uint32_t *inc;       // inc[c * stride + k] = c-th contact of particle k: partner slot | INC_LO_FLAG
    size_t    stride;    // particle capacity
    uint32_t  maxdeg;    // list positions stored per particle (>= INC_MAXDEG); longer lists are recomputed on demand
    uint32_t *list[ 2 ]; // active-particle queues (ping-pong between rounds); CTA b of the wavefront owns [b*chunk, (b+1)*chunk)
    uint32_t *counters;  // [0..2] rotating totals of still-active particles, [3] error flags, [4] rounds of the last resolve,
                         // [5] tiles with work left after the tile passes, [6..7] particle visits of the last resolve (64-bit),
                         // [8] particles with at least one contact in this substep, [9] densest tile in particles per 1024 cells (from k_collide_tile),
                         // [10] slab mode: the taint reached an owned particle, [11] rows the taint travelled into the window,
                         // [12] longest contact list that did not fit the stored lists (0: all fit)
};

// ---- K7, global-memory form: one thread per sorted particle --------------------------------------------------
// JACOBI = true : positions and velocities in one pass (velocities from frozen values).
// JACOBI = false: positions only; the contact list and the contact count go to the wavefront buffers.
// One particle, generic path: candidates straight from global memory, one thread does everything.
template <int STASH, bool JACOBI>
// (Arguments by value: this function is not inlined, and references would force the callers' copies of the kernel parameters
// into local memory for the whole kernel, not just around the rare calls.)
__device__ __noinline__ void collide_one_generic( const ParticleSet S, float4 *__restrict__ pv_out, const uint32_t *__restrict__ cell_start,
                                                     const GridDims G, const float dt, const WaveBuffers WB, const size_t k ) {
    uint32_t             cnt = 0;
    const Self           me  = load_self( S, k );
    const GlobalAccessor acc{ S.pv, S.rm, S.gid, cell_start };
    float4               out = make_float4( me.x, me.y, me.vx, me.vy );
    for_each_contact_in_list_order<GlobalAccessor, STASH>( acc, G, me, S.key[ k ], [ & ]( const uint32_t q ) {
        apply_push( acc, me, q, dt, out );
        if ( JACOBI ) {
            apply_impulse_frozen( acc, me, q, dt, out );
        } else {
            if ( cnt < WB.maxdeg )
                WB.inc[ (size_t)cnt * WB.stride + k ] = q | ( me.gid < acc.gid[ q ] ? INC_LO_FLAG : 0u );
            cnt++;
        }
    } );
    pv_out[ k ] = out;
    if ( !JACOBI ) {
        if ( cnt > PC_MASK ) {   // more simultaneous contacts than the state word can count
            atomicOr( &WB.counters[ 3 ], 1u );
            cnt = PC_MASK;
        }
        if ( cnt > WB.maxdeg )   // list longer than what is stored: tell the host, which lengthens the lists for the next substeps
            atomicMax( &WB.counters[ 12 ], cnt );
        WB.deg[ k ]   = (uint16_t)cnt;
        WB.state[ k ] = initial_state( G, S.key[ k ] );
    }
}

template <int STASH, bool JACOBI>
__global__ void __launch_bounds__( COLLIDE_THREADS ) k_collide_global( const ParticleSet S, float4 *__restrict__ pv_out,
                                                                       const uint32_t *__restrict__ cell_start, const size_t n,
                                                                       const GridDims G, const float dt, const WaveBuffers WB ) {
    const size_t k = (size_t)blockIdx.x * COLLIDE_THREADS + threadIdx.x;
    if ( k < n )
        collide_one_generic<STASH, JACOBI>( S, pv_out, cell_start, G, dt, WB, k );
}

// ---- K7, shared-memory tile form (sequential-equivalent mode) --------------------------------------------------
// One CTA per tile of TILE_R cell rows x wc cell columns.  Because storage is sorted row-major by cell, the tile and
// its one-cell halo are TILE_R + 2 contiguous slot runs: they are staged into shared memory with coalesced loads as
// one 16-byte record per particle (x, y, radius, API index) plus the mass, each particle being read by ~1.5 CTAs
// instead of by the 9 cells around it.  Each thread then owns one centre particle and scans the three contiguous
// shared-memory ranges of its 3x3 neighbourhood with a lean loop: one LDS.128, the reference's AABB reject on x (:353),
// on y (:356), then the circle test (:360-362).  The test is symmetric, so it is evaluated as (me - candidate) without
// looking at indices.  The few hits are then put into the reference's list order (pairs in which the particle is j
// by ascending i, then pairs in which it is i by scan order) and their push terms added in that order, so positions
// equal the reference's bit for bit; the ordered contact list goes to the wavefront buffers.
// Tiles or particles that do not fit the fixed budget take the generic path.
#ifndef PPC_TILE_MINBLOCKS
#define PPC_TILE_MINBLOCKS 4   // 32 registers, four CTAs per SM: measured 26% faster than two CTAs at 60 registers (uniform16m)
#endif
#ifndef PPC_TILE_THREADS
#define PPC_TILE_THREADS 512
#endif
#ifndef PPC_TILE_CAP
#define PPC_TILE_CAP 2304
#endif
#ifndef PPC_TILE_TARGET
#define PPC_TILE_TARGET 900    // centre particles per tile the host aims for
#endif
#ifndef PPC_TILE_UNROLL
#define PPC_TILE_UNROLL 2
#endif
#define PPC_PRAGMA( x ) _Pragma( #x )
#define PPC_UNROLL( n ) PPC_PRAGMA( unroll n )
constexpr int TILE_THREADS = PPC_TILE_THREADS;
constexpr int TILE_R       = 8;      // centre rows per tile
constexpr int TILE_MAX_WC  = 64;     // centre columns per tile (run-time value <= this)
constexpr int TILE_CAP     = PPC_TILE_CAP;   // staged particles (centre + halo)
constexpr int TILE_TARGET  = PPC_TILE_TARGET;
constexpr int TILE_STASH   = 16;     // hits per particle kept by the lean path
constexpr int TILE_BW      = TILE_MAX_WC + 4;   // boundary-table row stride

struct TileSmem {
    float4   cand[ TILE_CAP ];                    // x, y, radius, API index (bits)
    float    mass[ TILE_CAP ];   // INVERSE mass 1 / (float)(uint8_t)m
    uint8_t  ccol[ TILE_CAP ];   // cell column of the particle, relative to the first staged column
    uint32_t bound[ ( TILE_R + 2 ) * TILE_BW ];   // bound[j][i] = first staged index of cell c0 + i in staged row j
    uint32_t run_g0[ TILE_R + 2 ], run_s0[ TILE_R + 3 ], cen_s0[ TILE_R + 1 ];
    uint32_t fallback;
};

// Staged particle count of the tile with centre columns [cx0, cx1) and centre rows [cy0, cy0 + TILE_R): the sum of
// its TILE_R + 2 slot runs.  Called by all threads; the result is uniform.
__device__ __forceinline__ uint32_t tile_staged_count( TileSmem &T, const uint32_t *__restrict__ cell_start, const GridDims &G, const int cx0,
                                                       const int cx1, const int cy0 ) {
    const int c0 = max( cx0 - 1, 0 ), c1 = min( cx1 + 1, G.cols );
    __syncthreads();
    if ( threadIdx.x < TILE_R + 2 ) {
        const int gy = cy0 - 1 + (int)threadIdx.x;
        uint32_t  len = 0;
        if ( gy >= 0 && gy < G.rows )
            len = cell_start[ (size_t)gy * G.cols + c1 ] - cell_start[ (size_t)gy * G.cols + c0 ];
        T.run_s0[ threadIdx.x ] = len;
    }
    __syncthreads();
    uint32_t total = 0;
#pragma unroll
    for ( int j = 0; j < TILE_R + 2; j++ )
        total += T.run_s0[ j ];
    __syncthreads();
    return total;
}

// One (sub-)tile: centre columns [cx0, cx1), centre rows [cy0, cy0 + TILE_R).
__device__ __forceinline__ void collide_tile_body( TileSmem &T, const ParticleSet &S, float4 *__restrict__ pv_out,
                                                   const uint32_t *__restrict__ cell_start, const GridDims &G, const float dt,
                                                   const WaveBuffers &WB, const int cx0, const int cx1, const int cy0, uint32_t &n_active ) {
    const int c0 = max( cx0 - 1, 0 ), c1 = min( cx1 + 1, G.cols );      // staged columns [c0, c1)
    const int nb = c1 - c0 + 1;                                         // boundaries per staged row
    __syncthreads();                                                    // shared memory may still be in use by the previous sub-tile

    // ---- run table: staged row j = 0 .. TILE_R+1 is grid row cy0 - 1 + j ----
    if ( threadIdx.x < TILE_R + 2 ) {
        const int j = threadIdx.x, gy = cy0 - 1 + j;
        uint32_t  g0 = 0, g1 = 0;
        if ( gy >= 0 && gy < G.rows ) {
            g0 = cell_start[ (size_t)gy * G.cols + c0 ];
            g1 = cell_start[ (size_t)gy * G.cols + c1 ];
        }
        T.run_g0[ j ] = g0;
        T.run_s0[ j ] = g1 - g0;   // length for now
    }
    __syncthreads();
    if ( threadIdx.x == 0 ) {
        uint32_t run = 0;
        for ( int j = 0; j < TILE_R + 2; j++ ) {
            const uint32_t len = T.run_s0[ j ];
            T.run_s0[ j ]      = run;
            run += len;
        }
        T.run_s0[ TILE_R + 2 ] = run;
        T.fallback             = run > (uint32_t)TILE_CAP ? 1u : 0u;
    }
    __syncthreads();
    if ( T.run_s0[ TILE_R + 2 ] == 0 )
        return;
    // ---- boundary table ----
    for ( int e = threadIdx.x; e < ( TILE_R + 2 ) * nb; e += TILE_THREADS ) {
        const int j = e / nb, i = e - j * nb, gy = cy0 - 1 + j;
        uint32_t  v = T.run_s0[ j ];
        if ( gy >= 0 && gy < G.rows )
            v += cell_start[ (size_t)gy * G.cols + c0 + i ] - T.run_g0[ j ];
        T.bound[ j * TILE_BW + i ] = v;
    }
    __syncthreads();
    const int cb = cx0 - c0, ce = cx1 - c0;   // centre columns within the boundary table: [cb, ce)
    if ( threadIdx.x == 0 ) {
        uint32_t run = 0;
        for ( int j = 1; j <= TILE_R; j++ ) {
            T.cen_s0[ j - 1 ] = run;
            run += T.bound[ j * TILE_BW + ce ] - T.bound[ j * TILE_BW + cb ];
        }
        T.cen_s0[ TILE_R ] = run;
    }
    __syncthreads();
    const uint32_t centre = T.cen_s0[ TILE_R ];
    if ( centre == 0 )
        return;

    if ( T.fallback ) {   // tile does not fit shared memory: thread-per-particle from global memory
        for ( int j = 1; j <= TILE_R; j++ ) {
            const uint32_t s_b = T.bound[ j * TILE_BW + cb ], s_e = T.bound[ j * TILE_BW + ce ];
            const uint32_t g_b = T.run_g0[ j ] + ( s_b - T.run_s0[ j ] );
            for ( uint32_t t = threadIdx.x; t < s_e - s_b; t += TILE_THREADS )
                collide_one_generic<COLLIDE_STASH, false>( S, pv_out, cell_start, G, dt, WB, (size_t)g_b + t );
        }
        return;
    }

    // ---- stage the TILE_R + 2 runs (flat index: every thread busy, one exposed memory latency) ----
    for ( uint32_t t = threadIdx.x; t < T.run_s0[ TILE_R + 2 ]; t += TILE_THREADS ) {
        int j = 0;
        while ( j < TILE_R + 1 && t >= T.run_s0[ j + 1 ] )
            j++;
        const uint32_t g = T.run_g0[ j ] + ( t - T.run_s0[ j ] );
        const float4   p = S.pv[ g ];
        const float2   a = S.rm[ g ];
        T.cand[ t ]      = make_float4( p.x, p.y, a.x, __uint_as_float( S.gid[ g ] ) );
        T.mass[ t ]      = inv_mass_u8( a.y );   // inverse mass, as the resolve reads it (:141-147)
        int clo = 0, chi = nb - 1;               // column i with bound[j][i] <= t < bound[j][i+1]
        while ( chi - clo > 1 ) {
            const int mid = ( clo + chi ) >> 1;
            if ( t >= T.bound[ j * TILE_BW + mid ] )
                clo = mid;
            else
                chi = mid;
        }
        T.ccol[ t ] = (uint8_t)clo;
    }
    __syncthreads();

    // ---- one thread per centre particle ----
    for ( uint32_t t = threadIdx.x; t < centre; t += TILE_THREADS ) {
        int j = 1;
        while ( j < TILE_R && t >= T.cen_s0[ j ] )
            j++;
        const uint32_t self = T.bound[ j * TILE_BW + cb ] + ( t - T.cen_s0[ j - 1 ] );   // staged index
        const uint32_t gk   = T.run_g0[ j ] + ( self - T.run_s0[ j ] );                   // storage slot
        const float4   me4  = T.cand[ self ];
        const float    x = me4.x, y = me4.y, r = me4.z;
        const uint32_t gid = __float_as_uint( me4.w );
        const uint32_t key = S.key[ gk ];
        const int      gy = (int)( key / (uint32_t)G.cols ), gx = (int)( key - (uint32_t)gy * (uint32_t)G.cols );
        const int      lo = max( gx - 1, 0 ), hi = min( gx + 1, G.cols - 1 );

        // hits: staged index | staged row << 16, and a 64-bit key that sorts them into the reference's list order:
        //   pairs (i, me), i < me : key = i                                   (ascending i, all before the pairs (me, *))
        //   pairs (me, j), j > me : key = 2^48 | scan cell << 32 | ~j         (scan order, descending index inside a cell)
        uint32_t           stash[ TILE_STASH ];
        unsigned long long skey[ TILE_STASH ];
        uint32_t           nh = 0;
#pragma unroll 1
        for ( int d = 0; d < 3; d++ ) {
            const int ry = gy - 1 + d;
            if ( ry < 0 || ry >= G.rows )
                continue;
            const uint32_t *B  = &T.bound[ ( j - 1 + d ) * TILE_BW ];
            const uint32_t  rb = B[ lo - c0 ], re = B[ hi + 1 - c0 ];
            PPC_UNROLL( PPC_TILE_UNROLL )
            for ( uint32_t s = rb; s < re; s++ ) {
                const float4 c  = T.cand[ s ];
                const float  rs = __fadd_rn( r, c.z );
                const float  dx = __fsub_rn( x, c.x );
                if ( fabsf( dx ) > rs )   // :353
                    continue;
                const float dy = __fsub_rn( y, c.y );
                if ( fabsf( dy ) > rs )   // :356
                    continue;
                const float d2 = __fadd_rn( __fmul_rn( dx, dx ), __fmul_rn( dy, dy ) );
                if ( d2 <= __fmul_rn( rs, rs ) && s != self ) {   // :360-362
                    if ( nh < (uint32_t)TILE_STASH ) {
                        const uint32_t g = __float_as_uint( c.w );
                        stash[ nh ]      = s | ( (uint32_t)d << 16 );
                        skey[ nh ]       = g < gid ? (unsigned long long)g
                                                   : ( 1ull << 48 ) | ( (unsigned long long)( (uint32_t)d * 128u + T.ccol[ s ] ) << 32 ) | ( 0xffffffffu - g );
                    }
                    nh++;
                }
            }
        }
        if ( nh == 0 ) {
            const float4 old = S.pv[ gk ];
            pv_out[ gk ]     = make_float4( x, y, old.z, old.w );
            WB.deg[ gk ]     = 0;
            WB.state[ gk ]   = initial_state( G, key );
            continue;
        }
        n_active++;
        if ( nh > (uint32_t)TILE_STASH ) {   // crowded particle: generic path
            collide_one_generic<COLLIDE_STASH, false>( S, pv_out, cell_start, G, dt, WB, (size_t)gk );
            continue;
        }
        // list order: pairs (i, me) by ascending i, then pairs (me, j) by scan cell, descending index inside a cell
        float       X = x, Y = y;
        const float wme = T.mass[ self ];
#pragma unroll 1
        for ( uint32_t a = 0; a < nh; a++ ) {
            uint32_t           best = a;
            unsigned long long bkey = skey[ a ];
#pragma unroll 1
            for ( uint32_t b = a + 1; b < nh; b++ )
                if ( skey[ b ] < bkey ) {
                    bkey = skey[ b ];
                    best = b;
                }
            const uint32_t e = stash[ best ];
            stash[ best ]    = stash[ a ];
            skey[ best ]     = skey[ a ];
            stash[ a ]       = e;
            skey[ a ]        = bkey;
            const uint32_t s     = e & 0xffffu;
            const float4   c     = T.cand[ s ];
            const bool     me_lo = gid < __float_as_uint( c.w );
            const float    wq    = T.mass[ s ];
            if ( me_lo ) {
                const PushTerms P = push_terms( contact_geom( x, y, r, c.x, c.y, c.z, dt ), wme, wq );
                if ( P.push ) {
                    X = __fadd_rn( X, P.dpx_lo );
                    Y = __fadd_rn( Y, P.dpy_lo );
                }
            } else {
                const PushTerms P = push_terms( contact_geom( c.x, c.y, c.z, x, y, r, dt ), wq, wme );
                if ( P.push ) {
                    X = __fsub_rn( X, P.dpx_hi );
                    Y = __fsub_rn( Y, P.dpy_hi );
                }
            }
            const int jj = j - 1 + (int)( e >> 16 );   // staged row of the partner -> its storage slot
            WB.inc[ (size_t)a * WB.stride + gk ] = ( T.run_g0[ jj ] + ( s - T.run_s0[ jj ] ) ) | ( me_lo ? INC_LO_FLAG : 0u );
        }
        const float4 old = S.pv[ gk ];
        pv_out[ gk ]     = make_float4( X, Y, old.z, old.w );
        WB.deg[ gk ]     = (uint16_t)nh;
        WB.state[ gk ]   = initial_state( G, key );
    }
}

__global__ void __launch_bounds__( TILE_THREADS, PPC_TILE_MINBLOCKS ) k_collide_tile( const ParticleSet S, float4 *__restrict__ pv_out,
                                                                  const uint32_t *__restrict__ cell_start, const GridDims G, const float dt,
                                                                  const WaveBuffers WB, const int wc, const int tiles_x ) {
    extern __shared__ __align__( 16 ) unsigned char smem_raw[];
    TileSmem &T = *reinterpret_cast<TileSmem *>( smem_raw );
    const int tx = (int)( blockIdx.x % (unsigned)tiles_x ), ty = (int)( blockIdx.x / (unsigned)tiles_x );
    const int cx0 = tx * wc, cy0 = ty * TILE_R;
    const int cx1 = min( cx0 + wc, G.cols );
    // Particle density varies across the box (a polydisperse pile holds ten times more small particles per cell than
    // large ones): a tile that would overflow shared memory is cut into column strips that fit.
    const uint32_t total = tile_staged_count( T, cell_start, G, cx0, cx1, cy0 );
    if ( total == 0 )
        return;
    if ( threadIdx.x == 0 )   // densest tile, in particles per 1024 cells: sizes the wavefront tiles (a substep late)
        atomicMax( &WB.counters[ 9 ], (uint32_t)( ( (unsigned long long)total << 10 ) / (unsigned)( ( TILE_R + 2 ) * ( cx1 - cx0 + 2 ) ) ) );
    const int strips = (int)( ( total + ( TILE_CAP * 6 / 10 ) - 1 ) / ( TILE_CAP * 6 / 10 ) );
    const int sw     = max( ( cx1 - cx0 + strips - 1 ) / strips, 1 );
    uint32_t  n_active = 0;
    for ( int sx = cx0; sx < cx1; sx += sw )
        collide_tile_body( T, S, pv_out, cell_start, G, dt, WB, sx, min( sx + sw, cx1 ), cy0, n_active );
    // how many particles have contacts: the host uses it (a substep late) to decide whether tile passes pay off
#pragma unroll
    for ( int d = 16; d > 0; d >>= 1 )
        n_active += __shfl_xor_sync( 0xffffffffu, n_active, d );
    if ( ( threadIdx.x & 31u ) == 0 && n_active )
        atomicAdd( &WB.counters[ 8 ], n_active );
}

// ---- K7, shared-memory tile form with fine x-bins (default) ------------------------------------------------------------
// Same tile, same staging, same per-particle ordering and arithmetic as k_collide_tile above; what changes is WHICH staged
// particles a centre particle looks at.  The reference's 3x3 cell neighbourhood (cellSize = 2 maxRadius >= 12 px) holds
// ~30 candidates for r = 2 at area fraction 0.30 and ~165 for r = 1 at 0.40, of which 1-3 touch.  Contacts are limited by
// |dx| <= r_me + r_other and |dy| <= r_me + r_other (:352-357), so with R = the largest radius staged for this tile:
//   * each staged row (one cell row of the tile + halo) is counting-sorted in shared memory into fine x-bins of a quarter
//     cell; a particle scans only the bins overlapping [x - r - R, x + r + R] (a contiguous range of the row's index);
//   * the cell row above / below is scanned only when y - r - R / y + r + R reaches into it.
// Both filters are conservative (monotone float arithmetic plus a 0.02 px margin that covers every rounding step at
// coordinates below 65,536); the reference's own AABB + circle test then decides, so the pair set is unchanged -- with
// one extra check that the two cell columns differ by at most one, because the reference never pairs particles two
// columns apart even in the one-ulp corner where its float division puts a touching pair there.
#ifndef PPC_TILE2_CAP
#define PPC_TILE2_CAP 2048
#endif
#ifndef PPC_TILE2_TARGET
#define PPC_TILE2_TARGET 800
#endif
#ifndef PPC_TILE2_F
#define PPC_TILE2_F 4
#endif
constexpr int TILE2_CAP  = PPC_TILE2_CAP;         // staged particles
constexpr int TILE2_F    = PPC_TILE2_F;           // fine bins per cell
constexpr int TILE2_MAXC = 64;                    // staged columns (centre + halo)
constexpr int TILE2_NB   = TILE2_MAXC * TILE2_F;  // fine bins per staged row
constexpr int TILE2_BINS = ( TILE_R + 2 ) * TILE2_NB;
constexpr int TILE2_TARGET = PPC_TILE2_TARGET;

struct Tile2Smem {
    float4   cand[ TILE2_CAP ];    // x, y, radius, API index (bits), in storage order
    uint32_t fend[ TILE2_BINS ];   // after the scatter: end (exclusive) of fine bin f = row * TILE2_NB + bin in perm; its begin is fend[f - 1]
    uint16_t perm[ TILE2_CAP ];    // staged indices sorted by (staged row, fine bin)
    uint8_t  mass[ TILE2_CAP ];    // mass as the resolve reads it: truncated to uint8_t (:141-142)
    uint8_t  ccol[ TILE2_CAP ];    // cell column relative to the first staged column
    float    inv_lut[ 256 ];       // 1 / (float)k, the inverse masses of :146-147
    uint32_t run_g0[ TILE_R + 2 ], run_s0[ TILE_R + 3 ], cen_g0[ TILE_R + 1 ], cen_s0[ TILE_R + 1 ];
    uint32_t wsum[ 16 ];
    uint32_t rmax_bits, fallback;
};

__device__ __forceinline__ int fine_bin( const float x, const float x0, const float inv_w, const int nb ) {
    const float q = __fmul_rn( __fsub_rn( x, x0 ), inv_w );
    int         b = ( q >= 0.0f ) ? ( q < 16777216.0f ? __float2int_rz( q ) : nb ) : 0;   // NaN -> 0
    return b < nb ? b : nb - 1;
}

__device__ __forceinline__ void collide_tile2_body( Tile2Smem &T, const ParticleSet &S, float4 *__restrict__ pv_out,
                                                    const uint32_t *__restrict__ cell_start, const GridDims &G, const float dt,
                                                    const WaveBuffers &WB, const int cx0, const int cx1, const int cy0, uint32_t &n_active ) {
    const int c0 = max( cx0 - 1, 0 ), c1 = min( cx1 + 1, G.cols );      // staged columns [c0, c1)
    const int nb = ( c1 - c0 ) * TILE2_F;                               // fine bins per staged row
    __syncthreads();                                                    // shared memory may still be in use by the previous sub-tile

    // ---- run table: staged row j = 0 .. TILE_R+1 is grid row cy0 - 1 + j ----
    if ( threadIdx.x < TILE_R + 2 ) {
        const int j = threadIdx.x, gy = cy0 - 1 + j;
        uint32_t  g0 = 0, g1 = 0, m0 = 0, m1 = 0;
        if ( gy >= 0 && gy < G.rows ) {
            const size_t row = (size_t)gy * G.cols;
            g0 = cell_start[ row + c0 ], g1 = cell_start[ row + c1 ];
            m0 = cell_start[ row + cx0 ], m1 = cell_start[ row + cx1 ];
        }
        T.run_g0[ j ] = g0;
        T.run_s0[ j ] = g1 - g0;   // length for now
        if ( j >= 1 && j <= TILE_R ) {
            T.cen_g0[ j - 1 ] = m0;
            T.cen_s0[ j - 1 ] = m1 - m0;   // length for now
        }
    }
    for ( int f = threadIdx.x; f < ( TILE_R + 2 ) * TILE2_NB; f += TILE_THREADS )
        T.fend[ f ] = 0;
    __syncthreads();
    if ( threadIdx.x == 0 ) {
        uint32_t run = 0;
        for ( int j = 0; j < TILE_R + 2; j++ ) {
            const uint32_t len = T.run_s0[ j ];
            T.run_s0[ j ]      = run;
            run += len;
        }
        T.run_s0[ TILE_R + 2 ] = run;
        T.fallback             = run > (uint32_t)TILE2_CAP ? 1u : 0u;
        uint32_t cen = 0;
        for ( int j = 0; j < TILE_R; j++ ) {
            const uint32_t len = T.cen_s0[ j ];
            T.cen_s0[ j ]      = cen;
            cen += len;
        }
        T.cen_s0[ TILE_R ] = cen;
        T.rmax_bits        = 0u;
    }
    __syncthreads();
    const uint32_t staged = T.run_s0[ TILE_R + 2 ], centre = T.cen_s0[ TILE_R ];
    if ( staged == 0 || centre == 0 )
        return;

    if ( T.fallback ) {   // tile does not fit shared memory: thread-per-particle from global memory
        for ( int j = 0; j < TILE_R; j++ ) {
            const uint32_t len = T.cen_s0[ j + 1 ] - T.cen_s0[ j ];
            for ( uint32_t t = threadIdx.x; t < len; t += TILE_THREADS )
                collide_one_generic<COLLIDE_STASH, false>( S, pv_out, cell_start, G, dt, WB, (size_t)T.cen_g0[ j ] + t );
        }
        return;
    }

    // ---- stage the TILE_R + 2 runs; histogram of (row, fine bin); largest radius ----
    const float x0 = __fmul_rn( (float)c0, G.cell_size ), inv_w = __fdiv_rn( (float)TILE2_F, G.cell_size );
    float       rmax = 0.0f;
    for ( uint32_t t = threadIdx.x; t < staged; t += TILE_THREADS ) {
        int j = 0;
        while ( j < TILE_R + 1 && t >= T.run_s0[ j + 1 ] )
            j++;
        const uint32_t g = T.run_g0[ j ] + ( t - T.run_s0[ j ] );
        const float4   p = S.pv[ g ];
        const float2   a = S.rm[ g ];
        T.cand[ t ]      = make_float4( p.x, p.y, a.x, __uint_as_float( S.gid[ g ] ) );
        T.mass[ t ]      = (uint8_t)( (unsigned)__float2int_rz( a.y ) & 0xffu );
        T.ccol[ t ]      = (uint8_t)( S.key[ g ] - (uint32_t)( cy0 - 1 + j ) * (uint32_t)G.cols - (uint32_t)c0 );
        rmax             = fmaxf( rmax, a.x );
        atomicAdd( &T.fend[ j * TILE2_NB + fine_bin( p.x, x0, inv_w, nb ) ], 1u );
    }
    if ( threadIdx.x < 256 )
        T.inv_lut[ threadIdx.x ] = __fdiv_rn( 1.0f, (float)threadIdx.x );
#pragma unroll
    for ( int d = 16; d > 0; d >>= 1 )
        rmax = fmaxf( rmax, __shfl_xor_sync( 0xffffffffu, rmax, d ) );
    if ( ( threadIdx.x & 31u ) == 0 && rmax > 0.0f )
        atomicMax( &T.rmax_bits, __float_as_uint( rmax ) );   // positive floats order like their bit patterns
    __syncthreads();

    // ---- exclusive scan of the bin counts (each thread a contiguous chunk; warp scan; block scan) ----
    {
        constexpr int PER = ( TILE2_BINS + TILE_THREADS - 1 ) / TILE_THREADS;
        const int     f0  = threadIdx.x * PER;
        uint32_t      loc[ PER ], sum = 0;
#pragma unroll
        for ( int i = 0; i < PER; i++ ) {
            loc[ i ] = ( f0 + i < TILE2_BINS ) ? T.fend[ f0 + i ] : 0u;
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
        uint32_t ws = ( threadIdx.x & 31u ) < (unsigned)( TILE_THREADS / 32 ) ? T.wsum[ threadIdx.x & 31u ] : 0u;   // every warp scans the warp sums
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
            if ( f0 + i < TILE2_BINS )
                T.fend[ f0 + i ] = run;   // begin of the bin; the scatter below advances it to the end
            run += loc[ i ];
        }
    }
    __syncthreads();
    for ( uint32_t t = threadIdx.x; t < staged; t += TILE_THREADS ) {
        int j = 0;
        while ( j < TILE_R + 1 && t >= T.run_s0[ j + 1 ] )
            j++;
        T.perm[ atomicAdd( &T.fend[ j * TILE2_NB + fine_bin( T.cand[ t ].x, x0, inv_w, nb ) ], 1u ) ] = (uint16_t)t;
    }
    __syncthreads();
    const float R = __uint_as_float( T.rmax_bits );

    // ---- one thread per centre particle ----
    for ( uint32_t t = threadIdx.x; t < centre; t += TILE_THREADS ) {
        int j = 1;
        while ( j < TILE_R && t >= T.cen_s0[ j ] )
            j++;
        const uint32_t gk   = T.cen_g0[ j - 1 ] + ( t - T.cen_s0[ j - 1 ] );   // storage slot
        const uint32_t self = T.run_s0[ j ] + ( gk - T.run_g0[ j ] );          // staged index
        const float4   me4  = T.cand[ self ];
        const float    x = me4.x, y = me4.y, r = me4.z;
        const uint32_t gid  = __float_as_uint( me4.w );
        const int      mycol = T.ccol[ self ];
        const int      gy    = cy0 - 1 + j;
        const float    reach = __fadd_rn( __fadd_rn( r, R ), 0.02f );
        const int      b_lo  = fine_bin( __fsub_rn( x, reach ), x0, inv_w, nb ), b_hi = fine_bin( __fadd_rn( x, reach ), x0, inv_w, nb );
        const float    top   = __fmul_rn( (float)( gy + G.row0 ), G.cell_size ), bot = __fmul_rn( (float)( gy + G.row0 + 1 ), G.cell_size );
        const uint32_t st0   = ( gy == G.taint_lo || gy == G.taint_hi ) ? ST_TAINT : 0u;   // outermost ghost row of a slab window

        uint32_t stash[ TILE_STASH ];
        uint32_t nh = 0;
#pragma unroll 1
        for ( int d = 0; d < 3; d++ ) {
            const int ry = gy - 1 + d;
            if ( ry < 0 || ry >= G.rows )
                continue;
            if ( d == 0 && __fsub_rn( y, reach ) >= top )   // nothing in the row above can be within reach
                continue;
            if ( d == 2 && __fadd_rn( y, reach ) < bot )
                continue;
            const int      f_lo = ( j - 1 + d ) * TILE2_NB + b_lo, f_hi = ( j - 1 + d ) * TILE2_NB + b_hi;
            const uint32_t rb = f_lo > 0 ? T.fend[ f_lo - 1 ] : 0u, re = T.fend[ f_hi ];
            PPC_UNROLL( PPC_TILE_UNROLL )
            for ( uint32_t q = rb; q < re; q++ ) {
                const uint32_t s  = T.perm[ q ];
                const float4   c  = T.cand[ s ];
                const float    rs = __fadd_rn( r, c.z );
                const float    dx = __fsub_rn( x, c.x );
                if ( fabsf( dx ) > rs )   // :353
                    continue;
                const float dy = __fsub_rn( y, c.y );
                if ( fabsf( dy ) > rs )   // :356
                    continue;
                const float d2 = __fadd_rn( __fmul_rn( dx, dx ), __fmul_rn( dy, dy ) );
                if ( d2 <= __fmul_rn( rs, rs ) && s != self ) {   // :360-362
                    const int dc = (int)T.ccol[ s ] - mycol;
                    if ( dc < -1 || dc > 1 )   // two columns apart: outside the reference's 3x3 scan (:331-340)
                        continue;
                    if ( nh < (uint32_t)TILE_STASH )
                        stash[ nh ] = s | ( (uint32_t)d << 16 );
                    nh++;
                }
            }
        }
        if ( nh == 0 ) {
            const float4 old = S.pv[ gk ];
            pv_out[ gk ]     = make_float4( x, y, old.z, old.w );
            WB.deg[ gk ]     = 0;
            WB.state[ gk ]   = st0;
            continue;
        }
        n_active++;
        if ( nh > (uint32_t)TILE_STASH ) {   // crowded particle: generic path
            collide_one_generic<COLLIDE_STASH, false>( S, pv_out, cell_start, G, dt, WB, (size_t)gk );
            continue;
        }
        // list order: pairs (i, me) by ascending i, then pairs (me, j) by scan cell, descending index inside a cell.
        //   pairs (i, me), i < me : key = i
        //   pairs (me, j), j > me : key = 2^48 | scan cell << 32 | ~j
        // The keys are recomputed from the staged data when two hits are compared (a particle has a handful of hits), rather than
        // kept in a per-thread array: thread-local arrays live in local memory, which is most of this kernel's memory traffic.
        auto sort_key = [ & ]( const uint32_t e ) -> unsigned long long {
            const uint32_t s = e & 0xffffu, g = __float_as_uint( T.cand[ s ].w );
            return g < gid ? (unsigned long long)g
                           : ( 1ull << 48 ) | ( (unsigned long long)( ( e >> 16 ) * 128u + T.ccol[ s ] ) << 32 ) | ( 0xffffffffu - g );
        };
        float       X = x, Y = y;
        const float wme = T.inv_lut[ T.mass[ self ] ];
#pragma unroll 1
        for ( uint32_t a = 0; a < nh; a++ ) {
            uint32_t           best = a, e = stash[ a ];
            unsigned long long bkey = sort_key( e );
#pragma unroll 1
            for ( uint32_t b = a + 1; b < nh; b++ ) {
                const uint32_t           eb = stash[ b ];
                const unsigned long long kb = sort_key( eb );
                if ( kb < bkey ) {
                    bkey = kb;
                    best = b;
                    e    = eb;
                }
            }
            if ( best != a ) {
                stash[ best ] = stash[ a ];
                stash[ a ]    = e;
            }
            const uint32_t s     = e & 0xffffu;
            const float4   c     = T.cand[ s ];
            const bool     me_lo = gid < __float_as_uint( c.w );
            const float    wq    = T.inv_lut[ T.mass[ s ] ];
            {   // operands selected by role, one evaluation: lo = the pair's i
                const PushTerms P = push_terms( contact_geom( me_lo ? x : c.x, me_lo ? y : c.y, me_lo ? r : c.z, me_lo ? c.x : x, me_lo ? c.y : y,
                                                              me_lo ? c.z : r, dt ), me_lo ? wme : wq, me_lo ? wq : wme );
                if ( P.push ) {   // x + t and x - (-t) are the same float operation; the hi side subtracts its own terms (:152-155)
                    X = me_lo ? __fadd_rn( X, P.dpx_lo ) : __fsub_rn( X, P.dpx_hi );
                    Y = me_lo ? __fadd_rn( Y, P.dpy_lo ) : __fsub_rn( Y, P.dpy_hi );
                }
            }
            const int jj = j - 1 + (int)( e >> 16 );   // staged row of the partner -> its storage slot
            WB.inc[ (size_t)a * WB.stride + gk ] = ( T.run_g0[ jj ] + ( s - T.run_s0[ jj ] ) ) | ( me_lo ? INC_LO_FLAG : 0u );
        }
        const float4 old = S.pv[ gk ];
        pv_out[ gk ]     = make_float4( X, Y, old.z, old.w );
        WB.deg[ gk ]     = (uint16_t)nh;
        WB.state[ gk ]   = st0;
    }
}

__global__ void __launch_bounds__( TILE_THREADS, PPC_TILE_MINBLOCKS ) k_collide_tile2( const ParticleSet S, float4 *__restrict__ pv_out,
                                                                   const uint32_t *__restrict__ cell_start, const GridDims G, const float dt,
                                                                   const WaveBuffers WB, const int wc, const int tiles_x ) {
    extern __shared__ __align__( 16 ) unsigned char smem_raw[];
    Tile2Smem &T = *reinterpret_cast<Tile2Smem *>( smem_raw );
    const int tx = (int)( blockIdx.x % (unsigned)tiles_x ), ty = (int)( blockIdx.x / (unsigned)tiles_x );
    const int cx0 = tx * wc, cy0 = ty * TILE_R;
    const int cx1 = min( cx0 + wc, G.cols );
    // staged particle count of the whole tile; a tile that would overflow shared memory is cut into column strips that fit
    const int c0 = max( cx0 - 1, 0 ), c1 = min( cx1 + 1, G.cols );
    __syncthreads();
    if ( threadIdx.x < TILE_R + 2 ) {
        const int gy = cy0 - 1 + (int)threadIdx.x;
        uint32_t  len = 0;
        if ( gy >= 0 && gy < G.rows )
            len = cell_start[ (size_t)gy * G.cols + c1 ] - cell_start[ (size_t)gy * G.cols + c0 ];
        T.run_s0[ threadIdx.x ] = len;
    }
    __syncthreads();
    uint32_t total = 0;
#pragma unroll
    for ( int j = 0; j < TILE_R + 2; j++ )
        total += T.run_s0[ j ];
    if ( total == 0 )
        return;
    if ( threadIdx.x == 0 )   // densest tile, in particles per 1024 cells: sizes the wavefront tiles (a substep late)
        atomicMax( &WB.counters[ 9 ], (uint32_t)( ( (unsigned long long)total << 10 ) / (unsigned)( ( TILE_R + 2 ) * ( cx1 - cx0 + 2 ) ) ) );
    const int strips = (int)( ( total + ( TILE2_CAP * 6 / 10 ) - 1 ) / ( TILE2_CAP * 6 / 10 ) );
    const int sw     = max( ( cx1 - cx0 + strips - 1 ) / strips, 1 );
    uint32_t  n_active = 0;
    for ( int sx = cx0; sx < cx1; sx += sw )
        collide_tile2_body( T, S, pv_out, cell_start, G, dt, WB, sx, min( sx + sw, cx1 ), cy0, n_active );
#pragma unroll
    for ( int d = 16; d > 0; d >>= 1 )
        n_active += __shfl_xor_sync( 0xffffffffu, n_active, d );
    if ( ( threadIdx.x & 31u ) == 0 && n_active )
        atomicAdd( &WB.counters[ 8 ], n_active );
}

// ---- K7b: sequential-equivalent velocity resolve, as a wavefront over the pair dependency graph ----------------
// Cooperative launch (all CTAs co-resident).  Round r: every still-active particle looks at the head of its contact
// list; if it is the pair's i and the partner's head is this very pair -- and neither particle was advanced earlier
// in this round -- it fires the pair: impulse of :158-190 from the CURRENT velocities, both particles advance.
// Particles with contacts left go to the next round's list.  One grid barrier per round; the number of rounds is
// the depth of the reference's pair dependency graph.
__device__ __forceinline__ uint32_t head_contact( const ParticleSet &S, const GlobalAccessor &acc, const GridDims &G, const WaveBuffers &WB,
                                                  const uint32_t k, const uint32_t pc, const uint32_t deg ) {
    if ( deg <= WB.maxdeg )
        return WB.inc[ (size_t)pc * WB.stride + k ];
    return nth_contact( acc, G, load_self( S, k ), S.key[ k ], pc );
}

// One wavefront step for particle `me` in round r.  Returns true while the particle still has unresolved contacts.
__device__ __forceinline__ bool wavefront_visit( const ParticleSet &S, const GlobalAccessor &acc, float4 *pv_out, const GridDims &G,
                                                 const float dt, const WaveBuffers &WB, const uint32_t me, const uint32_t r ) {
    const uint32_t deg = WB.deg[ me ];
    if ( deg == 0 )
        return false;
    const uint32_t w  = __ldcg( &WB.state[ me ] );
    const uint32_t pc = w & PC_MASK;
    if ( pc >= deg )
        return false;
    if ( st_round( w ) == r )   // advanced earlier in this round by the partner's thread: look again next round
        return true;
    const uint32_t entry = head_contact( S, acc, G, WB, me, pc, deg );
    if ( !( entry & INC_LO_FLAG ) )   // the pair's j waits for the pair's i to execute it
        return true;
    const uint32_t q    = entry & ~INC_LO_FLAG;
    const uint32_t wq   = __ldcg( &WB.state[ q ] );
    const uint32_t pcq  = wq & PC_MASK;
    const uint32_t degq = WB.deg[ q ];
    if ( st_round( wq ) == r || pcq >= degq || head_contact( S, acc, G, WB, q, pcq, degq ) != me )
        return true;   // the partner still has earlier pairs to resolve
    const Self a = load_self( S, me ), b = load_self( S, q );   // frozen geometry (pass 1, :84-127)
    float2    *va = reinterpret_cast<float2 *>( pv_out + me ) + 1;
    float2    *vb = reinterpret_cast<float2 *>( pv_out + q ) + 1;
    float2     ua = __ldcg( va ), ub = __ldcg( vb );             // current velocities (:159-162)
    const ImpulseTerms T = impulse_terms( contact_geom( a.x, a.y, a.r, b.x, b.y, b.r, dt ), inv_mass_u8( a.m ), inv_mass_u8( b.m ), ua.x,
                                          ua.y, ub.x, ub.y );
    if ( T.impulse ) {
        ua.x = __fadd_rn( ua.x, T.dvx_lo );
        ua.y = __fadd_rn( ua.y, T.dvy_lo );
        ub.x = __fsub_rn( ub.x, T.dvx_hi );
        ub.y = __fsub_rn( ub.y, T.dvy_hi );
        *va  = ua;
        *vb  = ub;
    }
    const uint32_t taint = ( w | wq ) & ST_TAINT;
    WB.state[ me ] = taint | ( r << PC_BITS ) | ( pc + 1 );
    WB.state[ q ]  = taint | ( r << PC_BITS ) | ( pcq + 1 );
    return pc + 1 < deg;
}

// Each CTA owns a contiguous chunk of sorted particles and a private queue of those that still have contacts to
// resolve (queues never grow: particles only leave), so rounds need no global atomics except one add per CTA for
// the termination count.  Round 1 scans the chunk itself; later rounds walk the queue.
__global__ void __launch_bounds__( WAVE_THREADS ) k_resolve_wavefront( const ParticleSet S, float4 *pv_out,
                                                                       const uint32_t *__restrict__ cell_start, const size_t n,
                                                                       const GridDims G, const float dt, const WaveBuffers WB,
                                                                       const int after_tile_passes ) {
    namespace cg        = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    if ( after_tile_passes && __ldcg( &WB.counters[ 5 ] ) == 0 )   // the tile passes resolved everything
        return;
    const GlobalAccessor acc{ S.pv, S.rm, S.gid, cell_start };
    const unsigned       lane = threadIdx.x & 31u;
    __shared__ uint32_t  q_len[ 2 ];   // length of this CTA's queue being read / being written
    volatile uint32_t   *counters = WB.counters;

    const size_t chunk = ( n + gridDim.x - 1 ) / gridDim.x;
    const size_t first = (size_t)blockIdx.x * chunk;
    const size_t last  = first + chunk < n ? first + chunk : n;   // may be <= first for trailing CTAs
    if ( threadIdx.x == 0 )
        q_len[ 0 ] = q_len[ 1 ] = 0;
    __syncthreads();

    for ( uint32_t r = 1;; r++ ) {
        if ( r > 1 && counters[ ( r - 1 ) % 3 ] == 0 )   // nobody has contacts left
            break;
        if ( r >= ST_ROUND_LIMIT ) {     // round stamp would overflow: flag and stop
            if ( blockIdx.x == 0 && threadIdx.x == 0 )
                atomicOr( &WB.counters[ 3 ], 2u );
            break;
        }
        if ( blockIdx.x == 0 && threadIdx.x == 0 )
            counters[ ( r + 1 ) % 3 ] = 0;
        const uint32_t *src     = WB.list[ ( r - 1 ) & 1 ] + first;
        uint32_t       *dst     = WB.list[ r & 1 ] + first;
        uint32_t       *dst_len = &q_len[ r & 1 ];
        const size_t    count   = r == 1 ? ( last > first ? last - first : 0 ) : (size_t)q_len[ ( r - 1 ) & 1 ];

        for ( size_t t0 = ( threadIdx.x & ~31u ); t0 < count; t0 += WAVE_THREADS ) {
            const size_t t    = t0 + lane;
            bool         keep = false;
            uint32_t     me   = 0;
            if ( t < count ) {
                me   = r == 1 ? (uint32_t)( first + t ) : src[ t ];
                keep = wavefront_visit( S, acc, pv_out, G, dt, WB, me, r );
            }
            const unsigned active = __ballot_sync( 0xffffffffu, keep );
            if ( active ) {
                const unsigned leader = (unsigned)( __ffs( active ) - 1 );
                uint32_t       base   = 0;
                if ( lane == leader )
                    base = atomicAdd( dst_len, (uint32_t)__popc( active ) );   // shared-memory counter
                base = __shfl_sync( 0xffffffffu, base, leader );
                if ( keep )
                    dst[ base + (uint32_t)__popc( active & ( ( 1u << lane ) - 1u ) ) ] = me;
            }
        }
        __syncthreads();
        if ( threadIdx.x == 0 ) {
            if ( *dst_len )
                atomicAdd( &WB.counters[ r % 3 ], *dst_len );
            q_len[ ( r - 1 ) & 1 ] = 0;   // becomes the write queue of the next round
            if ( count )
                atomicAdd( reinterpret_cast<unsigned long long *>( WB.counters + 6 ), (unsigned long long)count );
            if ( blockIdx.x == 0 )
                WB.counters[ 4 ] = r;
        }
        grid.sync();
    }
}

// ---- K7b, event-driven global rounds on packed records -------------------------------------------------------------
// The polling rounds above touch every still-active particle in every round, six scattered sectors per visit.  Here
// (a) a particle is only examined when something changed for it: once at the start, and in the round after each of its
// own advances (the executor of a pair queues both particles).  A pair becomes ready exactly when the later of its two
// predecessor pairs fires, and that fire queues the particle they share, so nothing is missed;
// (b) everything an examination or a fire needs sits in two 32-byte records per particle, one sector each:
//     WaveRec {state, deg, first 6 contacts}  and  DynRec {x, y, r, 1/m (frozen), vx, vy (current)}.
// Firing rule and stamps are those of k_resolve_wavefront; a pair is claimed by a compare-and-swap of its i particle's
// state word to this round's stamp, so two threads that examine the two ends of a ready pair fire it once.
struct __align__( 16 ) WaveRec {
    uint32_t state;      // (round of last advance or claim << PC_BITS) | contacts already resolved
    uint32_t deg;
    uint32_t c[ 6 ];     // contacts 0..5: partner slot | INC_LO_FLAG; later ones stay in WaveBuffers::inc
};
struct __align__( 16 ) DynRec {
    float x, y, r, w;    // frozen pass-1 geometry (:84-127) and inverse mass (:141-147)
    float vx, vy;        // current velocity
    float pad0, pad1;
};
constexpr int WREC_INLINE = 6;

struct WaveRecs {
    WaveRec *wr;
    DynRec  *dr;
};

// Build the records of the particles that have contacts (one thread per particle).
__global__ void __launch_bounds__( 256 ) k_wave_pack( const ParticleSet S, const float4 *__restrict__ pv_out, const WaveBuffers WB,
                                                      const WaveRecs R, const size_t n ) {
    const size_t k = (size_t)blockIdx.x * 256 + threadIdx.x;
    if ( k >= n )
        return;
    const uint32_t d = WB.deg[ k ];
    if ( d == 0 )
        return;
    WaveRec w;
    w.state = WB.state[ k ] & ST_TAINT;
    w.deg   = d;
#pragma unroll
    for ( int c = 0; c < WREC_INLINE; c++ )
        w.c[ c ] = ( (uint32_t)c < d && (uint32_t)c < WB.maxdeg ) ? WB.inc[ (size_t)c * WB.stride + k ] : 0u;
    R.wr[ k ] = w;
    const float4 p = S.pv[ k ];
    const float2 a = S.rm[ k ];
    const float4 o = pv_out[ k ];
    DynRec       q;
    q.x = p.x, q.y = p.y, q.r = a.x, q.w = inv_mass_u8( a.y ), q.vx = o.z, q.vy = o.w, q.pad0 = q.pad1 = 0.0f;
    R.dr[ k ] = q;
}

__device__ __forceinline__ uint32_t rec_head( const ParticleSet &S, const GlobalAccessor &acc, const GridDims &G, const WaveBuffers &WB,
                                              const WaveRec &w, const uint32_t k, const uint32_t pc ) {
    if ( pc < (uint32_t)WREC_INLINE ) {   // selects, not w.c[ pc ]: a dynamically indexed member would put every loaded record into local memory
        uint32_t v = w.c[ 0 ];
        v          = pc == 1 ? w.c[ 1 ] : v;
        v          = pc == 2 ? w.c[ 2 ] : v;
        v          = pc == 3 ? w.c[ 3 ] : v;
        v          = pc == 4 ? w.c[ 4 ] : v;
        v          = pc == 5 ? w.c[ 5 ] : v;
        return v;
    }
    if ( w.deg <= WB.maxdeg )
        return WB.inc[ (size_t)pc * WB.stride + k ];
    return nth_contact( acc, G, load_self( S, k ), S.key[ k ], pc );
}

__device__ __forceinline__ WaveRec load_wave_rec( const WaveRec *p ) {
    const uint4 a = __ldcg( reinterpret_cast<const uint4 *>( p ) ), b = __ldcg( reinterpret_cast<const uint4 *>( p ) + 1 );
    WaveRec     w;
    w.state = a.x, w.deg = a.y, w.c[ 0 ] = a.z, w.c[ 1 ] = a.w, w.c[ 2 ] = b.x, w.c[ 3 ] = b.y, w.c[ 4 ] = b.z, w.c[ 5 ] = b.w;
    return w;
}

#ifndef PPC_EV_MINBLOCKS
#define PPC_EV_MINBLOCKS 4   // 64 registers: measured on the 16M pile 9.3 / 7.8 / 7.4 / 7.1 / 8.1 / 10.5 ms with 8 / 6 / 5 / 4 / 3 / 2 CTAs per SM (spills vs loads in flight)
#endif
#ifndef PPC_EV_THREADS
#define PPC_EV_THREADS 256
#endif
constexpr int EV_THREADS = PPC_EV_THREADS;
#ifndef PPC_EV_FOLLOW
#define PPC_EV_FOLLOW 6
#endif
constexpr int EV_FOLLOW  = PPC_EV_FOLLOW;   // most pairs one thread may fire in a row along a chain (the launch passes the actual limit)
constexpr int EV_BUF     = ( EV_FOLLOW + 2 ) * EV_THREADS * 2;   // per-CTA staging of queue pushes: one global atomic per flush

__global__ void __launch_bounds__( EV_THREADS, PPC_EV_MINBLOCKS ) k_resolve_events( const ParticleSet S, float4 *pv_out, const uint32_t *__restrict__ cell_start,
                                                                  const size_t n, const GridDims G, const float dt, const WaveBuffers WB,
                                                                  const WaveRecs R, const int follow_max ) {
    namespace cg        = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const GlobalAccessor acc{ S.pv, S.rm, S.gid, cell_start };
#ifdef PPC_EV_WARPBUF   // queue pushes staged per WARP: no CTA barrier inside a round
    constexpr int        WBUF = EV_BUF / ( EV_THREADS / 32 );
    __shared__ uint32_t  wbuf[ EV_THREADS / 32 ][ WBUF ];
    __shared__ uint32_t  wbuf_n[ EV_THREADS / 32 ];
    uint32_t            *buf   = wbuf[ threadIdx.x >> 5 ];
    uint32_t            &buf_n = wbuf_n[ threadIdx.x >> 5 ];
    if ( ( threadIdx.x & 31u ) == 0 )
        buf_n = 0;
#else
    __shared__ uint32_t  buf[ EV_BUF ];
    __shared__ uint32_t  buf_n, buf_base;
#endif
    volatile uint32_t   *counters = WB.counters;
    const size_t         tid_g = (size_t)blockIdx.x * EV_THREADS + threadIdx.x, nthreads = (size_t)gridDim.x * EV_THREADS;
    if ( threadIdx.x == 0 )
        buf_n = 0;
    __syncthreads();

    for ( uint32_t r = 1;; r++ ) {
        const uint32_t n_q = r == 1 ? 0u : counters[ ( r - 1 ) % 3 ];
        if ( r > 1 && n_q == 0 )
            break;
        if ( r >= ST_ROUND_LIMIT ) {
            if ( blockIdx.x == 0 && threadIdx.x == 0 )
                atomicOr( &WB.counters[ 3 ], 2u );
            break;
        }
        if ( blockIdx.x == 0 && threadIdx.x == 0 )
            counters[ ( r + 1 ) % 3 ] = 0;
        const uint32_t *src   = WB.list[ ( r - 1 ) & 1 ];
        uint32_t       *dst   = WB.list[ r & 1 ];
        uint32_t       *d_cnt = &WB.counters[ r % 3 ];
        const size_t    items = r == 1 ? n : (size_t)n_q;
        const size_t    iters = ( items + nthreads - 1 ) / nthreads;

        for ( size_t it = 0; it < iters; it++ ) {   // uniform trip count: the flush below uses CTA barriers
            const size_t i = it * nthreads + tid_g;
            if ( i < items ) {
                const uint32_t p = r == 1 ? (uint32_t)i : src[ i ];
                if ( r > 1 || WB.deg[ p ] != 0 ) {
                    const WaveRec  A   = load_wave_rec( R.wr + p );
                    const uint32_t pcA = A.state & PC_MASK;
                    if ( pcA < A.deg && st_round( A.state ) != r ) {
                        const uint32_t e = rec_head( S, acc, G, WB, A, p, pcA ), o = e & ~INC_LO_FLAG;
                        const WaveRec  B = load_wave_rec( R.wr + o );
                        const uint32_t pcB = B.state & PC_MASK;
                        if ( pcB < B.deg && st_round( B.state ) != r && ( rec_head( S, acc, G, WB, B, o, pcB ) & ~INC_LO_FLAG ) == p ) {
                            const bool     p_lo = ( e & INC_LO_FLAG ) != 0;
                            const uint32_t lo = p_lo ? p : o, hi = p_lo ? o : p;
                            const uint32_t st_lo = p_lo ? A.state : B.state, st_hi = p_lo ? B.state : A.state;
                            const uint32_t deg_lo = p_lo ? A.deg : B.deg, deg_hi = p_lo ? B.deg : A.deg;
                            const uint32_t claim = ( st_lo & ST_TAINT ) | ( r << PC_BITS ) | ( st_lo & PC_MASK );
                            if ( atomicCAS( &R.wr[ lo ].state, st_lo, claim ) == st_lo ) {   // this thread fires the pair
                                const float4 ga = __ldcg( reinterpret_cast<const float4 *>( R.dr + lo ) );
                                const float4 gb = __ldcg( reinterpret_cast<const float4 *>( R.dr + hi ) );
                                float2      *va = reinterpret_cast<float2 *>( R.dr + lo ) + 2, *vb = reinterpret_cast<float2 *>( R.dr + hi ) + 2;
                                float2       ua = __ldcg( va ), ub = __ldcg( vb );
                                const ImpulseTerms T =
                                    impulse_terms( contact_geom( ga.x, ga.y, ga.z, gb.x, gb.y, gb.z, dt ), ga.w, gb.w, ua.x, ua.y, ub.x, ub.y );
                                if ( T.impulse ) {
                                    ua.x = __fadd_rn( ua.x, T.dvx_lo );
                                    ua.y = __fadd_rn( ua.y, T.dvy_lo );
                                    ub.x = __fsub_rn( ub.x, T.dvx_hi );
                                    ub.y = __fsub_rn( ub.y, T.dvy_hi );
                                }
                                // the pair's j is done for this round
                                const uint32_t nhi = ( st_hi & PC_MASK ) + 1;
                                uint32_t       taint = ( st_lo | st_hi ) & ST_TAINT;   // of the chain's particle from here on
                                *vb                = ub;
                                R.wr[ hi ].state   = taint | ( r << PC_BITS ) | nhi;
                                if ( nhi < deg_hi )
                                    buf[ atomicAdd( &buf_n, 1u ) ] = hi;
                                else
                                    *( reinterpret_cast<float2 *>( pv_out + hi ) + 1 ) = ub;   // last contact done: final velocity

                                // Follow the chain from the pair's i: its list position, record and velocity stay in registers,
                                // and its state word already carries this round's stamp, so nobody else touches it.  Its next
                                // pair (cur, x) may fire right away iff x has not advanced in this round (its velocity was last
                                // written before the previous grid barrier, so it is visible) and x's head is that same pair --
                                // then no other thread can fire it: an examiner of x sees cur's stamp and backs off.
                                WaveRec C;   // contacts of the cursor particle, selected field by field: a reference chosen at run time
                                             // between A and B would force both records into local memory
                                C.state = 0u, C.deg = deg_lo;
#pragma unroll
                                for ( int ci = 0; ci < WREC_INLINE; ci++ )
                                    C.c[ ci ] = p_lo ? A.c[ ci ] : B.c[ ci ];
                                const uint32_t cur = lo, cdeg = deg_lo;
                                uint32_t       cpc = ( st_lo & PC_MASK ) + 1;
                                float2         uc  = ua;
                                for ( int follow = 0; follow < follow_max && cpc < cdeg; follow++ ) {
                                    const uint32_t e2 = rec_head( S, acc, G, WB, C, cur, cpc ), x = e2 & ~INC_LO_FLAG;
                                    const WaveRec  X   = load_wave_rec( R.wr + x );
                                    const uint32_t pcx = X.state & PC_MASK;
                                    if ( st_round( X.state ) == r || pcx >= X.deg || ( rec_head( S, acc, G, WB, X, x, pcx ) & ~INC_LO_FLAG ) != cur )
                                        break;
                                    if ( atomicCAS( &R.wr[ x ].state, X.state, ( X.state & ST_TAINT ) | ( r << PC_BITS ) | pcx ) != X.state )
                                        break;
                                    const float4 gx = __ldcg( reinterpret_cast<const float4 *>( R.dr + x ) );
                                    float2      *vx = reinterpret_cast<float2 *>( R.dr + x ) + 2;
                                    float2       ux = __ldcg( vx );
                                    if ( e2 & INC_LO_FLAG ) {   // cur is this pair's i
                                        const ImpulseTerms F = impulse_terms( contact_geom( ga.x, ga.y, ga.z, gx.x, gx.y, gx.z, dt ), ga.w, gx.w,
                                                                              uc.x, uc.y, ux.x, ux.y );
                                        if ( F.impulse ) {
                                            uc.x = __fadd_rn( uc.x, F.dvx_lo );
                                            uc.y = __fadd_rn( uc.y, F.dvy_lo );
                                            ux.x = __fsub_rn( ux.x, F.dvx_hi );
                                            ux.y = __fsub_rn( ux.y, F.dvy_hi );
                                        }
                                    } else {   // cur is this pair's j
                                        const ImpulseTerms F = impulse_terms( contact_geom( gx.x, gx.y, gx.z, ga.x, ga.y, ga.z, dt ), gx.w, ga.w,
                                                                              ux.x, ux.y, uc.x, uc.y );
                                        if ( F.impulse ) {
                                            ux.x = __fadd_rn( ux.x, F.dvx_lo );
                                            ux.y = __fadd_rn( ux.y, F.dvy_lo );
                                            uc.x = __fsub_rn( uc.x, F.dvx_hi );
                                            uc.y = __fsub_rn( uc.y, F.dvy_hi );
                                        }
                                    }
                                    cpc++;
                                    taint |= X.state & ST_TAINT;
                                    *vx              = ux;
                                    R.wr[ x ].state = taint | ( r << PC_BITS ) | ( pcx + 1 );
                                    if ( pcx + 1 < X.deg )
                                        buf[ atomicAdd( &buf_n, 1u ) ] = x;
                                    else
                                        *( reinterpret_cast<float2 *>( pv_out + x ) + 1 ) = ux;
                                }
                                *va              = uc;
                                R.wr[ cur ].state = taint | ( r << PC_BITS ) | cpc;
                                if ( cpc < cdeg )
                                    buf[ atomicAdd( &buf_n, 1u ) ] = cur;
                                else
                                    *( reinterpret_cast<float2 *>( pv_out + cur ) + 1 ) = uc;
                            }
                        }
                    }
                }
            }
            // pushes were staged in shared memory; flush with one global atomic when the buffer could overflow
#ifdef PPC_EV_WARPBUF
            __syncwarp();
            const uint32_t staged = buf_n;
            __syncwarp();
            if ( staged > (uint32_t)( WBUF - ( EV_FOLLOW + 2 ) * 32 ) || ( it + 1 == iters && staged ) ) {
                uint32_t base = 0;
                if ( ( threadIdx.x & 31u ) == 0 ) {
                    base  = atomicAdd( d_cnt, staged );
                    buf_n = 0;
                }
                base = __shfl_sync( 0xffffffffu, base, 0 );
                for ( uint32_t t = threadIdx.x & 31u; t < staged; t += 32 )
                    dst[ base + t ] = buf[ t ];
                __syncwarp();
            }
#else
            __syncthreads();
            const uint32_t staged = buf_n;   // read between two barriers: the same value in every thread
            __syncthreads();
            if ( staged > (uint32_t)( EV_BUF - ( EV_FOLLOW + 2 ) * EV_THREADS ) || ( it + 1 == iters && staged ) ) {
                if ( threadIdx.x == 0 ) {
                    buf_base = atomicAdd( d_cnt, staged );
                    buf_n    = 0;
                }
                __syncthreads();
                for ( uint32_t t = threadIdx.x; t < staged; t += EV_THREADS )
                    dst[ buf_base + t ] = buf[ t ];
                __syncthreads();
            }
#endif
        }
        if ( threadIdx.x == 0 && blockIdx.x == 0 )
            WB.counters[ 4 ] = r;
        grid.sync();
    }
}

// ---- K7b, tile passes: the same wavefront, run inside shared memory on one tile of cells at a time ------------------
// In a dense packing the dependency DAG is ~60 pairs deep and nearly every particle is active, so global rounds
// stream the whole state ~60 times with scattered access.  A tile pass instead loads the particles of a square tile of
// cells (contact lists, velocities, list positions, frozen geometry) into shared memory once, runs rounds there until
// nothing more can fire -- only pairs with BOTH particles inside the tile are eligible, everything else simply waits --
// and writes velocities and list positions back.  The firing rule is unchanged (head of both lists, not yet advanced
// this round), so any interleaving of tile passes and global rounds is still a topological order of the reference's
// dependency graph.  Passes alternate between four tilings shifted by half a tile; measured on the dense pile one pass
// resolves ~80% of the pairs and four passes >99.9%; k_resolve_wavefront then finishes the remainder.
constexpr int      WTILE_THREADS = 512;    // two particles per thread; three CTAs per SM overlap their barriers
constexpr int      WTILE_CAP     = 1024;   // particles per (sub-)tile: 64 KB of shared memory
constexpr int      WTILE_TW      = 30;     // tile edge in cells: divisible by 2 (half shift), 3 and 5 (odd subdivisions)
constexpr int      WTILE_MAXR    = WTILE_TW;   // rows per tile (run table)
constexpr uint16_t WTILE_NONLOCAL = 0x7fffu;
constexpr uint16_t WTILE_LO       = 0x8000u;
constexpr uint32_t WTILE_MOVED    = 0x80000000u;

struct WaveTileSmem {
    float    x[ WTILE_CAP ], y[ WTILE_CAP ], r[ WTILE_CAP ], w[ WTILE_CAP ];   // frozen geometry, inverse mass
    float2   v[ WTILE_CAP ];                                                    // current velocity
    uint32_t st[ WTILE_CAP ];                                                   // list position | WTILE_MOVED once advanced in this pass
    uint32_t qmark[ WTILE_CAP ];                                                // round in which the particle's head pair was queued
    uint16_t queue[ 2 ][ WTILE_CAP ];                                           // executors (the pair's i) of the pairs ready to fire
    uint16_t fired_hi[ WTILE_CAP ];                                             // partner of each pair fired in the current round
    uint16_t deg[ WTILE_CAP ];
    uint16_t list[ INC_MAXDEG ][ WTILE_CAP ];   // partner as tile-local index | WTILE_LO, or WTILE_NONLOCAL
    uint32_t run_g0[ WTILE_MAXR ], run_s0[ WTILE_MAXR + 1 ];
    uint32_t qlen[ 2 ], left;
};

// One (sub-)tile of a wavefront pass: cells [cxa, cxb) x [cya, cyb).  Returns, through left_tiles, whether any of its
// particles still has unresolved contacts (only counted on the last pass).
__device__ __forceinline__ void wave_tile_body( WaveTileSmem &T, const ParticleSet &S, float4 *pv_out, const uint32_t *__restrict__ cell_start,
                                                const GridDims &G, const float dt, const WaveBuffers &WB, const int cxa, const int cxb,
                                                const int cya, const int cyb, const int last_pass ) {
    __syncthreads();   // shared memory may still be in use by the previous sub-tile
    if ( cxa >= cxb || cya >= cyb )
        return;
    const int nrows = cyb - cya;
    if ( (int)threadIdx.x < nrows ) {
        const size_t row = (size_t)( cya + threadIdx.x ) * G.cols;
        const uint32_t g0 = cell_start[ row + cxa ], g1 = cell_start[ row + cxb ];
        T.run_g0[ threadIdx.x ] = g0;
        T.run_s0[ threadIdx.x ] = g1 - g0;   // length for now
    }
    __syncthreads();
    if ( threadIdx.x == 0 ) {
        uint32_t run = 0;
        for ( int j = 0; j < nrows; j++ ) {
            const uint32_t len = T.run_s0[ j ];
            T.run_s0[ j ]      = run;
            run += len;
        }
        T.run_s0[ nrows ] = run;
        T.qlen[ 0 ] = T.qlen[ 1 ] = 0;
        T.left                    = 0;
    }
    __syncthreads();
    const uint32_t count = T.run_s0[ nrows ];
    if ( count == 0 )
        return;
    if ( count > (uint32_t)WTILE_CAP ) {   // too crowded for shared memory: left to the global rounds
        if ( last_pass && threadIdx.x == 0 )
            atomicAdd( &WB.counters[ 5 ], 1u );
        return;
    }

    // ---- stage (flat index over the tile's particles) ----
    uint32_t left = 0;
    for ( uint32_t s = threadIdx.x; s < count; s += WTILE_THREADS ) {
        int lo = 0, hi = nrows;   // row j with run_s0[j] <= s < run_s0[j+1]
        while ( hi - lo > 1 ) {
            const int mid = ( lo + hi ) >> 1;
            if ( s >= T.run_s0[ mid ] )
                lo = mid;
            else
                hi = mid;
        }
        const int      j = lo;
        const uint32_t g = T.run_g0[ j ] + ( s - T.run_s0[ j ] );
        const uint32_t d = WB.deg[ g ];
        T.deg[ s ]       = (uint16_t)d;
        uint32_t pc      = d ? ( __ldcg( &WB.state[ g ] ) & PC_MASK ) : 0u;
        if ( d > (uint32_t)INC_MAXDEG ) {   // list not stored: this particle waits for the global rounds
            left += pc < d ? 1u : 0u;
            pc = d, T.deg[ s ] = 0;
        }
        T.st[ s ]    = pc < d ? pc : d;
        T.qmark[ s ] = 0;
        if ( pc < d ) {
            const float4 p = S.pv[ g ];
            const float2 a = S.rm[ g ];
            T.x[ s ] = p.x, T.y[ s ] = p.y, T.r[ s ] = a.x, T.w[ s ] = inv_mass_u8( a.y );
            T.v[ s ] = __ldcg( reinterpret_cast<const float2 *>( pv_out + g ) + 1 );
            for ( uint32_t c = pc; c < d; c++ ) {
                const uint32_t e = WB.inc[ (size_t)c * WB.stride + g ], q = e & ~INC_LO_FLAG;
                uint16_t       local = WTILE_NONLOCAL;
#pragma unroll
                for ( int dj = -1; dj <= 1; dj++ ) {   // a contact lies in the same or an adjacent cell row
                    const int jj = j + dj;
                    if ( jj >= 0 && jj < nrows && q >= T.run_g0[ jj ] && q < T.run_g0[ jj ] + ( T.run_s0[ jj + 1 ] - T.run_s0[ jj ] ) )
                        local = (uint16_t)( T.run_s0[ jj ] + ( q - T.run_g0[ jj ] ) );
                }
                T.list[ c ][ s ] = local == WTILE_NONLOCAL ? WTILE_NONLOCAL : (uint16_t)( local | ( ( e & INC_LO_FLAG ) ? WTILE_LO : 0 ) );
            }
        }
    }
    __syncthreads();

    // ---- rounds, event driven: a queue of the pairs that are ready, instead of polling every particle every round ----
    // ready(lo, hi): the pair is at the head of both lists.  A particle's head is unique, so the ready pairs of one round
    // never share a particle; pairs that BECOME ready when a round's pairs advance are found right after that round
    // (all advances visible) and queued, once, for the next round.
    auto head_ready = [ & ]( const uint32_t lo, const uint32_t hi ) -> bool {
        const uint32_t pl = T.st[ lo ] & PC_MASK, ph = T.st[ hi ] & PC_MASK;
        return pl < T.deg[ lo ] && ph < T.deg[ hi ] && T.list[ pl ][ lo ] == (uint16_t)( hi | WTILE_LO ) && T.list[ ph ][ hi ] == (uint16_t)lo;
    };
    for ( uint32_t t = threadIdx.x; t < count; t += WTILE_THREADS ) {   // initial scan
        const uint32_t pc = T.st[ t ] & PC_MASK;
        if ( pc >= T.deg[ t ] )
            continue;
        const uint16_t e = T.list[ pc ][ t ];
        if ( ( e & WTILE_LO ) && ( e & 0x7fffu ) != WTILE_NONLOCAL && head_ready( t, e & 0x7fffu ) )
            T.queue[ 0 ][ atomicAdd( &T.qlen[ 0 ], 1u ) ] = (uint16_t)t;
    }
    __syncthreads();
    for ( uint32_t round = 1;; round++ ) {
        const uint32_t cur = ( round - 1 ) & 1u, nxt = cur ^ 1u;
        const uint32_t nq  = T.qlen[ cur ];
        if ( nq == 0 )
            break;
        if ( threadIdx.x == 0 )
            T.qlen[ nxt ] = 0;
        for ( uint32_t i = threadIdx.x; i < nq; i += WTILE_THREADS ) {   // fire
            const uint32_t t = T.queue[ cur ][ i ];
            const uint32_t pc = T.st[ t ] & PC_MASK;
            const uint32_t q  = T.list[ pc ][ t ] & 0x7fffu;
            const uint32_t pcq = T.st[ q ] & PC_MASK;
            const ContactGeom  C = contact_geom( T.x[ t ], T.y[ t ], T.r[ t ], T.x[ q ], T.y[ q ], T.r[ q ], dt );
            float2             ua = T.v[ t ], ub = T.v[ q ];
            const ImpulseTerms I = impulse_terms( C, T.w[ t ], T.w[ q ], ua.x, ua.y, ub.x, ub.y );
            if ( I.impulse ) {
                ua.x = __fadd_rn( ua.x, I.dvx_lo );
                ua.y = __fadd_rn( ua.y, I.dvy_lo );
                ub.x = __fsub_rn( ub.x, I.dvx_hi );
                ub.y = __fsub_rn( ub.y, I.dvy_hi );
                T.v[ t ] = ua;
                T.v[ q ] = ub;
            }
            T.st[ t ]       = WTILE_MOVED | ( pc + 1 );
            T.st[ q ]       = WTILE_MOVED | ( pcq + 1 );
            T.fired_hi[ i ] = (uint16_t)q;
        }
        __syncthreads();
        for ( uint32_t i = threadIdx.x; i < 2 * nq; i += WTILE_THREADS ) {   // which pairs did that make ready?
            const uint32_t p  = ( i & 1u ) ? T.fired_hi[ i >> 1 ] : T.queue[ cur ][ i >> 1 ];
            const uint32_t pc = T.st[ p ] & PC_MASK;
            if ( pc >= T.deg[ p ] )
                continue;
            const uint16_t e = T.list[ pc ][ p ];
            if ( ( e & 0x7fffu ) == WTILE_NONLOCAL )
                continue;
            const uint32_t lo = ( e & WTILE_LO ) ? p : ( e & 0x7fffu ), hi = ( e & WTILE_LO ) ? ( e & 0x7fffu ) : p;
            if ( head_ready( lo, hi ) && atomicExch( &T.qmark[ lo ], round ) != round )
                T.queue[ nxt ][ atomicAdd( &T.qlen[ nxt ], 1u ) ] = (uint16_t)lo;
        }
        __syncthreads();
    }

    // ---- write back what moved ----
    for ( uint32_t s = threadIdx.x; s < count; s += WTILE_THREADS ) {
        const uint32_t w = T.st[ s ];
        if ( w & WTILE_MOVED ) {   // advanced at least once in this pass
            int lo = 0, hi = nrows;
            while ( hi - lo > 1 ) {
                const int mid = ( lo + hi ) >> 1;
                if ( s >= T.run_s0[ mid ] )
                    lo = mid;
                else
                    hi = mid;
            }
            const uint32_t g = T.run_g0[ lo ] + ( s - T.run_s0[ lo ] );
            WB.state[ g ]    = w & PC_MASK;
            *( reinterpret_cast<float2 *>( pv_out + g ) + 1 ) = T.v[ s ];
        }
        if ( ( w & PC_MASK ) < T.deg[ s ] )
            left++;
    }
    if ( last_pass ) {   // how much is left for the global rounds
        const unsigned any = __ballot_sync( 0xffffffffu, left != 0 );
        if ( any && ( threadIdx.x & 31u ) == 0 )
            atomicAdd( &T.left, 1u );
        __syncthreads();
        if ( threadIdx.x == 0 && T.left )
            atomicAdd( &WB.counters[ 5 ], T.left );
    }
}

// Work items of one tile pass.  Density varies across the box (a polydisperse pile holds ten times more small particles
// per cell than large ones), so a tile that would overflow shared memory is cut into f x f sub-tiles, f in {3, 5}.  f is
// ODD on purpose: the shifted passes move the tiling by tw/2 = (f/2) sub-tile widths, a half-integer, so every sub-tile
// boundary of one pass falls in the middle of a sub-tile of the shifted pass (and an odd subdivision never shares a
// boundary with a half-shifted odd one).  Each sub-tile is its own CTA, so their memory latencies overlap.
struct WaveItem {
    uint16_t tx, ty;
    uint8_t  f, kx, ky, pad;
};

// One warp per (pass, tile): count the tile's particles, choose f, append f*f items to the pass's list.
__global__ void __launch_bounds__( 256 ) k_wave_plan( const uint32_t *__restrict__ cell_start, const GridDims G, const int tw, const int passes,
                                                      WaveItem *__restrict__ items, const uint32_t items_stride, uint32_t *__restrict__ n_items,
                                                      uint32_t *__restrict__ left_flag ) {
    const unsigned lane = threadIdx.x & 31u;
    const size_t   w    = ( (size_t)blockIdx.x * 256 + threadIdx.x ) >> 5;
    const int      txn = ( G.cols + tw / 2 + tw - 1 ) / tw, tyn = ( G.rows + tw / 2 + tw - 1 ) / tw;   // enough for every shift
    const size_t   per_pass = (size_t)txn * tyn;
    if ( w >= per_pass * passes )
        return;
    const int pass = (int)( w / per_pass ), tile = (int)( w % per_pass );
    const int tx = tile % txn, ty = tile / txn;
    const int off_x = ( pass & 1 ) ? tw / 2 : 0, off_y = ( ( pass >> 1 ) & 1 ) ? tw / 2 : 0;
    const int cxa = max( tx * tw - off_x, 0 ), cxb = min( tx * tw - off_x + tw, G.cols );
    const int cya = max( ty * tw - off_y, 0 ), cyb = min( ty * tw - off_y + tw, G.rows );
    uint32_t  total = 0;
    if ( cxa < cxb )
        for ( int y = cya + (int)lane; y < cyb; y += 32 )
            total += cell_start[ (size_t)y * G.cols + cxb ] - cell_start[ (size_t)y * G.cols + cxa ];
#pragma unroll
    for ( int d = 16; d > 0; d >>= 1 )
        total += __shfl_xor_sync( 0xffffffffu, total, d );
    if ( total == 0 )
        return;
    const uint32_t f    = total <= (uint32_t)( WTILE_CAP * 85 / 100 ) ? 1u : ( total <= 9u * ( WTILE_CAP / 2 ) ? 3u : 5u );
    uint32_t       base = 0;
    if ( lane == 0 )
        base = atomicAdd( &n_items[ pass ], f * f );
    base = __shfl_sync( 0xffffffffu, base, 0 );
    if ( lane < f * f ) {
        WaveItem it;
        it.tx = (uint16_t)tx, it.ty = (uint16_t)ty, it.f = (uint8_t)f, it.kx = (uint8_t)( lane % f ), it.ky = (uint8_t)( lane / f ), it.pad = 0;
        if ( base + lane < items_stride )
            items[ (size_t)pass * items_stride + base + lane ] = it;
        else
            atomicAdd( left_flag, 1u );   // list full (cannot happen with the host's bound): leave the work to the global rounds
    }
}

__global__ void __launch_bounds__( WTILE_THREADS, 3 ) k_wave_tile( const ParticleSet S, float4 *pv_out, const uint32_t *__restrict__ cell_start,
                                                                   const GridDims G, const float dt, const WaveBuffers WB, const int tw,
                                                                   const int pass, const WaveItem *__restrict__ items,
                                                                   const uint32_t items_stride, const uint32_t *__restrict__ n_items,
                                                                   const int last_pass ) {
    if ( blockIdx.x >= min( n_items[ pass ], items_stride ) )
        return;
    extern __shared__ __align__( 16 ) unsigned char smem_raw[];
    WaveTileSmem  &T  = *reinterpret_cast<WaveTileSmem *>( smem_raw );
    const WaveItem it = items[ (size_t)pass * items_stride + blockIdx.x ];
    const int      off_x = ( pass & 1 ) ? tw / 2 : 0, off_y = ( ( pass >> 1 ) & 1 ) ? tw / 2 : 0;
    const int      sw = tw / it.f;                                       // tw is a multiple of 30
    const int      ox = it.tx * tw - off_x + it.kx * sw, oy = it.ty * tw - off_y + it.ky * sw;   // unclipped origin of the sub-tile
    wave_tile_body( T, S, pv_out, cell_start, G, dt, WB, max( ox, 0 ), min( ox + sw, G.cols ), max( oy, 0 ), min( oy + sw, G.rows ), last_pass );
}

// ---- candidate-pair list in the reference's order (parity exporter; not on the hot path) -----------------------
// Pass 1 counts, per API index i, the pairs (i, j > i); after an exclusive scan pass 2 writes them at offset[i]
// in scan order, which reproduces the list of src/particles_solver.c:311-383 entry for entry.
__global__ void __launch_bounds__( COLLIDE_THREADS ) k_pairs_count( const ParticleSet S, const uint32_t *__restrict__ cell_start, const size_t n,
                                                                    const GridDims G, uint32_t *__restrict__ count_by_gid ) {
    const size_t k = (size_t)blockIdx.x * COLLIDE_THREADS + threadIdx.x;
    if ( k >= n )
        return;
    const Self           me = load_self( S, k );
    const GlobalAccessor acc{ S.pv, S.rm, S.gid, cell_start };
    const uint32_t       key = S.key[ k ];
    const int            cy = (int)( key / (uint32_t)G.cols ), cx = (int)( key - (uint32_t)cy * (uint32_t)G.cols );
    uint32_t             cnt = 0;
    for_each_candidate( acc, G, cx, cy, [ & ]( const uint32_t q ) {
        const Cand c = acc.geom( q );
        if ( c.gid > me.gid && in_contact( me, c ) )
            cnt++;
    } );
    count_by_gid[ me.gid ] = cnt;
}

__global__ void __launch_bounds__( COLLIDE_THREADS ) k_pairs_emit( const ParticleSet S, const uint32_t *__restrict__ cell_start, const size_t n,
                                                                   const GridDims G, const uint32_t *__restrict__ offset_by_gid,
                                                                   int32_t *__restrict__ pairs, const size_t max_pairs ) {
    const size_t k = (size_t)blockIdx.x * COLLIDE_THREADS + threadIdx.x;
    if ( k >= n )
        return;
    const Self           me = load_self( S, k );
    const GlobalAccessor acc{ S.pv, S.rm, S.gid, cell_start };
    const uint32_t       key = S.key[ k ];
    const int            cy = (int)( key / (uint32_t)G.cols ), cx = (int)( key - (uint32_t)cy * (uint32_t)G.cols );
    size_t               at = offset_by_gid[ me.gid ];
    for_each_candidate( acc, G, cx, cy, [ & ]( const uint32_t q ) {
        const Cand c = acc.geom( q );
        if ( c.gid > me.gid && in_contact( me, c ) ) {
            if ( at < max_pairs ) {
                pairs[ 2 * at ]     = (int32_t)me.gid;
                pairs[ 2 * at + 1 ] = (int32_t)c.gid;
            }
            at++;
        }
    } );
}

// The pair set as the HOT kernel found it: every contact-list entry in which the particle is the pair's i (k_collide_tile2 /
// k_collide_tile / k_collide_global wrote them for the wavefront).  out_count[0] = pairs, [1] = particles whose list was
// longer than the stored positions (their pairs are incomplete here).
__global__ void __launch_bounds__( COLLIDE_THREADS ) k_pairs_from_lists( const uint32_t *__restrict__ gid, const WaveBuffers WB, const size_t n,
                                                                         int32_t *__restrict__ pairs, const uint32_t max_pairs,
                                                                         uint32_t *__restrict__ out_count ) {
    const size_t k = (size_t)blockIdx.x * COLLIDE_THREADS + threadIdx.x;
    if ( k >= n )
        return;
    const uint32_t deg = WB.deg[ k ];
    if ( deg > WB.maxdeg )
        atomicAdd( &out_count[ 1 ], 1u );
    for ( uint32_t c = 0; c < deg && c < WB.maxdeg; c++ ) {
        const uint32_t e = WB.inc[ (size_t)c * WB.stride + k ];
        if ( !( e & INC_LO_FLAG ) )
            continue;
        const uint32_t at = atomicAdd( &out_count[ 0 ], 1u );
        if ( at < max_pairs ) {
            pairs[ 2 * at ]     = (int32_t)gid[ k ];
            pairs[ 2 * at + 1 ] = (int32_t)gid[ e & ~INC_LO_FLAG ];
        }
    }
}

}   // namespace ppc
