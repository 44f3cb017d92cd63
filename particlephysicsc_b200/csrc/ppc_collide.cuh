// ppc_collide.cuh -- narrow phase + contact resolution (K7), and the candidate-pair exporter used by parity tests.
//
// Input is the cell-sorted particle set of this substep (storage order = ascending (cell key, API index)) plus the
// cell-start table.  Each thread owns ONE particle and gathers the contributions of all its contacts, so there are
// no atomics and no write conflicts: the result is deterministic, run to run and across GPU counts.
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
//  * Velocity impulses (:158-190) are evaluated Jacobi-style: every pair reads the velocities the resolve started
//    with, whereas the reference lets earlier pairs of the list feed later ones (Gauss-Seidel).  Summation order is
//    the same list order.  oracle/ppc_oracle.c:ppc_oracle_resolve_jacobi is the CPU model of exactly this.
#pragma once
#include "ppc_common.cuh"
#include "ppc_math.cuh"

namespace ppc {

constexpr int COLLIDE_THREADS = 128;
constexpr int COLLIDE_STASH   = 24;   // contacts kept per thread before falling back to the streaming path

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
    __device__ __forceinline__ void cell_range( const uint32_t c, uint32_t &b, uint32_t &e ) const {
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

template <class Acc>
__device__ __forceinline__ void apply_contact( const Acc &acc, const Self &me, const uint32_t q, const float dt, float4 &out ) {
    const Cand c = acc.geom( q );
    float      qvx, qvy, qm;
    acc.dyn( q, qvx, qvy, qm );
    if ( me.gid < c.gid ) {   // this particle is the pair's i
        const PairTerms T = pair_terms( me.x, me.y, me.r, me.m, me.vx, me.vy, c.x, c.y, c.r, qm, qvx, qvy, dt );
        if ( T.push ) {
            out.x = __fadd_rn( out.x, T.dpx_lo );
            out.y = __fadd_rn( out.y, T.dpy_lo );
        }
        if ( T.impulse ) {
            out.z = __fadd_rn( out.z, T.dvx_lo );
            out.w = __fadd_rn( out.w, T.dvy_lo );
        }
    } else {   // this particle is the pair's j
        const PairTerms T = pair_terms( c.x, c.y, c.r, qm, qvx, qvy, me.x, me.y, me.r, me.m, me.vx, me.vy, dt );
        if ( T.push ) {
            out.x = __fsub_rn( out.x, T.dpx_hi );
            out.y = __fsub_rn( out.y, T.dpy_hi );
        }
        if ( T.impulse ) {
            out.z = __fsub_rn( out.z, T.dvx_hi );
            out.w = __fsub_rn( out.w, T.dvy_hi );
        }
    }
}

// Resolve every contact of one particle.  Returns its (p_x, p_y, v_x, v_y) after the resolve.
template <class Acc, int STASH>
__device__ float4 resolve_particle( const Acc &acc, const GridDims &G, const Self &me, const uint32_t key, const float dt ) {
    const int cy = (int)( key / (uint32_t)G.cols ), cx = (int)( key - (uint32_t)cy * (uint32_t)G.cols );
    float4    out = make_float4( me.x, me.y, me.vx, me.vy );

    // One sweep in reference order.  Contacts in which this particle is i go to the front of the stash in the
    // order met (already the list order); contacts in which it is j go to the back and are sorted by index.
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
        // pairs (i, me) with i < me come first in the reference list, by ascending i
        for ( int a = 0; a < n_j; a++ ) {
            int best = STASH - n_j + a;
            for ( int b = best + 1; b < STASH; b++ )
                if ( stash_g[ b ] < stash_g[ best ] )
                    best = b;
            const int      at = STASH - n_j + a;
            const uint32_t tq = stash_q[ best ], tg = stash_g[ best ];
            stash_q[ best ] = stash_q[ at ];
            stash_g[ best ] = stash_g[ at ];
            stash_q[ at ]   = tq;
            stash_g[ at ]   = tg;
            apply_contact( acc, me, tq, dt, out );
        }
        // then pairs (me, j) in scan order
        for ( int a = 0; a < n_i; a++ )
            apply_contact( acc, me, stash_q[ a ], dt, out );
        return out;
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
        apply_contact( acc, me, best_q, dt, out );
        have_last = true;
        last      = best_g;
    }
    for_each_candidate( acc, G, cx, cy, [ & ]( const uint32_t q ) {
        const Cand c = acc.geom( q );
        if ( c.gid <= me.gid || !in_contact( me, c ) )
            return;
        apply_contact( acc, me, q, dt, out );
    } );
    return out;
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

// ---- K7, global-memory form: one thread per sorted particle --------------------------------------------------
template <int STASH>
__global__ void __launch_bounds__( COLLIDE_THREADS ) k_collide_global( const ParticleSet S, float4 *__restrict__ pv_out,
                                                                       const uint32_t *__restrict__ cell_start, const size_t n,
                                                                       const GridDims G, const float dt ) {
    const size_t k = (size_t)blockIdx.x * COLLIDE_THREADS + threadIdx.x;
    if ( k >= n )
        return;
    const Self           me = load_self( S, k );
    const GlobalAccessor acc{ S.pv, S.rm, S.gid, cell_start };
    pv_out[ k ] = resolve_particle<GlobalAccessor, STASH>( acc, G, me, S.key[ k ], dt );
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

}   // namespace ppc
