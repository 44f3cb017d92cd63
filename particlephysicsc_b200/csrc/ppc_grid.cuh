// ppc_grid.cuh -- streaming kernels of one substep: integrate + wall-constrain + cell key + histogram (K1),
// maximum radius (K2), counting-sort scatter (K3/K5), canonicalise + gather (K6), and the host<->device packers.
//
// Storage model.  Particles live on the device in STORAGE order, which after the first substep is the canonical
// sorted order of the reference's grid: ascending (cell key, API index).  Each particle carries its API index
// (gid).  Because particles move much less than one cell per substep, storage order stays almost sorted from
// one substep to the next, so every pass below streams HBM sequentially and the per-cell atomics of the
// counting sort are warp-aggregated (neighbouring lanes share a cell).
#pragma once
#include "ppc_common.cuh"
#include "ppc_math.cuh"

namespace ppc {

constexpr int STREAM_THREADS = 256;

struct StepParams {
    float dt_update;      // time step of the pending particlesUpdate (if do_integrate)
    int   do_integrate;   // reference src/particles_solver.c:18-28
    int   do_colour;      // reference src/particles_solver.c:31-39 (only on substeps whose result is drawn)
    int   do_walls_keys;  // reference src/particles_solver.c:207-228 and :280-291
    int   do_histogram;   // count particles per cell key (off in slab mode: the window's histogram is built after the exchange)
    int   defer_state;    // K1 writes only the key (and the colour): k_canonical_gather repeats the integrate + wall pass on the record it
                          // moves anyway and writes the result once (16 B per particle less written here, the same 16 B read there either way)
    float wall_w, wall_h; // (float)(int)window_width / height
    GridDims grid;
};

// ---- K1: integrate + wall-constrain + cell key + warp-aggregated cell histogram ------------------------------
// One thread per storage slot; 128-bit load/store of (p_x, p_y, v_x, v_y).  Algorithmic traffic: 20 B read
// (pv + radius; mass only when colouring) + 4 B key (+ 16 B write unless the gather applies the state, P.defer_state).
// With cell_rank given, every particle also keeps what its cell's counter read when it was counted: its (arbitrary but unique) rank
// inside the cell, so the counting sort needs no second round of atomics (k_cell_place).
__global__ void __launch_bounds__( STREAM_THREADS ) k_integrate_walls_keys( ParticleSet S, uint32_t *__restrict__ cell_count, const size_t n,
                                                                            const StepParams P, uint32_t *__restrict__ cell_rank ) {
    const size_t   k     = (size_t)blockIdx.x * STREAM_THREADS + threadIdx.x;
    const bool     valid = k < n;
    const unsigned lane  = threadIdx.x & 31u;
    uint32_t       key   = 0xffffffffu;
    if ( valid ) {
        float4       s = S.pv[ k ];
        const float2 a = S.rm[ k ];
        if ( P.do_integrate ) {
            integrate( s.x, s.y, s.z, s.w, P.dt_update );
            if ( P.do_colour )
                S.col[ k ] = energy_colour( s.z, s.w, a.y );
        }
        if ( P.do_walls_keys ) {
            walls( s.x, s.y, s.z, s.w, a.x, P.wall_w, P.wall_h );
            const int cx = cell_coord( s.x, P.grid.cell_size, P.grid.cols );
            const int cy = cell_coord( s.y, P.grid.cell_size, P.grid.rows );
            key          = (uint32_t)( cy * P.grid.cols + cx );   // reference :291
            S.key[ k ]   = key;
        }
        if ( !P.defer_state )
            S.pv[ k ] = s;
    }
    if ( P.do_walls_keys && P.do_histogram ) {   // whole warp arrives here: lanes of one cell elect a leader that adds their count
        const unsigned peers  = __match_any_sync( 0xffffffffu, key );
        const unsigned leader = (unsigned)( __ffs( peers ) - 1 );
        uint32_t       base   = 0;
        if ( valid && lane == leader )
            base = atomicAdd( &cell_count[ key ], (uint32_t)__popc( peers ) );
        if ( cell_rank ) {
            base = __shfl_sync( 0xffffffffu, base, leader );
            if ( valid )
                cell_rank[ k ] = base + (uint32_t)__popc( peers & ( ( 1u << lane ) - 1u ) );
        }
    }
}

// ---- K2: maximum radius (reference :231-236), only when the radius set changed -------------------------------
// Radii are positive finite floats, so their bit patterns order like unsigned integers; out must be pre-set to
// the bits of 6.0f (the reference's floor).
__global__ void __launch_bounds__( STREAM_THREADS ) k_max_radius( const float2 *__restrict__ rm, const size_t n, uint32_t *__restrict__ out ) {
    float m = 0.0f;
    for ( size_t k = (size_t)blockIdx.x * STREAM_THREADS + threadIdx.x; k < n; k += (size_t)gridDim.x * STREAM_THREADS ) {
        const float r = rm[ k ].x;
        if ( r > m )   // same comparison as the reference: NaN never wins
            m = r;
    }
#pragma unroll
    for ( int d = 16; d > 0; d >>= 1 ) {
        const float o = __shfl_xor_sync( 0xffffffffu, m, d );
        if ( o > m )
            m = o;
    }
    if ( ( threadIdx.x & 31u ) == 0 && m > 0.0f )
        atomicMax( out, __float_as_uint( m ) );
}

// ---- K5: counting-sort scatter ------------------------------------------------------------------------------
// slot_of[pos] = storage slot that lands at sorted position pos; positions inside a cell are handed out by a
// warp-aggregated atomic, so the order inside a cell is arbitrary here and fixed by k_canonical_gather.
// The atomic counts the cell's own histogram entry DOWN: when the scatter is done the histogram is all zeros again, which is
// what the next substep's K1 needs -- no separate fill counters, no clearing pass over the cell table (two per substep before).
__global__ void __launch_bounds__( STREAM_THREADS ) k_cell_scatter( const uint32_t *__restrict__ key, uint32_t *__restrict__ cell_count,
                                                                    const uint32_t *__restrict__ cell_start, uint32_t *__restrict__ slot_of,
                                                                    const size_t n ) {
    const size_t   k     = (size_t)blockIdx.x * STREAM_THREADS + threadIdx.x;
    const bool     valid = k < n;
    const unsigned lane  = threadIdx.x & 31u;
    const uint32_t c     = valid ? key[ k ] : 0xffffffffu;
    const unsigned peers  = __match_any_sync( 0xffffffffu, c );
    const unsigned leader = (unsigned)( __ffs( peers ) - 1 );
    uint32_t       base   = 0;
    if ( valid && lane == leader )
        base = atomicSub( &cell_count[ c ], (uint32_t)__popc( peers ) ) - (uint32_t)__popc( peers );
    base = __shfl_sync( peers, base, leader );
    if ( valid )
        slot_of[ cell_start[ c ] + base + (uint32_t)__popc( peers & ( ( 1u << lane ) - 1u ) ) ] = (uint32_t)k;
}

// K5 when K1 left ranks: sorted position = cell start + rank, no atomics.  The particle of rank 0 clears its cell's histogram entry
// (every non-empty cell has exactly one), so the histogram is all zeros again for the next substep.
__global__ void __launch_bounds__( STREAM_THREADS ) k_cell_place( const uint32_t *__restrict__ key, const uint32_t *__restrict__ cell_rank,
                                                                  uint32_t *__restrict__ cell_count, const uint32_t *__restrict__ cell_start,
                                                                  uint32_t *__restrict__ slot_of, const size_t n ) {
    const size_t k = (size_t)blockIdx.x * STREAM_THREADS + threadIdx.x;
    if ( k >= n )
        return;
    const uint32_t c = key[ k ], r = cell_rank[ k ];
    slot_of[ cell_start[ c ] + r ] = (uint32_t)k;
    if ( r == 0 )
        cell_count[ c ] = 0u;
}

// ---- K6: canonicalise + gather ------------------------------------------------------------------------------
// For sorted position pos holding storage slot src: the particle's final position is cell_start[c] + (number of
// particles of the same cell with a smaller API index).  That makes storage the stable ascending (key, index)
// order -- the canonical form of the reference's chains (:293-294) -- whatever order the scatter or the radix
// sort left inside the cell.  Then the whole particle record moves from the old set to the new one.
// With rec_a / mcol given ("resolve"=3) the gather also leaves what the tile kernel stages with bulk copies (ppc_gs.cuh):
//   rec_a = {x, y, radius, API index}   mcol = mass byte (:141-142) | cell column byte << 8
__global__ void __launch_bounds__( STREAM_THREADS ) k_canonical_gather( const ParticleSet In, const ParticleSet Out,
                                                                        const uint32_t *__restrict__ slot_of,
                                                                        const uint32_t *__restrict__ cell_start, const size_t n,
                                                                        const int move_colour, const uint32_t unranked_cell,
                                                                        float4 *__restrict__ rec_a, uint16_t *__restrict__ mcol, const uint32_t cols,
                                                                        const StepParams P ) {
    const size_t pos = (size_t)blockIdx.x * STREAM_THREADS + threadIdx.x;
    if ( pos >= n )
        return;
    const uint32_t src = slot_of[ pos ];
    const uint32_t c   = In.key[ src ];
    const uint32_t g   = In.gid[ src ];
    const uint32_t b = cell_start[ c ], e = cell_start[ c + 1 ];
    uint32_t       rank = 0;
    if ( c == unranked_cell )   // a slab window's drop bin: thousands of slots nobody reads, any order will do
        rank = (uint32_t)pos - b;
    else
        for ( uint32_t q = b; q < e; q++ )
            rank += ( In.gid[ slot_of[ q ] ] < g ) ? 1u : 0u;
    const uint32_t dst = b + rank;
    float4         s4  = In.pv[ src ];
    const float2   a2  = In.rm[ src ];
    if ( P.defer_state ) {   // K1 computed exactly this to find the key and kept it to itself
        if ( P.do_integrate )
            integrate( s4.x, s4.y, s4.z, s4.w, P.dt_update );
        walls( s4.x, s4.y, s4.z, s4.w, a2.x, P.wall_w, P.wall_h );
    }
    if ( rec_a ) {
        const uint32_t cx = c - ( c / cols ) * cols;
        rec_a[ dst ] = make_float4( s4.x, s4.y, a2.x, __uint_as_float( g ) );
        mcol[ dst ]  = (uint16_t)( ( (unsigned)__float2int_rz( a2.y ) & 0xffu ) | ( ( cx & 0xffu ) << 8 ) );
    }
    Out.pv[ dst ]  = s4;
    Out.rm[ dst ]  = a2;
    Out.gid[ dst ] = g;
    Out.key[ dst ] = c;
    if ( move_colour )
        Out.col[ dst ] = In.col[ src ];
}

// ---- host -> device: pack API-order SoA staging into storage slots [first, first+count) ------------------------
// New particles enter storage at slot == API index (appended, unsorted); the next grid build sorts them in.
__global__ void __launch_bounds__( STREAM_THREADS ) k_pack_upload( const ParticleSet S, const float *__restrict__ px, const float *__restrict__ py,
                                                                   const float *__restrict__ vx, const float *__restrict__ vy,
                                                                   const float *__restrict__ mass, const float *__restrict__ radius,
                                                                   const uchar4 *__restrict__ col, const size_t first, const size_t count ) {
    const size_t t = (size_t)blockIdx.x * STREAM_THREADS + threadIdx.x;
    if ( t >= count )
        return;
    const size_t k = first + t;   // staging arrays are indexed by API index as well
    S.pv[ k ]  = make_float4( px[ k ], py[ k ], vx[ k ], vy[ k ] );
    S.rm[ k ]  = make_float2( radius[ k ], mass[ k ] );
    S.gid[ k ] = (uint32_t)k;
    S.key[ k ] = 0u;
    S.col[ k ] = col[ k ];
}

// ---- device -> host: scatter storage order back to API order into the staging arrays --------------------------
constexpr unsigned PPC_POS = 1u, PPC_VEL = 2u, PPC_COLOR = 4u;
__global__ void __launch_bounds__( STREAM_THREADS ) k_unpack_download( const ParticleSet S, float *__restrict__ px, float *__restrict__ py,
                                                                       float *__restrict__ vx, float *__restrict__ vy,
                                                                       uchar4 *__restrict__ col, const size_t n, const unsigned mask ) {
    const size_t k = (size_t)blockIdx.x * STREAM_THREADS + threadIdx.x;
    if ( k >= n )
        return;
    const uint32_t g = S.gid[ k ];
    const float4   s = S.pv[ k ];
    if ( mask & PPC_POS ) {
        px[ g ] = s.x;
        py[ g ] = s.y;
    }
    if ( mask & PPC_VEL ) {
        vx[ g ] = s.z;
        vy[ g ] = s.w;
    }
    if ( mask & PPC_COLOR )
        col[ g ] = S.col[ k ];
}

// "resolve"=3 keeps the pre-resolve positions in the tile kernel's staged records; the pair exporter wants them as pv
__global__ void __launch_bounds__( STREAM_THREADS ) k_rec_to_pv( const float4 *__restrict__ rec_a, float4 *__restrict__ pv, const size_t n ) {
    const size_t k = (size_t)blockIdx.x * STREAM_THREADS + threadIdx.x;
    if ( k >= n )
        return;
    const float4 a = rec_a[ k ];
    pv[ k ] = make_float4( a.x, a.y, 0.0f, 0.0f );
}

// "resolve"=3 fell back to the exact resolve: positions back to what the pair search saw; the velocities were not touched
__global__ void __launch_bounds__( STREAM_THREADS ) k_rec_to_positions( const float4 *__restrict__ rec_a, float4 *__restrict__ pv, const size_t n ) {
    const size_t k = (size_t)blockIdx.x * STREAM_THREADS + threadIdx.x;
    if ( k >= n )
        return;
    const float4 a = rec_a[ k ];
    *reinterpret_cast<float2 *>( pv + k ) = make_float2( a.x, a.y );
}

// keys by API index / sorted order, for the parity exporters
__global__ void __launch_bounds__( STREAM_THREADS ) k_export_keys( const ParticleSet S, uint32_t *__restrict__ keys_by_gid, const size_t n ) {
    const size_t k = (size_t)blockIdx.x * STREAM_THREADS + threadIdx.x;
    if ( k < n )
        keys_by_gid[ S.gid[ k ] ] = S.key[ k ];
}

}   // namespace ppc
