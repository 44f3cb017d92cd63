// ppc_radix.cuh -- hand-written stable LSD radix sort of (key, value) uint32 pairs (no CUB/Thrust).
//
// Broad-phase use: keys are cell keys cy*cols+cx (reference src/particles_solver.c:291, at most 25 bits for the
// 65,535-pixel world), values are storage slots.  A stable sort by key of slots that are in API-index order yields
// exactly the canonical (key, index) order the reference's per-cell chains encode (src/particles_solver.c:293-294
// push-front => descending index per chain; canonical form = ascending).
//
// 8-bit digits, ceil(key_bits / 8) passes.  Per pass:
//   k_radix_hist    : each CTA counts the digits of its 2048-element tile (warp-aggregated shared-memory
//                     histogram) and writes hist[digit][tile]
//   exclusive_scan  : over hist laid out digit-major => global base of every (digit, tile)
//   k_radix_scatter : each CTA re-reads its tile, ranks every element among equal digits IN TILE ORDER
//                     (warp match + per-warp digit counters + cross-warp prefix) and scatters key and value.
// Stability: element order inside a tile is (warp, round, lane) = ascending index; tiles are ordered by the
// digit-major scan.
#pragma once
#include "ppc_common.cuh"
#include "ppc_scan.cuh"

namespace ppc {

constexpr int RADIX_THREADS = 256;
constexpr int RADIX_WARPS   = RADIX_THREADS / 32;
constexpr int RADIX_ITEMS   = 8;                              // rounds per warp
constexpr int RADIX_TILE    = RADIX_THREADS * RADIX_ITEMS;    // 2048
constexpr int RADIX_BITS    = 8;
constexpr int RADIX_BINS    = 1 << RADIX_BITS;

__global__ void __launch_bounds__( RADIX_THREADS ) k_radix_hist( const uint32_t *__restrict__ keys, uint32_t *__restrict__ hist,
                                                                 const size_t n, const unsigned num_tiles, const int shift ) {
    __shared__ uint32_t bins[ RADIX_BINS ];
    bins[ threadIdx.x ] = 0;   // RADIX_THREADS == RADIX_BINS
    __syncthreads();
    const size_t   base = (size_t)blockIdx.x * RADIX_TILE;
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for ( int r = 0; r < RADIX_ITEMS; r++ ) {
        const size_t   idx   = base + (size_t)r * RADIX_THREADS + threadIdx.x;
        const bool     valid = idx < n;
        const uint32_t d     = valid ? ( ( keys[ idx ] >> shift ) & ( RADIX_BINS - 1 ) ) : ( 0x10000u | lane );
        const unsigned peers = __match_any_sync( 0xffffffffu, d );
        if ( valid && lane == (unsigned)( __ffs( peers ) - 1 ) )
            atomicAdd( &bins[ d ], (uint32_t)__popc( peers ) );
    }
    __syncthreads();
    hist[ (size_t)threadIdx.x * num_tiles + blockIdx.x ] = bins[ threadIdx.x ];
}

// vals_in == nullptr means "value = element index" (first pass of an index sort).
__global__ void __launch_bounds__( RADIX_THREADS ) k_radix_scatter( const uint32_t *__restrict__ keys_in,
                                                                    const uint32_t *__restrict__ vals_in,
                                                                    uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                                                                    const uint32_t *__restrict__ hist_scanned, const size_t n,
                                                                    const unsigned num_tiles, const int shift ) {
    __shared__ uint32_t wcount[ RADIX_WARPS ][ RADIX_BINS ];
    const unsigned      lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for ( int q = threadIdx.x; q < RADIX_WARPS * RADIX_BINS; q += RADIX_THREADS )
        ( &wcount[ 0 ][ 0 ] )[ q ] = 0;
    __syncthreads();

    // Tile order: warp w owns elements [w*256, (w+1)*256) of the tile, visited in 8 rounds of 32 lanes.
    const size_t base = (size_t)blockIdx.x * RADIX_TILE + (size_t)warp * ( 32 * RADIX_ITEMS );
    uint32_t     key[ RADIX_ITEMS ], val[ RADIX_ITEMS ], rank[ RADIX_ITEMS ];
#pragma unroll
    for ( int r = 0; r < RADIX_ITEMS; r++ ) {
        const size_t idx   = base + (size_t)r * 32 + lane;
        const bool   valid = idx < n;
        key[ r ]           = valid ? keys_in[ idx ] : 0u;
        val[ r ]           = valid ? ( vals_in ? vals_in[ idx ] : (uint32_t)idx ) : 0u;
        const uint32_t d   = valid ? ( ( key[ r ] >> shift ) & ( RADIX_BINS - 1 ) ) : ( 0x10000u | lane );
        const unsigned peers  = __match_any_sync( 0xffffffffu, d );
        const unsigned leader = (unsigned)( __ffs( peers ) - 1 );
        uint32_t       prev   = 0;
        if ( valid && lane == leader ) {   // only this warp touches wcount[warp][*]; one leader per digit per round
            prev                = wcount[ warp ][ d ];
            wcount[ warp ][ d ] = prev + (uint32_t)__popc( peers );
        }
        prev      = __shfl_sync( peers, prev, leader );
        rank[ r ] = prev + (uint32_t)__popc( peers & ( ( 1u << lane ) - 1u ) );
        __syncwarp();
    }
    __syncthreads();
    {   // thread d: exclusive prefix of digit d over the warps, seeded with the global base of (d, tile)
        const unsigned d   = threadIdx.x;
        uint32_t       run = hist_scanned[ (size_t)d * num_tiles + blockIdx.x ];
#pragma unroll
        for ( int w = 0; w < RADIX_WARPS; w++ ) {
            const uint32_t c = wcount[ w ][ d ];
            wcount[ w ][ d ] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for ( int r = 0; r < RADIX_ITEMS; r++ ) {
        const size_t idx = base + (size_t)r * 32 + lane;
        if ( idx < n ) {
            const uint32_t d    = ( key[ r ] >> shift ) & ( RADIX_BINS - 1 );
            const uint32_t dest = wcount[ warp ][ d ] + rank[ r ];
            keys_out[ dest ]    = key[ r ];
            vals_out[ dest ]    = val[ r ];
        }
    }
}

static inline size_t radix_scratch_words( size_t n ) {
    const size_t hist = (size_t)RADIX_BINS * div_up( n, RADIX_TILE );
    return hist + scan_scratch_words( hist ) + 8;
}

static inline int radix_passes( int key_bits ) { return key_bits <= 0 ? 1 : ( key_bits + RADIX_BITS - 1 ) / RADIX_BITS; }

// Stable sort of n (key, value) pairs on the low key_bits bits.  values_in may be nullptr (value = index).
// Ping-pongs between (keys_a, vals_a) and (keys_b, vals_b); keys_a holds the input and is overwritten.
// Returns 0 if the result ends in the a buffers, 1 if in the b buffers.
static inline int radix_sort_pairs( uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, bool iota_values,
                                    size_t n, int key_bits, uint32_t *scratch, cudaStream_t s ) {
    if ( n == 0 )
        return 0;
    const unsigned tiles     = div_up( n, RADIX_TILE );
    const size_t   hist_len  = (size_t)RADIX_BINS * tiles;
    uint32_t      *hist      = scratch;
    uint32_t      *scan_tmp  = scratch + hist_len;
    const int      passes    = radix_passes( key_bits );
    uint32_t      *kin = keys_a, *vin = vals_a, *kout = keys_b, *vout = vals_b;
    for ( int p = 0; p < passes; p++ ) {
        const int shift = p * RADIX_BITS;
        PPC_LAUNCH( k_radix_hist, tiles, RADIX_THREADS, 0, s, kin, hist, n, tiles, shift );
        exclusive_scan( hist, hist, hist_len, scan_tmp, s );
        PPC_LAUNCH( k_radix_scatter, tiles, RADIX_THREADS, 0, s, kin, ( p == 0 && iota_values ) ? (const uint32_t *)nullptr : vin, kout,
                    vout, hist, n, tiles, shift );
        uint32_t *t;
        t = kin, kin = kout, kout = t;
        t = vin, vin = vout, vout = t;
    }
    return ( passes & 1 );
}

}   // namespace ppc
