// ppc_scan.cuh -- hand-written exclusive prefix sum over uint32 (no CUB/Thrust).
//
// Used for the cell-start table (counts per cell -> first sorted slot of each cell; the sorted-range form of the
// reference's per-cell chains, src/particles_solver.c:265-295), for the per-tile digit offsets of the radix sort
// and for the pair-list offsets of the debug exporter.
//
// Three launches: per-tile reduce -> single-CTA scan of the tile sums -> per-tile scan + offset.  Tiles are
// 4096 elements (256 threads x 16 consecutive items, read as 4 x 128-bit loads per thread); HBM traffic is
// 2 reads + 1 write of the array.
#pragma once
#include "ppc_common.cuh"

namespace ppc {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS   = 16;
constexpr int SCAN_TILE    = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_inclusive_scan( uint32_t v ) {
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for ( int d = 1; d < 32; d <<= 1 ) {
        const uint32_t t = __shfl_up_sync( 0xffffffffu, v, d );
        if ( lane >= (unsigned)d )
            v += t;
    }
    return v;
}

// Exclusive scan of one value per thread across a 256-thread CTA.  sh must hold >= 9 words.  Ends with a barrier,
// so sh may be reused immediately.
__device__ __forceinline__ uint32_t block_exclusive_scan( const uint32_t v, uint32_t *sh, uint32_t &total ) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t incl = warp_inclusive_scan( v );
    if ( lane == 31 )
        sh[ warp ] = incl;
    __syncthreads();
    if ( warp == 0 ) {
        const uint32_t w  = lane < ( SCAN_THREADS / 32 ) ? sh[ lane ] : 0u;
        const uint32_t wi = warp_inclusive_scan( w );
        if ( lane < ( SCAN_THREADS / 32 ) )
            sh[ lane ] = wi - w;
        if ( lane == ( SCAN_THREADS / 32 ) - 1 )
            sh[ SCAN_THREADS / 32 ] = wi;
    }
    __syncthreads();
    const uint32_t res = sh[ warp ] + incl - v;
    total              = sh[ SCAN_THREADS / 32 ];
    __syncthreads();
    return res;
}

__device__ __forceinline__ void scan_load_items( const uint32_t *__restrict__ in, const size_t n, const size_t first,
                                                 uint32_t ( &item )[ SCAN_ITEMS ] ) {
    if ( first + SCAN_ITEMS <= n ) {
        const uint4 *p = reinterpret_cast<const uint4 *>( in + first );   // first is a multiple of 16 elements
#pragma unroll
        for ( int q = 0; q < SCAN_ITEMS / 4; q++ ) {
            const uint4 v     = p[ q ];
            item[ 4 * q ]     = v.x;
            item[ 4 * q + 1 ] = v.y;
            item[ 4 * q + 2 ] = v.z;
            item[ 4 * q + 3 ] = v.w;
        }
    } else {
#pragma unroll
        for ( int q = 0; q < SCAN_ITEMS; q++ )
            item[ q ] = ( first + q < n ) ? in[ first + q ] : 0u;
    }
}

__global__ void __launch_bounds__( SCAN_THREADS ) k_scan_reduce( const uint32_t *__restrict__ in, uint32_t *__restrict__ tile_sums,
                                                                 const size_t n ) {
    __shared__ uint32_t sh[ 16 ];
    const size_t        first = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t            item[ SCAN_ITEMS ];
    scan_load_items( in, n, first, item );
    uint32_t s = 0;
#pragma unroll
    for ( int q = 0; q < SCAN_ITEMS; q++ )
        s += item[ q ];
    uint32_t total;
    (void)block_exclusive_scan( s, sh, total );
    if ( threadIdx.x == 0 )
        tile_sums[ blockIdx.x ] = total;
}

// One CTA turns the tile sums into exclusive tile offsets, in place.
__global__ void __launch_bounds__( SCAN_THREADS ) k_scan_tiles( uint32_t *__restrict__ tile_sums, const uint32_t num_tiles ) {
    __shared__ uint32_t sh[ 16 ];
    uint32_t            carry = 0;
    for ( uint32_t base = 0; base < num_tiles; base += SCAN_THREADS ) {
        const uint32_t idx = base + threadIdx.x;
        const uint32_t v   = idx < num_tiles ? tile_sums[ idx ] : 0u;
        uint32_t       total;
        const uint32_t ex = block_exclusive_scan( v, sh, total );
        if ( idx < num_tiles )
            tile_sums[ idx ] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__( SCAN_THREADS ) k_scan_apply( const uint32_t *in, uint32_t *out,
                                                                const uint32_t *__restrict__ tile_offsets, const size_t n ) {
    __shared__ uint32_t sh[ 16 ];
    const size_t        first = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t            item[ SCAN_ITEMS ];
    scan_load_items( in, n, first, item );
    uint32_t s = 0;
#pragma unroll
    for ( int q = 0; q < SCAN_ITEMS; q++ )
        s += item[ q ];
    uint32_t total;
    uint32_t run = block_exclusive_scan( s, sh, total ) + tile_offsets[ blockIdx.x ];
    if ( first + SCAN_ITEMS <= n ) {
        uint4 *p = reinterpret_cast<uint4 *>( out + first );
#pragma unroll
        for ( int q = 0; q < SCAN_ITEMS / 4; q++ ) {
            uint4 v;
            v.x = run;
            run += item[ 4 * q ];
            v.y = run;
            run += item[ 4 * q + 1 ];
            v.z = run;
            run += item[ 4 * q + 2 ];
            v.w = run;
            run += item[ 4 * q + 3 ];
            p[ q ] = v;
        }
    } else {
#pragma unroll
        for ( int q = 0; q < SCAN_ITEMS; q++ ) {
            if ( first + q < n )
                out[ first + q ] = run;
            run += item[ q ];
        }
    }
}

// ---- single-pass variant: decoupled look-back --------------------------------------------------------------------------
// One launch, one read and one write of the array.  CTAs take tile numbers from a ticket counter (so a CTA only ever waits for tiles
// whose CTAs are already running), publish their tile's sum, and find their exclusive prefix by looking back over the published
// descriptors of the preceding tiles, 32 at a time: sums are added up until a tile is met that already knows its inclusive prefix.
// A descriptor is ONE 64-bit word {value, epoch << 2 | state}: written and read whole, so no fence is needed between value and flag,
// and never cleared: the epoch (a counter of launches on this descriptor array) tells this launch's words from stale ones.
constexpr uint32_t SCAN_AGGREGATE = 1u, SCAN_INCLUSIVE = 2u;

__global__ void __launch_bounds__( SCAN_THREADS ) k_scan_lookback( const uint32_t *in, uint32_t *out, const size_t n, unsigned long long *desc,
                                                                   uint32_t *ticket, const uint32_t epoch, const uint32_t tiles ) {
    __shared__ uint32_t sh[ 16 ];
    __shared__ uint32_t s_tile, s_prefix;
    if ( threadIdx.x == 0 )
        s_tile = atomicAdd( ticket, 1u );
    __syncthreads();
    const uint32_t tile  = s_tile;
    const size_t   first = (size_t)tile * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t       item[ SCAN_ITEMS ];
    scan_load_items( in, n, first, item );
    uint32_t s = 0;
#pragma unroll
    for ( int q = 0; q < SCAN_ITEMS; q++ )
        s += item[ q ];
    uint32_t total;
    uint32_t run = block_exclusive_scan( s, sh, total );
    volatile unsigned long long *vd = desc;
    const unsigned long long     tag = (unsigned long long)( epoch << 2 ) << 32;
    if ( threadIdx.x < 32 ) {
        const unsigned lane = threadIdx.x;
        if ( lane == 0 ) {
            vd[ tile ] = tag | ( (unsigned long long)( tile == 0 ? SCAN_INCLUSIVE : SCAN_AGGREGATE ) << 32 ) | total;
            if ( tile == tiles - 1 )
                *ticket = 0u;   // every ticket of this launch has been taken: ready for the next one
        }
        uint32_t prefix = 0;
        if ( tile != 0 ) {
            int base = (int)tile - 1;
            for ( ;; ) {
                const int          p = base - (int)lane;
                unsigned long long d;
                do {   // a tile before the first one counts as "inclusive prefix 0"
                    d = p >= 0 ? vd[ p ] : ( tag | ( (unsigned long long)SCAN_INCLUSIVE << 32 ) );
                } while ( __any_sync( 0xffffffffu, ( d >> 34 ) != (unsigned long long)epoch || ( ( d >> 32 ) & 3u ) == 0u ) );
                const unsigned incl = __ballot_sync( 0xffffffffu, ( ( d >> 32 ) & 3u ) == SCAN_INCLUSIVE );
                const unsigned stop = incl ? (unsigned)( __ffs( (int)incl ) - 1 ) : 31u;   // nearest tile that knows its inclusive prefix
                uint32_t       v    = lane <= stop ? (uint32_t)d : 0u;
#pragma unroll
                for ( int k = 16; k > 0; k >>= 1 )
                    v += __shfl_xor_sync( 0xffffffffu, v, k );
                prefix += v;
                if ( incl )
                    break;
                base -= 32;
            }
            if ( lane == 0 )
                vd[ tile ] = tag | ( (unsigned long long)SCAN_INCLUSIVE << 32 ) | ( prefix + total );
        }
        if ( lane == 0 )
            s_prefix = prefix;
    }
    __syncthreads();
    run += s_prefix;
    if ( first + SCAN_ITEMS <= n ) {
        uint4 *p = reinterpret_cast<uint4 *>( out + first );
#pragma unroll
        for ( int q = 0; q < SCAN_ITEMS / 4; q++ ) {
            uint4 v;
            v.x = run;
            run += item[ 4 * q ];
            v.y = run;
            run += item[ 4 * q + 1 ];
            v.z = run;
            run += item[ 4 * q + 2 ];
            v.w = run;
            run += item[ 4 * q + 3 ];
            p[ q ] = v;
        }
    } else {
#pragma unroll
        for ( int q = 0; q < SCAN_ITEMS; q++ ) {
            if ( first + q < n )
                out[ first + q ] = run;
            run += item[ q ];
        }
    }
}

// desc: div_up(n, SCAN_TILE) 64-bit words, zero-filled when allocated; ticket: one zero-initialised word; epoch: 1, 2, 3 ... per call
// on this descriptor array (below 2^30).
static inline void exclusive_scan_single_pass( const uint32_t *in, uint32_t *out, size_t n, unsigned long long *desc, uint32_t *ticket,
                                               uint32_t epoch, cudaStream_t s ) {
    if ( n == 0 )
        return;
    const unsigned tiles = div_up( n, SCAN_TILE );
    PPC_LAUNCH( k_scan_lookback, tiles, SCAN_THREADS, 0, s, in, out, n, desc, ticket, epoch, tiles );
}

// Number of uint32 words of scratch exclusive_scan needs for n elements.
static inline size_t scan_scratch_words( size_t n ) { return div_up( n, SCAN_TILE ) + 1; }

// out[i] = sum of in[0..i) for i in [0, n).  in and out may alias.  To get the grand total as well, scan n+1
// elements with in[n] = 0.
static inline void exclusive_scan( const uint32_t *in, uint32_t *out, size_t n, uint32_t *scratch, cudaStream_t s ) {
    if ( n == 0 )
        return;
    const unsigned tiles = div_up( n, SCAN_TILE );
    PPC_LAUNCH( k_scan_reduce, tiles, SCAN_THREADS, 0, s, in, scratch, n );
    PPC_LAUNCH( k_scan_tiles, 1, SCAN_THREADS, 0, s, scratch, tiles );
    PPC_LAUNCH( k_scan_apply, tiles, SCAN_THREADS, 0, s, in, out, scratch, n );
}

}   // namespace ppc
