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
