/*
 * slab_multi_gpu.cu -- the slab decomposition driven from C++ through the C ABI alone (no Python, no torch, no NCCL):
 * one process, one slab store per slab, slab k on GPU k % (number of GPUs); the exchange buffers move with
 * cudaMemcpyPeerAsync on the receiving store's stream.  (bench.py does the same with one PROCESS per GPU and NCCL send/recv; the C calls are identical.)
 *
 * The field is also run whole on GPU 0 through the reference's own entry points, and the two results are compared
 * particle by particle: the program exits 0 only if every position and velocity is bit-identical.
 *
 *   nvcc -O2 -std=c++17 -Iinclude examples/slab_multi_gpu.cu -Lparticlephysicsc_b200/lib -lparticlephysicsc_b200 \
 *        -Xlinker -rpath -Xlinker $PWD/particlephysicsc_b200/lib -o slab_multi_gpu && ./slab_multi_gpu [slabs] [particles] [substeps]
 */
#include "ppc_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK( call )                                                                                     \
    do {                                                                                               \
        cudaError_t e__ = ( call );                                                                    \
        if ( e__ != cudaSuccess ) {                                                                    \
            fprintf( stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString( e__ ), __FILE__, __LINE__ ); \
            exit( 2 );                                                                                 \
        }                                                                                              \
    } while ( 0 )

static uint64_t g_state = 0x5EED0007ull;
static uint64_t next_u64() {
    uint64_t z = ( g_state += 0x9E3779B97F4A7C15ull );
    z          = ( z ^ ( z >> 30 ) ) * 0xBF58476D1CE4E5B9ull;
    z          = ( z ^ ( z >> 27 ) ) * 0x94D049BB133111EBull;
    return z ^ ( z >> 31 );
}
static float unit() { return (float)( next_u64() >> 40 ) * ( 1.0f / 16777216.0f ); }

struct Slab {
    struct Particles_t p;
    int                device;
    void              *send[ 2 ] = { nullptr, nullptr }, *recv[ 2 ] = { nullptr, nullptr };   // [0] lower rows, [1] upper rows
    size_t             cap[ 4 ]  = { 0, 0, 0, 0 };
    uint64_t           counts[ 2 ] = { 0, 0 };
};

static void ensure( Slab &s, void **buf, size_t &cap, size_t records ) {
    if ( records <= cap )
        return;
    CK( cudaSetDevice( s.device ) );
    if ( *buf )
        CK( cudaFree( *buf ) );
    cap = records + records / 4 + 1024;
    CK( cudaMalloc( buf, cap * PPC_SLAB_RECORD_BYTES ) );
}

int main( int argc, char **argv ) {
    const int    slabs = argc > 1 ? atoi( argv[ 1 ] ) : 4;
    const size_t n     = argc > 2 ? (size_t)atol( argv[ 2 ] ) : 400000;
    const int    steps = argc > 3 ? atoi( argv[ 3 ] ) : 8;
    const float  radius = 2.0f, phi = 0.30f, dt = 1.0f / 144.0f / 8.0f;
    const int    halo = 8;
    int          gpus = 0;
    CK( cudaGetDeviceCount( &gpus ) );
    const uint16_t W = (uint16_t)std::lround( std::sqrt( (double)n * 3.14159265 * radius * radius / phi ) ), H = W;

    // the field: uniform random, moving
    std::vector<float> px( n ), py( n ), vx( n ), vy( n ), mass( n ), rad( n, radius );
    for ( size_t i = 0; i < n; i++ ) {
        px[ i ] = radius + unit() * ( W - 2 * radius ), py[ i ] = radius + unit() * ( H - 2 * radius );
        vx[ i ] = ( unit() - 0.5f ) * 400.0f, vy[ i ] = ( unit() - 0.5f ) * 400.0f;
        mass[ i ] = (float)( 5 + next_u64() % 6 );
    }

    // ---- the whole field on GPU 0, through the reference's entry points ----
    ppc_set_device( 0 );
    struct Particles_t whole = particlesCreate( n );
    ppc_set_sync_policy( &whole, PPC_SYNC_MANUAL );
    memcpy( whole.p_x, px.data(), n * 4 ), memcpy( whole.p_y, py.data(), n * 4 ), memcpy( whole.v_x, vx.data(), n * 4 );
    memcpy( whole.v_y, vy.data(), n * 4 ), memcpy( whole.mass, mass.data(), n * 4 ), memcpy( whole.radius, rad.data(), n * 4 );
    whole.length = n;
    ppc_host_modified( &whole );
    for ( int s = 0; s < steps; s++ ) {
        particlesUpdate( &whole, dt );
        particlesCollisionsCheck_Grid( &whole, W, H, dt );
    }
    ppc_download( &whole, PPC_DL_POS | PPC_DL_VEL );

    // ---- the same field as slabs of cell rows ----
    ppc_grid_info grid;
    if ( ppc_grid_dims( W, H, radius, &grid ) != 0 || grid.rows < slabs ) {
        fprintf( stderr, "cannot cut %d cell rows into %d slabs\n", grid.rows, slabs );
        return 2;
    }
    std::vector<Slab> S( slabs );
    for ( int k = 0; k < slabs; k++ ) {
        const int row_begin = grid.rows * k / slabs, row_end = grid.rows * ( k + 1 ) / slabs;
        std::vector<uint32_t> ids;
        for ( size_t i = 0; i < n; i++ ) {   // ownership by the cell row of the initial position (the reference's own division, src/particles_solver.c:286-289)
            int cy = (int)( py[ i ] / grid.cell_size );
            cy     = std::min( std::max( cy, 0 ), grid.rows - 1 );
            if ( cy >= row_begin && cy < row_end )
                ids.push_back( (uint32_t)i );
        }
        S[ k ].device = k % gpus;
        ppc_set_device( S[ k ].device );
        S[ k ].p = particlesCreate( ids.size() * 2 + 65536 );
        for ( size_t t = 0; t < ids.size(); t++ ) {
            const uint32_t i = ids[ t ];
            S[ k ].p.p_x[ t ] = px[ i ], S[ k ].p.p_y[ t ] = py[ i ], S[ k ].p.v_x[ t ] = vx[ i ], S[ k ].p.v_y[ t ] = vy[ i ];
            S[ k ].p.mass[ t ] = mass[ i ], S[ k ].p.radius[ t ] = rad[ i ];
        }
        S[ k ].p.length = ids.size();
        if ( ppc_slab_init( &S[ k ].p, ids.data(), W, H, radius, row_begin, row_end, halo ) != 0 ) {
            fprintf( stderr, "ppc_slab_init failed for slab %d\n", k );
            return 2;
        }
    }
    for ( int s = 0; s < steps; s++ ) {
        for ( auto &sl : S ) {   // integrate + walls, count what goes to each neighbour
            if ( ppc_slab_begin( &sl.p, dt, 0, sl.counts ) != 0 ) {
                fprintf( stderr, "halo of %d rows is too thin for this field\n", halo );
                return 3;
            }
            ensure( sl, &sl.send[ 0 ], sl.cap[ 0 ], sl.counts[ 0 ] ), ensure( sl, &sl.send[ 1 ], sl.cap[ 1 ], sl.counts[ 1 ] );
            ppc_slab_pack( &sl.p, sl.send[ 0 ], sl.send[ 1 ], 0 );
        }
        for ( auto &sl : S )
            ppc_synchronize( &sl.p );
        for ( int k = 0; k < slabs; k++ ) {   // slab k receives what k-1 sent upwards and what k+1 sent downwards
            Slab          &sl = S[ k ];
            const uint64_t from_lo = k > 0 ? S[ k - 1 ].counts[ 1 ] : 0, from_hi = k + 1 < slabs ? S[ k + 1 ].counts[ 0 ] : 0;
            ensure( sl, &sl.recv[ 0 ], sl.cap[ 2 ], from_lo ), ensure( sl, &sl.recv[ 1 ], sl.cap[ 3 ], from_hi );
            // on the receiving store's own stream (ppc_stream), so that ppc_slab_finish is ordered behind the copies; the senders
            // were synchronised above
            cudaStream_t stream = (cudaStream_t)ppc_stream( &sl.p );
            if ( from_lo )
                CK( cudaMemcpyPeerAsync( sl.recv[ 0 ], sl.device, S[ k - 1 ].send[ 1 ], S[ k - 1 ].device, from_lo * PPC_SLAB_RECORD_BYTES, stream ) );
            if ( from_hi )
                CK( cudaMemcpyPeerAsync( sl.recv[ 1 ], sl.device, S[ k + 1 ].send[ 0 ], S[ k + 1 ].device, from_hi * PPC_SLAB_RECORD_BYTES, stream ) );
            ppc_slab_finish( &sl.p, dt, sl.recv[ 0 ], from_lo, sl.recv[ 1 ], from_hi, 0, nullptr );
        }
    }

    // ---- compare ----
    size_t seen = 0, bad = 0;
    for ( auto &sl : S ) {
        std::vector<uint32_t> ids( sl.p.size );
        if ( ppc_slab_download( &sl.p, ids.data(), PPC_DL_POS | PPC_DL_VEL | PPC_DL_IDS ) != 0 ) {
            fprintf( stderr, "a slab reports that its halo was too thin\n" );
            return 3;
        }
        for ( size_t t = 0; t < sl.p.length; t++, seen++ ) {
            const uint32_t i = ids[ t ];
            bad += memcmp( &sl.p.p_x[ t ], &whole.p_x[ i ], 4 ) != 0 || memcmp( &sl.p.p_y[ t ], &whole.p_y[ i ], 4 ) != 0 ||
                   memcmp( &sl.p.v_x[ t ], &whole.v_x[ i ], 4 ) != 0 || memcmp( &sl.p.v_y[ t ], &whole.v_y[ i ], 4 ) != 0;
        }
        ppc_slab_info_t info;
        ppc_slab_info( &sl.p, &info );
        printf( "slab rows [%d, %d) on GPU %d: %llu particles, inexactness mark travelled %u of %d halo rows\n", info.own_row_begin, info.own_row_end,
                sl.device, (unsigned long long)info.n_owned, info.taint_rows, info.halo_rows );
    }
    printf( "%zu particles over %d slabs on %d GPU(s), %d substeps: %zu owned in total, %zu differ from the single-GPU run\n", n, slabs, gpus, steps, seen, bad );
    for ( auto &sl : S ) {
        CK( cudaSetDevice( sl.device ) );
        for ( int d = 0; d < 2; d++ ) {
            cudaFree( sl.send[ d ] );
            cudaFree( sl.recv[ d ] );
        }
        particlesFree( &sl.p );
    }
    particlesFree( &whole );
    return ( seen == n && bad == 0 ) ? 0 : 1;
}
