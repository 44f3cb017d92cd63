/*
 * headless_scene.c -- the reference's debug scene (src/main.c:36-38,80-98) without a window.
 *
 * Same calls, same order as the reference's main loop: every frame one random position / mass / radius is added
 * `per_frame` times through particlesAddParticle, then particlesUpdate + particlesCollisionsCheck_Grid run once with a
 * fixed dt = 1/144 (the reference uses the last frame time, capped at 1/60; src/main.c:87-90).  GetRandomValue is replaced
 * by a seeded splitmix64 so runs are reproducible.  Links against libparticlephysicsc_b200.so OR against the reference's own
 * objects: the source uses nothing but the reference's entry points, which is the point of the drop-in.
 *
 *   gcc -O2 -Iinclude examples/headless_scene.c -Lparticlephysicsc_b200/lib -lparticlephysicsc_b200 \
 *       -Wl,-rpath,$PWD/particlephysicsc_b200/lib -o headless_scene && ./headless_scene 4500 10
 */
#include "ppc_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <time.h>

/* The B200 library's frame copy-back moves what the draw loop reads (p_x, p_y, colors); velocities come on request.  The symbol is
 * weak so that this file still links against the reference's own objects, where the host arrays ARE the state. */
extern void ppc_download( struct Particles_t *p, unsigned mask ) __attribute__( ( weak ) );
/* With PPC_RESOLVE=3 in the environment (the fast resolve): how many substeps the library ran on the exact path
 * because a tile did not fit (particles are spawned ten at a time on one spot here) */
extern unsigned long long ppc_gs_fallbacks( struct Particles_t *p ) __attribute__( ( weak ) );

static uint64_t g_state = 0x5EED0001ull;
static uint64_t next_u64( void ) {
    uint64_t z = ( g_state += 0x9E3779B97F4A7C15ull );
    z          = ( z ^ ( z >> 30 ) ) * 0xBF58476D1CE4E5B9ull;
    z          = ( z ^ ( z >> 27 ) ) * 0x94D049BB133111EBull;
    return z ^ ( z >> 31 );
}
static int random_value( int lo, int hi ) { return lo + (int)( next_u64() % (uint64_t)( hi - lo + 1 ) ); } /* inclusive, like GetRandomValue */

int main( int argc, char **argv ) {
    const int      frames = argc > 1 ? atoi( argv[ 1 ] ) : 4500; /* test_limit, src/main.c:37 */
    const int      per_frame = argc > 2 ? atoi( argv[ 2 ] ) : 10; /* test_num_particles, src/main.c:38 */
    const uint16_t W = 16 * 75, H = 9 * 75;                      /* WINDOW_SIZE_FACTOR, includes/constants.h:6 */
    const float    dt = 1.0f / 144.0f;

    struct Particles_t particles = particlesCreate( 50000 ); /* NUM_INIT_SIZE, includes/constants.h:24 */
    struct timespec    t0, t1;
    clock_gettime( CLOCK_MONOTONIC, &t0 );
    double particle_steps = 0;
    for ( int f = 0; f < frames; f++ ) {
        Vector2 pos = { (float)random_value( 0, W ), (float)random_value( 0, H ) }; /* src/main.c:81 */
        int     m = random_value( 5, 10 ), r = random_value( 2, 8 );               /* src/main.c:82-83 */
        particlesAddParticle( &particles, pos, (float)m, (float)r, (uint8_t)per_frame );
        if ( particles.length > 0 ) { /* src/main.c:92-98 */
            particlesUpdate( &particles, dt );
            particlesCollisionsCheck_Grid( &particles, W, H, dt );
            particle_steps += (double)particles.length;
        }
        /* particlesDraw( &particles, tex ) would read p_x, p_y, radius, colors here (src/particles_arrays.c:317-320) */
    }
    clock_gettime( CLOCK_MONOTONIC, &t1 );
    const double secs = ( t1.tv_sec - t0.tv_sec ) + 1e-9 * ( t1.tv_nsec - t0.tv_nsec );
    double       sx = 0, sy = 0, ke = 0;
    if ( ppc_download )
        ppc_download( &particles, PPC_DL_VEL );
    for ( size_t k = 0; k < particles.length; k++ ) {
        sx += particles.p_x[ k ];
        sy += particles.p_y[ k ];
        ke += 0.5 * particles.mass[ k ] * ( (double)particles.v_x[ k ] * particles.v_x[ k ] + (double)particles.v_y[ k ] * particles.v_y[ k ] );
    }
    printf( "frames %d particles %zu size %zu  sum_x %.6f sum_y %.6f ke %.6f  %.3f s  %.3e particle-substeps/s\n", frames, particles.length,
            particles.size, sx, sy, ke, secs, particle_steps / secs );
    if ( ppc_gs_fallbacks && getenv( "PPC_RESOLVE" ) )
        printf( "PPC_RESOLVE=%s: %llu of %d substeps ran on the exact resolve instead\n", getenv( "PPC_RESOLVE" ), ppc_gs_fallbacks( &particles ), frames );
    particlesFree( &particles );
    return 0;
}
