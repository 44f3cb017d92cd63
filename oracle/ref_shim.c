/*
 * ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Companion of oracle/_ref/libref.so (the unmodified reference sources compiled by oracle/Makefile).
 * Loaded with RTLD_GLOBAL *before* libref.so it does two jobs:
 *
 *  1. Interposes particlesResolveVelocity (reference: src/particles_solver.c:58).  In the -fPIC build
 *     the call at src/particles_solver.c:386 goes through the PLT, so this definition receives the
 *     candidate-pair list exactly as the reference built it (src/particles_solver.c:301-383), can copy
 *     it out, optionally permute it (noise-floor experiments) and then forwards
 *     to the real function whose address the loader hands us via ppc_capture_set_real().
 *  2. Defines the five Raylib draw/image calls that src/particles_arrays.c:297-307,337 references, as
 *     aborting stubs, so libref.so loads without Raylib.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <raylib.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct Particles_t;
struct Pair {
    int i, j;
};

typedef void ( *resolve_fn )( struct Particles_t *, struct Pair *, size_t, float );

static resolve_fn   g_real        = NULL;
static int          g_capture     = 0;
static int          g_order_mode  = 0; /* 0 as built, 1 reversed, 2 shuffled */
static uint64_t     g_rng         = 0x9E3779B97F4A7C15ull;
static struct Pair *g_pairs       = NULL;
static size_t       g_pairs_count = 0;
static size_t       g_pairs_cap   = 0;
static size_t       g_calls       = 0;

static uint64_t splitmix64( void ) {
    uint64_t z = ( g_rng += 0x9E3779B97F4A7C15ull );
    z          = ( z ^ ( z >> 30 ) ) * 0xBF58476D1CE4E5B9ull;
    z          = ( z ^ ( z >> 27 ) ) * 0x94D049BB133111EBull;
    return z ^ ( z >> 31 );
}

void   ppc_capture_set_real( void *fn ) { g_real = (resolve_fn)fn; }
void   ppc_capture_enable( int on ) { g_capture = on; }
void   ppc_capture_set_order( int mode, uint64_t seed ) {
    g_order_mode = mode;
    g_rng        = seed;
}
size_t ppc_capture_count( void ) { return g_pairs_count; }
size_t ppc_capture_calls( void ) { return g_calls; }
void   ppc_capture_reset( void ) {
    g_pairs_count = 0;
    g_calls       = 0;
}
/* copies min(count, max_pairs) pairs as int32 (i, j) tuples */
size_t ppc_capture_copy( int32_t *dst, size_t max_pairs ) {
    size_t n = g_pairs_count < max_pairs ? g_pairs_count : max_pairs;
    if ( n )
        memcpy( dst, g_pairs, n * sizeof( struct Pair ) );
    return n;
}

void particlesResolveVelocity( struct Particles_t *particles, struct Pair *collisions, size_t count, float time_step ) {
    g_calls++;
    if ( g_capture ) {
        if ( count > g_pairs_cap ) {
            free( g_pairs );
            g_pairs_cap = count + count / 2 + 16;
            g_pairs     = (struct Pair *)malloc( g_pairs_cap * sizeof( struct Pair ) );
            if ( !g_pairs ) {
                fprintf( stderr, "ref_shim: out of memory capturing %zu pairs\n", count );
                abort();
            }
        }
        if ( count )
            memcpy( g_pairs, collisions, count * sizeof( struct Pair ) );
        g_pairs_count = count;
    }
    if ( g_order_mode == 1 ) {
        for ( size_t a = 0, b = count; a + 1 < b; a++ ) {
            b--;
            struct Pair t   = collisions[ a ];
            collisions[ a ] = collisions[ b ];
            collisions[ b ] = t;
        }
    } else if ( g_order_mode == 2 ) {
        for ( size_t a = count; a > 1; a-- ) {
            size_t      b       = (size_t)( splitmix64() % a );
            struct Pair t       = collisions[ a - 1 ];
            collisions[ a - 1 ] = collisions[ b ];
            collisions[ b ]     = t;
        }
    }
    if ( !g_real ) /* linked (not ctypes-loaded) use: the next definition in load order is the reference's own */
        g_real = (resolve_fn)dlsym( RTLD_NEXT, "particlesResolveVelocity" );
    if ( !g_real ) {
        fprintf( stderr, "ref_shim: real particlesResolveVelocity not bound (call ppc_capture_set_real)\n" );
        abort();
    }
    g_real( particles, collisions, count, time_step );
}

/* --- Raylib symbols referenced by src/particles_arrays.c; never called by the oracle harness --- */
static void no_raylib( const char *what ) {
    fprintf( stderr, "ref_shim: %s called but Raylib is not present in the oracle build\n", what );
    abort();
}
Image GenImageColor( int width, int height, Color color ) {
    (void)width, (void)height, (void)color;
    no_raylib( "GenImageColor" );
    return ( Image ){ 0 };
}
void ImageDrawCircle( Image *dst, int centerX, int centerY, int radius, Color color ) {
    (void)dst, (void)centerX, (void)centerY, (void)radius, (void)color;
    no_raylib( "ImageDrawCircle" );
}
Texture2D LoadTextureFromImage( Image image ) {
    (void)image;
    no_raylib( "LoadTextureFromImage" );
    return ( Texture2D ){ 0 };
}
void UnloadImage( Image image ) {
    (void)image;
    no_raylib( "UnloadImage" );
}
void DrawTextureEx( Texture2D texture, Vector2 position, float rotation, float scale, Color tint ) {
    (void)texture, (void)position, (void)rotation, (void)scale, (void)tint;
    no_raylib( "DrawTextureEx" );
}
