/*
 * ppc_oracle.c -- CPU restatement of the ParticlePhysicsC solver hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by or executed from the
 * product library (particlephysicsc_b200/).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it, and only as the checker or as the
 * CPU figure reported next to the GPU number.
 *
 * Parity pin: the reference ships no tests or golden vectors, so this restatement is
 * pinned against the reference ITSELF: oracle/Makefile compiles the unmodified reference sources into
 * oracle/_ref/libref.so and tests/test_oracle_vs_reference.py checks every function below bit-for-bit
 * against it on seeded inputs; tests/golden/ holds vectors generated from libref.so by
 * tests/golden/make_golden.py so the check also runs where /root/reference is absent.
 *
 * Every function cites the reference lines it restates (paths relative to the reference root).
 * The file is compiled with -ffp-contract=off: the reference Makefile's -std=c11 implies the same for
 * gcc, so no expression below may be fused into an FMA.
 *
 * The data layout is the reference's own: seven independent float/uint8x4 arrays (struct Particles_t,
 * includes/structs.h:9-22).  All functions take the raw arrays so they can be driven from ctypes.
 */
#include "ppc_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* includes/constants.h:10-20.  The literal TYPES matter: the unsuffixed ones are doubles and promote
 * the expressions they appear in. */
#define K_SOLVER_TUNER 0.3f
#define K_EPSILON 1e-8f
#define K_DEPTH_CAP 0.05f
#define K_SLOP_FACTOR 0.01f
#define K_BAUMGARTE 0.2f
#define K_GRAVITY 9.81f
#define K_AIR_RESISTENCE 0.2
#define K_RESTITUTION 0.9
#define K_DAMPING_WALL 0.9
#define K_ENERGY_MIN 1e-4

/* ------------------------------------------------------------------------------------------------
 * Colour: src/utils.c:7-16 (clamp to a byte), src/utils.c:18-27 (clamp_f), src/utils.c:29-45 (hue
 * rotation of RED = {230,41,55}, Raylib literal) as used by src/particles_solver.c:31-39.
 * ---------------------------------------------------------------------------------------------- */
static unsigned char byte_of( float intensity ) {
    if ( intensity < 0 )
        return 0;
    if ( intensity > 255 )
        return 255;
    return (unsigned char)intensity;
}

static void energy_colour( float vx, float vy, float mass, uint8_t *rgba ) {
    /* src/particles_solver.c:31-32 : float speed, then a double expression narrowed to float */
    float speed = vx * vx + vy * vy;
    float t     = (float)( ( (double)( mass * speed ) + K_ENERGY_MIN ) / ( 1000000 - K_ENERGY_MIN ) );
    /* :34 clamp_f(t, 0, 1) -- NaN falls through both comparisons */
    if ( t < 0.0f )
        t = 0.0f;
    else if ( t > 1.0f )
        t = 1.0f;
    /* :35-36 cube root through double pow with a float-rounded exponent */
    t = (float)pow( (double)t, (double)( 1.0f / 3.0f ) );
    /* :38 hue angle, float arithmetic */
    float hue = 240 * ( 1 - t );
    /* src/utils.c:31-35 */
    float arg   = hue * 3.14159265f / 180;
    float c     = (float)cos( (double)arg );
    float s     = (float)sin( (double)arg );
    float r13   = sqrtf( 1.0f / 3.0f );
    float omc   = 1.0f - c;
    float third = 1.0f / 3.0f;
    /* src/utils.c:37-39: three distinct entry values, cyclically arranged */
    float diag = c + omc * third;
    float offm = third * omc - r13 * s;
    float offp = third * omc + r13 * s;
    /* src/utils.c:41-43 with in = RED */
    const float R = 230, G = 41, B = 55;
    rgba[ 0 ] = byte_of( R * diag + G * offm + B * offp );
    rgba[ 1 ] = byte_of( R * offp + G * diag + B * offm );
    rgba[ 2 ] = byte_of( R * offm + G * offp + B * diag );
    rgba[ 3 ] = 255; /* src/particles_solver.c:39 */
}

/* ------------------------------------------------------------------------------------------------
 * particlesUpdate: src/particles_solver.c:16-43
 * ---------------------------------------------------------------------------------------------- */
void ppc_oracle_update( float *px, float *py, float *vx, float *vy, const float *mass, uint8_t *rgba, size_t n, float dt,
                        int do_colour ) {
    const double t  = (double)dt;
    const double ax = 0.5 * 0 * t * t;                 /* :20, kept as written: ((0.5*0)*dt)*dt */
    const double ay = 0.5 * (double)K_GRAVITY * t * t; /* :21 */
    const double dr = 1 - K_AIR_RESISTENCE * t;        /* :24-25 */
    const float  g  = K_GRAVITY * dt;                  /* :27-28, float product */
    for ( size_t k = 0; k < n; k++ ) {
        px[ k ] = (float)( (double)px[ k ] + ( ax + (double)( vx[ k ] * dt ) ) );
        py[ k ] = (float)( (double)py[ k ] + ( ay + (double)( vy[ k ] * dt ) ) );
        vx[ k ] = (float)( (double)vx[ k ] * dr );
        vy[ k ] = (float)( (double)vy[ k ] * dr );
        vx[ k ] += g; /* gravity on x as well: reference behaviour, :27 */
        vy[ k ] += g;
        if ( do_colour && rgba )
            energy_colour( vx[ k ], vy[ k ], mass[ k ], rgba + 4 * k );
    }
}

/* ------------------------------------------------------------------------------------------------
 * Wall pass: src/particles_solver.c:207-228.  All four tests use the values loaded at the top.
 * ---------------------------------------------------------------------------------------------- */
void ppc_oracle_walls( float *px, float *py, float *vx, float *vy, const float *radius, size_t n, uint16_t W, uint16_t H ) {
    for ( size_t k = 0; k < n; k++ ) {
        const float x = px[ k ], y = py[ k ], r = radius[ k ];
        const float xr = (float)(int)W - r, yb = (float)(int)H - r;
        if ( x <= r ) {
            px[ k ] = r;
            vx[ k ] = (float)( (double)vx[ k ] * -K_DAMPING_WALL );
        }
        if ( x >= xr ) {
            px[ k ] = xr;
            vx[ k ] = (float)( (double)vx[ k ] * -K_DAMPING_WALL );
        }
        if ( y <= r ) {
            py[ k ] = r;
            vy[ k ] = (float)( (double)vy[ k ] * -K_DAMPING_WALL );
        }
        if ( y >= yb ) {
            py[ k ] = yb;
            vy[ k ] = (float)( (double)vy[ k ] * -K_DAMPING_WALL );
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Grid geometry: src/particles_solver.c:231-249
 * ---------------------------------------------------------------------------------------------- */
void ppc_oracle_grid_dims( const float *radius, size_t n, uint16_t W, uint16_t H, float *cell_size, int *cols, int *rows ) {
    float big = 6.0f; /* :231 */
    for ( size_t k = 0; k < n; k++ )
        if ( radius[ k ] > big )
            big = radius[ k ];
    const float cs = 2.0f * big; /* :244 */
    *cell_size     = cs;
    *cols          = (int)ceilf( (float)W / cs ); /* :247 */
    *rows          = (int)ceilf( (float)H / cs ); /* :248 */
}

/* src/particles_solver.c:280-291 -- one IEEE division, truncation toward zero, clamp.  The float->int
 * conversion of NaN / out-of-range values is undefined in C; x86 (cvttss2si) yields INT_MIN, which the
 * clamp turns into cell 0.  Stated explicitly here so the oracle does not depend on the compiler. */
static int cell_coord( float p, float cs, int limit ) {
    const float q = p / cs;
    int         c;
    if ( q >= -2147483648.0f && q < 2147483648.0f )
        c = (int)q;
    else
        c = INT32_MIN;
    if ( c < 0 )
        c = 0;
    if ( c >= limit )
        c = limit - 1;
    return c;
}

/* Cell keys (:280-291), then the canonical form of the per-cell chains of :265-295: a stable ascending
 * (key, index) order plus a cell-start table (cells + 1 entries).  Reference chains hold the same
 * members in DESCENDING index order (push-front, :293-294). */
void ppc_oracle_grid_build( const float *px, const float *py, size_t n, float cell_size, int cols, int rows, uint32_t *keys,
                            uint32_t *order, uint32_t *cell_start ) {
    const size_t cells = (size_t)cols * (size_t)rows;
    memset( cell_start, 0, ( cells + 1 ) * sizeof( uint32_t ) );
    for ( size_t k = 0; k < n; k++ ) {
        const int cx = cell_coord( px[ k ], cell_size, cols );
        const int cy = cell_coord( py[ k ], cell_size, rows );
        keys[ k ]    = (uint32_t)( cy * cols + cx ); /* :291 */
        cell_start[ keys[ k ] + 1 ]++;
    }
    for ( size_t c = 0; c < cells; c++ )
        cell_start[ c + 1 ] += cell_start[ c ];
    if ( order ) {
        uint32_t *fill = (uint32_t *)malloc( ( cells ? cells : 1 ) * sizeof( uint32_t ) );
        memcpy( fill, cell_start, cells * sizeof( uint32_t ) );
        for ( size_t k = 0; k < n; k++ )
            order[ fill[ keys[ k ] ]++ ] = (uint32_t)k;
        free( fill );
    }
}

/* ------------------------------------------------------------------------------------------------
 * Pair search: src/particles_solver.c:311-383.  Emits (i, j) in the reference's list order: i ascending,
 * neighbour cells dy-major / dx-minor, members of a cell in descending index.
 * Returns the number of pairs; *out is malloc'd (caller frees with ppc_oracle_free).
 * ---------------------------------------------------------------------------------------------- */
size_t ppc_oracle_pairs( const float *px, const float *py, const float *radius, size_t n, int cols, int rows, const uint32_t *keys,
                         const uint32_t *order, const uint32_t *cell_start, int32_t **out ) {
    size_t   cap = n * 4 < 16 ? 16 : n * 4, cnt = 0;
    int32_t *pairs = (int32_t *)malloc( cap * 2 * sizeof( int32_t ) );
    for ( size_t i = 0; i < n; i++ ) {
        const float xi = px[ i ], yi = py[ i ], ri = radius[ i ];
        const int   cx = (int)( keys[ i ] % (uint32_t)cols ), cy = (int)( keys[ i ] / (uint32_t)cols );
        for ( int oy = -1; oy <= 1; oy++ ) {
            const int ny = cy + oy;
            if ( ny < 0 || ny >= rows )
                continue;
            for ( int ox = -1; ox <= 1; ox++ ) {
                const int nx = cx + ox;
                if ( nx < 0 || nx >= cols )
                    continue;
                const size_t c = (size_t)ny * (size_t)cols + (size_t)nx;
                for ( uint32_t s = cell_start[ c + 1 ]; s-- > cell_start[ c ]; ) {
                    const size_t j = order[ s ];
                    if ( j <= i ) /* :344 */
                        continue;
                    const float rs = ri + radius[ j ];
                    const float dx = xi - px[ j ];
                    if ( dx > rs || dx < -rs ) /* :353 */
                        continue;
                    const float dy = yi - py[ j ];
                    if ( dy > rs || dy < -rs ) /* :356 */
                        continue;
                    const float d2 = dx * dx + dy * dy;
                    if ( d2 <= rs * rs ) { /* :360-362, inclusive */
                        if ( cnt == cap ) {
                            cap *= 2;
                            pairs = (int32_t *)realloc( pairs, cap * 2 * sizeof( int32_t ) );
                        }
                        pairs[ 2 * cnt ]     = (int32_t)i;
                        pairs[ 2 * cnt + 1 ] = (int32_t)j;
                        cnt++;
                    }
                }
            }
        }
    }
    *out = pairs;
    return cnt;
}

void ppc_oracle_free( void *p ) { free( p ); }

/* ------------------------------------------------------------------------------------------------
 * Contact geometry, frozen before anything moves: src/particles_solver.c:84-127 (pass 1).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    float nx, ny;    /* unit normal from j towards i (:96-97) */
    float raw;       /* penetration depth (:103) */
    float corrected; /* clamped share resolved positionally (:106-113) */
    float vbias;     /* Baumgarte velocity term (:116-127) */
} contact_t;

static contact_t contact_of( float xi, float yi, float ri, float xj, float yj, float rj, float dt ) {
    contact_t   c;
    const float dx = xi - xj, dy = yi - yj;
    float       d2 = dx * dx + dy * dy;
    d2             = fmaxf( d2, K_EPSILON ); /* :88 */
    const float d  = sqrtf( d2 );
    if ( d < K_EPSILON ) { /* :92, unreachable because d >= 1e-4 */
        c.nx = 1.0f;
        c.ny = 0.0f;
    } else {
        c.nx = dx / d;
        c.ny = dy / d;
    }
    c.raw       = ( ri + rj ) - d;
    c.corrected = 0.0f;
    c.vbias     = 0.0f;
    if ( c.raw > 0.0f ) {
        const float cap = K_DEPTH_CAP * fminf( ri, rj );
        float       cor = c.raw * K_SOLVER_TUNER;
        cor             = fmaxf( -cap, fminf( cor, cap ) );
        c.corrected     = cor;
        const float residual = c.raw - cor;
        const float slop     = K_SLOP_FACTOR * fmaxf( ri, rj );
        if ( residual > slop )
            c.vbias = K_BAUMGARTE * ( residual - slop ) / dt; /* :121 */
    }
    return c;
}

/* masses are read through uint8_t in pass 2 (:141-142) */
static float inv_mass_u8( float m ) { return 1.0f / (float)(uint8_t)m; }

/* ------------------------------------------------------------------------------------------------
 * particlesResolveVelocity, sequential: src/particles_solver.c:58-195.
 * ---------------------------------------------------------------------------------------------- */
void ppc_oracle_resolve_sequential( float *px, float *py, float *vx, float *vy, const float *mass, const float *radius, size_t n,
                                    const int32_t *pairs, size_t npairs, float dt ) {
    if ( npairs == 0 || n < 2 ) /* :60 */
        return;
    contact_t *ct = (contact_t *)malloc( npairs * sizeof( contact_t ) );
    for ( size_t k = 0; k < npairs; k++ ) {
        const int i = pairs[ 2 * k ], j = pairs[ 2 * k + 1 ];
        ct[ k ] = contact_of( px[ i ], py[ i ], radius[ i ], px[ j ], py[ j ], radius[ j ], dt );
    }
    for ( size_t k = 0; k < npairs; k++ ) {
        const int   i = pairs[ 2 * k ], j = pairs[ 2 * k + 1 ];
        const float wi = inv_mass_u8( mass[ i ] ), wj = inv_mass_u8( mass[ j ] );
        const float wsum = wi + wj;
        const float nx = ct[ k ].nx, ny = ct[ k ].ny;
        if ( ct[ k ].raw > 0.0f ) { /* :145-156 */
            const float fi = ( wi / wsum ) * ct[ k ].corrected;
            const float fj = ( wj / wsum ) * ct[ k ].corrected;
            px[ i ] += fi * nx;
            py[ i ] += fi * ny;
            px[ j ] -= fj * nx;
            py[ j ] -= fj * ny;
        }
        const float rx = vx[ i ] - vx[ j ], ry = vy[ i ] - vy[ j ]; /* :159-166, current velocities */
        const float vn = rx * nx + ry * ny;
        if ( vn > 0.0f && ct[ k ].vbias == 0.0f ) /* :170 */
            continue;
        const float want = (float)( -K_RESTITUTION * (double)vn + (double)ct[ k ].vbias ); /* :175 */
        const float jmag = ( want - vn ) / wsum;                                          /* :180 */
        const float ix = jmag * nx, iy = jmag * ny;
        vx[ i ] += wi * ix; /* :187-190 */
        vy[ i ] += wi * iy;
        vx[ j ] -= wj * ix;
        vy[ j ] -= wj * iy;
    }
    free( ct );
}

/* ------------------------------------------------------------------------------------------------
 * Jacobi model of the B200 narrow phase -- NOT reference behaviour.  Same contact geometry and the
 * same per-pair formulas as above, but every pair reads the velocities as they were when the resolve
 * started, and each particle sums its contributions in the order the reference list would have
 * touched it (pairs where it is j, by ascending i, then pairs where it is i, in list order).  Position
 * pushes therefore equal the sequential ones bit for bit; velocities are the device semantics that the
 * GPU kernel must reproduce exactly (tests/test_gpu_parity.py).
 * ---------------------------------------------------------------------------------------------- */
void ppc_oracle_resolve_jacobi( float *px, float *py, float *vx, float *vy, const float *mass, const float *radius, size_t n,
                                const int32_t *pairs, size_t npairs, float dt ) {
    if ( npairs == 0 || n < 2 )
        return;
    float *vx0 = (float *)malloc( n * sizeof( float ) ), *vy0 = (float *)malloc( n * sizeof( float ) );
    memcpy( vx0, vx, n * sizeof( float ) );
    memcpy( vy0, vy, n * sizeof( float ) );
    /* every pair's terms come from the frozen state; the list order then fixes the summation order */
    contact_t *ct = (contact_t *)malloc( npairs * sizeof( contact_t ) );
    for ( size_t k = 0; k < npairs; k++ ) {
        const int i = pairs[ 2 * k ], j = pairs[ 2 * k + 1 ];
        ct[ k ] = contact_of( px[ i ], py[ i ], radius[ i ], px[ j ], py[ j ], radius[ j ], dt );
    }
    for ( size_t k = 0; k < npairs; k++ ) {
        const int   i = pairs[ 2 * k ], j = pairs[ 2 * k + 1 ];
        const float wi = inv_mass_u8( mass[ i ] ), wj = inv_mass_u8( mass[ j ] );
        const float wsum = wi + wj;
        const float nx = ct[ k ].nx, ny = ct[ k ].ny;
        if ( ct[ k ].raw > 0.0f ) {
            const float fi = ( wi / wsum ) * ct[ k ].corrected;
            const float fj = ( wj / wsum ) * ct[ k ].corrected;
            px[ i ] += fi * nx;
            py[ i ] += fi * ny;
            px[ j ] -= fj * nx;
            py[ j ] -= fj * ny;
        }
        const float rx = vx0[ i ] - vx0[ j ], ry = vy0[ i ] - vy0[ j ]; /* frozen velocities */
        const float vn = rx * nx + ry * ny;
        if ( vn > 0.0f && ct[ k ].vbias == 0.0f )
            continue;
        const float want = (float)( -K_RESTITUTION * (double)vn + (double)ct[ k ].vbias );
        const float jmag = ( want - vn ) / wsum;
        const float ix = jmag * nx, iy = jmag * ny;
        vx[ i ] += wi * ix;
        vy[ i ] += wi * iy;
        vx[ j ] -= wj * ix;
        vy[ j ] -= wj * iy;
    }
    free( ct );
    free( vx0 );
    free( vy0 );
}

/* ------------------------------------------------------------------------------------------------
 * Tile Gauss-Seidel model of the B200 "resolve"=3 path -- NOT reference behaviour.
 *
 * It is the reference's pass 2 (src/particles_solver.c:131-191) with the velocity impulses applied in a DIFFERENT,
 * deterministic pair order; everything else is the reference's: the same candidate pairs, the same frozen contact
 * geometry (pass 1, :84-127), position pushes summed in the reference's list order (so positions equal the
 * sequential ones bit for bit), and every impulse computed by the same formulas from the CURRENT velocities of its
 * two particles (:159-190).  The order is the one the device kernel (csrc/ppc_gs.cuh) executes:
 *
 *   1. The grid is cut into tiles of tile_r cell rows x wc cell columns, coloured (tx & 1) + 2 (ty & 1).  A tile
 *      whose staged particle count (tile + one-cell halo) exceeds strip_limit is cut into column strips, greedily from
 *      the left, each as wide as keeps its own staged count within strip_limit; its CTA runs them one after the other.
 *   2. A pair belongs to ONE sub-tile (tile, strip): the one holding both particles, else the earlier of the two
 *      in the order (colour of the tile; inside a tile the strip index).  Sub-tiles run in that order; tiles of one
 *      colour touch disjoint particles, so their mutual order is immaterial.
 *   3. Inside a sub-tile pairs are VISITED grouped by the pair of cells they connect.  Nine phases: same cell; east
 *      neighbour, base column even / odd; south, south-east, south-west neighbour, base row even / odd (base =
 *      the smaller cell in (row, column) order; parities of the grid-global coordinates).  The groups of a phase
 *      share no particle.
 *   4. Inside a group pairs are ordered by their ANCHOR particle (the one inside the sub-tile if the other is in
 *      its halo, else the one in the base cell, else -- same cell -- the lower index), ascending index, and then by
 *      the pair's position in the anchor's contact list (the order the reference's pair list touches that particle:
 *      pairs in which it is j by ascending i, then pairs in which it is i in list order).
 *   5. Visited in that order, every pair takes the lowest pair-colour (0..31) not yet used at either of its particles by
 *      an earlier pair of the sub-tile; the sub-tile's pairs RUN colour by colour (greedy edge colouring: ~8 dependent steps
 *      per dense sub-tile instead of the ~14 levels that list scheduling of the visiting order itself needs).
 *
 * tests/test_gpu_gs.py checks the device against this model bit for bit, and this model against the reference
 * statistically (the north-star's 1 % bar, with the reference's own reordering spread beside it).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t sub;    /* colour << 28 | strip (the tile itself is immaterial: see 2. above, but kept for locality) */
    uint32_t tile;
    uint32_t phase;
    uint32_t base;   /* global key of the base cell */
    uint32_t anchor; /* index of the anchor particle */
    uint32_t rank;   /* position in the anchor's contact list */
    uint32_t k;      /* index into the pair list */
    uint32_t colour; /* pair-colour inside the sub-tile (5. below), filled in after the first sort */
    uint32_t visit;  /* position in the visiting order */
} gs_key_t;

static int gs_key_cmp( const void *pa, const void *pb ) {
    const gs_key_t *a = (const gs_key_t *)pa, *b = (const gs_key_t *)pb;
#define GS_CMP( f )                                                                                                  \
    if ( a->f != b->f )                                                                                              \
        return a->f < b->f ? -1 : 1;
    GS_CMP( sub ) GS_CMP( tile ) GS_CMP( phase ) GS_CMP( base ) GS_CMP( anchor ) GS_CMP( rank )
    return 0;
}

static int gs_colour_cmp( const void *pa, const void *pb ) {
    const gs_key_t *a = (const gs_key_t *)pa, *b = (const gs_key_t *)pb;
    GS_CMP( sub ) GS_CMP( tile ) GS_CMP( colour ) GS_CMP( visit )
#undef GS_CMP
    return 0;
}

void ppc_oracle_tilegs_order( const int32_t *pairs, size_t npairs, size_t n, const uint32_t *keys, const uint32_t *cell_start, int cols,
                              int rows, int row0, int wc, int tile_r, int strip_limit, uint32_t *perm_out ) {
    if ( npairs == 0 )
        return;
    const int tiles_x = ( cols + wc - 1 ) / wc;
    const int ty0 = row0 / tile_r, tiles_y = ( row0 + rows + tile_r - 1 ) / tile_r - ty0;
    /* strip of every tile column (1. above): greedily from the left, each strip as wide as keeps its staged particle count
     * (strip + one halo column either side, tile rows + one halo row either side) within strip_limit, at least one column */
    uint8_t *strip_of = (uint8_t *)calloc( (size_t)tiles_x * tiles_y * wc, 1 );
    for ( int ty = 0; ty < tiles_y; ty++ )
        for ( int tx = 0; tx < tiles_x; tx++ ) {
            const int cx0 = tx * wc, cx1 = cx0 + wc < cols ? cx0 + wc : cols;
            const int c0 = cx0 > 0 ? cx0 - 1 : 0, c1 = cx1 + 1 < cols ? cx1 + 1 : cols;
            const int cy0 = ( ty0 + ty ) * tile_r - row0;
            uint32_t  P[ 70 ]; /* wc + 3 <= 67 */
            for ( int c = 0; c <= c1 - c0; c++ ) {
                P[ c ] = 0;
                for ( int gy = cy0 - 1; gy <= cy0 + tile_r; gy++ )
                    if ( gy >= 0 && gy < rows )
                        P[ c ] += cell_start[ (size_t)gy * cols + c0 + c ];
            }
            const uint32_t total = P[ c1 - c0 ] - P[ 0 ];
            uint8_t       *so    = strip_of + ( (size_t)ty * tiles_x + tx ) * wc;
            int            strip = 0;
            for ( int sx = cx0; sx < cx1; strip++ ) {
                int ex = cx1;
                if ( total > (uint32_t)strip_limit ) {
                    ex = sx + 1;
                    while ( ex < cx1 && P[ ( ex + 2 < c1 ? ex + 2 : c1 ) - c0 ] - P[ ( sx - 1 > c0 ? sx - 1 : c0 ) - c0 ] <= (uint32_t)strip_limit )
                        ex++;
                }
                for ( int c = sx; c < ex; c++ )
                    so[ c - cx0 ] = (uint8_t)strip;
                sx = ex;
            }
        }
    /* contact-list ranks (4. above) */
    uint32_t *cnt_j = (uint32_t *)calloc( n ? n : 1, sizeof( uint32_t ) ), *first_i = (uint32_t *)malloc( ( n ? n : 1 ) * sizeof( uint32_t ) );
    uint32_t *rank_i = (uint32_t *)malloc( npairs * sizeof( uint32_t ) ), *rank_j = (uint32_t *)malloc( npairs * sizeof( uint32_t ) );
    for ( size_t k = 0; k < npairs; k++ )
        rank_j[ k ] = cnt_j[ pairs[ 2 * k + 1 ] ]++; /* the list is ascending in i, so this counts the pairs (i' < i, j) */
    for ( size_t k = 0; k < npairs; k++ ) {
        const int i = pairs[ 2 * k ];
        if ( k == 0 || pairs[ 2 * ( k - 1 ) ] != i )
            first_i[ i ] = (uint32_t)k;
        rank_i[ k ] = cnt_j[ i ] + ( (uint32_t)k - first_i[ i ] );
    }
    gs_key_t *K = (gs_key_t *)malloc( npairs * sizeof( gs_key_t ) );
    for ( size_t k = 0; k < npairs; k++ ) {
        const int      i = pairs[ 2 * k ], j = pairs[ 2 * k + 1 ];
        const int      cxi = (int)( keys[ i ] % (uint32_t)cols ), cyi = (int)( keys[ i ] / (uint32_t)cols );
        const int      cxj = (int)( keys[ j ] % (uint32_t)cols ), cyj = (int)( keys[ j ] / (uint32_t)cols );
        const int      txi = cxi / wc, tyi = ( cyi + row0 ) / tile_r - ty0, txj = cxj / wc, tyj = ( cyj + row0 ) / tile_r - ty0;
        const uint32_t ti = (uint32_t)( tyi * tiles_x + txi ), tj = (uint32_t)( tyj * tiles_x + txj );
        const uint32_t si = strip_of[ (size_t)ti * wc + ( cxi - txi * wc ) ], sj = strip_of[ (size_t)tj * wc + ( cxj - txj * wc ) ];
        const uint32_t ci = (uint32_t)( ( txi & 1 ) + 2 * ( ( tyi + ty0 ) & 1 ) ), cj = (uint32_t)( ( txj & 1 ) + 2 * ( ( tyj + ty0 ) & 1 ) );
        /* owner sub-tile (2.) and whether each particle lies inside it */
        int owner_is_i;
        if ( ti == tj )
            owner_is_i = si <= sj;
        else
            owner_is_i = ci < cj;
        const uint32_t ot = owner_is_i ? ti : tj, os = owner_is_i ? si : sj, oc = owner_is_i ? ci : cj;
        const int      i_in = ( ti == ot && si == os ), j_in = ( tj == ot && sj == os );
        /* cells, base cell, phase (3.) */
        const int gyi = cyi + row0, gyj = cyj + row0;
        const int i_base = ( gyi < gyj ) || ( gyi == gyj && cxi <= cxj );
        const int by = i_base ? gyi : gyj, bx = i_base ? cxi : cxj, oy = i_base ? gyj : gyi, ox = i_base ? cxj : cxi;
        uint32_t  phase;
        if ( oy == by && ox == bx )
            phase = 0;
        else if ( oy == by )
            phase = 1 + (uint32_t)( bx & 1 );
        else if ( ox == bx )
            phase = 3 + (uint32_t)( by & 1 );
        else if ( ox == bx + 1 )
            phase = 5 + (uint32_t)( by & 1 );
        else
            phase = 7 + (uint32_t)( by & 1 );
        /* anchor (4.) */
        int anchor_is_i;
        if ( i_in && j_in ) {
            if ( gyi == gyj && cxi == cxj )
                anchor_is_i = 1; /* same cell: the lower index, and i < j */
            else
                anchor_is_i = i_base;
        } else
            anchor_is_i = i_in;
        K[ k ].sub    = ( oc << 28 ) | os;
        K[ k ].tile   = ot;
        K[ k ].phase  = phase;
        K[ k ].base   = (uint32_t)( by - row0 ) * (uint32_t)cols + (uint32_t)bx;
        K[ k ].anchor = (uint32_t)( anchor_is_i ? i : j );
        K[ k ].rank   = anchor_is_i ? rank_i[ k ] : rank_j[ k ];
        K[ k ].k      = (uint32_t)k;
    }
    qsort( K, npairs, sizeof( gs_key_t ), gs_key_cmp );
    if ( getenv( "PPC_ORACLE_GS_STATS" ) ) { /* diagnostics: depth of the levelled schedule vs a greedy edge colouring, per sub-tile */
        uint32_t *lvl = (uint32_t *)calloc( n, sizeof( uint32_t ) ), *used = (uint32_t *)calloc( n, sizeof( uint32_t ) );
        double    sum_l = 0, sum_c = 0, sum_p = 0;
        uint32_t  max_l = 0, max_c = 0, groups = 0;
        for ( size_t k = 0; k < npairs; ) {
            size_t e = k;
            while ( e < npairs && K[ e ].sub == K[ k ].sub && K[ e ].tile == K[ k ].tile )
                e++;
            uint32_t gl = 0, gc = 0;
            for ( size_t q = k; q < e; q++ ) {
                const int i = pairs[ 2 * K[ q ].k ], j = pairs[ 2 * K[ q ].k + 1 ];
                const uint32_t l = ( lvl[ i ] > lvl[ j ] ? lvl[ i ] : lvl[ j ] ) + 1;
                lvl[ i ] = lvl[ j ] = l;
                gl = l > gl ? l : gl;
                uint32_t m = used[ i ] | used[ j ], c = 0;
                while ( c < 32 && ( ( m >> c ) & 1u ) )
                    c++;
                used[ i ] |= 1u << c, used[ j ] |= 1u << c;
                gc = c + 1 > gc ? c + 1 : gc;
            }
            for ( size_t q = k; q < e; q++ ) {
                const int i = pairs[ 2 * K[ q ].k ], j = pairs[ 2 * K[ q ].k + 1 ];
                lvl[ i ] = lvl[ j ] = 0, used[ i ] = used[ j ] = 0;
            }
            sum_l += gl, sum_c += gc, sum_p += (double)( e - k ), groups++;
            max_l = gl > max_l ? gl : max_l, max_c = gc > max_c ? gc : max_c;
            k = e;
        }
        fprintf( stderr, "tilegs stats: %u sub-tiles with pairs, %.1f pairs each; levels mean %.2f max %u; greedy colours mean %.2f max %u\n", groups,
                 sum_p / groups, sum_l / groups, max_l, sum_c / groups, max_c );
        free( lvl );
        free( used );
    }
    /* 5. pair-colours: in the visiting order just established, every pair of a sub-tile takes the lowest colour (0..31) that no
     *    earlier pair of that sub-tile has used at either of its particles; the sub-tile's pairs then run colour by colour (the
     *    pairs of one colour share no particle: their mutual order is immaterial, kept as visited).  A pair that finds all 32
     *    colours taken keeps colour 32: the device flags such a sub-tile and runs the substep on the exact resolve instead. */
    {
        uint32_t *used = (uint32_t *)calloc( n ? n : 1, sizeof( uint32_t ) );
        for ( size_t k = 0; k < npairs; ) {
            size_t e = k;
            while ( e < npairs && K[ e ].sub == K[ k ].sub && K[ e ].tile == K[ k ].tile )
                e++;
            for ( size_t q = k; q < e; q++ ) {
                const int      i = pairs[ 2 * K[ q ].k ], j = pairs[ 2 * K[ q ].k + 1 ];
                const uint32_t m = used[ i ] | used[ j ];
                uint32_t       c = 0;
                while ( c < 32 && ( ( m >> c ) & 1u ) )
                    c++;
                if ( c < 32 )
                    used[ i ] |= 1u << c, used[ j ] |= 1u << c;
                K[ q ].colour = c;
                K[ q ].visit  = (uint32_t)q;
            }
            for ( size_t q = k; q < e; q++ )
                used[ pairs[ 2 * K[ q ].k ] ] = used[ pairs[ 2 * K[ q ].k + 1 ] ] = 0;
            k = e;
        }
        free( used );
        if ( !getenv( "PPC_ORACLE_GS_VISITING_ORDER" ) ) /* (diagnostics: run the pairs in the visiting order itself, as round 2 first did) */
            qsort( K, npairs, sizeof( gs_key_t ), gs_colour_cmp );
    }
    for ( size_t k = 0; k < npairs; k++ )
        perm_out[ k ] = K[ k ].k;
    free( K );
    free( rank_i );
    free( rank_j );
    free( cnt_j );
    free( first_i );
    free( strip_of );
}

/* Positions in the reference's list order, velocities in the order perm[] (a permutation of the pair list). */
void ppc_oracle_resolve_permuted( float *px, float *py, float *vx, float *vy, const float *mass, const float *radius, size_t n,
                                  const int32_t *pairs, size_t npairs, float dt, const uint32_t *perm ) {
    if ( npairs == 0 || n < 2 )
        return;
    contact_t *ct = (contact_t *)malloc( npairs * sizeof( contact_t ) );
    for ( size_t k = 0; k < npairs; k++ ) {
        const int i = pairs[ 2 * k ], j = pairs[ 2 * k + 1 ];
        ct[ k ] = contact_of( px[ i ], py[ i ], radius[ i ], px[ j ], py[ j ], radius[ j ], dt );
    }
    for ( size_t k = 0; k < npairs; k++ ) { /* :145-156 in list order */
        const int   i = pairs[ 2 * k ], j = pairs[ 2 * k + 1 ];
        const float wi = inv_mass_u8( mass[ i ] ), wj = inv_mass_u8( mass[ j ] );
        const float wsum = wi + wj;
        if ( ct[ k ].raw > 0.0f ) {
            const float fi = ( wi / wsum ) * ct[ k ].corrected;
            const float fj = ( wj / wsum ) * ct[ k ].corrected;
            px[ i ] += fi * ct[ k ].nx;
            py[ i ] += fi * ct[ k ].ny;
            px[ j ] -= fj * ct[ k ].nx;
            py[ j ] -= fj * ct[ k ].ny;
        }
    }
    for ( size_t q = 0; q < npairs; q++ ) { /* :158-190 in the permuted order */
        const size_t k = perm[ q ];
        const int    i = pairs[ 2 * k ], j = pairs[ 2 * k + 1 ];
        const float  wi = inv_mass_u8( mass[ i ] ), wj = inv_mass_u8( mass[ j ] );
        const float  wsum = wi + wj;
        const float  nx = ct[ k ].nx, ny = ct[ k ].ny;
        const float  rx = vx[ i ] - vx[ j ], ry = vy[ i ] - vy[ j ];
        const float  vn = rx * nx + ry * ny;
        if ( vn > 0.0f && ct[ k ].vbias == 0.0f )
            continue;
        const float want = (float)( -K_RESTITUTION * (double)vn + (double)ct[ k ].vbias );
        const float jmag = ( want - vn ) / wsum;
        const float ix = jmag * nx, iy = jmag * ny;
        vx[ i ] += wi * ix;
        vy[ i ] += wi * iy;
        vx[ j ] -= wj * ix;
        vy[ j ] -= wj * iy;
    }
    free( ct );
}

static int g_gs_wc = 16, g_gs_tile_r = 8, g_gs_strip_limit = 921;
void       ppc_oracle_set_tilegs( int wc, int tile_r, int strip_limit ) {
    g_gs_wc          = wc;
    g_gs_tile_r      = tile_r;
    g_gs_strip_limit = strip_limit;
}

/* ------------------------------------------------------------------------------------------------
 * particlesCollisionsCheck_Grid: src/particles_solver.c:197-393, composed from the pieces above.
 * mode 0 = sequential resolve (reference), 1 = Jacobi model, 2 = no resolve (walls + grid only),
 * 3 = tile Gauss-Seidel model (parameters from ppc_oracle_set_tilegs).
 * Optional outputs (may be NULL): keys[n], order[n], cell_start[cells+1] (caller sizes it with
 * ppc_oracle_grid_dims first), *pairs_out / *npairs_out.
 * ---------------------------------------------------------------------------------------------- */
void ppc_oracle_collisions_check_grid( float *px, float *py, float *vx, float *vy, const float *mass, const float *radius, size_t n,
                                       uint16_t W, uint16_t H, float dt, int mode, uint32_t *keys_out, uint32_t *order_out,
                                       uint32_t *cell_start_out, int32_t **pairs_out, size_t *npairs_out ) {
    if ( npairs_out )
        *npairs_out = 0;
    if ( pairs_out )
        *pairs_out = NULL;
    if ( n < 2 ) /* :204-205: nothing at all happens, not even the wall pass */
        return;
    ppc_oracle_walls( px, py, vx, vy, radius, n, W, H );
    float cs;
    int   cols, rows;
    ppc_oracle_grid_dims( radius, n, W, H, &cs, &cols, &rows );
    const size_t cells = (size_t)cols * (size_t)rows;
    uint32_t    *keys  = keys_out ? keys_out : (uint32_t *)malloc( n * sizeof( uint32_t ) );
    uint32_t    *order = order_out ? order_out : (uint32_t *)malloc( n * sizeof( uint32_t ) );
    uint32_t    *start = cell_start_out ? cell_start_out : (uint32_t *)malloc( ( cells + 1 ) * sizeof( uint32_t ) );
    ppc_oracle_grid_build( px, py, n, cs, cols, rows, keys, order, start );
    int32_t *pairs  = NULL;
    size_t   npairs = ppc_oracle_pairs( px, py, radius, n, cols, rows, keys, order, start, &pairs );
    if ( mode == 0 )
        ppc_oracle_resolve_sequential( px, py, vx, vy, mass, radius, n, pairs, npairs, dt );
    else if ( mode == 1 )
        ppc_oracle_resolve_jacobi( px, py, vx, vy, mass, radius, n, pairs, npairs, dt );
    else if ( mode == 3 && npairs ) {
        uint32_t *perm = (uint32_t *)malloc( npairs * sizeof( uint32_t ) );
        ppc_oracle_tilegs_order( pairs, npairs, n, keys, start, cols, rows, 0, g_gs_wc, g_gs_tile_r, g_gs_strip_limit, perm );
        ppc_oracle_resolve_permuted( px, py, vx, vy, mass, radius, n, pairs, npairs, dt, perm );
        free( perm );
    }
    if ( npairs_out )
        *npairs_out = npairs;
    if ( pairs_out )
        *pairs_out = pairs;
    else
        free( pairs );
    if ( !keys_out )
        free( keys );
    if ( !order_out )
        free( order );
    if ( !cell_start_out )
        free( start );
}

/* One substep as main.c runs it (src/main.c:92-98): update, then collisions. */
void ppc_oracle_substep( float *px, float *py, float *vx, float *vy, const float *mass, const float *radius, uint8_t *rgba, size_t n,
                         uint16_t W, uint16_t H, float dt, int mode, int do_colour ) {
    if ( n == 0 ) /* src/main.c:92 */
        return;
    ppc_oracle_update( px, py, vx, vy, mass, rgba, n, dt, do_colour );
    ppc_oracle_collisions_check_grid( px, py, vx, vy, mass, radius, n, W, H, dt, mode, NULL, NULL, NULL, NULL, NULL );
}

/* ------------------------------------------------------------------------------------------------
 * Statistics used by the 1000-substep comparison (BASELINE.json north_star): kinetic energy, momentum,
 * residual overlap (sum of positive penetration depths over all overlapping pairs) and contact count.
 * Not reference code: a measurement on a state, evaluated identically for both sides.
 * out[0]=KE, out[1]=px, out[2]=py, out[3]=overlap sum, out[4]=overlapping pair count.
 * ---------------------------------------------------------------------------------------------- */
void ppc_oracle_stats( const float *px, const float *py, const float *vx, const float *vy, const float *mass, const float *radius,
                       size_t n, uint16_t W, uint16_t H, double *out ) {
    double ke = 0, mx = 0, my = 0, ov = 0, cnt = 0;
    for ( size_t k = 0; k < n; k++ ) {
        const double m = mass[ k ];
        ke += 0.5 * m * ( (double)vx[ k ] * vx[ k ] + (double)vy[ k ] * vy[ k ] );
        mx += m * vx[ k ];
        my += m * vy[ k ];
    }
    if ( n >= 2 ) {
        float cs;
        int   cols, rows;
        ppc_oracle_grid_dims( radius, n, W, H, &cs, &cols, &rows );
        const size_t cells = (size_t)cols * (size_t)rows;
        uint32_t    *keys  = (uint32_t *)malloc( n * sizeof( uint32_t ) );
        uint32_t    *order = (uint32_t *)malloc( n * sizeof( uint32_t ) );
        uint32_t    *start = (uint32_t *)malloc( ( cells + 1 ) * sizeof( uint32_t ) );
        ppc_oracle_grid_build( px, py, n, cs, cols, rows, keys, order, start );
        int32_t *pairs  = NULL;
        size_t   npairs = ppc_oracle_pairs( px, py, radius, n, cols, rows, keys, order, start, &pairs );
        for ( size_t k = 0; k < npairs; k++ ) {
            const int    i = pairs[ 2 * k ], j = pairs[ 2 * k + 1 ];
            const double dx = (double)px[ i ] - px[ j ], dy = (double)py[ i ] - py[ j ];
            const double depth = (double)radius[ i ] + radius[ j ] - sqrt( dx * dx + dy * dy );
            if ( depth > 0 ) {
                ov += depth;
                cnt += 1;
            }
        }
        free( pairs );
        free( keys );
        free( order );
        free( start );
    }
    out[ 0 ] = ke;
    out[ 1 ] = mx;
    out[ 2 ] = my;
    out[ 3 ] = ov;
    out[ 4 ] = cnt;
}
