/*
 * ppc_b200.h -- C ABI of libparticlephysicsc_b200.so, the B200-native drop-in for the solver hot path of
 * isekar0/ParticlePhysicsC.
 *
 * Part 1 declares the reference's OWN entry points with their own signatures; the library exports exactly
 * these symbols, so the reference's main.c and particlesDraw link against it unchanged (the reference has no
 * FFI layer: its boundary is the plain C link interface).  Each declaration cites the
 * reference interface it replaces (paths relative to the reference root).
 *
 * Part 2 declares the additional ppc_* entry points a headless harness needs (when to copy back, parity
 * exporters, timing).  Nothing in this header mentions torch or CUDA types; pointers and sizes only.
 *
 * Include order: when building against the reference's own headers (includes/structs.h, <raylib.h>) include
 * those FIRST; this header then reuses their types.  Stand-alone (tests, headless harness) it defines
 * layout-identical Color / Vector2 / struct Particles_t / struct Pair itself.
 */
#ifndef PPC_B200_H
#define PPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- types: reference includes/structs.h:9-26 and Raylib's Color / Vector2 -------------------------------- */
#ifndef RAYLIB_H
#define PPC_B200_OWN_RAYLIB_TYPES 1
typedef struct Color {
    unsigned char r, g, b, a;
} Color;
typedef struct Vector2 {
    float x, y;
} Vector2;
#endif

#ifndef STRUCTS_H
#define PPC_B200_OWN_STRUCTS 1
struct Particles_t { /* 72 bytes; field order is part of the ABI (includes/structs.h:9-22) */
    float *v_x, *v_y;
    float *p_x, *p_y;
    Color *colors;
    float *mass;
    float *radius;
    size_t length; /* particles processed by the solver: [0, length) */
    size_t size;   /* capacity of the seven arrays */
};
struct Pair { /* includes/structs.h:24-26 */
    int i, j;
};
#endif

/* =====================================================================================================
 * Part 1 -- the reference's entry points (drop-in).
 * ===================================================================================================== */

/* includes/particles_arrays.h:17, src/particles_arrays.c:145-194.  Seven host arrays of num_particles_init
 * elements (page-locked so copy-back is a straight DMA), zero-filled; length = 0.  Returned BY VALUE.
 * Allocation failure: message on stderr and exit(EXIT_FAILURE), as the reference. */
struct Particles_t particlesCreate( size_t num_particles_init );

/* includes/particles_arrays.h:16, src/particles_arrays.c:134-143.  Frees the seven arrays and the device
 * mirror; does not touch the struct's fields (as the reference). */
void particlesFree( struct Particles_t *particles );

/* includes/particles_arrays.h:18, src/particles_arrays.c:257-278.  Bug-compatible: grows x2 when
 * length + num_particles >= size; each copy increments length FIRST and writes slot `length`, so slot 0 is
 * never written and the newest particle is dormant until the next add.  Colour BLUE, zero velocity. */
void particlesAddParticle( struct Particles_t *particles, const Vector2 particle_i_position, const float particle_i_mass,
                           const float particle_i_radius, uint8_t num_particles );

/* includes/particles_arrays.h:19, src/particles_arrays.c:280-291.  Zeroes slot `length`, decrements length. */
void particlesRemoveParticle( struct Particles_t *particles );

/* includes/particles_solver.h:9, src/particles_solver.c:16-43.  Kinematic update + drag + gravity kick (+ colour). */
void particlesUpdate( struct Particles_t *particles, const float time_step );

/* includes/particles_solver.h:10, src/particles_solver.c:197-393.  Wall pass, uniform grid, pair search, contact
 * resolution.  No-op when length < 2 (src/particles_solver.c:204-205). */
void particlesCollisionsCheck_Grid( struct Particles_t *particles, const uint16_t window_width, const uint16_t window_height,
                                    float time_step );

/* =====================================================================================================
 * Part 2 -- harness extensions.
 * ===================================================================================================== */

/* When the host arrays of struct Particles_t are made current. */
enum {
    PPC_SYNC_EVERY_CALL = 0, /* after every particlesUpdate / particlesCollisionsCheck_Grid (fully transparent) */
    PPC_SYNC_FRAME      = 1, /* after every particlesCollisionsCheck_Grid: what main.c needs (default)          */
    PPC_SYNC_MANUAL     = 2  /* headless: only ppc_download() copies back ("only when a frame is drawn")        */
};
/* which arrays ppc_download / the automatic copy-back move */
enum { PPC_DL_POS = 1, PPC_DL_VEL = 2, PPC_DL_COLOR = 4, PPC_DL_ALL = 7 };

void ppc_set_device( int cuda_device );                               /* before the first particlesCreate; default 0 or $PPC_DEVICE */
void ppc_set_sync_policy( struct Particles_t *p, int policy );
void ppc_set_download_mask( struct Particles_t *p, unsigned mask );   /* arrays moved by the automatic copy-back; default by policy:
                                                                       * FRAME = PPC_DL_POS | PPC_DL_COLOR, what particlesDraw reads
                                                                       * (src/particles_arrays.c:317-320; velocities stay on the device until
                                                                       * asked for -- particlesRemoveParticle does), EVERY_CALL = PPC_DL_ALL */
void ppc_request_colors( struct Particles_t *p );                     /* MANUAL policy: the next particlesUpdate also computes colours */
void ppc_host_modified( struct Particles_t *p );                      /* caller wrote the host arrays directly: re-upload at next call */
void ppc_download( struct Particles_t *p, unsigned mask );            /* make host arrays current now (blocking) */
void ppc_synchronize( struct Particles_t *p );                        /* wait for all queued device work */
/* The same copy-back without blocking: the state as of this call is copied to the host arrays while the substeps that
 * follow run (second CUDA stream); the host arrays are valid after ppc_download_wait().  This is "a frame is drawn"
 * overlapped with the computation of the next one. */
void ppc_download_async( struct Particles_t *p, unsigned mask );
void ppc_download_wait( struct Particles_t *p );

/* n x { particlesUpdate(dt); particlesCollisionsCheck_Grid(W, H, dt); } without touching host memory;
 * colours are computed on the last substep iff color_last. */
void ppc_run_substeps( struct Particles_t *p, uint16_t window_width, uint16_t window_height, float time_step, int n, int color_last );

/* Named integer options; returns 0 on success, -1 for an unknown name.
 *   "sort"     0 = counting sort on the cell key (default), 1 = LSD radix sort
 *   "stash"    per-thread contact stash: 1 = on (default), 0 = off (streaming path only; test hook)
 *   "resolve"  1 = sequential-equivalent wavefront: replays the reference's pair order, bit-exact (default),
 *              3 = tile Gauss-Seidel: the reference's sequential pass 2 (src/particles_solver.c:131-191) with the pair
 *                  list in a different, deterministic order (by shared-memory tile, then by cell pair); positions stay
 *                  bit-exact, velocities agree with the reference as well as the reference agrees with itself under a
 *                  reordering of its list; one fused kernel, no dependency rounds,
 *              2 = Jacobi (one pass, diverges in dense packings), 0 = walls + grid only
 *              The environment variable PPC_RESOLVE (0..3) sets the initial value of every store created afterwards: for programs
 *              that only call the reference's six entry points.
 *   "gs_wc"    "resolve"=3: tile width in cells (0 = from the particle density, default).  Part of the pair order: the
 *              ranks of a multi-GPU field must all use the same value
 *   "gs_dump"  "resolve"=3: 1 = every substep also dumps the pairs it resolves (ppc_debug_hot_pairs); test hook
 *   "gs_strip" "resolve"=3: staged particles above which a tile is cut into column strips (0 = the kernel's capacity, default)
 *   "gs_merge" "resolve"=3: 1 = the velocity passes of the four tile colours run in ONE launch, a tile starting as soon as its
 *              neighbouring tiles of lower colour are done; 0 = one launch per colour (default: measured within 3 % of each
 *              other either way).  "gs_slice": records per slice of
 *              the velocity kernel's record stream (0 = chosen from the pair density).  Neither changes any result
 *   "scan_single"  1 = the cell-start table is scanned in one launch (decoupled look-back; default); 0 = reduce / scan / apply
 *   "cell_ranks"   1 = the histogram pass also leaves every particle's rank inside its cell, so the counting sort places the
 *              particles without a second round of atomics (default); 0 = the scatter hands the positions out itself
 *   "defer_state"  1 = the first pass of a substep writes only cell keys and the sorting gather applies the integrate + wall
 *              pass to the record it moves (default: 16 bytes per particle less traffic); 0 = the first pass writes the state
 *   "collide"  0 = auto (shared-memory tile kernel with fine x-bins), 1 = global-memory thread-per-particle kernel,
 *              3 = shared-memory tile kernel scanning the whole 3x3 neighbourhood
 *   "slab_strict"  slab mode: 1 = ppc_slab_begin fails (-2) when the substep just finished was not exact (default); 0 = it only
 *              counts such substeps (ppc_slab_info), for a harness that prefers to widen the halo and redo a phase
 *   "wave_mode"    global rounds of the wavefront: 1 = event driven on packed records while more than a quarter of the
 *              particles are in contact, else polling (default); 2 = always event driven; 0 = always polling
 *   "wave_passes"  shared-memory tile passes ahead of the (polling) global rounds: 0 = none (default); k = k passes while
 *              more than a quarter of the particles are in contact; -k = always k passes.  Experimental: correct, but
 *              slower than the event-driven rounds on every workload measured so far */
int ppc_set_option( struct Particles_t *p, const char *name, long long value );

/* ---- parity exporters: state of the LAST grid build (all return 0 on success, -1 if no grid was built) ---- */
typedef struct ppc_grid_info {
    float    max_radius, cell_size; /* src/particles_solver.c:231-244 */
    int      cols, rows;            /* :247-248 */
    uint64_t cells, n;
} ppc_grid_info;
int     ppc_debug_grid_info( struct Particles_t *p, ppc_grid_info *out );
int     ppc_debug_keys( struct Particles_t *p, uint32_t *keys_by_index /* n */ );      /* :280-291 */
int     ppc_debug_sorted_order( struct Particles_t *p, uint32_t *order /* n */ );      /* stable ascending (key, index) */
int     ppc_debug_cell_start( struct Particles_t *p, uint32_t *cell_start /* cells+1 */ );
/* candidate pairs (i, j) in the reference's list order (:311-383); returns the count, writes min(count, max_pairs) */
int64_t ppc_debug_pairs( struct Particles_t *p, int32_t *pairs, size_t max_pairs );
/* The pair set exactly as the HOT narrow-phase kernel of the last substep found it, (i, j) with i < j, unordered: rebuilt from
 * the contact lists k_collide_tile2 wrote ("resolve"=1) or dumped by k_collide_gs while it resolved them ("resolve"=3 with
 * "gs_dump").  Returns the count; -1 nothing to report; -2 a contact list exceeded the stored lists (pairs incomplete). */
int64_t ppc_debug_hot_pairs( struct Particles_t *p, int32_t *pairs, size_t max_pairs );
/* tile geometry of the last "resolve"=3 substep: {tile width, tile height (cells), strip threshold, strip capacity (particles)} */
int     ppc_debug_gs_params( struct Particles_t *p, int out[ 4 ] );
/* "resolve"=3 holds a sub-tile's particles, contacts and dependency levels in shared memory.  A substep in which some sub-tile
 * does not fit (a region far denser than the rest of the field) is run on the exact resolve instead, for the whole field, so
 * the result stays a function of the state alone.  Returns how many substeps of this store did that. */
unsigned long long ppc_gs_fallbacks( struct Particles_t *p );
/* ... by cause (one substep may have several): [0] a one-column strip of a tile holds more particles than shared memory takes,
 * [1] more contacts in one sub-tile than its contact pool holds, [2] more than 127 owned contacts in one cell or more than 32
 * pair-colours around one pair, [3] the pair-record buffer is full */
void    ppc_gs_fallback_reasons( struct Particles_t *p, unsigned long long out[ 4 ] );
/* last sequential-equivalent resolve: rounds = depth of the reference's pair dependency graph, visits = particle visits */
int     ppc_debug_wave_stats( struct Particles_t *p, unsigned long long *rounds, unsigned long long *visits );

/* ---- timing ---- */
void              *ppc_stream( struct Particles_t *p );                /* the cudaStream_t all work is queued on */
unsigned long long ppc_kernel_launches( void );                        /* kernels launched by this library so far */
void               ppc_profile( struct Particles_t *p, int enable );   /* per-stage CUDA-event timing (serialises) */
int                ppc_profile_stages( void );
const char        *ppc_profile_stage_name( int stage );
int                ppc_profile_read( struct Particles_t *p, int stage, double *total_ms, unsigned long long *launches );
void               ppc_profile_reset( struct Particles_t *p );
const char        *ppc_version( void );

/* =====================================================================================================
 * Part 3 -- multi-GPU slab decomposition (no reference counterpart).
 *
 * The field is cut into slabs of whole cell rows, one store (one process, one GPU) per slab.  A substep is
 *     ppc_slab_begin   integrate + walls of the owned particles; how many records go to each neighbour
 *     ppc_slab_pack    write those records (32 bytes each) into two device buffers of the caller
 *     ppc_slab_interior  (optional) the part of the integrate pass that cannot produce records, while the buffers travel
 *     -- the caller moves the buffers to the neighbouring ranks (NCCL send/recv, or a device copy in one process) --
 *     ppc_slab_finish  take the received records in, build the grid over owned + ghost rows, narrow phase, resolve
 * The particles of a window's outermost ghost row miss their outward contacts; the resolve marks them and everything that
 * resolves a pair with a marked particle.  While the mark does not reach an owned particle, owned positions AND velocities
 * are bit-identical to a single-GPU run of the whole field; ppc_slab_begin / ppc_slab_download return -2 when it did in the
 * substep just finished (halo_rows too small; ppc_slab_info reports how far the mark travelled).  -1 on misuse.
 * ===================================================================================================== */
#define PPC_SLAB_RECORD_BYTES 32

/* grid of the whole field for a given maximum radius (src/particles_solver.c:231-249): rows are what gets partitioned */
int ppc_grid_dims( uint16_t window_width, uint16_t window_height, float max_radius, ppc_grid_info *out );

/* Turn the store into one slab.  The host arrays hold this rank's particles in [0, length), global_ids[k] their index
 * in the whole field (it fixes the pair order on every rank alike); max_radius is the maximum over the WHOLE field. */
int ppc_slab_init( struct Particles_t *p, const uint32_t *global_ids, uint16_t window_width, uint16_t window_height, float max_radius,
                   int own_row_begin, int own_row_end, int halo_rows );
/* "resolve"=3 on slabs: every rank sets "gs_wc" to ppc_gs_auto_tile_width( particles, cells of the WHOLE field ) -- the tile order
 * is part of the result -- and needs at least ppc_slab_gs_halo() ghost rows (9 to 16 on the side where an odd tile row borders, 1
 * to 8 on the other; ppc_slab_finish returns -4 otherwise).  Owned positions and velocities are then bit-identical to the
 * single-GPU run by construction; there is no inexactness mark in this mode.
 * "slab_nb_lo_rows" / "slab_nb_hi_rows" (ppc_set_option, before ppc_slab_init): cell rows the neighbouring ranks own; ppc_slab_init
 * returns -5 for a halo deeper than that (ghost rows only come from the direct neighbours). */
int ppc_slab_gs_halo( int own_row_begin, int own_row_end, int rows_total );
int ppc_gs_auto_tile_width( double particles, double cells );
int ppc_slab_begin( struct Particles_t *p, float time_step, int colour, uint64_t send_counts[ 2 ] /* to lower / upper rows */ );
/* with_header: messages of a fixed, agreed size -- record 0 of each buffer holds the number of records that follow, so the
 * receiver needs no count in advance (ppc_slab_finish then takes the buffers' CAPACITIES and reports the counts received) */
int ppc_slab_pack( struct Particles_t *p, void *send_lo, void *send_hi /* device buffers of send_counts (+1) records */, int with_header );
/* Optional, between ppc_slab_pack and ppc_slab_finish: ppc_slab_begin integrates only the owned particles near the borders (the
 * ones that can end up in a message); this queues the rest on the store's stream.  A host that starts the transfer of the packed
 * buffers on ANOTHER stream (after an event recorded behind ppc_slab_pack) and then calls this overlaps the transfer with most of
 * the integrate pass; ppc_slab_finish calls it itself if the host did not. */
int ppc_slab_interior( struct Particles_t *p );
/* returns the number of owned particles after the substep */
int64_t ppc_slab_finish( struct Particles_t *p, float time_step, const void *recv_lo, uint64_t n_lo, const void *recv_hi, uint64_t n_hi,
                         int with_header, uint64_t received[ 2 ] /* may be NULL */ );
/* owned particles -> host arrays [0, length) in storage order (ascending cell key, global index); mask = PPC_DL_* bits
 * (a drawn frame needs PPC_DL_POS | PPC_DL_COLOR only; mass and radius change owner, never value) */
enum { PPC_DL_IDS = 8, PPC_DL_MASS_RADIUS = 16 };
int ppc_slab_download( struct Particles_t *p, uint32_t *global_ids, unsigned mask );
/* non-blocking form (see ppc_download_async); ppc_download_wait() completes it; global_ids should be page-locked */
int ppc_slab_download_async( struct Particles_t *p, uint32_t *global_ids, unsigned mask );

typedef struct ppc_slab_info_t {
    uint64_t n_owned, own_first, n_window; /* owned particles = sorted slots [own_first, own_first + n_owned) of n_window */
    int      row_lo, row_hi;               /* window of the last grid build (owned + ghost rows) */
    int      own_row_begin, own_row_end, halo_rows, cols, rows;
    uint32_t rounds;                       /* dependency rounds of the last resolve */
    uint32_t taint_rows;                   /* rows the inexactness of the window's outermost rows travelled inwards in the last resolve */
    uint32_t violations;                   /* substeps since ppc_slab_init in which it reached an owned particle ("slab_strict" 0 counts instead of failing) */
} ppc_slab_info_t;
int ppc_slab_info( struct Particles_t *p, ppc_slab_info_t *out );
/* candidate pairs (i, j) whose i (lower global index) this rank owns; every pair of the field appears on exactly one rank */
int64_t ppc_slab_debug_pairs( struct Particles_t *p, int32_t *pairs, size_t max_pairs );

/* =====================================================================================================
 * Part 4 -- state snapshots (the reference has no on-disk format; this one is ours).
 *   header: 8 x uint64 little endian = magic "PPCB2001", version 1, n, window width, window height, 0, 0, 0
 *   body:   p_x[n], p_y[n], v_x[n], v_y[n], mass[n], radius[n] (float32), colors[n] (4 x uint8): the seven arrays
 *           of struct Particles_t (includes/structs.h:9-22) in API order, 28 n bytes.
 * Used to keep settled piles and to replay parity cases.
 * ===================================================================================================== */
/* makes the host arrays current, then writes them; 0 on success, -2 cannot open, -3 short write */
int ppc_snapshot_write( struct Particles_t *p, uint16_t window_width, uint16_t window_height, const char *path );
/* fills the host arrays of a store created with particlesCreate( >= n ) and sets length = n; returns n,
 * -2 cannot open, -3 not a snapshot / truncated, -(n + 16) when the store holds fewer than n particles */
int64_t ppc_snapshot_read( struct Particles_t *p, const char *path, uint16_t *window_width, uint16_t *window_height );

#ifdef __cplusplus
}
#endif
#endif /* PPC_B200_H */
