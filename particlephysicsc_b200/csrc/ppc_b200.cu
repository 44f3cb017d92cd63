// ppc_b200.cu -- the C-ABI shim: the reference's entry points (include/ppc_b200.h part 1) over a device-resident
// particle store, plus the headless-harness extensions (part 2).
//
// The caller-visible struct Particles_t keeps its 72-byte layout and its seven host arrays (reference
// includes/structs.h:9-22); it has no room for a handle, so the device mirror lives in a side table keyed by the
// address of the `mass` array (the app holds exactly one instance, reference src/main.c:57).
//
// Data movement at the boundary:
//   host -> device : only the slots that entered the processed range [0, length) since the last solver call
//   device -> host : p_x, p_y, v_x, v_y, colors when the sync policy says the host arrays must be current
//   per substep    : nothing
//
// There is no CPU fallback: every entry point needs a CUDA device and aborts loudly without one.
#include "../../include/ppc_b200.h"

#include "ppc_collide.cuh"
#include "ppc_common.cuh"
#include "ppc_grid.cuh"
#include "ppc_gs.cuh"
#include "ppc_radix.cuh"
#include "ppc_scan.cuh"
#include "ppc_slab.cuh"

#include <math.h>
#include <string.h>
#include <vector>

namespace ppc {
unsigned long long g_launches = 0;

enum Stage { ST_INTEGRATE = 0, ST_SCAN, ST_SORT, ST_GATHER, ST_COLLIDE, ST_WAVEFRONT, ST_UPLOAD, ST_DOWNLOAD, ST_MISC, ST_SLAB, ST_COUNT };
static const char *k_stage_names[ ST_COUNT ] = { "integrate_walls_keys", "cell_scan", "cell_sort", "canonical_gather",
                                                 "collide",              "resolve_wavefront", "upload", "download", "misc", "slab_exchange_prep" };

struct Store {
    // identity / host side
    const void *host_key   = nullptr;   // particles->mass
    int         device     = 0;         // the CUDA device this store lives on (ppc_set_device at creation time)
    bool        tile_attr_set = false, wtile_attr_set = false;   // dynamic shared-memory limits of the tile kernels set on that device
    bool        owns_host  = false;
    size_t      cap        = 0;   // capacity of host and device arrays (particles->size)
    size_t      n          = 0;   // particles resident on the device: API indices [0, n)
    bool        dirty_all  = true;
    float       max_radius = 6.0f;   // reference src/particles_solver.c:231
    bool        max_radius_dirty = true;
    bool        lists_unverified = true;   // particles were uploaded since the contact-list length was last checked synchronously
    // policy
    int      sync_policy   = PPC_SYNC_FRAME;
    unsigned dl_mask       = 0;   // 0: by policy -- FRAME copies what the draw loop reads (p_x, p_y, colors: 12 B per particle, reference
                                  // src/particles_arrays.c:317-320), EVERY_CALL (fully transparent) copies the velocities as well
    bool     colour_wanted = false;
    bool     col_valid     = false;   // set[cur].col holds the colours of the latest particlesUpdate
    int      opt_sort = 0, opt_stash = 1, opt_resolve = 1, opt_collide = 0, opt_wave_passes = 0, opt_wave_mode = 1, opt_slab_strict = 1;   // resolve: 0 off, 1 sequential-equivalent, 2 Jacobi, 3 tile Gauss-Seidel
    int      opt_nb_lo_rows = 0, opt_nb_hi_rows = 0;   // slab mode: cell rows the lower / upper neighbour owns (0: not told)
    int      opt_defer_state = 1;   // K1 writes keys only and the gather applies the integrate + wall pass (0: K1 writes the state, as round 1 did)
    int      opt_gs_wc = 0, opt_gs_dump = 0, opt_gs_strip = 0;   // tile width of the "resolve"=3 kernel (0: from the particle density); dump the pairs it resolves
    // pending particlesUpdate (fused into the next collisions check)
    bool  pending = false;
    float pending_dt = 0.0f;
    bool  pending_colour = false;
    // device
    cudaStream_t stream = nullptr;
    // asynchronous copy-back ("a frame is drawn" while the next one is computed): host copies run on a second stream
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t  ev_unpacked = nullptr, ev_copied = nullptr;
    bool         copy_pending = false;
    ParticleSet  set[ 2 ] = {};
    int          cur = 0;
    float       *api_f[ 6 ] = {};   // px, py, vx, vy, mass, radius in API order (staging for upload / download)
    uchar4      *api_col = nullptr;
    uint32_t    *api_ids = nullptr;   // slab mode: global indices of the copied-back particles
    uint32_t    *slot_of = nullptr, *sort_keys_b = nullptr, *sort_vals_b = nullptr, *sort_keys_a = nullptr;
    uint32_t    *cell_count = nullptr, *cell_start = nullptr;
    int          opt_slab_fused = 1;   // slab mode: one fused first pass (k_slab_integrate) instead of K1 + count + pack + localize sweeps over the owned particles
    int          opt_slab_overlap = 1; // slab mode: the fused first pass does the border slots first, the interior while the records travel
    size_t       slab_band_lo = 0, slab_band_hi = 0;   // owned slots (relative to own_first) of the rows 2 * halo from either border, at the last grid build
    size_t       slab_int_first = 0, slab_int_count = 0;   // interior part of the current substep's first pass, still to be launched
    StepParams   slab_int_P = {};
    int          slab_int_row_lo = 0, slab_int_row_hi = 0;
    uint32_t     slab_int_drop = 0;
    bool         slab_fused_now = false;   // this substep's ppc_slab_begin ran the fused pass
    int          opt_scan_single = 1;  // cell-start scan in one launch (decoupled look-back); 0: reduce / scan / apply (three launches)
    unsigned long long *scan_desc = nullptr;   // [0]: ticket counter, [1..]: tile descriptors of the single-pass scan
    size_t       scan_desc_cap = 0;
    uint32_t     scan_epoch = 0;
    bool         ranks_valid = false;  // K1 of this substep left per-cell ranks in sort_keys_a
    int          opt_cell_ranks = 1;   // 0: the scatter hands out the positions inside a cell with its own atomics (round 1)
    bool         hist_clean = false;   // cell_count is all zeros (the counting-sort scatter counts it back down): K1 need not clear it
    size_t       cells_cap = 0;
    uint32_t    *scratch = nullptr;
    size_t       scratch_words = 0;
    uint32_t    *d_word = nullptr;   // one word for small reductions
    // "resolve"=3 (tile Gauss-Seidel, ppc_gs.cuh): the staged records, error flags, tile width of the last substep, pair dump
    float4      *rec_a = nullptr;
    uint16_t    *gs_mcol = nullptr;
    uint4       *gs_chunks = nullptr;                       // pair records of the current substep, one chunk per sub-tile
    uint32_t     gs_chunk_cap = 0;                          // 16-byte units
    uint2       *gs_desc = nullptr;                         // chunk of every (tile, strip)
    size_t       gs_desc_cap = 0;
    uint32_t    *gs_flags = nullptr, *gs_dump = nullptr;   // gs_flags[0] error bits, [1] pairs dumped, [2] chunk units handed out
    uint32_t     gs_dump_cap = 0;
    int          gs_wc = 0, gs_strip = 0;
    bool         gs_attr_set = false;
    uint32_t    *h_gs_flag = nullptr;                       // pinned copy of gs_flags[0] after the narrow phase of the current substep
    cudaEvent_t  ev_gs = nullptr;
    unsigned long long gs_fallbacks = 0;                    // substeps the tile kernel could not hold and the exact resolve ran instead
    uint32_t     gs_pairs_per_tile = 0;                     // owned pairs per tile (with pairs or not) of the last settled substep: sizes k_gs_resolve's slices
    int          opt_gs_merge = 0;                          // 1: the four tile colours of k_gs_resolve in ONE launch (tickets + per-tile done flags).  Measured:
                                                            // velocity pass 0.840 -> 0.819 ms on the 16M pile, 0.097 -> 0.106 ms on the dilute uniform field:
                                                            // the colour boundaries are not what bounds the pass, so the default stays four plain launches
    uint32_t    *gs_done = nullptr;                         // per tile: epoch of the substep in which its velocity pass finished
    uint32_t     gs_epoch = 0;
    int          opt_gs_slice = 0;                          // pair records per slice of k_gs_resolve (0: from gs_pairs_per_tile)
    size_t       gs_tiles = 0;
    unsigned long long gs_fallback_reason[ 4 ] = { 0, 0, 0, 0 };   // ... by cause: GS_ERR_CAP, _POOL, _DEPTH, _CHUNK
    bool         gs_pending = false;                        // the verdict of the last "resolve"=3 substep has not been looked at yet (gs_settle)
    size_t       gs_pending_n = 0;                          // ... and what the exact resolve needs if the verdict is "did not fit"
    GridDims     gs_pending_grid = {};
    float        gs_pending_dt = 0.0f;
    WaveBuffers  wave = {};          // sequential-equivalent resolve (allocated on first use)
    WaveRecs     recs = {};          // packed records of the event-driven rounds
    int          events_grid = 0;
    size_t       wave_cap = 0;
    int          wave_grid = 0;
    WaveItem    *wave_items = nullptr;   // work items of the tile passes (4 lists)
    uint32_t    *wave_n_items = nullptr;
    size_t       wave_items_stride = 0;
    uint32_t    *h_active_dma = nullptr;   // pinned [5]: where the device's report lands behind every resolve (see take_activity_report)
    uint32_t    *h_active = nullptr;   // host copy [5] of the report as last looked at: particles with contacts, densest tile (per 1024 cells), -, -, longest unstored contact list, as of a recent substep
    // slab mode (multi-GPU decomposition): this store holds the particles of one slab of cell rows
    bool      slab = false;
    SlabRows  slab_rows = {};
    GridDims  slab_world = {};               // the whole field's grid
    float     slab_wall_w = 0, slab_wall_h = 0;
    size_t    own_first = 0, n_own = 0;      // owned particles = slots [own_first, own_first + n_own) of the current set
    int       row_lo = 0, row_hi = 0;        // window of the last grid build
    uint32_t *slab_counts = nullptr;         // device [8]: send counts lo/hi, lost, -, pack cursors lo/hi
    uint32_t *h_slab = nullptr;              // pinned [8]
    uint32_t  slab_own_first = 0, slab_own_end = 0, slab_taint_rows = 0, slab_violations = 0;
    int       wave_hops = 1;                 // contacts an error can travel per wavefront round of the engine last used (0: unbounded)
    // last grid
    bool     grid_valid = false, pre_valid = false;
    GridDims grid = {};
    // profiling
    bool               profiling = false;
    cudaEvent_t        ev0 = nullptr, ev1 = nullptr;
    double             stage_ms[ ST_COUNT ] = {};
    unsigned long long stage_launches[ ST_COUNT ] = {};
};

static std::vector<Store *> g_stores;
static int                  g_device = -1;   // device of the stores created next (ppc_set_device / $PPC_DEVICE / 0)

static void select_device() {
    if ( g_device < 0 ) {
        const char *e = getenv( "PPC_DEVICE" );
        g_device      = e ? atoi( e ) : 0;
    }
    int count = 0;
    cudaError_t err = cudaGetDeviceCount( &count );
    if ( err != cudaSuccess || count == 0 ) {
        fprintf( stderr, "particlephysicsc_b200: no CUDA device available (%s); this library has no CPU fallback\n",
                 err == cudaSuccess ? "device count is 0" : cudaGetErrorString( err ) );
        abort();
    }
    PPC_CK( cudaSetDevice( g_device ) );
}

template <class T>
static T *dev_alloc( size_t count ) {
    T *p = nullptr;
    PPC_CK( cudaMalloc( &p, ( count ? count : 1 ) * sizeof( T ) ) );
    return p;
}

struct StageTimer {
    Store *st;
    int    stage;
    unsigned long long l0;
    StageTimer( Store *s, int g ) : st( s ), stage( g ), l0( g_launches ) {
        if ( st->profiling )
            PPC_CK( cudaEventRecord( st->ev0, st->stream ) );
    }
    ~StageTimer() {
        if ( st->profiling ) {
            PPC_CK( cudaEventRecord( st->ev1, st->stream ) );
            PPC_CK( cudaEventSynchronize( st->ev1 ) );
            float ms = 0;
            PPC_CK( cudaEventElapsedTime( &ms, st->ev0, st->ev1 ) );
            st->stage_ms[ stage ] += ms;
            st->stage_launches[ stage ] += g_launches - l0;
        }
    }
};

// ---- device storage -------------------------------------------------------------------------------------------
static void free_wave_buffers( Store *st ) {
    cudaFree( st->wave.deg );
    cudaFree( st->wave.state );
    cudaFree( st->wave.inc );
    cudaFree( st->wave.list[ 0 ] );
    cudaFree( st->wave.list[ 1 ] );
    cudaFree( st->wave.counters );
    cudaFree( st->recs.wr );
    cudaFree( st->recs.dr );
    st->recs     = WaveRecs{};
    st->wave     = WaveBuffers{};
    st->wave_cap = 0;
}

// Contact lists are stored with a fixed number of positions per particle (16, or 64 for stores up to 4M particles).  A particle
// with more contacts than that still resolves correctly -- its list entries are recomputed on demand -- but very slowly, so
// when a substep reports such lists (heavily overlapping fields: many particles spawned on one spot) the lists are
// lengthened for the following substeps, as far as 4 GB of list storage go.
// The wavefront reports its activity (particles in contact, longest unstored list, ...) through five pinned words that a
// device-to-host copy rewrites behind every resolve while the host, which queues substeps far ahead of the device, looks at them
// when it queues the next one.  Waiting for the copy would stall the queue, and a report ordered by an event is as old as the
// host is ahead (measured: the engine choice below then runs a whole batch on the report of the batch's first substep).  So the
// words are read as they are: each is an independent, naturally aligned 32-bit value written whole by the copy engine, so the
// host sees an older or a newer report, never a torn one -- and no consumer's RESULT depends on which: both wavefront engines
// are exact, and list growth only ever acts on the largest length seen.
static void take_activity_report( Store *st ) {
    const volatile uint32_t *src = st->h_active_dma;
    const uint32_t pending_grow = st->h_active[ 4 ];
    for ( int k = 0; k < 5; k++ )
        st->h_active[ k ] = src[ k ];
    if ( pending_grow > st->h_active[ 4 ] )
        st->h_active[ 4 ] = pending_grow;
}

static void grow_contact_lists( Store *st ) {
    take_activity_report( st );
    const uint32_t seen = st->h_active[ 4 ];   // a substep or two old
    if ( !st->wave.deg || seen <= st->wave.maxdeg )
        return;
    uint32_t want = st->wave.maxdeg;
    while ( want < seen && want < 4096u )
        want *= 2;
    const size_t limit = ( (size_t)4 << 30 ) / ( sizeof( uint32_t ) * ( st->wave_cap ? st->wave_cap : 1 ) );
    if ( want > limit )
        want = (uint32_t)limit;
    if ( want <= st->wave.maxdeg )
        return;
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    cudaFree( st->wave.inc );
    st->wave.maxdeg   = want;
    st->wave.inc      = dev_alloc<uint32_t>( st->wave_cap * (size_t)want );
    st->h_active[ 4 ] = 0;
}

static void ensure_wave_buffers( Store *st ) {
    if ( st->wave_cap >= st->cap && st->wave.deg ) {
        grow_contact_lists( st );
        return;
    }
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    free_wave_buffers( st );
    const size_t cap    = st->cap ? st->cap : 1;
    st->wave.deg        = dev_alloc<uint16_t>( cap );
    st->wave.state      = dev_alloc<uint32_t>( cap );
    // Contact lists: 16 positions per particle always; 64 while that stays under 1 GB, so that heavily overlapping
    // scenes (the reference's debug scene stacks 10 coincident copies per add, src/main.c:80-85) keep stored lists.
    st->wave.maxdeg     = cap <= ( (size_t)1 << 22 ) ? 64u : (uint32_t)INC_MAXDEG;
    st->wave.inc        = dev_alloc<uint32_t>( cap * (size_t)st->wave.maxdeg );
    st->wave.stride     = cap;
    st->wave.list[ 0 ]  = dev_alloc<uint32_t>( cap );
    st->wave.list[ 1 ]  = dev_alloc<uint32_t>( cap );
    st->wave.counters   = dev_alloc<uint32_t>( 16 );
    st->wave_cap        = cap;
    PPC_CK( cudaMemsetAsync( st->wave.counters, 0, 16 * sizeof( uint32_t ), st->stream ) );
    int per_sm = 0, sms = 0, dev = 0;
    PPC_CK( cudaGetDevice( &dev ) );
    PPC_CK( cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, dev ) );
    PPC_CK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm, k_resolve_wavefront, WAVE_THREADS, 0 ) );
    if ( per_sm < 1 ) {
        fprintf( stderr, "particlephysicsc_b200: wavefront kernel cannot be made resident\n" );
        abort();
    }
    st->wave_grid = per_sm * sms;
    st->recs.wr   = dev_alloc<WaveRec>( cap );
    st->recs.dr   = dev_alloc<DynRec>( cap );
    PPC_CK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm, k_resolve_events, EV_THREADS, 0 ) );
    if ( per_sm < 1 ) {
        fprintf( stderr, "particlephysicsc_b200: event wavefront kernel cannot be made resident\n" );
        abort();
    }
    st->events_grid = per_sm * sms;   // latency bound: as many loads in flight as the SM holds
}

static void check_gs_flags( Store *st ) {   // after a stream synchronise
    if ( st->gs_flags ) {
        uint32_t flags = 0;
        PPC_CK( cudaMemcpy( &flags, st->gs_flags, sizeof( uint32_t ), cudaMemcpyDeviceToHost ) );
        if ( flags ) {
            fprintf( stderr, "particlephysicsc_b200: \"resolve\"=3 cannot run this field (%s%s%s%s); use the default resolve\n",
                     ( flags & GS_ERR_CAP ) ? "a one-column strip of a tile holds more particles than shared memory takes " : "",
                     ( flags & GS_ERR_POOL ) ? "more contacts in one tile than its contact pool holds " : "",
                     ( flags & GS_ERR_DEPTH ) ? "more than 127 owned contacts in one cell, or a chain of more than 255 dependent pairs inside one tile " : "",
                     ( flags & GS_ERR_CHUNK ) ? "pair-record buffer full" : "" );
            abort();
        }
    }
}

static void check_device_flags( Store *st ) {   // after a stream synchronise
    check_gs_flags( st );
    if ( !st->wave.counters )
        return;
    uint32_t flags = 0;
    PPC_CK( cudaMemcpy( &flags, st->wave.counters + 3, sizeof( uint32_t ), cudaMemcpyDeviceToHost ) );
    if ( flags ) {
        fprintf( stderr, "particlephysicsc_b200: resolve failed (%s%s)\n", ( flags & 1u ) ? "a particle has more than 4095 simultaneous contacts " : "",
                 ( flags & 2u ) ? "pair dependency chain longer than 2^19 - 1 rounds" : "" );
        abort();
    }
}

static void free_gs_buffers( Store *st ) {
    cudaFree( st->rec_a );
    cudaFree( st->gs_mcol );
    cudaFree( st->gs_chunks );
    cudaFree( st->gs_desc );
    cudaFree( st->gs_done );
    st->gs_done = nullptr;
    cudaFree( st->gs_dump );
    st->rec_a       = nullptr;
    st->gs_mcol     = nullptr;
    st->gs_chunks   = nullptr;
    st->gs_desc     = nullptr;
    st->gs_desc_cap = 0;
    st->gs_dump     = nullptr;
    st->gs_dump_cap = 0;
}

// Buffers of the tile Gauss-Seidel path, on first use: 18 bytes of staged records per particle plus the pair records.
static void ensure_gs_buffers( Store *st ) {
    if ( !st->gs_flags ) {
        st->gs_flags = dev_alloc<uint32_t>( 4 );
        PPC_CK( cudaMemsetAsync( st->gs_flags, 0, 4 * sizeof( uint32_t ), st->stream ) );
    }
    if ( !st->rec_a ) {
        st->rec_a   = dev_alloc<float4>( st->cap );
        st->gs_mcol = dev_alloc<uint16_t>( st->cap + 16 );   // bulk copies of the 16-bit words round their ends up to 16 bytes
        // pair records: 16 bytes per pair plus a small header per sub-tile; a dense pile has ~3 pairs per particle
        const size_t units = st->cap * 6 + ( (size_t)1 << 20 );
        st->gs_chunk_cap   = (uint32_t)( units < 0xfffffff0u ? units : 0xfffffff0u );
        st->gs_chunks      = dev_alloc<uint4>( st->gs_chunk_cap );
    }
    if ( st->opt_gs_dump && !st->gs_dump ) {
        st->gs_dump_cap = (uint32_t)( st->cap * 8 < 0x7fffffffu ? st->cap * 8 : 0x7fffffffu );
        st->gs_dump     = dev_alloc<uint32_t>( 2 * (size_t)st->gs_dump_cap );
    }
}

static void free_particle_arrays( Store *st ) {
    free_wave_buffers( st );
    free_gs_buffers( st );
    for ( int s = 0; s < 2; s++ ) {
        cudaFree( st->set[ s ].pv );
        cudaFree( st->set[ s ].rm );
        cudaFree( st->set[ s ].gid );
        cudaFree( st->set[ s ].key );
        cudaFree( st->set[ s ].col );
        st->set[ s ] = ParticleSet{};
    }
    for ( int a = 0; a < 6; a++ ) {
        cudaFree( st->api_f[ a ] );
        st->api_f[ a ] = nullptr;
    }
    cudaFree( st->api_col );
    cudaFree( st->api_ids );
    cudaFree( st->slot_of );
    cudaFree( st->sort_keys_a );
    cudaFree( st->sort_keys_b );
    cudaFree( st->sort_vals_b );
    cudaFree( st->scratch );
    st->api_col = nullptr;
    st->api_ids = nullptr;
    st->slot_of = st->sort_keys_a = st->sort_keys_b = st->sort_vals_b = st->scratch = nullptr;
    st->scratch_words = 0;
}

// ---- asynchronous copy-back -----------------------------------------------------------------------------------------
// The unpack kernel (storage order -> staging arrays) runs on the compute stream; the device-to-host copies of the staging
// arrays run on a second stream, so they overlap the substeps that follow.  Anything that touches the staging arrays or
// the host arrays again waits first.
static void copy_wait( Store *st ) {
    if ( st->copy_pending ) {
        PPC_CK( cudaStreamSynchronize( st->copy_stream ) );
        st->copy_pending = false;
    }
}
// Call after the unpack kernel was queued on the compute stream: returns the stream the host copies go to.
static cudaStream_t copy_stream_after_unpack( Store *st, bool async ) {
    if ( !async )
        return st->stream;
    PPC_CK( cudaEventRecord( st->ev_unpacked, st->stream ) );
    PPC_CK( cudaStreamWaitEvent( st->copy_stream, st->ev_unpacked, 0 ) );
    return st->copy_stream;
}
static void copy_started( Store *st, bool async ) {
    if ( async ) {
        PPC_CK( cudaEventRecord( st->ev_copied, st->copy_stream ) );
        st->copy_pending = true;
    }
}
// Before the staging arrays are overwritten by a new unpack: the previous copies must have read them.
static void staging_reuse( Store *st, bool async ) {
    if ( !st->copy_pending )
        return;
    if ( async )
        PPC_CK( cudaStreamWaitEvent( st->stream, st->ev_copied, 0 ) );   // device-side wait; the host does not block
    else
        copy_wait( st );
}

// (Re)allocate the per-particle device arrays for `cap` particles, keeping the first `keep` storage slots.
static void resize_particle_arrays( Store *st, size_t cap, size_t keep ) {
    copy_wait( st );
    ParticleSet old[ 2 ] = { st->set[ 0 ], st->set[ 1 ] };
    for ( int s = 0; s < 2; s++ ) {
        st->set[ s ].pv  = dev_alloc<float4>( cap );
        st->set[ s ].rm  = dev_alloc<float2>( cap );
        st->set[ s ].gid = dev_alloc<uint32_t>( cap );
        st->set[ s ].key = dev_alloc<uint32_t>( cap );
        st->set[ s ].col = dev_alloc<uchar4>( cap );
    }
    if ( keep && old[ st->cur ].pv ) {   // only the current set holds live data between calls
        const ParticleSet &o = old[ st->cur ], &d = st->set[ st->cur ];
        PPC_CK( cudaMemcpyAsync( d.pv, o.pv, keep * sizeof( float4 ), cudaMemcpyDeviceToDevice, st->stream ) );
        PPC_CK( cudaMemcpyAsync( d.rm, o.rm, keep * sizeof( float2 ), cudaMemcpyDeviceToDevice, st->stream ) );
        PPC_CK( cudaMemcpyAsync( d.gid, o.gid, keep * sizeof( uint32_t ), cudaMemcpyDeviceToDevice, st->stream ) );
        PPC_CK( cudaMemcpyAsync( d.key, o.key, keep * sizeof( uint32_t ), cudaMemcpyDeviceToDevice, st->stream ) );
        PPC_CK( cudaMemcpyAsync( d.col, o.col, keep * sizeof( uchar4 ), cudaMemcpyDeviceToDevice, st->stream ) );
        PPC_CK( cudaStreamSynchronize( st->stream ) );
    }
    for ( int s = 0; s < 2; s++ ) {
        cudaFree( old[ s ].pv );
        cudaFree( old[ s ].rm );
        cudaFree( old[ s ].gid );
        cudaFree( old[ s ].key );
        cudaFree( old[ s ].col );
    }
    for ( int a = 0; a < 6; a++ ) {
        cudaFree( st->api_f[ a ] );
        st->api_f[ a ] = dev_alloc<float>( cap );
    }
    cudaFree( st->api_col );
    cudaFree( st->api_ids );
    cudaFree( st->slot_of );
    cudaFree( st->sort_keys_a );
    cudaFree( st->sort_keys_b );
    cudaFree( st->sort_vals_b );
    st->api_col     = dev_alloc<uchar4>( cap );
    st->api_ids     = dev_alloc<uint32_t>( cap );
    st->slot_of     = dev_alloc<uint32_t>( cap );
    st->sort_keys_a = dev_alloc<uint32_t>( cap );
    st->sort_keys_b = dev_alloc<uint32_t>( cap );
    st->sort_vals_b = dev_alloc<uint32_t>( cap );
    st->cap         = cap;
    st->pre_valid   = false;
    free_wave_buffers( st );
    free_gs_buffers( st );
}

static void ensure_scratch( Store *st, size_t words ) {
    if ( words <= st->scratch_words )
        return;
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    cudaFree( st->scratch );
    st->scratch       = dev_alloc<uint32_t>( words );
    st->scratch_words = words;
}

static void ensure_cells( Store *st, size_t cells ) {
    if ( cells + 1 <= st->cells_cap )
        return;
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    cudaFree( st->cell_count );
    cudaFree( st->cell_start );
    st->cells_cap  = cells + 1 + cells / 8;
    st->cell_count = dev_alloc<uint32_t>( st->cells_cap );
    st->cell_start = dev_alloc<uint32_t>( st->cells_cap );
    st->hist_clean = false;
}

static Store *new_store( const struct Particles_t *p, bool owns_host ) {
    select_device();
    Store *st     = new Store();
    st->device    = g_device;
    st->host_key  = p->mass;
    st->owns_host = owns_host;
    PPC_CK( cudaStreamCreateWithFlags( &st->stream, cudaStreamNonBlocking ) );
    PPC_CK( cudaStreamCreateWithFlags( &st->copy_stream, cudaStreamNonBlocking ) );
    PPC_CK( cudaEventCreateWithFlags( &st->ev_unpacked, cudaEventDisableTiming ) );
    PPC_CK( cudaEventCreateWithFlags( &st->ev_copied, cudaEventDisableTiming ) );
    PPC_CK( cudaEventCreate( &st->ev0 ) );
    PPC_CK( cudaEventCreate( &st->ev1 ) );
    st->d_word = dev_alloc<uint32_t>( 4 );
    PPC_CK( cudaHostAlloc( (void **)&st->h_active_dma, 5 * sizeof( uint32_t ), cudaHostAllocDefault ) );
    st->h_active      = (uint32_t *)malloc( 5 * sizeof( uint32_t ) );
    st->h_active_dma[ 0 ] = st->h_active[ 0 ] = 0xffffffffu;   // unknown yet: assume dense
    for ( int k = 1; k < 5; k++ )
        st->h_active_dma[ k ] = 0;
    st->h_active[ 1 ] = st->h_active[ 2 ] = st->h_active[ 3 ] = st->h_active[ 4 ] = 0;
    resize_particle_arrays( st, p->size, 0 );
    // A program that only knows the reference's six entry points (main.c, examples/headless_scene.c) chooses the resolve from outside
    if ( const char *e = getenv( "PPC_RESOLVE" ) ) {
        const int v = atoi( e );
        if ( v < 0 || v > 3 ) {
            fprintf( stderr, "particlephysicsc_b200: PPC_RESOLVE=%s (expected 0..3)\n", e );
            abort();
        }
        st->opt_resolve = v;
    }
    g_stores.push_back( st );
    return st;
}

// Find the device mirror of a caller's struct; adopt arrays we did not allocate (a harness that filled its own).
static void gs_settle( Store *st );
static Store *store_of( struct Particles_t *p ) {
    for ( Store *st : g_stores )
        if ( st->host_key == p->mass ) {
            PPC_CK( cudaSetDevice( st->device ) );   // several stores on several GPUs in one process (or a host framework that
            if ( st->gs_pending )                    // switches devices between calls): always work on this store's device
                gs_settle( st );                     // every entry point sees a state whose last substep is settled
            return st;
        }
    Store *st     = new_store( p, false );
    st->dirty_all = true;
    return st;
}

// ---- host <-> device ------------------------------------------------------------------------------------------
static float *const *host_float_arrays( struct Particles_t *p, float *out[ 6 ] ) {
    out[ 0 ] = p->p_x, out[ 1 ] = p->p_y, out[ 2 ] = p->v_x, out[ 3 ] = p->v_y, out[ 4 ] = p->mass, out[ 5 ] = p->radius;
    return out;
}

// Upload API indices [first, first+count) from the host arrays into storage slots of the same numbers.
static void upload_range( Store *st, struct Particles_t *p, size_t first, size_t count ) {
    if ( count == 0 )
        return;
    copy_wait( st );   // the staging arrays are shared with the copy-back
    st->lists_unverified = true;
    StageTimer T( st, ST_UPLOAD );
    float     *h[ 6 ];
    host_float_arrays( p, h );
    for ( int a = 0; a < 6; a++ )
        PPC_CK( cudaMemcpyAsync( st->api_f[ a ] + first, h[ a ] + first, count * sizeof( float ), cudaMemcpyHostToDevice, st->stream ) );
    PPC_CK( cudaMemcpyAsync( st->api_col + first, p->colors + first, count * sizeof( uchar4 ), cudaMemcpyHostToDevice, st->stream ) );
    PPC_LAUNCH( k_pack_upload, div_up( count, STREAM_THREADS ), STREAM_THREADS, 0, st->stream, st->set[ st->cur ], st->api_f[ 0 ],
                st->api_f[ 1 ], st->api_f[ 2 ], st->api_f[ 3 ], st->api_f[ 4 ], st->api_f[ 5 ], st->api_col, first, count );
    // the host buffers may be pageable (adopted arrays) or rewritten by the caller right after we return
    PPC_CK( cudaStreamSynchronize( st->stream ) );
}

static void refresh_max_radius( Store *st ) {
    if ( !st->max_radius_dirty )
        return;
    StageTimer     T( st, ST_MISC );
    const float    floor_r = 6.0f;   // reference src/particles_solver.c:231
    uint32_t       bits;
    memcpy( &bits, &floor_r, 4 );
    PPC_CK( cudaMemcpyAsync( st->d_word, &bits, 4, cudaMemcpyHostToDevice, st->stream ) );
    if ( st->n ) {
        const unsigned blocks = st->n < (size_t)148 * 8 * STREAM_THREADS ? div_up( st->n, STREAM_THREADS ) : 148u * 8u;
        PPC_LAUNCH( k_max_radius, blocks, STREAM_THREADS, 0, st->stream, st->set[ st->cur ].rm, st->n, st->d_word );
    }
    PPC_CK( cudaMemcpyAsync( &bits, st->d_word, 4, cudaMemcpyDeviceToHost, st->stream ) );
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    memcpy( &st->max_radius, &bits, 4 );
    st->max_radius_dirty = false;
}

// Bring the device mirror in line with what the caller did to the host struct since the last call.
static void absorb_host_changes( Store *st, struct Particles_t *p ) {
    if ( st->slab ) {
        fprintf( stderr, "particlephysicsc_b200: this store is one slab of a multi-GPU field; drive it with ppc_slab_begin / _pack / _finish\n" );
        abort();
    }
    if ( p->size != st->cap ) {   // arrays were re-sized behind our back (adopted arrays only)
        PPC_CK( cudaStreamSynchronize( st->stream ) );
        resize_particle_arrays( st, p->size, 0 );
        st->dirty_all = true;
    }
    if ( p->length > st->cap ) {
        fprintf( stderr, "particlephysicsc_b200: length %zu exceeds size %zu\n", p->length, p->size );
        abort();
    }
    if ( !st->dirty_all && p->length < st->n )   // shrunk without particlesRemoveParticle: host arrays are the truth
        st->dirty_all = true;
    if ( st->dirty_all ) {
        st->pending = false;
        st->n       = p->length;
        upload_range( st, p, 0, st->n );
        st->col_valid        = true;   // the host's colours
        st->dirty_all        = false;
        st->grid_valid       = false;
        st->pre_valid        = false;
        st->max_radius_dirty = true;
    } else if ( p->length > st->n ) {
        const size_t first = st->n, count = p->length - st->n;
        upload_range( st, p, first, count );
        st->n = p->length;
        if ( !st->max_radius_dirty )
            for ( size_t k = first; k < first + count; k++ )   // reference :231-236, incrementally
                if ( p->radius[ k ] > st->max_radius )
                    st->max_radius = p->radius[ k ];
        st->pre_valid = false;
    }
}

static unsigned auto_download_mask( const Store *st ) {
    if ( st->dl_mask )
        return st->dl_mask;
    return st->sync_policy == PPC_SYNC_EVERY_CALL ? (unsigned)PPC_DL_ALL : (unsigned)( PPC_DL_POS | PPC_DL_COLOR );
}

// ---- the substep ----------------------------------------------------------------------------------------------
static void launch_integrate( Store *st, const StepParams &P ) {
    StageTimer T( st, ST_INTEGRATE );
    if ( P.do_walls_keys && P.do_histogram && !st->hist_clean )
        PPC_CK( cudaMemsetAsync( st->cell_count, 0, st->cells_cap * sizeof( uint32_t ), st->stream ) );
    if ( P.do_walls_keys && P.do_histogram )
        st->hist_clean = false;
    // counting-sort mode: K1 also leaves every particle's rank inside its cell (in the radix engine's idle key buffer)
    uint32_t *ranks = ( P.do_walls_keys && P.do_histogram && st->opt_sort == 0 && st->opt_cell_ranks ) ? st->sort_keys_a : nullptr;
    st->ranks_valid = ranks != nullptr;
    PPC_LAUNCH( k_integrate_walls_keys, div_up( st->n, STREAM_THREADS ), STREAM_THREADS, 0, st->stream, st->set[ st->cur ], st->cell_count,
                st->n, P, ranks );
}

// particlesUpdate on its own (no collisions check follows, or the host wants to look at the result).
static void flush_pending_update( Store *st ) {
    if ( !st->pending )
        return;
    st->pending = false;
    if ( st->n == 0 )
        return;
    StepParams P   = {};
    P.dt_update    = st->pending_dt;
    P.do_integrate = 1;
    P.do_colour    = st->pending_colour ? 1 : 0;
    launch_integrate( st, P );
    st->col_valid = st->pending_colour;
    st->pre_valid = false;
}

static int bits_for( uint32_t cells ) {
    int b = 0;
    while ( b < 32 && ( ( cells - 1 ) >> b ) != 0 )
        b++;
    return b > 0 ? b : 1;
}

// Cell-start table, counting / radix sort and canonical gather over the first n storage slots of the current set, whose
// keys lie in [0, bins) and whose histogram is in cell_count (K4-K6).  The sorted copy becomes the current set.
static ParticleSet offset_view( const ParticleSet &S, const size_t first ) {
    return ParticleSet{ S.pv + first, S.rm + first, S.gid + first, S.key + first, S.col + first };
}

static void grid_stage( Store *st, const size_t n, const size_t bins, const int cols, const size_t first = 0, const uint32_t unranked_cell = 0xffffffffu,
                        const StepParams *deferred = nullptr ) {
    const ParticleSet  A = offset_view( st->set[ st->cur ], first );   // input: n slots starting at `first`
    const ParticleSet &B = st->set[ st->cur ^ 1 ];
    {
        StageTimer T( st, ST_SCAN );                                                   // K4 cell-start table
        if ( st->opt_scan_single ) {   // one launch (decoupled look-back)
            const size_t tiles = div_up( bins + 1, SCAN_TILE );
            if ( tiles + 1 > st->scan_desc_cap ) {
                PPC_CK( cudaStreamSynchronize( st->stream ) );
                cudaFree( st->scan_desc );
                st->scan_desc_cap = tiles + 1 + tiles / 8;
                st->scan_desc     = dev_alloc<unsigned long long>( st->scan_desc_cap + 1 );
                PPC_CK( cudaMemsetAsync( st->scan_desc, 0, ( st->scan_desc_cap + 1 ) * sizeof( unsigned long long ), st->stream ) );
                st->scan_epoch = 0;
            }
            if ( ++st->scan_epoch >= ( 1u << 30 ) ) {   // the epoch field is 30 bits: start over on clean descriptors
                PPC_CK( cudaMemsetAsync( st->scan_desc, 0, ( st->scan_desc_cap + 1 ) * sizeof( unsigned long long ), st->stream ) );
                st->scan_epoch = 1;
            }
            exclusive_scan_single_pass( st->cell_count, st->cell_start, bins + 1, st->scan_desc + 1, reinterpret_cast<uint32_t *>( st->scan_desc ),
                                        st->scan_epoch, st->stream );
        } else
            exclusive_scan( st->cell_count, st->cell_start, bins + 1, st->scratch, st->stream );
    }
    const uint32_t *slot_of = st->slot_of;
    {
        StageTimer T( st, ST_SORT );                                                   // K5
        if ( st->opt_sort == 1 ) {
            PPC_CK( cudaMemcpyAsync( st->sort_keys_a, A.key, n * sizeof( uint32_t ), cudaMemcpyDeviceToDevice, st->stream ) );
            const int where = radix_sort_pairs( st->sort_keys_a, st->slot_of, st->sort_keys_b, st->sort_vals_b, true, n,
                                                bits_for( (uint32_t)bins ), st->scratch, st->stream );
            slot_of         = where ? st->sort_vals_b : st->slot_of;
        } else if ( st->ranks_valid ) {   // the first pass ranked the particles inside their cells: no atomics here
            PPC_LAUNCH( k_cell_place, div_up( n, STREAM_THREADS ), STREAM_THREADS, 0, st->stream, A.key, (const uint32_t *)( st->sort_keys_a + first ),
                        st->cell_count, st->cell_start, st->slot_of, n );
            st->hist_clean  = true;
            st->ranks_valid = false;
        } else {   // counts the histogram back down to zero
            PPC_LAUNCH( k_cell_scatter, div_up( n, STREAM_THREADS ), STREAM_THREADS, 0, st->stream, A.key, st->cell_count, st->cell_start,
                        st->slot_of, n );
            st->hist_clean = true;
        }
    }
    {
        StageTimer T( st, ST_GATHER );                                                 // K6
        const bool gs = st->opt_resolve == 3;   // also leave the records the tile kernel stages
        StepParams P  = {};
        if ( deferred )
            P = *deferred;   // K1 kept the integrated, wall-constrained state to itself: the gather repeats it on the record it moves
        PPC_LAUNCH( k_canonical_gather, div_up( n, STREAM_THREADS ), STREAM_THREADS, 0, st->stream, A, B, slot_of, st->cell_start, n,
                    st->col_valid ? 1 : 0, unranked_cell, gs ? st->rec_a : (float4 *)nullptr, gs ? st->gs_mcol : (uint16_t *)nullptr,
                    (uint32_t)cols, P );
    }
    st->cur ^= 1;   // B is now the sorted, pre-resolve state
}

// Narrow phase + contact resolution over the first n slots of the (sorted) current set on grid G (K7, K7b).
// "resolve"=3: narrow phase, position pushes and the tile Gauss-Seidel velocity resolve in one kernel (ppc_gs.cuh), one launch
// per tile colour.  Reads the staged records the gather wrote, writes the final state into the sorted set's pv array.
static int gs_auto_width( const double particles, const double cells ) {
    // as wide as keeps a tile of average density, with its one-cell halo, just inside the kernel's staging capacity (measured: filling
    // the shared memory beats one particle per thread -- uniform16m 1.71 -> 1.44 ms, pile16m 4.2 -> 4.0 ms); denser regions are cut
    // into strips by the kernel itself
    const double per_cell = cells > 0 ? particles / cells : 0.0;
    return (int)( (double)GS_STAGE_TARGET / ( ( GS_R + 2 ) * ( per_cell > 0.05 ? per_cell : 0.05 ) ) ) - 2;
}

static int gs_tile_width( const Store *st, const size_t n, const GridDims &G ) {
    int wc = st->opt_gs_wc;
    if ( wc <= 0 )
        wc = gs_auto_width( (double)n, (double)G.cells );
    const int wc_max = GS_THREADS / ( GS_R + 2 ) - 2 < GS_MAXC - 2 ? GS_THREADS / ( GS_R + 2 ) - 2 : GS_MAXC - 2;   // one thread per staged cell
    return wc < 2 ? 2 : ( wc > wc_max ? wc_max : wc );   // >= 2 cells, so tiles of one colour touch disjoint particles
}

// The verdict "every sub-tile fit" is looked at by gs_settle, a substep late at most: the host queues the velocity kernels and
// returns, so it is never more than one narrow phase behind the GPU and the GPU never waits for it.
static void resolve_gs( Store *st, const size_t n, const GridDims &G, const float dt ) {
    if ( !st->gs_attr_set ) {
        PPC_CK( cudaFuncSetAttribute( k_gs_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof( GsSmem ) ) );
        st->gs_attr_set = true;
    }
    GsLaunch L = {};
    L.wc       = st->gs_wc = gs_tile_width( st, n, G );
    L.tiles_x  = ( G.cols + L.wc - 1 ) / L.wc;
    L.ty_first = G.row0 / GS_R;
    L.strip_limit = st->gs_strip = ( st->opt_gs_strip > 0 && st->opt_gs_strip < GS_STRIP_LIMIT ) ? st->opt_gs_strip : GS_STRIP_LIMIT;
    const int    tiles_y = ( G.row0 + G.rows + GS_R - 1 ) / GS_R - L.ty_first;
    const size_t tiles   = (size_t)L.tiles_x * tiles_y;
    if ( tiles * GS_MAXSTRIPS > st->gs_desc_cap ) {
        PPC_CK( cudaStreamSynchronize( st->stream ) );
        cudaFree( st->gs_desc );
        st->gs_desc_cap = tiles * GS_MAXSTRIPS + tiles * GS_MAXSTRIPS / 8;
        st->gs_desc     = dev_alloc<uint2>( st->gs_desc_cap );
        cudaFree( st->gs_done );
        st->gs_done = dev_alloc<uint32_t>( st->gs_desc_cap / GS_MAXSTRIPS + 1 );
        PPC_CK( cudaMemsetAsync( st->gs_done, 0, ( st->gs_desc_cap / GS_MAXSTRIPS + 1 ) * sizeof( uint32_t ), st->stream ) );
        st->gs_epoch = 0;
    }
    const GsPairs PR{ st->gs_chunks, st->gs_flags + 2, st->gs_chunk_cap, st->gs_desc };
    GsDump        D{ nullptr, nullptr, 0u };
    PPC_CK( cudaMemsetAsync( st->gs_flags + 1, 0, 3 * sizeof( uint32_t ), st->stream ) );   // dump count, record cursor, ticket counter
    if ( st->opt_gs_dump )
        D = GsDump{ st->gs_dump, st->gs_flags + 1, st->gs_dump_cap };
    {
        StageTimer T( st, ST_COLLIDE );   // narrow phase, position pushes, pair records: all tiles at once
        PPC_LAUNCH( k_gs_pairs, (unsigned)tiles, GS_THREADS, sizeof( GsSmem ), st->stream, (const float4 *)st->rec_a, (const uint16_t *)st->gs_mcol,
                    st->set[ st->cur ].pv, st->cell_start, G, dt, L, PR, st->gs_flags, D );
    }
    // Did every sub-tile fit the kernel's shared-memory capacities?  The answer travels to the host behind the narrow phase; the
    // velocity kernels are queued meanwhile and do nothing if it is no, so the GPU never waits for the host.
    if ( !st->h_gs_flag ) {
        PPC_CK( cudaHostAlloc( (void **)&st->h_gs_flag, 3 * sizeof( uint32_t ), cudaHostAllocDefault ) );
        PPC_CK( cudaEventCreateWithFlags( &st->ev_gs, cudaEventDisableTiming ) );
    }
    PPC_CK( cudaMemcpyAsync( st->h_gs_flag, st->gs_flags, 3 * sizeof( uint32_t ), cudaMemcpyDeviceToHost, st->stream ) );   // flag, -, record units handed out
    st->gs_tiles = tiles;
    // dense fields (a thousand pairs and more per tile) stream their records in long slices; dilute ones keep the CTA small
    const int slice = st->opt_gs_slice > 0 ? st->opt_gs_slice : ( st->gs_pairs_per_tile >= 600u ? 512 : ( st->gs_pairs_per_tile >= 100u ? 256 : 64 ) );
    L.slice_log2    = 4;
    while ( L.slice_log2 < 10 && ( 1 << ( L.slice_log2 + 1 ) ) <= slice )
        L.slice_log2++;   // 16 .. 1024 records, a power of two
    PPC_CK( cudaEventRecord( st->ev_gs, st->stream ) );
    {
        StageTimer T( st, ST_WAVEFRONT );     // velocity impulses: tiles of one colour before those of the next
        const size_t smem = ( (size_t)GSB_NBUF << L.slice_log2 ) * sizeof( uint4 );
        if ( st->opt_gs_merge ) {             // one launch: a tile starts as soon as its neighbours of lower colour are done
            L.merged  = 1;
            L.tiles_y = tiles_y;
            L.epoch   = ++st->gs_epoch;
            uint32_t total = 0;
            for ( int colour = 0; colour < 4; colour++ ) {
                const int ntx = ( L.tiles_x - ( colour & 1 ) + 1 ) / 2, ty_start = ( ( colour >> 1 ) - L.ty_first ) & 1;
                const int nty = ( tiles_y - ty_start + 1 ) / 2;
                L.col_begin[ colour ]    = total;
                L.col_ntx[ colour ]      = ntx > 0 ? ntx : 1;
                L.col_ty_start[ colour ] = ty_start;
                total += ( ntx > 0 && nty > 0 ) ? (uint32_t)( ntx * nty ) : 0u;
            }
            L.col_begin[ 4 ] = total;   // == tiles
            PPC_LAUNCH( k_gs_resolve, total, GSB_THREADS, smem, st->stream, (const uint16_t *)st->gs_mcol, st->set[ st->cur ].pv, L, PR,
                        (const uint32_t *)st->gs_flags, st->gs_flags + 3, st->gs_done );
        } else
            for ( int colour = 0; colour < 4; colour++ ) {
                L.colour   = colour;
                L.ntx      = ( L.tiles_x - ( colour & 1 ) + 1 ) / 2;
                L.ty_start = ( ( colour >> 1 ) - L.ty_first ) & 1;
                const int nty = ( tiles_y - L.ty_start + 1 ) / 2;
                if ( L.ntx <= 0 || nty <= 0 )
                    continue;
                PPC_LAUNCH( k_gs_resolve, (unsigned)( L.ntx * nty ), GSB_THREADS, smem, st->stream, (const uint16_t *)st->gs_mcol, st->set[ st->cur ].pv, L, PR,
                            (const uint32_t *)st->gs_flags, st->gs_flags + 3, st->gs_done );
            }
    }
    st->gs_pending      = true;
    st->gs_pending_n    = n;
    st->gs_pending_grid = G;
    st->gs_pending_dt   = dt;
}

static void resolve_exact( Store *st, const size_t n, const GridDims &G, const float dt, const int mode );

// Look at the verdict of the last "resolve"=3 substep; called before anything else touches the state (the next substep, a copy-back,
// an exporter ...: store_of and collide_step).  Nothing has been queued behind that substep's velocity kernels yet.
static void gs_settle( Store *st ) {
    if ( !st->gs_pending )
        return;
    st->gs_pending = false;
    PPC_CK( cudaEventSynchronize( st->ev_gs ) );
    if ( st->gs_tiles )
        st->gs_pairs_per_tile = (uint32_t)( st->h_gs_flag[ 2 ] / st->gs_tiles );   // (units include the chunk headers: fine for a threshold)
    if ( *st->h_gs_flag == 0u )
        return;
    if ( st->slab )    // a rank cannot fall back on its own: the single-GPU run it must equal would have done so for the whole field
        return;        // (the flag stays set; the next synchronising call reports it)
    // Some sub-tile overflowed (a region far denser than the rest: an impact front, a clump of overlapping particles).  That substep
    // runs on the exact resolve instead -- for the WHOLE field, so the result is a function of the state alone: positions back to what
    // the pair search saw (the velocities are untouched: the velocity kernels saw the flag and did nothing), then the reference's order.
    const size_t n = st->gs_pending_n;
    st->gs_fallbacks++;
    for ( int b = 0; b < 4; b++ )
        if ( ( *st->h_gs_flag >> b ) & 1u )
            st->gs_fallback_reason[ b ]++;
    PPC_CK( cudaMemsetAsync( st->gs_flags, 0, sizeof( uint32_t ), st->stream ) );
    PPC_LAUNCH( k_rec_to_positions, div_up( n, STREAM_THREADS ), STREAM_THREADS, 0, st->stream, (const float4 *)st->rec_a, st->set[ st->cur ].pv, n );
    resolve_exact( st, n, st->gs_pending_grid, st->gs_pending_dt, 1 );
}

static void resolve_stage( Store *st, const size_t n, const GridDims &G, const float dt ) {
    if ( st->opt_resolve == 3 )
        resolve_gs( st, n, G, dt );
    else
        resolve_exact( st, n, G, dt, st->opt_resolve );
}

// mode: the "resolve" option (1 exact wavefront, 2 Jacobi, 0 none)
static void resolve_exact( Store *st, const size_t n, const GridDims &G, const float dt, const int mode ) {
    ParticleSet &S = st->set[ st->cur ], &O = st->set[ st->cur ^ 1 ];
    if ( mode ) {
        const bool jacobi = mode == 2;
        if ( !jacobi ) {
            ensure_wave_buffers( st );
            PPC_CK( cudaMemsetAsync( st->wave.counters, 0, 3 * sizeof( uint32_t ), st->stream ) );
            PPC_CK( cudaMemsetAsync( st->wave.counters + 4, 0, 9 * sizeof( uint32_t ), st->stream ) );
        }
        WaveBuffers WB = st->wave;
        for ( int attempt = 0; attempt < 2; attempt++ ) {
            StageTimer T( st, ST_COLLIDE );                                            // K7: narrow phase + position pushes
            const unsigned blocks = div_up( n, COLLIDE_THREADS );
            if ( jacobi && st->opt_stash )
                PPC_LAUNCH( ( k_collide_global<COLLIDE_STASH, true> ), blocks, COLLIDE_THREADS, 0, st->stream, S, O.pv, st->cell_start, n, G, dt, WB );
            else if ( jacobi )
                PPC_LAUNCH( ( k_collide_global<0, true> ), blocks, COLLIDE_THREADS, 0, st->stream, S, O.pv, st->cell_start, n, G, dt, WB );
            else if ( st->opt_collide != 1 && st->opt_stash ) {   // shared-memory tile kernels
                if ( !st->tile_attr_set ) {   // function attributes are per device: once per store
                    PPC_CK( cudaFuncSetAttribute( k_collide_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof( TileSmem ) ) );
                    PPC_CK( cudaFuncSetAttribute( k_collide_tile2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof( Tile2Smem ) ) );
                    st->tile_attr_set = true;
                }
                const bool   bins     = st->opt_collide != 3;   // fine x-bins (default) or the plain 3x3 scan
                const double per_cell = (double)n / (double)G.cells;
                int          wc       = (int)( (double)( bins ? TILE2_TARGET : TILE_TARGET ) / ( TILE_R * ( per_cell > 0.05 ? per_cell : 0.05 ) ) );
                const int    wc_max   = bins ? TILE2_MAXC - 2 : TILE_MAX_WC;
                wc                    = wc < 4 ? 4 : ( wc > wc_max ? wc_max : wc );
                const int tiles_x = ( G.cols + wc - 1 ) / wc, tiles_y = ( G.rows + TILE_R - 1 ) / TILE_R;
                if ( bins )
                    PPC_LAUNCH( k_collide_tile2, (unsigned)( tiles_x * tiles_y ), TILE_THREADS, sizeof( Tile2Smem ), st->stream, S, O.pv,
                                st->cell_start, G, dt, WB, wc, tiles_x );
                else
                    PPC_LAUNCH( k_collide_tile, (unsigned)( tiles_x * tiles_y ), TILE_THREADS, sizeof( TileSmem ), st->stream, S, O.pv,
                                st->cell_start, G, dt, WB, wc, tiles_x );
            } else if ( st->opt_stash )
                PPC_LAUNCH( ( k_collide_global<COLLIDE_STASH, false> ), blocks, COLLIDE_THREADS, 0, st->stream, S, O.pv, st->cell_start, n, G, dt, WB );
            else
                PPC_LAUNCH( ( k_collide_global<0, false> ), blocks, COLLIDE_THREADS, 0, st->stream, S, O.pv, st->cell_start, n, G, dt, WB );
            // First grid build after particles were uploaded: look at once whether some contact list did not fit the stored
            // positions (many particles spawned on one spot); if so lengthen the lists and redo the narrow phase now rather
            // than resolve this substep through the slow on-demand path.  (Later on the same check runs a substep late, async.)
            if ( jacobi || attempt == 1 || !st->lists_unverified )
                break;
            st->lists_unverified = false;
            uint32_t seen = 0;
            PPC_CK( cudaMemcpyAsync( &seen, st->wave.counters + 12, sizeof( uint32_t ), cudaMemcpyDeviceToHost, st->stream ) );
            PPC_CK( cudaStreamSynchronize( st->stream ) );
            if ( seen <= st->wave.maxdeg )
                break;
            const uint32_t before = st->wave.maxdeg;
            st->h_active[ 4 ]     = seen;
            grow_contact_lists( st );
            if ( st->wave.maxdeg == before )
                break;   // no room to lengthen them
            WB = st->wave;
            PPC_CK( cudaMemsetAsync( st->wave.counters + 8, 0, 5 * sizeof( uint32_t ), st->stream ) );
        }
        if ( !jacobi ) {
            StageTimer T( st, ST_WAVEFRONT );                                          // K7b: velocities, in the reference's pair order
            float4         *pv_out = O.pv;
            const uint32_t *cs     = st->cell_start;
            // tile passes pay off when most particles are in contact (deep dependency chains); sparse contact sets
            // are cheaper on the compact queues of the global rounds.  The count is a few substeps old.
            take_activity_report( st );
            const bool      dense  = *st->h_active == 0xffffffffu || (double)*st->h_active > 0.25 * (double)n;
            const int       passes = ( st->opt_wave_passes < 0 ) ? -st->opt_wave_passes : ( dense ? st->opt_wave_passes : 0 );
            PPC_CK( cudaMemcpyAsync( st->h_active_dma, st->wave.counters + 8, 5 * sizeof( uint32_t ), cudaMemcpyDeviceToHost, st->stream ) );
            st->wave_hops = passes > 0 ? 0 : 1;
            if ( passes > 0 ) {   // shared-memory tile passes over four half-shifted tilings
                if ( !st->wtile_attr_set ) {
                    PPC_CK( cudaFuncSetAttribute( k_wave_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof( WaveTileSmem ) ) );
                    st->wtile_attr_set = true;
                }
                const int    tw  = WTILE_TW;   // fixed; dense tiles subdivide themselves (see k_wave_plan)
                const int    txn = ( G.cols + tw / 2 + tw - 1 ) / tw, tyn = ( G.rows + tw / 2 + tw - 1 ) / tw;
                const size_t tiles = (size_t)txn * tyn;
                // at most f*f items per tile, and a tile only subdivides when it holds many particles
                size_t       stride = tiles + n / 90 + 64;
                if ( stride > tiles * 25 )
                    stride = tiles * 25;
                if ( stride > st->wave_items_stride ) {
                    PPC_CK( cudaStreamSynchronize( st->stream ) );
                    cudaFree( st->wave_items );
                    st->wave_items        = dev_alloc<WaveItem>( stride * 4 );
                    st->wave_items_stride = stride;
                }
                if ( !st->wave_n_items )
                    st->wave_n_items = dev_alloc<uint32_t>( 8 );
                const int np = passes > 4 ? 4 : passes;   // the four tilings repeat after four passes
                PPC_CK( cudaMemsetAsync( st->wave_n_items, 0, 8 * sizeof( uint32_t ), st->stream ) );
                PPC_LAUNCH( k_wave_plan, div_up( tiles * np * 32, 256 ), 256, 0, st->stream, cs, G, tw, np, st->wave_items,
                            (uint32_t)st->wave_items_stride, st->wave_n_items, st->wave.counters + 5 );
                for ( int pass = 0; pass < passes; pass++ )
                    PPC_LAUNCH( k_wave_tile, (unsigned)st->wave_items_stride, WTILE_THREADS, sizeof( WaveTileSmem ), st->stream, S, pv_out, cs,
                                G, dt, WB, tw, pass & 3, st->wave_items, (uint32_t)st->wave_items_stride, st->wave_n_items,
                                pass == passes - 1 ? 1 : 0 );
            }
            size_t n_arg = n;
            // packed records pay off when most particles are in contact; sparse contact sets are cheaper on the polling
            // kernel's compact per-CTA queues (no pack pass).  "wave_mode": 1 = choose by density (default), 2 = always events.
            if ( ( st->opt_wave_mode == 2 || ( st->opt_wave_mode == 1 && dense ) ) && passes == 0 ) {
                const WaveRecs R = st->recs;
                st->wave_hops    = 2;   // a thread fires a star of pairs around one particle per round (chain following)
                PPC_LAUNCH( k_wave_pack, div_up( n, 256 ), 256, 0, st->stream, S, (const float4 *)pv_out, WB, R, n );
                int       grid = st->events_grid;
                const int want = (int)div_up( n, EV_THREADS );
                if ( grid > want )
                    grid = want;
                // Chain following trades rounds for divergence: on large fields a round is throughput bound and short chains win
                // (16M pile: 7.9 / 6.6 / 6.5 / 6.9 / 7.2 ms with 0 / 1 / 2 / 4 / 6 pairs followed); a small, deeply dependent field
                // (the reference's debug scene: 1,500 rounds of ~18 us) is bound by the number of rounds and wants long ones.
                int   follow = n > ( (size_t)1 << 21 ) ? 2 : EV_FOLLOW;
                void *args[] = { (void *)&S, (void *)&pv_out, (void *)&cs, (void *)&n_arg, (void *)&G, (void *)&dt, (void *)&WB, (void *)&R, (void *)&follow };
                PPC_CK( cudaLaunchCooperativeKernel( (const void *)k_resolve_events, dim3( grid ), dim3( EV_THREADS ), args, 0, st->stream ) );
                ++g_launches;
            } else {   // polling rounds (after tile passes, or "wave_mode"=0)
                int       grid = st->wave_grid;
                const int want = (int)div_up( n, WAVE_THREADS );
                if ( grid > want )
                    grid = want;
                int   after   = passes > 0 ? 1 : 0;
                void *args[] = { (void *)&S, (void *)&pv_out, (void *)&cs, (void *)&n_arg, (void *)&G, (void *)&dt, (void *)&WB, (void *)&after };
                PPC_CK( cudaLaunchCooperativeKernel( (const void *)k_resolve_wavefront, dim3( grid ), dim3( WAVE_THREADS ), args, 0, st->stream ) );
                ++g_launches;
            }
        }
        float4 *t = S.pv;   // post-resolve state becomes current; the pre-resolve one stays readable in the other
        S.pv      = O.pv;   // set until the next substep (ppc_debug_pairs uses it)
        O.pv      = t;
    }
}

// One collisions check (reference src/particles_solver.c:197-393), with the pending particlesUpdate fused in.
static void collide_step( Store *st, uint16_t W, uint16_t H, float dt ) {
    gs_settle( st );   // the previous substep's verdict, before this one queues anything
    const size_t n = st->n;
    if ( n < 2 ) {   // reference :204-205: not even the wall pass runs
        flush_pending_update( st );
        return;
    }
    refresh_max_radius( st );
    GridDims G;
    G.cell_size = 2.0f * st->max_radius;                                // :244
    G.cols      = (int)ceilf( (float)W / G.cell_size );                 // :247
    G.rows      = (int)ceilf( (float)H / G.cell_size );                 // :248
    if ( G.cols < 1 || G.rows < 1 ) {
        fprintf( stderr, "particlephysicsc_b200: window %ux%u gives an empty grid\n", (unsigned)W, (unsigned)H );
        abort();
    }
    G.cells = (uint32_t)G.cols * (uint32_t)G.rows;
    G.row0  = 0;
    G.taint_lo = G.taint_hi = -1;
    ensure_cells( st, G.cells );
    ensure_scratch( st, radix_scratch_words( n ) + scan_scratch_words( (size_t)G.cells + 1 ) + scan_scratch_words( n + 1 ) + 64 );

    StepParams P    = {};
    P.do_integrate  = st->pending ? 1 : 0;
    P.dt_update     = st->pending_dt;
    P.do_colour     = ( st->pending && st->pending_colour ) ? 1 : 0;
    P.do_walls_keys = 1;
    P.do_histogram  = 1;
    P.wall_w        = (float)(int)W;
    P.wall_h        = (float)(int)H;
    P.grid          = G;
    P.defer_state   = st->opt_defer_state ? 1 : 0;   // K1 writes keys only; the gather applies the integrate + wall pass to what it moves
    if ( st->pending )
        st->col_valid = st->pending_colour;
    st->pending     = false;

    launch_integrate( st, P );                                                         // K1 (+K3 histogram)
    if ( st->opt_resolve == 3 )
        ensure_gs_buffers( st );
    grid_stage( st, n, G.cells, G.cols, 0, 0xffffffffu, P.defer_state ? &P : nullptr );
    st->grid       = G;
    st->grid_valid = true;
    resolve_stage( st, n, G, dt );
    st->pre_valid = true;
}

static void download( Store *st, struct Particles_t *p, unsigned mask, bool async = false ) {
    gs_settle( st );   // (the frame copy-back inside particlesCollisionsCheck_Grid comes here without passing store_of again)
    flush_pending_update( st );
    const size_t n = st->n < p->length ? st->n : p->length;
    if ( n && mask ) {
        StageTimer T( st, ST_DOWNLOAD );
        staging_reuse( st, async );
        PPC_LAUNCH( k_unpack_download, div_up( n, STREAM_THREADS ), STREAM_THREADS, 0, st->stream, st->set[ st->cur ], st->api_f[ 0 ],
                    st->api_f[ 1 ], st->api_f[ 2 ], st->api_f[ 3 ], st->api_col, n, mask );
        const cudaStream_t cs = copy_stream_after_unpack( st, async );
        if ( mask & PPC_DL_POS ) {
            PPC_CK( cudaMemcpyAsync( p->p_x, st->api_f[ 0 ], n * sizeof( float ), cudaMemcpyDeviceToHost, cs ) );
            PPC_CK( cudaMemcpyAsync( p->p_y, st->api_f[ 1 ], n * sizeof( float ), cudaMemcpyDeviceToHost, cs ) );
        }
        if ( mask & PPC_DL_VEL ) {
            PPC_CK( cudaMemcpyAsync( p->v_x, st->api_f[ 2 ], n * sizeof( float ), cudaMemcpyDeviceToHost, cs ) );
            PPC_CK( cudaMemcpyAsync( p->v_y, st->api_f[ 3 ], n * sizeof( float ), cudaMemcpyDeviceToHost, cs ) );
        }
        if ( mask & PPC_DL_COLOR )
            PPC_CK( cudaMemcpyAsync( p->colors, st->api_col, n * sizeof( uchar4 ), cudaMemcpyDeviceToHost, cs ) );
        copy_started( st, async );
    }
    if ( async )
        return;
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    check_device_flags( st );
}

static void *pinned_alloc( size_t bytes ) {
    void *h = nullptr;
    if ( cudaHostAlloc( &h, bytes ? bytes : 4, cudaHostAllocDefault ) != cudaSuccess || !h ) {
        fprintf( stderr, "Failed to initialize particles" );   // reference src/particles_arrays.c:149
        exit( EXIT_FAILURE );
    }
    memset( h, 0, bytes );
    return h;
}

}   // namespace ppc

using namespace ppc;

// =================================================================================================================
// Part 1: the reference's entry points
// =================================================================================================================
extern "C" struct Particles_t particlesCreate( size_t num_particles_init ) {
    select_device();
    struct Particles_t p;
    p.mass   = (float *)pinned_alloc( num_particles_init * sizeof( float ) );
    p.radius = (float *)pinned_alloc( num_particles_init * sizeof( float ) );
    p.colors = (Color *)pinned_alloc( num_particles_init * sizeof( Color ) );
    p.p_x    = (float *)pinned_alloc( num_particles_init * sizeof( float ) );
    p.p_y    = (float *)pinned_alloc( num_particles_init * sizeof( float ) );
    p.v_x    = (float *)pinned_alloc( num_particles_init * sizeof( float ) );
    p.v_y    = (float *)pinned_alloc( num_particles_init * sizeof( float ) );
    p.length = 0;
    p.size   = num_particles_init;
    Store *st     = new_store( &p, true );
    st->dirty_all = false;   // nothing to upload yet: length == 0
    return p;
}

extern "C" void particlesFree( struct Particles_t *particles ) {
    for ( size_t k = 0; k < g_stores.size(); k++ ) {
        Store *st = g_stores[ k ];
        if ( st->host_key != particles->mass )
            continue;
        PPC_CK( cudaSetDevice( st->device ) );
        PPC_CK( cudaStreamSynchronize( st->stream ) );
        copy_wait( st );
        cudaStreamDestroy( st->copy_stream );
        cudaEventDestroy( st->ev_unpacked );
        cudaEventDestroy( st->ev_copied );
        free_particle_arrays( st );
        cudaFree( st->cell_count );
        cudaFree( st->cell_start );
        cudaFree( st->scan_desc );
        cudaFree( st->d_word );
        cudaFree( st->slab_counts );
        if ( st->h_slab )
            cudaFreeHost( st->h_slab );
        cudaFreeHost( st->h_active_dma );
        free( st->h_active );
        cudaFree( st->wave_items );
        cudaFree( st->wave_n_items );
        cudaFree( st->gs_flags );
        if ( st->h_gs_flag ) {
            cudaFreeHost( st->h_gs_flag );
            cudaEventDestroy( st->ev_gs );
        }
        cudaEventDestroy( st->ev0 );
        cudaEventDestroy( st->ev1 );
        cudaStreamDestroy( st->stream );
        if ( st->owns_host ) {
            cudaFreeHost( particles->mass );
            cudaFreeHost( particles->radius );
            cudaFreeHost( particles->colors );
            cudaFreeHost( particles->p_x );
            cudaFreeHost( particles->p_y );
            cudaFreeHost( particles->v_x );
            cudaFreeHost( particles->v_y );
        }
        delete st;
        g_stores.erase( g_stores.begin() + k );
        return;
    }
    // not ours and never seen by the solver: behave like the reference (src/particles_arrays.c:134-143)
    free( particles->mass );
    free( particles->radius );
    free( particles->colors );
    free( particles->p_x );
    free( particles->p_y );
    free( particles->v_x );
    free( particles->v_y );
}

// reference src/particles_arrays.c:196-255: double the seven arrays
static void grow_arrays( Store *st, struct Particles_t *p ) {
    size_t new_size = (size_t)( (int)p->size ) * 2;   // ARRAY_GROWTH_FACTOR, includes/constants.h:25
    if ( new_size < 16 )
        new_size = 16;
    if ( !st->owns_host ) {
        fprintf( stderr, "particlephysicsc_b200: cannot grow arrays that particlesCreate did not allocate\n" );
        abort();
    }
    // an asynchronous copy-back may still be writing the host arrays: let it land before they are copied and freed
    copy_wait( st );
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    void  **slots[ 7 ] = { (void **)&p->mass, (void **)&p->radius, (void **)&p->colors, (void **)&p->p_x,
                           (void **)&p->p_y,  (void **)&p->v_x,    (void **)&p->v_y };
    for ( int a = 0; a < 7; a++ ) {
        void *fresh = pinned_alloc( new_size * 4 );
        memcpy( fresh, *slots[ a ], p->size * 4 );
        cudaFreeHost( *slots[ a ] );
        *slots[ a ] = fresh;
    }
    p->size      = new_size;
    st->host_key = p->mass;
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    resize_particle_arrays( st, new_size, st->n );
}

extern "C" void particlesAddParticle( struct Particles_t *particles, const Vector2 particle_i_position, const float particle_i_mass,
                                      const float particle_i_radius, uint8_t num_particles ) {
    Store *st = store_of( particles );
    // reference :260-262 grows once; keep growing until the write below fits (the reference would overrun)
    if ( particles->length + num_particles >= particles->size )
        grow_arrays( st, particles );
    while ( particles->length + num_particles >= particles->size )
        grow_arrays( st, particles );
    const Color blue = { 0, 121, 241, 255 };   // Raylib BLUE
    while ( num_particles-- ) {                // reference :265-277
        particles->length += 1;
        const size_t index        = particles->length;
        particles->mass[ index ]   = particle_i_mass;
        particles->radius[ index ] = particle_i_radius;
        particles->colors[ index ] = blue;
        particles->p_y[ index ]    = (float)particle_i_position.y;
        particles->v_x[ index ]    = 0.0f;
        particles->v_y[ index ]    = 0.0f;
        particles->p_x[ index ]    = (float)particle_i_position.x;
    }
    // the device picks the new slots up at the next solver call (absorb_host_changes)
}

extern "C" void particlesRemoveParticle( struct Particles_t *particles ) {
    Store *st = store_of( particles );
    // The particle leaving the processed range keeps its latest state in the host arrays (it re-enters on the
    // next add, reference :266-269), so bring the host up to date first, then rebuild the mirror from it.
    if ( !st->dirty_all )
        download( st, particles, PPC_DL_ALL );
    const Color  raywhite = { 245, 245, 245, 255 };
    const size_t index    = particles->length;   // reference :281-290
    particles->mass[ index ]   = 0.0f;
    particles->radius[ index ] = 0.0f;
    particles->colors[ index ] = raywhite;
    particles->p_y[ index ]    = 0.0f;
    particles->v_x[ index ]    = 0.0f;
    particles->v_y[ index ]    = 0.0f;
    particles->p_x[ index ]    = 0.0f;
    particles->length -= 1;
    st->dirty_all = true;
}

extern "C" void particlesUpdate( struct Particles_t *particles, const float time_step ) {
    Store *st = store_of( particles );
    absorb_host_changes( st, particles );
    flush_pending_update( st );   // two updates in a row: the first one runs on its own
    st->pending        = true;
    st->pending_dt     = time_step;
    st->pending_colour = st->sync_policy != PPC_SYNC_MANUAL || st->colour_wanted;
    st->colour_wanted  = false;
    if ( st->sync_policy == PPC_SYNC_EVERY_CALL )
        download( st, particles, auto_download_mask( st ) );
}

extern "C" void particlesCollisionsCheck_Grid( struct Particles_t *particles, const uint16_t window_width, const uint16_t window_height,
                                               float time_step ) {
    Store *st = store_of( particles );
    absorb_host_changes( st, particles );
    collide_step( st, window_width, window_height, time_step );
    if ( st->sync_policy != PPC_SYNC_MANUAL )
        download( st, particles, auto_download_mask( st ) );
}

// =================================================================================================================
// Part 2: harness extensions
// =================================================================================================================
extern "C" void ppc_set_device( int cuda_device ) { g_device = cuda_device; }
extern "C" void ppc_set_sync_policy( struct Particles_t *p, int policy ) { store_of( p )->sync_policy = policy; }
extern "C" void ppc_set_download_mask( struct Particles_t *p, unsigned mask ) { store_of( p )->dl_mask = mask; }
extern "C" void ppc_request_colors( struct Particles_t *p ) { store_of( p )->colour_wanted = true; }
extern "C" void ppc_host_modified( struct Particles_t *p ) { store_of( p )->dirty_all = true; }
extern "C" void ppc_download( struct Particles_t *p, unsigned mask ) {
    Store *st = store_of( p );
    copy_wait( st );
    absorb_host_changes( st, p );
    download( st, p, mask );
}
extern "C" void ppc_download_async( struct Particles_t *p, unsigned mask ) {
    Store *st = store_of( p );
    absorb_host_changes( st, p );
    download( st, p, mask, true );
}
extern "C" void ppc_download_wait( struct Particles_t *p ) { copy_wait( store_of( p ) ); }
extern "C" void ppc_synchronize( struct Particles_t *p ) {
    Store *st = store_of( p );
    copy_wait( st );
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    check_device_flags( st );
}

extern "C" void ppc_run_substeps( struct Particles_t *p, uint16_t window_width, uint16_t window_height, float time_step, int n,
                                  int color_last ) {
    Store *st = store_of( p );
    absorb_host_changes( st, p );
    for ( int s = 0; s < n; s++ ) {
        if ( st->n == 0 )   // reference src/main.c:92
            break;
        flush_pending_update( st );
        st->pending        = true;
        st->pending_dt     = time_step;
        st->pending_colour = color_last && s == n - 1;
        collide_step( st, window_width, window_height, time_step );
    }
}

extern "C" int ppc_set_option( struct Particles_t *p, const char *name, long long value ) {
    Store *st = store_of( p );
    if ( !strcmp( name, "sort" ) )
        st->opt_sort = (int)value;
    else if ( !strcmp( name, "stash" ) )
        st->opt_stash = (int)value;
    else if ( !strcmp( name, "resolve" ) )
        st->opt_resolve = (int)value;
    else if ( !strcmp( name, "collide" ) )
        st->opt_collide = (int)value;
    else if ( !strcmp( name, "wave_passes" ) )
        st->opt_wave_passes = (int)value;
    else if ( !strcmp( name, "wave_mode" ) )
        st->opt_wave_mode = (int)value;
    else if ( !strcmp( name, "slab_strict" ) )
        st->opt_slab_strict = (int)value;
    else if ( !strcmp( name, "slab_nb_lo_rows" ) )
        st->opt_nb_lo_rows = (int)value;
    else if ( !strcmp( name, "slab_nb_hi_rows" ) )
        st->opt_nb_hi_rows = (int)value;
    else if ( !strcmp( name, "gs_wc" ) )
        st->opt_gs_wc = (int)value;
    else if ( !strcmp( name, "gs_strip" ) )
        st->opt_gs_strip = (int)value;
    else if ( !strcmp( name, "gs_dump" ) )
        st->opt_gs_dump = (int)value;
    else if ( !strcmp( name, "gs_merge" ) )
        st->opt_gs_merge = (int)value;
    else if ( !strcmp( name, "gs_slice" ) )
        st->opt_gs_slice = (int)value;
    else if ( !strcmp( name, "slab_overlap" ) )
        st->opt_slab_overlap = (int)value;
    else if ( !strcmp( name, "slab_fused" ) )
        st->opt_slab_fused = (int)value;
    else if ( !strcmp( name, "scan_single" ) )
        st->opt_scan_single = (int)value;
    else if ( !strcmp( name, "cell_ranks" ) )
        st->opt_cell_ranks = (int)value;
    else if ( !strcmp( name, "defer_state" ) )
        st->opt_defer_state = (int)value;
    else
        return -1;
    return 0;
}

extern "C" int ppc_debug_grid_info( struct Particles_t *p, ppc_grid_info *out ) {
    Store *st = store_of( p );
    if ( !st->grid_valid )
        return -1;
    out->max_radius = st->max_radius;
    out->cell_size  = st->grid.cell_size;
    out->cols       = st->grid.cols;
    out->rows       = st->grid.rows;
    out->cells      = st->grid.cells;
    out->n          = st->n;
    return 0;
}

extern "C" int ppc_debug_keys( struct Particles_t *p, uint32_t *keys_by_index ) {
    Store *st = store_of( p );
    if ( !st->grid_valid )
        return -1;
    PPC_LAUNCH( k_export_keys, div_up( st->n, STREAM_THREADS ), STREAM_THREADS, 0, st->stream, st->set[ st->cur ], st->slot_of, st->n );
    PPC_CK( cudaMemcpyAsync( keys_by_index, st->slot_of, st->n * sizeof( uint32_t ), cudaMemcpyDeviceToHost, st->stream ) );
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    return 0;
}

extern "C" int ppc_debug_sorted_order( struct Particles_t *p, uint32_t *order ) {
    Store *st = store_of( p );
    if ( !st->grid_valid )
        return -1;
    PPC_CK( cudaMemcpyAsync( order, st->set[ st->cur ].gid, st->n * sizeof( uint32_t ), cudaMemcpyDeviceToHost, st->stream ) );
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    return 0;
}

extern "C" int ppc_debug_cell_start( struct Particles_t *p, uint32_t *cell_start ) {
    Store *st = store_of( p );
    if ( !st->grid_valid )
        return -1;
    PPC_CK( cudaMemcpyAsync( cell_start, st->cell_start, ( (size_t)st->grid.cells + 1 ) * sizeof( uint32_t ), cudaMemcpyDeviceToHost,
                             st->stream ) );
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    return 0;
}

extern "C" int64_t ppc_debug_pairs( struct Particles_t *p, int32_t *pairs, size_t max_pairs ) {
    Store *st = store_of( p );
    if ( !st->grid_valid || !st->pre_valid )
        return -1;
    const size_t n   = st->n;
    ParticleSet  pre = st->set[ st->cur ];
    if ( st->opt_resolve )
        pre.pv = st->set[ st->cur ^ 1 ].pv;   // positions the pair search saw (before the resolve moved anything)
    if ( st->opt_resolve == 3 )               // ... which in this mode live in the staged records
        PPC_LAUNCH( k_rec_to_pv, div_up( n, STREAM_THREADS ), STREAM_THREADS, 0, st->stream, (const float4 *)st->rec_a, pre.pv, n );
    uint32_t *count  = st->sort_keys_a;       // n + 1 words needed; sort_keys_a has cap >= n, use scratch when tight
    uint32_t *offset = st->sort_keys_b;
    uint32_t *tmp    = nullptr;
    if ( n + 1 > st->cap ) {
        const size_t padded = ( n + 1 + 63 ) & ~(size_t)63;   // the scan reads and writes 128-bit vectors: keep both bases aligned
        tmp    = dev_alloc<uint32_t>( 2 * padded );
        count  = tmp;
        offset = tmp + padded;
    }
    PPC_CK( cudaMemsetAsync( count, 0, ( n + 1 ) * sizeof( uint32_t ), st->stream ) );
    PPC_LAUNCH( k_pairs_count, div_up( n, COLLIDE_THREADS ), COLLIDE_THREADS, 0, st->stream, pre, st->cell_start, n, st->grid, count );
    exclusive_scan( count, offset, n + 1, st->scratch, st->stream );
    uint32_t total = 0;
    PPC_CK( cudaMemcpyAsync( &total, offset + n, sizeof( uint32_t ), cudaMemcpyDeviceToHost, st->stream ) );
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    if ( pairs && max_pairs && total ) {
        const size_t keep  = total < max_pairs ? total : max_pairs;
        int32_t     *d_out = (int32_t *)dev_alloc<int32_t>( 2 * keep );
        PPC_LAUNCH( k_pairs_emit, div_up( n, COLLIDE_THREADS ), COLLIDE_THREADS, 0, st->stream, pre, st->cell_start, n, st->grid, offset,
                    d_out, keep );
        PPC_CK( cudaMemcpyAsync( pairs, d_out, 2 * keep * sizeof( int32_t ), cudaMemcpyDeviceToHost, st->stream ) );
        PPC_CK( cudaStreamSynchronize( st->stream ) );
        cudaFree( d_out );
    }
    cudaFree( tmp );
    return (int64_t)total;
}

extern "C" int ppc_debug_wave_stats( struct Particles_t *p, unsigned long long *rounds, unsigned long long *visits ) {
    Store *st = store_of( p );
    if ( !st->wave.counters )
        return -1;
    uint32_t w[ 8 ];
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    PPC_CK( cudaMemcpy( w, st->wave.counters, sizeof( w ), cudaMemcpyDeviceToHost ) );
    *rounds = w[ 4 ];
    *visits = ( (unsigned long long)w[ 7 ] << 32 ) | w[ 6 ];
    return 0;
}

// The pairs the HOT narrow-phase kernel of the last substep found, as (i, j) API indices with i < j, in no particular order:
// "resolve"=1 rebuilds them from the contact lists k_collide_tile2 wrote for the wavefront; "resolve"=3 returns what
// k_collide_gs dumped while resolving (option "gs_dump" must be on during that substep).  Returns the number of pairs,
// -1 when there is nothing to report, -2 when a contact list was longer than the stored lists (pairs incomplete).
extern "C" int64_t ppc_debug_hot_pairs( struct Particles_t *p, int32_t *pairs, size_t max_pairs ) {
    Store *st = store_of( p );
    if ( !st->grid_valid || !st->pre_valid )
        return -1;
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    check_device_flags( st );
    if ( st->opt_resolve == 3 ) {
        if ( !st->gs_dump )
            return -1;
        uint32_t total = 0;
        PPC_CK( cudaMemcpy( &total, st->gs_flags + 1, sizeof( uint32_t ), cudaMemcpyDeviceToHost ) );
        if ( total > st->gs_dump_cap )
            return -2;
        const size_t keep = total < max_pairs ? total : max_pairs;
        if ( pairs && keep )
            PPC_CK( cudaMemcpy( pairs, st->gs_dump, 2 * keep * sizeof( int32_t ), cudaMemcpyDeviceToHost ) );
        return (int64_t)total;
    }
    if ( st->opt_resolve != 1 || !st->wave.deg )
        return -1;
    const size_t n = st->n;
    uint32_t    *d_cnt = dev_alloc<uint32_t>( 2 );
    int32_t     *d_out = dev_alloc<int32_t>( 2 * ( max_pairs ? max_pairs : 1 ) );
    PPC_CK( cudaMemsetAsync( d_cnt, 0, 2 * sizeof( uint32_t ), st->stream ) );
    PPC_LAUNCH( k_pairs_from_lists, div_up( n, COLLIDE_THREADS ), COLLIDE_THREADS, 0, st->stream, st->set[ st->cur ].gid, st->wave, n, d_out,
                (uint32_t)( max_pairs < 0xffffffffu ? max_pairs : 0xffffffffu ), d_cnt );
    uint32_t h[ 2 ] = { 0, 0 };
    PPC_CK( cudaMemcpyAsync( h, d_cnt, sizeof( h ), cudaMemcpyDeviceToHost, st->stream ) );
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    const size_t keep = h[ 0 ] < max_pairs ? h[ 0 ] : max_pairs;
    if ( pairs && keep )
        PPC_CK( cudaMemcpy( pairs, d_out, 2 * keep * sizeof( int32_t ), cudaMemcpyDeviceToHost ) );
    cudaFree( d_cnt );
    cudaFree( d_out );
    return h[ 1 ] ? -2 : (int64_t)h[ 0 ];
}

// Tile geometry of the last "resolve"=3 substep: out = {tile width in cells, tile height in cells, staged particles above which
// a tile is cut into strips, staged-particle capacity of a strip} -- the parameters of the oracle's model of that order.
extern "C" int ppc_debug_gs_params( struct Particles_t *p, int out[ 4 ] ) {
    Store *st = store_of( p );
    if ( st->opt_resolve != 3 || st->gs_wc == 0 )
        return -1;
    out[ 0 ] = st->gs_wc, out[ 1 ] = GS_R, out[ 2 ] = st->gs_strip, out[ 3 ] = GS_CAP;
    return 0;
}

// Substeps of this store on which "resolve"=3 fell back to the exact resolve because a sub-tile overflowed the tile kernel.
extern "C" unsigned long long ppc_gs_fallbacks( struct Particles_t *p ) { return store_of( p )->gs_fallbacks; }
// ... by cause (a substep may have several): a one-column strip of a tile held more particles than shared memory takes; more contacts
// in one sub-tile than its contact pool holds; more than 127 owned contacts in one cell or more than 32 pair-colours around one pair;
// pair-record buffer full.
extern "C" void ppc_gs_fallback_reasons( struct Particles_t *p, unsigned long long out[ 4 ] ) {
    const Store *st = store_of( p );
    for ( int b = 0; b < 4; b++ )
        out[ b ] = st->gs_fallback_reason[ b ];
}

// =================================================================================================================
// Part 3: multi-GPU slab decomposition (one store per rank; the caller moves the exchange buffers, e.g. with
// torch.distributed / NCCL send-recv between neighbouring ranks)
// =================================================================================================================
static GridDims world_grid( uint16_t W, uint16_t H, float max_radius ) {
    GridDims G;
    const float m = max_radius > 6.0f ? max_radius : 6.0f;            // reference src/particles_solver.c:231-236
    G.cell_size   = 2.0f * m;                                           // :244
    G.cols        = (int)ceilf( (float)W / G.cell_size );               // :247
    G.rows        = (int)ceilf( (float)H / G.cell_size );               // :248
    G.cells       = (uint32_t)G.cols * (uint32_t)G.rows;
    G.row0        = 0;
    G.taint_lo = G.taint_hi = -1;
    return G;
}

extern "C" int ppc_grid_dims( uint16_t window_width, uint16_t window_height, float max_radius, ppc_grid_info *out ) {
    const GridDims G = world_grid( window_width, window_height, max_radius );
    if ( G.cols < 1 || G.rows < 1 )
        return -1;
    out->max_radius = max_radius > 6.0f ? max_radius : 6.0f;
    out->cell_size  = G.cell_size;
    out->cols       = G.cols;
    out->rows       = G.rows;
    out->cells      = G.cells;
    out->n          = 0;
    return 0;
}

// "resolve"=3 on a slab.  The tile order is anchored at the whole field's cell (0, 0), so a rank computes the tiles of its window
// exactly as the single-GPU run does, PROVIDED the window holds every tile an owned particle's velocity depends on.  Tiles of even
// tile rows (colours 0, 1) depend on their own tile row plus one cell row either side; tiles of odd tile rows (colours 2, 3) on the
// complete tile rows above and below as well.  Returns the ghost rows that takes for the owned rows [own_row_begin, own_row_end).
extern "C" int ppc_slab_gs_halo( int own_row_begin, int own_row_end, int rows_total ) {
    int need = 1;
    if ( own_row_begin > 0 ) {
        const int t0 = own_row_begin / GS_R;
        const int first = ( t0 & 1 ) ? GS_R * t0 - GS_R - 1 : GS_R * t0 - 1;   // first cell row needed
        const int lo = own_row_begin - ( first > 0 ? first : 0 );
        need = lo > need ? lo : need;
    }
    if ( own_row_end < rows_total ) {
        const int t1 = ( own_row_end - 1 ) / GS_R;
        const int last = ( t1 & 1 ) ? GS_R * t1 + 2 * GS_R : GS_R * t1 + GS_R;   // last cell row needed
        const int hi = ( last < rows_total - 1 ? last : rows_total - 1 ) - ( own_row_end - 1 );
        need = hi > need ? hi : need;
    }
    return need;
}

// Tile width "resolve"=3 chooses for a whole field of `particles` particles on `cells` cells: what the ranks of a slab
// decomposition must all set as "gs_wc" to reproduce the single-GPU pair order.
extern "C" int ppc_gs_auto_tile_width( double particles, double cells ) {
    Store tmp;
    tmp.opt_gs_wc = gs_auto_width( particles, cells );
    GridDims G = {};
    G.cells = 1;
    return gs_tile_width( &tmp, 1, G );
}

extern "C" int ppc_slab_init( struct Particles_t *p, const uint32_t *global_ids, uint16_t window_width, uint16_t window_height,
                              float max_radius, int own_row_begin, int own_row_end, int halo_rows ) {
    Store         *st = store_of( p );
    const GridDims G  = world_grid( window_width, window_height, max_radius );
    if ( G.cols < 1 || G.rows < 1 || own_row_begin < 0 || own_row_end > G.rows || own_row_begin >= own_row_end || halo_rows < 1 || p->length > p->size )
        return -1;
    // Ghost rows come from the two neighbouring ranks only, and a neighbour sends what it OWNS: a halo deeper than a neighbour's
    // slab would leave the outer part of the window empty without anything noticing ("slab_nb_lo_rows" / "slab_nb_hi_rows").
    if ( ( own_row_begin > 0 && st->opt_nb_lo_rows > 0 && halo_rows > st->opt_nb_lo_rows ) ||
         ( own_row_end < G.rows && st->opt_nb_hi_rows > 0 && halo_rows > st->opt_nb_hi_rows ) )
        return -5;
    if ( p->size != st->cap ) {
        PPC_CK( cudaStreamSynchronize( st->stream ) );
        resize_particle_arrays( st, p->size, 0 );
    }
    st->slab        = true;
    st->slab_world  = G;
    st->slab_wall_w = (float)(int)window_width;
    st->slab_wall_h = (float)(int)window_height;
    st->max_radius  = max_radius > 6.0f ? max_radius : 6.0f;
    st->max_radius_dirty = false;
    st->slab_rows   = SlabRows{ own_row_begin, own_row_end, halo_rows, own_row_begin > 0 ? 1 : 0, own_row_end < G.rows ? 1 : 0, G.cols };
    if ( !st->slab_counts )
        st->slab_counts = dev_alloc<uint32_t>( 8 );
    if ( !st->h_slab )
        PPC_CK( cudaHostAlloc( (void **)&st->h_slab, 16 * sizeof( uint32_t ), cudaHostAllocDefault ) );
    const size_t n = p->length;
    copy_wait( st );
    if ( n ) {
        StageTimer T( st, ST_UPLOAD );
        float     *h[ 6 ];
        host_float_arrays( p, h );
        for ( int a = 0; a < 6; a++ )
            PPC_CK( cudaMemcpyAsync( st->api_f[ a ], h[ a ], n * sizeof( float ), cudaMemcpyHostToDevice, st->stream ) );
        PPC_CK( cudaMemcpyAsync( st->api_col, p->colors, n * sizeof( uchar4 ), cudaMemcpyHostToDevice, st->stream ) );
        PPC_CK( cudaMemcpyAsync( st->slot_of, global_ids, n * sizeof( uint32_t ), cudaMemcpyHostToDevice, st->stream ) );
        PPC_LAUNCH( k_slab_upload, div_up( n, STREAM_THREADS ), STREAM_THREADS, 0, st->stream, st->set[ st->cur ], st->api_f[ 0 ], st->api_f[ 1 ],
                    st->api_f[ 2 ], st->api_f[ 3 ], st->api_f[ 4 ], st->api_f[ 5 ], st->api_col, st->slot_of, n );
        PPC_CK( cudaStreamSynchronize( st->stream ) );
    }
    st->own_first  = 0;
    st->n_own      = n;
    st->n          = n;
    st->slab_violations = 0;
    st->dirty_all  = false;
    st->pending    = false;
    st->col_valid  = true;
    st->grid_valid = st->pre_valid = false;
    return 0;
}

// Did the last resolve keep the owned particles exact?  The resolve marks every particle whose result may differ from the
// whole-field run (ST_TAINT, csrc/ppc_collide.cuh) and reports when the mark reaches an owned particle.  Stream must be idle.
static int slab_halo_ok( Store *st ) {
    check_gs_flags( st );
    if ( !st->wave.counters )
        return 1;
    uint32_t w[ 12 ];   // one read: error flags [3], taint reached an owned particle [10], rows it travelled [11]
    PPC_CK( cudaMemcpy( w, st->wave.counters, sizeof( w ), cudaMemcpyDeviceToHost ) );
    if ( w[ 3 ] ) {
        fprintf( stderr, "particlephysicsc_b200: resolve failed (%s%s)\n", ( w[ 3 ] & 1u ) ? "a particle has more than 4095 simultaneous contacts " : "",
                 ( w[ 3 ] & 2u ) ? "pair dependency chain too long" : "" );
        abort();
    }
    if ( st->opt_resolve != 1 || !( st->slab_rows.has_lo || st->slab_rows.has_hi ) )
        return 1;
    if ( st->wave_hops == 0 )
        return 0;   // shared-memory tile passes do not carry the mark
    st->slab_taint_rows = w[ 11 ];
    return w[ 10 ] == 0u ? 1 : 0;
}

extern "C" int ppc_slab_begin( struct Particles_t *p, float time_step, int colour, uint64_t send_counts[ 2 ] ) {
    Store *st = store_of( p );
    if ( !st->slab )
        return -1;
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    if ( st->grid_valid && !slab_halo_ok( st ) ) {   // the inexactness mark reached an owned particle in the substep just finished
        st->slab_violations++;
        if ( st->opt_slab_strict )
            return -2;
    }
    const size_t n = st->n_own;
    const ParticleSet V = offset_view( st->set[ st->cur ], st->own_first );
    StepParams   P = {};
    P.do_integrate = 1, P.dt_update = time_step, P.do_colour = colour ? 1 : 0, P.do_walls_keys = 1, P.do_histogram = 0;
    P.wall_w = st->slab_wall_w, P.wall_h = st->slab_wall_h, P.grid = st->slab_world;
    st->slab_fused_now = st->opt_slab_fused != 0 && st->opt_sort == 0;
    if ( st->slab_fused_now ) {
        // One pass over the owned particles: integrate, walls, window-local key + histogram + rank, neighbour counts and pack lists
        const SlabRows &R      = st->slab_rows;
        const int       row_lo = R.own_begin - R.halo > 0 ? R.own_begin - R.halo : 0;
        const int       row_hi = R.own_end + R.halo < st->slab_world.rows ? R.own_end + R.halo : st->slab_world.rows;
        const size_t    bins   = (size_t)st->slab_world.cols * (size_t)( row_hi - row_lo ) + 1;   // + the drop bin
        ensure_cells( st, bins );
        StageTimer T( st, ST_INTEGRATE );
        PPC_CK( cudaMemsetAsync( st->slab_counts, 0, 8 * sizeof( uint32_t ), st->stream ) );
        if ( !st->hist_clean )
            PPC_CK( cudaMemsetAsync( st->cell_count, 0, st->cells_cap * sizeof( uint32_t ), st->stream ) );
        st->hist_clean = false;
        // Border slots first (within 2 * halo rows of a border at the last grid build; everything before the first grid build): their
        // counts are what the host waits for, and their records can travel while ppc_slab_interior integrates the rest.
        size_t b_lo = n, b_hi = n;   // owned slots [0, b_lo) and [b_hi, n) are the border parts
        if ( st->opt_slab_overlap && st->grid_valid && ( R.has_lo || R.has_hi ) && st->slab_band_lo <= st->slab_band_hi && st->slab_band_hi <= n )
            b_lo = R.has_lo ? st->slab_band_lo : 0, b_hi = R.has_hi ? st->slab_band_hi : n;
        st->slab_int_first = b_lo, st->slab_int_count = b_hi > b_lo ? b_hi - b_lo : 0;
        st->slab_int_P = P, st->slab_int_row_lo = row_lo, st->slab_int_row_hi = row_hi, st->slab_int_drop = (uint32_t)( bins - 1 );
        const size_t parts[ 2 ][ 2 ] = { { 0, b_lo < n ? b_lo : n }, { b_hi, b_hi < n ? n - b_hi : 0 } };
        for ( int part = 0; part < 2; part++ )
            if ( parts[ part ][ 1 ] )
                PPC_LAUNCH( k_slab_integrate, div_up( parts[ part ][ 1 ], STREAM_THREADS ), STREAM_THREADS, 0, st->stream, V, parts[ part ][ 0 ],
                            parts[ part ][ 1 ], P, R, row_lo, row_hi, (uint32_t)( bins - 1 ), st->cell_count, st->sort_keys_a + st->own_first,
                            st->slab_counts, st->sort_keys_b, st->sort_vals_b, 0 );
        PPC_CK( cudaMemcpyAsync( st->h_slab, st->slab_counts, 4 * sizeof( uint32_t ), cudaMemcpyDeviceToHost, st->stream ) );
        PPC_CK( cudaStreamSynchronize( st->stream ) );
        st->col_valid = colour != 0;
    } else {
        if ( n ) {
            StageTimer T( st, ST_INTEGRATE );
            PPC_LAUNCH( k_integrate_walls_keys, div_up( n, STREAM_THREADS ), STREAM_THREADS, 0, st->stream, V, st->cell_count, n, P, (uint32_t *)nullptr );
        }
        st->col_valid = colour != 0;
        StageTimer T( st, ST_SLAB );
        PPC_CK( cudaMemsetAsync( st->slab_counts, 0, 8 * sizeof( uint32_t ), st->stream ) );
        if ( n )
            PPC_LAUNCH( k_slab_count, div_up( n, STREAM_THREADS ), STREAM_THREADS, 0, st->stream, V.key, n, st->slab_rows, st->slab_counts );
        PPC_CK( cudaMemcpyAsync( st->h_slab, st->slab_counts, 4 * sizeof( uint32_t ), cudaMemcpyDeviceToHost, st->stream ) );
        PPC_CK( cudaStreamSynchronize( st->stream ) );
    }
    if ( st->h_slab[ 2 ] ) {
        fprintf( stderr, "particlephysicsc_b200: %u particles moved more than the %d halo rows in one substep\n", st->h_slab[ 2 ], st->slab_rows.halo );
        abort();
    }
    send_counts[ 0 ] = st->h_slab[ 0 ];
    send_counts[ 1 ] = st->h_slab[ 1 ];
    st->pre_valid    = false;
    return 0;
}

// The rest of the fused first pass: the owned particles that cannot reach a send zone.  Called by the host once the packed records
// are on their way (so the transfer overlaps this pass), or by ppc_slab_finish if the host did not.
static void slab_interior( Store *st ) {
    if ( !st->slab_int_count )
        return;
    StageTimer T( st, ST_INTEGRATE );
    const ParticleSet V = offset_view( st->set[ st->cur ], st->own_first );
    PPC_LAUNCH( k_slab_integrate, div_up( st->slab_int_count, STREAM_THREADS ), STREAM_THREADS, 0, st->stream, V, st->slab_int_first, st->slab_int_count,
                st->slab_int_P, st->slab_rows, st->slab_int_row_lo, st->slab_int_row_hi, st->slab_int_drop, st->cell_count,
                st->sort_keys_a + st->own_first, st->slab_counts, st->sort_keys_b, st->sort_vals_b, 1 );
    st->slab_int_count = 0;
}

extern "C" int ppc_slab_interior( struct Particles_t *p ) {
    Store *st = store_of( p );
    if ( !st->slab )
        return -1;
    slab_interior( st );
    return 0;
}

extern "C" int ppc_slab_pack( struct Particles_t *p, void *send_lo, void *send_hi, int with_header ) {
    Store *st = store_of( p );
    if ( !st->slab )
        return -1;
    const size_t n = st->n_own;
    StageTimer   T( st, ST_SLAB );
    SlabRec     *lo = (SlabRec *)send_lo, *hi = (SlabRec *)send_hi;
    if ( with_header ) {   // fixed-size messages: record 0 carries the number of records that follow
        if ( st->slab_rows.has_lo )
            PPC_CK( cudaMemcpyAsync( lo, st->h_slab + 0, sizeof( uint32_t ), cudaMemcpyHostToDevice, st->stream ) );
        if ( st->slab_rows.has_hi )
            PPC_CK( cudaMemcpyAsync( hi, st->h_slab + 1, sizeof( uint32_t ), cudaMemcpyHostToDevice, st->stream ) );
        lo += 1, hi += 1;
    }
    if ( st->slab_fused_now ) {   // the first pass listed the slots to send
        const ParticleSet V = offset_view( st->set[ st->cur ], st->own_first );
        if ( st->h_slab[ 0 ] )
            PPC_LAUNCH( k_slab_pack_list, div_up( st->h_slab[ 0 ], STREAM_THREADS ), STREAM_THREADS, 0, st->stream, V, (const uint32_t *)st->sort_keys_b,
                        (size_t)st->h_slab[ 0 ], lo );
        if ( st->h_slab[ 1 ] )
            PPC_LAUNCH( k_slab_pack_list, div_up( st->h_slab[ 1 ], STREAM_THREADS ), STREAM_THREADS, 0, st->stream, V, (const uint32_t *)st->sort_vals_b,
                        (size_t)st->h_slab[ 1 ], hi );
    } else if ( n && ( st->h_slab[ 0 ] || st->h_slab[ 1 ] ) )
        PPC_LAUNCH( k_slab_pack, div_up( n, STREAM_THREADS ), STREAM_THREADS, 0, st->stream, offset_view( st->set[ st->cur ], st->own_first ), n,
                    st->slab_rows, lo, hi, st->slab_counts + 4 );
    return 0;
}

extern "C" int64_t ppc_slab_finish( struct Particles_t *p, float time_step, const void *recv_lo, uint64_t n_lo, const void *recv_hi, uint64_t n_hi,
                                    int with_header, uint64_t received[ 2 ] ) {
    Store *st = store_of( p );
    if ( !st->slab )
        return -1;
    const SlabRows &R     = st->slab_rows;
    const GridDims &GW    = st->slab_world;
    const size_t    n_loc = st->n_own + n_lo + n_hi;
    if ( st->own_first + n_loc > st->cap ) {
        fprintf( stderr, "particlephysicsc_b200: slab store too small: %zu owned + %llu received particles at offset %zu, capacity %zu\n", st->n_own,
                 (unsigned long long)( n_lo + n_hi ), st->own_first, st->cap );
        abort();
    }
    if ( st->opt_resolve == 3 && ( st->opt_gs_wc <= 0 || R.halo < ppc_slab_gs_halo( R.own_begin, R.own_end, GW.rows ) ) )
        return -4;   // the tile order must be the whole field's ("gs_wc") and the window must hold every tile an owned particle depends on
    const int row_lo = R.own_begin - R.halo > 0 ? R.own_begin - R.halo : 0;
    const int row_hi = R.own_end + R.halo < GW.rows ? R.own_end + R.halo : GW.rows;
    GridDims  G;
    G.cell_size = GW.cell_size, G.cols = GW.cols, G.rows = row_hi - row_lo, G.cells = (uint32_t)G.cols * (uint32_t)G.rows, G.row0 = row_lo;
    G.taint_lo = row_lo > 0 ? 0 : -1;                    // the window's first / last row misses contacts beyond the window
    G.taint_hi = row_hi < GW.rows ? G.rows - 1 : -1;
    const size_t bins = (size_t)G.cells + 1;   // + the drop bin
    ensure_cells( st, bins );
    ensure_scratch( st, radix_scratch_words( n_loc ) + scan_scratch_words( bins + 1 ) + scan_scratch_words( n_loc + 1 ) + 64 );
    if ( st->opt_resolve == 3 )
        ensure_gs_buffers( st );
    slab_interior( st );   // (if the host has not started it already)
    const ParticleSet V = offset_view( st->set[ st->cur ], st->own_first );
    {
        StageTimer T( st, ST_SLAB );
        if ( n_lo )
            PPC_LAUNCH( k_slab_append, div_up( n_lo, STREAM_THREADS ), STREAM_THREADS, 0, st->stream, V, st->n_own, (const SlabRec *)recv_lo, (size_t)n_lo, GW,
                        with_header );
        if ( n_hi )
            PPC_LAUNCH( k_slab_append, div_up( n_hi, STREAM_THREADS ), STREAM_THREADS, 0, st->stream, V, st->n_own + n_lo, (const SlabRec *)recv_hi, (size_t)n_hi, GW,
                        with_header );
        st->h_slab[ 3 ] = (uint32_t)n_lo, st->h_slab[ 7 ] = (uint32_t)n_hi;
        if ( with_header && n_lo )
            PPC_CK( cudaMemcpyAsync( st->h_slab + 3, recv_lo, sizeof( uint32_t ), cudaMemcpyDeviceToHost, st->stream ) );
        if ( with_header && n_hi )
            PPC_CK( cudaMemcpyAsync( st->h_slab + 7, recv_hi, sizeof( uint32_t ), cudaMemcpyDeviceToHost, st->stream ) );
        if ( st->slab_fused_now ) {   // the owned particles already carry local keys and ranks and are in the histogram: only what arrived is left
            const size_t n_rec = n_lo + n_hi;
            if ( n_rec )
                PPC_LAUNCH( k_slab_localize, div_up( n_rec, STREAM_THREADS ), STREAM_THREADS, 0, st->stream, V.key + st->n_own, n_rec, G.cols, row_lo, row_hi,
                            G.cells, st->cell_count, st->sort_keys_a + st->own_first + st->n_own );
            st->ranks_valid = true;
        } else {
            PPC_CK( cudaMemsetAsync( st->cell_count, 0, ( bins + 1 ) * sizeof( uint32_t ), st->stream ) );
            if ( n_loc )
                PPC_LAUNCH( k_slab_localize, div_up( n_loc, STREAM_THREADS ), STREAM_THREADS, 0, st->stream, V.key, n_loc, G.cols, row_lo, row_hi, G.cells,
                            st->cell_count, (uint32_t *)nullptr );
            st->ranks_valid = false;
        }
    }
    if ( n_loc )
        grid_stage( st, n_loc, bins, G.cols, st->own_first, G.cells );
    else {
        PPC_CK( cudaMemsetAsync( st->cell_start, 0, ( bins + 1 ) * sizeof( uint32_t ), st->stream ) );
        st->cur ^= 1;
    }
    // owned rows and the window end, as slot numbers of the sorted storage
    const size_t c_own0 = (size_t)( R.own_begin - row_lo ) * G.cols, c_own1 = (size_t)( R.own_end - row_lo ) * G.cols;
    PPC_CK( cudaMemcpyAsync( st->h_slab + 4, st->cell_start + c_own0, sizeof( uint32_t ), cudaMemcpyDeviceToHost, st->stream ) );
    PPC_CK( cudaMemcpyAsync( st->h_slab + 5, st->cell_start + c_own1, sizeof( uint32_t ), cudaMemcpyDeviceToHost, st->stream ) );
    PPC_CK( cudaMemcpyAsync( st->h_slab + 6, st->cell_start + G.cells, sizeof( uint32_t ), cudaMemcpyDeviceToHost, st->stream ) );
    // end of the first owned row / begin of the last owned row (taint scan, below)
    PPC_CK( cudaMemcpyAsync( st->h_slab + 8, st->cell_start + c_own0 + G.cols, sizeof( uint32_t ), cudaMemcpyDeviceToHost, st->stream ) );
    PPC_CK( cudaMemcpyAsync( st->h_slab + 9, st->cell_start + c_own1 - G.cols, sizeof( uint32_t ), cudaMemcpyDeviceToHost, st->stream ) );
    // first slot of the rows 2 * halo inside either border: what the next substep's first pass treats as border / interior
    const int    band_lo_row = R.own_begin + 2 * R.halo < R.own_end ? R.own_begin + 2 * R.halo : R.own_end;
    const int    band_hi_row = R.own_end - 2 * R.halo > R.own_begin ? R.own_end - 2 * R.halo : R.own_begin;
    PPC_CK( cudaMemcpyAsync( st->h_slab + 10, st->cell_start + (size_t)( band_lo_row - row_lo ) * G.cols, sizeof( uint32_t ), cudaMemcpyDeviceToHost, st->stream ) );
    PPC_CK( cudaMemcpyAsync( st->h_slab + 11, st->cell_start + (size_t)( band_hi_row - row_lo ) * G.cols, sizeof( uint32_t ), cudaMemcpyDeviceToHost, st->stream ) );
    PPC_CK( cudaMemcpyAsync( st->h_slab + 12, st->slab_counts + 2, sizeof( uint32_t ), cudaMemcpyDeviceToHost, st->stream ) );
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    if ( st->h_slab[ 12 ] ) {   // (the interior part of the first pass reports here)
        fprintf( stderr, "particlephysicsc_b200: %u particles moved more than the %d halo rows in one substep\n", st->h_slab[ 12 ], R.halo );
        abort();
    }
    st->slab_band_lo = st->h_slab[ 10 ] - st->h_slab[ 4 ], st->slab_band_hi = st->h_slab[ 11 ] - st->h_slab[ 4 ];
    const size_t n_work = st->h_slab[ 6 ];
    if ( received )
        received[ 0 ] = st->h_slab[ 3 ], received[ 1 ] = st->h_slab[ 7 ];
    if ( with_header && ( st->h_slab[ 3 ] > n_lo || st->h_slab[ 7 ] > n_hi ) )
        return -3;   // a neighbour had more records than the agreed message size: what arrived is incomplete
    st->own_first  = st->h_slab[ 4 ];
    st->n_own      = st->h_slab[ 5 ] - st->h_slab[ 4 ];
    st->n          = n_work;
    st->grid       = G;
    st->grid_valid = true;
    st->row_lo = row_lo, st->row_hi = row_hi;
    st->slab_own_first = (uint32_t)st->own_first, st->slab_own_end = (uint32_t)( st->own_first + st->n_own );
    if ( n_work >= 1 )
        resolve_stage( st, n_work, G, time_step );
    if ( n_work >= 1 && st->opt_resolve == 1 && st->wave_hops > 0 && ( G.taint_lo >= 0 || G.taint_hi >= 0 ) ) {
        StageTimer     T( st, ST_SLAB );
        const uint32_t lo_end = G.taint_lo >= 0 ? st->h_slab[ 8 ] : 0u, hi_begin = G.taint_hi >= 0 ? st->h_slab[ 9 ] : (uint32_t)n_work;
        const uint32_t items  = lo_end + ( (uint32_t)n_work - ( hi_begin > lo_end ? hi_begin : lo_end ) );
        const bool     events = st->wave_hops == 2;   // which engine ran, hence where the state words are
        if ( items )
            PPC_LAUNCH( k_slab_taint_scan, div_up( items, STREAM_THREADS ), STREAM_THREADS, 0, st->stream,
                        events ? &st->recs.wr->state : st->wave.state, events ? (uint32_t)( sizeof( WaveRec ) / 4 ) : 1u, st->wave.deg,
                        st->set[ st->cur ].key, G, lo_end, hi_begin > lo_end ? hi_begin : lo_end, (uint32_t)n_work, st->slab_own_first,
                        st->slab_own_end, st->wave.counters );
    }
    st->pre_valid = true;
    p->length     = st->n_own <= p->size ? st->n_own : p->size;
    return (int64_t)st->n_own;
}

static int slab_download( Store *st, struct Particles_t *p, uint32_t *global_ids, unsigned mask, bool async ) {
    const size_t n = st->n_own;
    if ( n > p->size )
        return -3;
    if ( n && mask ) {
        StageTimer T( st, ST_DOWNLOAD );
        staging_reuse( st, async );
        PPC_LAUNCH( k_slab_download, div_up( n, STREAM_THREADS ), STREAM_THREADS, 0, st->stream, st->set[ st->cur ], st->own_first, n, st->api_f[ 0 ],
                    st->api_f[ 1 ], st->api_f[ 2 ], st->api_f[ 3 ], st->api_f[ 4 ], st->api_f[ 5 ], st->api_col, st->api_ids );
        const cudaStream_t cs = copy_stream_after_unpack( st, async );
        const size_t       fb = n * sizeof( float );
        if ( mask & PPC_DL_POS ) {
            PPC_CK( cudaMemcpyAsync( p->p_x, st->api_f[ 0 ], fb, cudaMemcpyDeviceToHost, cs ) );
            PPC_CK( cudaMemcpyAsync( p->p_y, st->api_f[ 1 ], fb, cudaMemcpyDeviceToHost, cs ) );
        }
        if ( mask & PPC_DL_VEL ) {
            PPC_CK( cudaMemcpyAsync( p->v_x, st->api_f[ 2 ], fb, cudaMemcpyDeviceToHost, cs ) );
            PPC_CK( cudaMemcpyAsync( p->v_y, st->api_f[ 3 ], fb, cudaMemcpyDeviceToHost, cs ) );
        }
        if ( mask & PPC_DL_MASS_RADIUS ) {
            PPC_CK( cudaMemcpyAsync( p->mass, st->api_f[ 4 ], fb, cudaMemcpyDeviceToHost, cs ) );
            PPC_CK( cudaMemcpyAsync( p->radius, st->api_f[ 5 ], fb, cudaMemcpyDeviceToHost, cs ) );
        }
        if ( mask & PPC_DL_COLOR )
            PPC_CK( cudaMemcpyAsync( p->colors, st->api_col, n * sizeof( uchar4 ), cudaMemcpyDeviceToHost, cs ) );
        if ( ( mask & PPC_DL_IDS ) && global_ids )
            PPC_CK( cudaMemcpyAsync( global_ids, st->api_ids, n * sizeof( uint32_t ), cudaMemcpyDeviceToHost, cs ) );
        copy_started( st, async );
    }
    p->length = n;
    return 0;
}

extern "C" int ppc_slab_download( struct Particles_t *p, uint32_t *global_ids, unsigned mask ) {
    Store *st = store_of( p );
    if ( !st->slab )
        return -1;
    copy_wait( st );
    const int rc = slab_download( st, p, global_ids, mask, false );
    if ( rc != 0 )
        return rc;
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    return slab_halo_ok( st ) ? 0 : -2;
}

// The same, without waiting: the copies overlap the substeps that follow; ppc_download_wait() makes the host arrays valid.
// (global_ids must be page-locked memory, or the copy of the indices blocks.)
extern "C" int ppc_slab_download_async( struct Particles_t *p, uint32_t *global_ids, unsigned mask ) {
    Store *st = store_of( p );
    if ( !st->slab )
        return -1;
    return slab_download( st, p, global_ids, mask, true );
}

extern "C" int ppc_slab_info( struct Particles_t *p, ppc_slab_info_t *out ) {
    Store *st = store_of( p );
    if ( !st->slab )
        return -1;
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    out->n_owned = st->n_own, out->own_first = st->own_first, out->n_window = st->n;
    out->row_lo = st->row_lo, out->row_hi = st->row_hi, out->own_row_begin = st->slab_rows.own_begin, out->own_row_end = st->slab_rows.own_end;
    out->halo_rows = st->slab_rows.halo, out->cols = st->slab_world.cols, out->rows = st->slab_world.rows;
    uint32_t rounds = 0;
    if ( st->wave.counters && st->opt_resolve == 1 )
        PPC_CK( cudaMemcpy( &rounds, st->wave.counters + 4, sizeof( uint32_t ), cudaMemcpyDeviceToHost ) );
    out->rounds = rounds;
    slab_halo_ok( st );
    out->taint_rows = st->slab_taint_rows;
    out->violations = st->slab_violations;
    return 0;
}

extern "C" int64_t ppc_slab_debug_pairs( struct Particles_t *p, int32_t *pairs, size_t max_pairs ) {
    Store *st = store_of( p );
    if ( !st->slab || !st->grid_valid || !st->pre_valid )
        return -1;
    const size_t n = st->n_own;
    if ( n == 0 )
        return 0;
    ParticleSet pre = st->set[ st->cur ];
    if ( st->opt_resolve )
        pre.pv = st->set[ st->cur ^ 1 ].pv;   // positions the pair search saw
    if ( st->opt_resolve == 3 )
        PPC_LAUNCH( k_rec_to_pv, div_up( st->n, STREAM_THREADS ), STREAM_THREADS, 0, st->stream, (const float4 *)st->rec_a, pre.pv, st->n );
    const size_t padded = ( n + 1 + 63 ) & ~(size_t)63;
    uint32_t    *tmp    = dev_alloc<uint32_t>( 2 * padded );
    uint32_t    *count = tmp, *offset = tmp + padded;
    PPC_CK( cudaMemsetAsync( count, 0, ( n + 1 ) * sizeof( uint32_t ), st->stream ) );
    ensure_scratch( st, scan_scratch_words( n + 1 ) + 64 );
    PPC_LAUNCH( ( k_slab_pairs<false> ), div_up( n, COLLIDE_THREADS ), COLLIDE_THREADS, 0, st->stream, pre, st->cell_start, st->own_first, n, st->grid,
                count, (int32_t *)nullptr, (size_t)0 );
    exclusive_scan( count, offset, n + 1, st->scratch, st->stream );
    uint32_t total = 0;
    PPC_CK( cudaMemcpyAsync( &total, offset + n, sizeof( uint32_t ), cudaMemcpyDeviceToHost, st->stream ) );
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    if ( pairs && max_pairs && total ) {
        const size_t keep  = total < max_pairs ? total : max_pairs;
        int32_t     *d_out = dev_alloc<int32_t>( 2 * keep );
        PPC_LAUNCH( ( k_slab_pairs<true> ), div_up( n, COLLIDE_THREADS ), COLLIDE_THREADS, 0, st->stream, pre, st->cell_start, st->own_first, n,
                    st->grid, offset, d_out, keep );
        PPC_CK( cudaMemcpyAsync( pairs, d_out, 2 * keep * sizeof( int32_t ), cudaMemcpyDeviceToHost, st->stream ) );
        PPC_CK( cudaStreamSynchronize( st->stream ) );
        cudaFree( d_out );
    }
    cudaFree( tmp );
    return (int64_t)total;
}

// =================================================================================================================
// Part 4: state snapshots (the reference has no on-disk format, so this one is ours):
//   header  8 x uint64, little endian: magic "PPCB2001", version 1, n, window width, window height, 3 x reserved (0)
//   body    p_x[n], p_y[n], v_x[n], v_y[n], mass[n], radius[n] as float32, then colors[n] as 4 x uint8   (28 n bytes)
// Arrays are in API order (index i of every array is particle i), exactly the seven arrays of struct Particles_t.
// =================================================================================================================
static const uint64_t k_snapshot_magic = 0x3130303242435050ull;   // "PPCB2001"

extern "C" int ppc_snapshot_write( struct Particles_t *p, uint16_t window_width, uint16_t window_height, const char *path ) {
    Store *st = store_of( p );
    if ( st->slab )
        return -1;
    absorb_host_changes( st, p );
    download( st, p, PPC_DL_ALL );   // mass and radius never change on the device
    FILE *f = fopen( path, "wb" );
    if ( !f )
        return -2;
    const uint64_t n = p->length;
    const uint64_t header[ 8 ] = { k_snapshot_magic, 1, n, window_width, window_height, 0, 0, 0 };
    bool           ok = fwrite( header, sizeof( header ), 1, f ) == 1;
    float         *h[ 6 ] = { p->p_x, p->p_y, p->v_x, p->v_y, p->mass, p->radius };
    for ( int a = 0; a < 6 && ok; a++ )
        ok = n == 0 || fwrite( h[ a ], sizeof( float ), n, f ) == n;
    ok = ok && ( n == 0 || fwrite( p->colors, 4, n, f ) == n );
    ok = ( fclose( f ) == 0 ) && ok;
    return ok ? 0 : -3;
}

extern "C" int64_t ppc_snapshot_read( struct Particles_t *p, const char *path, uint16_t *window_width, uint16_t *window_height ) {
    Store *st = store_of( p );
    if ( st->slab )
        return -1;
    FILE *f = fopen( path, "rb" );
    if ( !f )
        return -2;
    uint64_t header[ 8 ];
    if ( fread( header, sizeof( header ), 1, f ) != 1 || header[ 0 ] != k_snapshot_magic || header[ 1 ] != 1 ) {
        fclose( f );
        return -3;
    }
    const uint64_t n = header[ 2 ];
    if ( n > p->size ) {   // the caller's store is too small; say how large it has to be
        fclose( f );
        return -(int64_t)n - 16;
    }
    PPC_CK( cudaStreamSynchronize( st->stream ) );
    float *h[ 6 ] = { p->p_x, p->p_y, p->v_x, p->v_y, p->mass, p->radius };
    bool   ok = true;
    for ( int a = 0; a < 6 && ok; a++ )
        ok = n == 0 || fread( h[ a ], sizeof( float ), n, f ) == n;
    ok = ok && ( n == 0 || fread( p->colors, 4, n, f ) == n );
    fclose( f );
    if ( !ok )
        return -3;
    p->length     = n;
    st->dirty_all = true;   // the host arrays are the truth now
    if ( window_width )
        *window_width = (uint16_t)header[ 3 ];
    if ( window_height )
        *window_height = (uint16_t)header[ 4 ];
    return (int64_t)n;
}

extern "C" void              *ppc_stream( struct Particles_t *p ) { return (void *)store_of( p )->stream; }
extern "C" unsigned long long ppc_kernel_launches( void ) { return g_launches; }
extern "C" void               ppc_profile( struct Particles_t *p, int enable ) { store_of( p )->profiling = enable != 0; }
extern "C" int                ppc_profile_stages( void ) { return ST_COUNT; }
extern "C" const char        *ppc_profile_stage_name( int stage ) { return stage >= 0 && stage < ST_COUNT ? k_stage_names[ stage ] : ""; }
extern "C" int ppc_profile_read( struct Particles_t *p, int stage, double *total_ms, unsigned long long *launches ) {
    Store *st = store_of( p );
    if ( stage < 0 || stage >= ST_COUNT )
        return -1;
    *total_ms = st->stage_ms[ stage ];
    *launches = st->stage_launches[ stage ];
    return 0;
}
extern "C" void ppc_profile_reset( struct Particles_t *p ) {
    Store *st = store_of( p );
    for ( int s = 0; s < ST_COUNT; s++ ) {
        st->stage_ms[ s ]       = 0;
        st->stage_launches[ s ] = 0;
    }
}
extern "C" const char *ppc_version( void ) { return "particlephysicsc_b200 0.1 (sm_100a)"; }
