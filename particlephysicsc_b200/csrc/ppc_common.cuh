// ppc_common.cuh -- shared declarations of the B200 solver library (error handling, launch bookkeeping, views).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

namespace ppc {

// Every CUDA call is checked; a failure is fatal and loud (there is no CPU fallback and no silent partial step,
// unlike the reference's silent early returns at src/particles_solver.c:67-68,256-263,305-309).
#define PPC_CK( call )                                                                                              \
    do {                                                                                                            \
        cudaError_t e__ = ( call );                                                                                 \
        if ( e__ != cudaSuccess ) {                                                                                 \
            fprintf( stderr, "particlephysicsc_b200: CUDA error %s (%s) at %s:%d\n", cudaGetErrorName( e__ ),      \
                     cudaGetErrorString( e__ ), __FILE__, __LINE__ );                                               \
            abort();                                                                                                \
        }                                                                                                           \
    } while ( 0 )

// Kernel launches are counted (bench.py reports the number as gpu_launches) and checked.
extern unsigned long long g_launches;
#define PPC_LAUNCH( kernel, grid, block, smem, stream, ... )                                                        \
    do {                                                                                                            \
        kernel<<<( grid ), ( block ), ( smem ), ( stream )>>>( __VA_ARGS__ );                                       \
        ++::ppc::g_launches;                                                                                        \
        PPC_CK( cudaGetLastError() );                                                                               \
    } while ( 0 )

static inline unsigned div_up( size_t a, size_t b ) { return (unsigned)( ( a + b - 1 ) / b ); }

// One set of per-particle device arrays, in STORAGE order (cell-sorted after the first substep).
//   pv  : (p_x, p_y, v_x, v_y)  -- the 16 B that change every substep
//   rm  : (radius, mass)        -- constant per particle, travels with it when storage is re-sorted
//   gid : API index (the index into the caller's struct Particles_t arrays)
//   key : cell key cy*cols+cx of the last grid build (reference src/particles_solver.c:291)
//   col : RGBA of the last drawn substep (reference src/particles_solver.c:38-39)
struct ParticleSet {
    float4   *pv;
    float2   *rm;
    uint32_t *gid;
    uint32_t *key;
    uchar4   *col;
};

// Uniform grid of one substep (reference src/particles_solver.c:231-249).
struct GridDims {
    float    cell_size;
    int      cols, rows;
    uint32_t cells;
    int      row0;   // grid row of the whole field that row 0 of this grid is (slab windows; 0 for a whole-field grid)
    int      taint_lo, taint_hi;   // slab windows: rows whose particles miss contacts beyond the window (-1: none)
};

}   // namespace ppc
