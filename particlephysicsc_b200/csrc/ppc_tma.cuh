// ppc_tma.cuh -- the Blackwell/Hopper asynchronous copy engine (TMA) in its 1-D bulk form, plus the mbarrier that its
// completion is signalled on.  Used by the tile kernel (ppc_gs.cuh) to stage the contiguous slot runs of a tile and its
// halo into shared memory without spending a single LSU instruction on them: one elected thread posts the expected byte
// count and issues one `cp.async.bulk` per run; everybody waits on the barrier's phase parity.  SASS: UBLKCP + SYNCS.
//
// Requirements of cp.async.bulk: source, destination and size are multiples of 16 bytes (the staged records are 16 B).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ppc {

__device__ __forceinline__ uint32_t smem_addr( const void *p ) { return (uint32_t)__cvta_generic_to_shared( p ); }

__device__ __forceinline__ void mbar_init( uint64_t *bar, const uint32_t arrivals ) {
    asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"( smem_addr( bar ) ), "r"( arrivals ) : "memory" );
    asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );   // make the initialised barrier visible to the async proxy
}

// One arrival that also announces `bytes` of asynchronous copies to come in this phase.
__device__ __forceinline__ void mbar_arrive_expect_tx( uint64_t *bar, const uint32_t bytes ) {
    asm volatile( "mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"( smem_addr( bar ) ), "r"( bytes ) : "memory" );
}

// global -> shared bulk copy; completes `bytes` of the barrier's transaction count when the data has landed.
__device__ __forceinline__ void tma_load_1d( void *dst_smem, const void *src_gmem, const uint32_t bytes, uint64_t *bar ) {
    asm volatile( "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"( smem_addr( dst_smem ) ),
                  "l"( src_gmem ), "r"( bytes ), "r"( smem_addr( bar ) )
                  : "memory" );
}

// Wait until the phase with the given parity has completed (all arrivals in, all announced bytes landed).
__device__ __forceinline__ void mbar_wait( uint64_t *bar, const uint32_t parity ) {
    asm volatile( "{\n\t"
                  ".reg .pred p;\n\t"
                  "PPC_MBAR_WAIT:\n\t"
                  "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n\t"
                  "@p bra PPC_MBAR_DONE;\n\t"
                  "bra PPC_MBAR_WAIT;\n\t"
                  "PPC_MBAR_DONE:\n\t"
                  "}" ::"r"( smem_addr( bar ) ),
                  "r"( parity )
                  : "memory" );
}

// Order this thread's earlier generic-proxy accesses (ordinary loads / stores) before later asynchronous-proxy ones
// (bulk copies): shared memory about to be overwritten by a copy, global memory written by ordinary stores and then
// read by a copy.
__device__ __forceinline__ void fence_proxy_async() { asm volatile( "fence.proxy.async;" ::: "memory" ); }

}   // namespace ppc
