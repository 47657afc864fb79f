// Shared-memory staging primitives of the TMA-staged cull kernel (dpcu_cull.cu::cullStagedKernel):
// mbarrier + 1-D bulk tensor-memory-accelerator copies (cp.async.bulk -> SASS UBLKCP) for the
// contiguous AABB streams, 16-byte cp.async (LDGSTS) for the gathered matrix rows.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

namespace dpcu
{
  __device__ __forceinline__ uint32_t smemAddr( void const *p )
  {
    return static_cast<uint32_t>( __cvta_generic_to_shared( p ) );
  }

  __device__ __forceinline__ void mbarInit( uint64_t *bar, uint32_t arrivals )
  {
    asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"( smemAddr( bar ) ), "r"( arrivals ) : "memory" );
  }
  __device__ __forceinline__ void mbarInitFence()
  {
    asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
  }
  // one arrival that also announces `bytes` of asynchronous copies to come
  __device__ __forceinline__ void mbarArriveExpectTx( uint64_t *bar, uint32_t bytes )
  {
    asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"( smemAddr( bar ) ), "r"( bytes ) : "memory" );
  }
  __device__ __forceinline__ void mbarWait( uint64_t *bar, uint32_t parity )
  {
    asm volatile(
      "{\n"
      "  .reg .pred p;\n"
      "WAIT_%=:\n"
      "  mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "  @p bra DONE_%=;\n"
      "  bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"( smemAddr( bar ) ),
      "r"( parity )
      : "memory" );
  }

  // global -> shared bulk copy by the TMA unit; completion is signalled on `bar` as `bytes` of
  // transaction count.  dst, src and bytes are multiples of 16.
  __device__ __forceinline__ void tmaLoad1d( void *dst, void const *src, uint32_t bytes, uint64_t *bar )
  {
    asm volatile( "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"( smemAddr( dst ) ),
                  "l"( src ), "r"( bytes ), "r"( smemAddr( bar ) )
                  : "memory" );
  }

  // 16-byte asynchronous copy global -> shared (allocating in L1: neighbouring rows of one matrix
  // share 32-byte sectors)
  __device__ __forceinline__ void cpAsync16( void *dst, void const *src )
  {
    asm volatile( "cp.async.ca.shared.global [%0], [%1], 16;" ::"r"( smemAddr( dst ) ), "l"( src ) : "memory" );
  }
  __device__ __forceinline__ void cpAsyncCommit()
  {
    asm volatile( "cp.async.commit_group;" ::: "memory" );
  }
  template <int N>
  __device__ __forceinline__ void cpAsyncWait()
  {
    asm volatile( "cp.async.wait_group %0;" ::"n"( N ) : "memory" );
  }
}
