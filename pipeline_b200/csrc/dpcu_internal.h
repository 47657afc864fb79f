// Internal helpers shared by the libdpcu.so translation units (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/dpcu.h"

// NVTX ranges around the entry points of the hot path (the reference brackets the same places
// with dp::util::ProfileEntry, dp/culling/cpu/src/ManagerImpl.cpp:469, dp/sg/xbar/src/SceneTree.cpp:157).  Header-only
// NVTX v3: without a profiler attached a range is one predictable branch.
#include <nvtx3/nvToolsExt.h>

namespace dpcu
{
  // thread-local message behind dpcuGetLastError()
  char *lastErrorBuffer();
  int   fail( int code, char const *fmt, ... );
  int   failCuda( cudaError_t err, char const *call, char const *file, int line );

  // fails loudly when no device is usable: there is no CPU fallback anywhere in this library
  int   requireDevice();

  struct Range
  {
    explicit Range( char const *name ) { nvtxRangePushA( name ); }
    ~Range() { nvtxRangePop(); }
    Range( Range const & ) = delete;
    Range & operator=( Range const & ) = delete;
  };

  struct DeviceGuard
  {
    explicit DeviceGuard( int device ) : m_prev( -1 )
    {
      cudaGetDevice( &m_prev );
      if ( m_prev != device ) cudaSetDevice( device );
      m_device = device;
    }
    ~DeviceGuard()
    {
      if ( m_prev >= 0 && m_prev != m_device ) cudaSetDevice( m_prev );
    }
    int m_prev, m_device;
  };

  // growable device allocation (capacity doubles, contents preserved on request)
  struct DeviceArray
  {
    void  *ptr = nullptr;
    size_t capacity = 0;
    int reserve( size_t bytes, bool keep, cudaStream_t stream );
    void release();
  };

  // growable pinned staging buffer for host -> device uploads of pageable caller memory
  struct PinnedArray
  {
    void  *ptr = nullptr;
    size_t capacity = 0;
    int reserve( size_t bytes );
    void release();
  };

  inline size_t divUp( size_t a, size_t b ) { return ( a + b - 1 ) / b; }

  // "The work last submitted for this object", as an event: unlike a remembered cudaStream_t it
  // stays valid after the caller destroyed the stream the work ran on.
  struct StreamFence
  {
    cudaEvent_t event = nullptr;
    bool        pending = false;
    cudaError_t record( cudaStream_t s )
    {
      cudaError_t e = cudaSuccess;
      if ( !event ) e = cudaEventCreateWithFlags( &event, cudaEventDisableTiming );
      if ( e == cudaSuccess ) e = cudaEventRecord( event, s );
      pending = ( e == cudaSuccess );
      return e;
    }
    // make stream s wait for the recorded work (device-side ordering, no host block)
    cudaError_t orderBefore( cudaStream_t s ) const { return pending ? cudaStreamWaitEvent( s, event, 0 ) : cudaSuccess; }
    void hostWait() const { if ( pending ) cudaEventSynchronize( event ); }
    void destroy() { if ( event ) cudaEventDestroy( event ); event = nullptr; pending = false; }
  };
}

struct dpcuStream
{
  cudaStream_t stream;
  int          blocking;
  int          priority;
  int          device;
};

struct dpcuEvent
{
  cudaEvent_t event;
  unsigned    flags;
};

#define DPCU_CUDA( call )                                                              \
  do {                                                                                 \
    cudaError_t dpcu_err_ = ( call );                                                  \
    if ( dpcu_err_ != cudaSuccess )                                                    \
      return dpcu::failCuda( dpcu_err_, #call, __FILE__, __LINE__ );                   \
  } while ( 0 )

#define DPCU_TRY( call )                                                               \
  do {                                                                                 \
    int dpcu_rc_ = ( call );                                                           \
    if ( dpcu_rc_ != DPCU_OK ) return dpcu_rc_;                                        \
  } while ( 0 )

#define DPCU_REQUIRE( cond, msg )                                                      \
  do {                                                                                 \
    if ( !( cond ) ) return dpcu::fail( DPCU_ERR_INVALID_VALUE, "%s: %s", __func__, msg ); \
  } while ( 0 )
