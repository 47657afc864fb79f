// dp/cuda layer of the C ABI: device selection, device / pinned buffers, streams, events.
// C equivalents of the reference's host RAII wrappers (dp/cuda/{Device,Buffer,BufferHost,
// Stream,Event}.h and src/*.cpp); errors become status codes + dpcuGetLastError() instead of
// the exceptions CUDA_VERIFY throws (dp/cuda/Config.h:47-57).
#include "dpcu_internal.h"

namespace dpcu
{
  char *lastErrorBuffer()
  {
    static thread_local char buffer[1024] = { 0 };
    return buffer;
  }

  int fail( int code, char const *fmt, ... )
  {
    va_list ap;
    va_start( ap, fmt );
    vsnprintf( lastErrorBuffer(), 1024, fmt, ap );
    va_end( ap );
    return code;
  }

  int failCuda( cudaError_t err, char const *call, char const *file, int line )
  {
    // same wording as the reference's cudaVerify (dp/cuda/Config.h:47-55)
    snprintf( lastErrorBuffer(), 1024, "Error on executing %s in file <%s> line %d: %s : %s",
              call, file, line, cudaGetErrorName( err ), cudaGetErrorString( err ) );
    cudaGetLastError();   // clear the sticky-less error so later calls report their own
    if ( err == cudaErrorMemoryAllocation ) return DPCU_ERR_OUT_OF_MEMORY;
    return int( err ) > 0 ? int( err ) : DPCU_ERR_INVALID_VALUE;
  }

  int requireDevice()
  {
    int n = 0;
    cudaError_t err = cudaGetDeviceCount( &n );
    if ( err != cudaSuccess || n <= 0 )
    {
      cudaGetLastError();
      return fail( DPCU_ERR_NO_DEVICE, "no usable CUDA device (%s); libdpcu has no CPU fallback",
                   err != cudaSuccess ? cudaGetErrorString( err ) : "device count is 0" );
    }
    return DPCU_OK;
  }

  int DeviceArray::reserve( size_t bytes, bool keep, cudaStream_t stream )
  {
    if ( bytes <= capacity ) return DPCU_OK;
    size_t newCap = capacity ? capacity : 256;
    while ( newCap < bytes ) newCap += newCap / 2 + 256;
    newCap = ( newCap + 255 ) & ~size_t( 255 );
    void *np = nullptr;
    DPCU_CUDA( cudaMalloc( &np, newCap ) );
    if ( ptr )
    {
      if ( keep && capacity )
      {
        cudaError_t e = cudaMemcpyAsync( np, ptr, capacity, cudaMemcpyDeviceToDevice, stream );
        if ( e == cudaSuccess ) e = cudaStreamSynchronize( stream );
        if ( e != cudaSuccess ) { cudaFree( np ); return failCuda( e, "cudaMemcpyAsync(grow)", __FILE__, __LINE__ ); }
      }
      else
      {
        cudaStreamSynchronize( stream );   // nobody may still read the old block
      }
      cudaFree( ptr );
    }
    ptr = np;
    capacity = newCap;
    return DPCU_OK;
  }

  void DeviceArray::release()
  {
    if ( ptr ) cudaFree( ptr );
    ptr = nullptr;
    capacity = 0;
  }

  int PinnedArray::reserve( size_t bytes )
  {
    if ( bytes <= capacity ) return DPCU_OK;
    size_t newCap = capacity ? capacity : 4096;
    while ( newCap < bytes ) newCap += newCap / 2 + 4096;
    void *np = nullptr;
    DPCU_CUDA( cudaHostAlloc( &np, newCap, cudaHostAllocDefault ) );
    if ( ptr ) cudaFreeHost( ptr );
    ptr = np;
    capacity = newCap;
    return DPCU_OK;
  }

  void PinnedArray::release()
  {
    if ( ptr ) cudaFreeHost( ptr );
    ptr = nullptr;
    capacity = 0;
  }
}

struct dpcuBuffer
{
  void  *devicePointer;
  size_t size;
};

struct dpcuHostBuffer
{
  void    *pointer;
  size_t   size;
  unsigned flags;
};

extern "C"
{
  const char *dpcuGetLastError( void ) { return dpcu::lastErrorBuffer(); }
  int dpcuGetVersion( void ) { return 0x00010000; }

  // ------------------------------------------------------------------ device
  int dpcuDeviceCount( int *count )
  {
    DPCU_REQUIRE( count, "count is NULL" );
    *count = 0;
    DPCU_TRY( dpcu::requireDevice() );
    DPCU_CUDA( cudaGetDeviceCount( count ) );
    return DPCU_OK;
  }

  int dpcuDeviceSelect( int device )
  {
    DPCU_TRY( dpcu::requireDevice() );
    DPCU_CUDA( cudaSetDevice( device ) );
    return DPCU_OK;
  }

  int dpcuDeviceCurrent( int *device )
  {
    DPCU_REQUIRE( device, "device is NULL" );
    DPCU_TRY( dpcu::requireDevice() );
    DPCU_CUDA( cudaGetDevice( device ) );
    return DPCU_OK;
  }

  int dpcuDeviceSynchronize( void )
  {
    DPCU_TRY( dpcu::requireDevice() );
    DPCU_CUDA( cudaDeviceSynchronize() );
    return DPCU_OK;
  }

  int dpcuDeviceInfo( int device, char *name, size_t nameBytes, int *smCount, size_t *globalMemBytes,
                      int *ccMajor, int *ccMinor )
  {
    DPCU_TRY( dpcu::requireDevice() );
    cudaDeviceProp prop;
    DPCU_CUDA( cudaGetDeviceProperties( &prop, device ) );
    if ( name && nameBytes ) { strncpy( name, prop.name, nameBytes - 1 ); name[nameBytes - 1] = 0; }
    if ( smCount ) *smCount = prop.multiProcessorCount;
    if ( globalMemBytes ) *globalMemBytes = prop.totalGlobalMem;
    if ( ccMajor ) *ccMajor = prop.major;
    if ( ccMinor ) *ccMinor = prop.minor;
    return DPCU_OK;
  }

  int dpcuDeviceEnablePeerAccess( int device, int peer )
  {
    DPCU_TRY( dpcu::requireDevice() );
    int can = 0;
    DPCU_CUDA( cudaDeviceCanAccessPeer( &can, device, peer ) );
    if ( !can ) return dpcu::fail( DPCU_ERR_UNSUPPORTED, "device %d cannot access peer %d", device, peer );
    dpcu::DeviceGuard guard( device );
    cudaError_t e = cudaDeviceEnablePeerAccess( peer, 0 );
    if ( e == cudaErrorPeerAccessAlreadyEnabled ) { cudaGetLastError(); return DPCU_OK; }
    DPCU_CUDA( e );
    return DPCU_OK;
  }

  // ------------------------------------------------------------------ device buffer
  int dpcuBufferCreate( dpcuBuffer **out, size_t bytes )
  {
    DPCU_REQUIRE( out, "out is NULL" );
    *out = nullptr;
    DPCU_TRY( dpcu::requireDevice() );
    void *p = nullptr;
    if ( bytes ) DPCU_CUDA( cudaMalloc( &p, bytes ) );
    *out = new dpcuBuffer{ p, bytes };
    return DPCU_OK;
  }

  int dpcuBufferDestroy( dpcuBuffer *buffer )
  {
    if ( !buffer ) return DPCU_OK;
    cudaError_t e = buffer->devicePointer ? cudaFree( buffer->devicePointer ) : cudaSuccess;
    delete buffer;
    DPCU_CUDA( e );
    return DPCU_OK;
  }

  int dpcuBufferSize( const dpcuBuffer *buffer, size_t *bytes )
  {
    DPCU_REQUIRE( buffer && bytes, "NULL argument" );
    *bytes = buffer->size;
    return DPCU_OK;
  }

  int dpcuBufferDevicePointer( const dpcuBuffer *buffer, void **devicePointer )
  {
    DPCU_REQUIRE( buffer && devicePointer, "NULL argument" );
    *devicePointer = buffer->devicePointer;
    return DPCU_OK;
  }

  int dpcuBufferUpload( dpcuBuffer *buffer, size_t offset, const void *host, size_t bytes, dpcuStream *stream )
  {
    DPCU_REQUIRE( buffer && ( host || !bytes ), "NULL argument" );
    DPCU_REQUIRE( offset + bytes <= buffer->size, "range exceeds buffer size" );   // DP_ASSERT in Buffer.cpp
    if ( !bytes ) return DPCU_OK;
    char *dst = static_cast<char *>( buffer->devicePointer ) + offset;
    if ( stream ) DPCU_CUDA( cudaMemcpyAsync( dst, host, bytes, cudaMemcpyHostToDevice, stream->stream ) );
    else          DPCU_CUDA( cudaMemcpy( dst, host, bytes, cudaMemcpyHostToDevice ) );
    return DPCU_OK;
  }

  int dpcuBufferDownload( const dpcuBuffer *buffer, size_t offset, void *host, size_t bytes, dpcuStream *stream )
  {
    DPCU_REQUIRE( buffer && ( host || !bytes ), "NULL argument" );
    DPCU_REQUIRE( offset + bytes <= buffer->size, "range exceeds buffer size" );
    if ( !bytes ) return DPCU_OK;
    char const *src = static_cast<char const *>( buffer->devicePointer ) + offset;
    if ( stream ) DPCU_CUDA( cudaMemcpyAsync( host, src, bytes, cudaMemcpyDeviceToHost, stream->stream ) );
    else          DPCU_CUDA( cudaMemcpy( host, src, bytes, cudaMemcpyDeviceToHost ) );
    return DPCU_OK;
  }

  int dpcuBufferFill( dpcuBuffer *buffer, int byteValue, size_t bytes, size_t offset )
  {
    DPCU_REQUIRE( buffer, "NULL argument" );
    DPCU_REQUIRE( offset + bytes <= buffer->size, "range exceeds buffer size" );
    if ( bytes ) DPCU_CUDA( cudaMemset( static_cast<char *>( buffer->devicePointer ) + offset, byteValue, bytes ) );
    return DPCU_OK;
  }

  int dpcuBufferFillAsync( dpcuBuffer *buffer, int byteValue, size_t bytes, size_t offset, dpcuStream *stream )
  {
    DPCU_REQUIRE( buffer && stream, "NULL argument" );
    DPCU_REQUIRE( offset + bytes <= buffer->size, "range exceeds buffer size" );
    if ( bytes ) DPCU_CUDA( cudaMemsetAsync( static_cast<char *>( buffer->devicePointer ) + offset, byteValue, bytes, stream->stream ) );
    return DPCU_OK;
  }

  // ------------------------------------------------------------------ pinned host buffer
  int dpcuHostBufferCreate( dpcuHostBuffer **out, size_t bytes, unsigned flags )
  {
    DPCU_REQUIRE( out, "out is NULL" );
    *out = nullptr;
    DPCU_TRY( dpcu::requireDevice() );
    unsigned cf = cudaHostAllocDefault;
    if ( flags & DPCU_HOST_PORTABLE )      cf |= cudaHostAllocPortable;
    if ( flags & DPCU_HOST_MAPPED )        cf |= cudaHostAllocMapped;
    if ( flags & DPCU_HOST_WRITECOMBINED ) cf |= cudaHostAllocWriteCombined;
    void *p = nullptr;
    if ( bytes ) DPCU_CUDA( cudaHostAlloc( &p, bytes, cf ) );
    *out = new dpcuHostBuffer{ p, bytes, flags };
    return DPCU_OK;
  }

  int dpcuHostBufferDestroy( dpcuHostBuffer *buffer )
  {
    if ( !buffer ) return DPCU_OK;
    cudaError_t e = buffer->pointer ? cudaFreeHost( buffer->pointer ) : cudaSuccess;
    delete buffer;
    DPCU_CUDA( e );
    return DPCU_OK;
  }

  int dpcuHostBufferPointer( const dpcuHostBuffer *buffer, void **hostPointer )
  {
    DPCU_REQUIRE( buffer && hostPointer, "NULL argument" );
    *hostPointer = buffer->pointer;
    return DPCU_OK;
  }

  int dpcuHostBufferSize( const dpcuHostBuffer *buffer, size_t *bytes )
  {
    DPCU_REQUIRE( buffer && bytes, "NULL argument" );
    *bytes = buffer->size;
    return DPCU_OK;
  }

  // ------------------------------------------------------------------ stream
  int dpcuStreamCreate( dpcuStream **out, int blocking, int priority )
  {
    DPCU_REQUIRE( out, "out is NULL" );
    *out = nullptr;
    DPCU_TRY( dpcu::requireDevice() );
    cudaStream_t s;
    DPCU_CUDA( cudaStreamCreateWithPriority( &s, blocking ? cudaStreamDefault : cudaStreamNonBlocking, priority ) );
    int device = 0;
    cudaGetDevice( &device );
    *out = new dpcuStream{ s, blocking, priority, device };
    return DPCU_OK;
  }

  int dpcuStreamDestroy( dpcuStream *stream )
  {
    if ( !stream ) return DPCU_OK;
    cudaError_t e = cudaStreamDestroy( stream->stream );
    delete stream;
    DPCU_CUDA( e );
    return DPCU_OK;
  }

  int dpcuStreamSynchronize( dpcuStream *stream )
  {
    DPCU_REQUIRE( stream, "stream is NULL" );
    DPCU_CUDA( cudaStreamSynchronize( stream->stream ) );
    return DPCU_OK;
  }

  int dpcuStreamIsCompleted( dpcuStream *stream, int *completed )
  {
    DPCU_REQUIRE( stream && completed, "NULL argument" );
    cudaError_t e = cudaStreamQuery( stream->stream );
    if ( e == cudaErrorNotReady ) { cudaGetLastError(); *completed = 0; return DPCU_OK; }
    DPCU_CUDA( e );
    *completed = 1;
    return DPCU_OK;
  }

  int dpcuStreamWaitEvent( dpcuStream *stream, dpcuEvent *event )
  {
    DPCU_REQUIRE( stream && event, "NULL argument" );
    DPCU_CUDA( cudaStreamWaitEvent( stream->stream, event->event, 0 ) );
    return DPCU_OK;
  }

  int dpcuStreamNative( dpcuStream *stream, void **cudaStream )
  {
    DPCU_REQUIRE( stream && cudaStream, "NULL argument" );
    *cudaStream = stream->stream;
    return DPCU_OK;
  }

  // ------------------------------------------------------------------ event
  int dpcuEventCreate( dpcuEvent **out, unsigned flags )
  {
    DPCU_REQUIRE( out, "out is NULL" );
    *out = nullptr;
    DPCU_TRY( dpcu::requireDevice() );
    unsigned cf = cudaEventDefault;
    if ( flags & DPCU_EVENT_BLOCKING_SYNC )  cf |= cudaEventBlockingSync;
    if ( flags & DPCU_EVENT_DISABLE_TIMING ) cf |= cudaEventDisableTiming;
    cudaEvent_t e;
    DPCU_CUDA( cudaEventCreateWithFlags( &e, cf ) );
    *out = new dpcuEvent{ e, flags };
    return DPCU_OK;
  }

  int dpcuEventDestroy( dpcuEvent *event )
  {
    if ( !event ) return DPCU_OK;
    cudaError_t e = cudaEventDestroy( event->event );
    delete event;
    DPCU_CUDA( e );
    return DPCU_OK;
  }

  int dpcuEventRecord( dpcuEvent *event, dpcuStream *stream )
  {
    DPCU_REQUIRE( event, "event is NULL" );
    DPCU_CUDA( cudaEventRecord( event->event, stream ? stream->stream : 0 ) );
    return DPCU_OK;
  }

  int dpcuEventSynchronize( dpcuEvent *event )
  {
    DPCU_REQUIRE( event, "event is NULL" );
    DPCU_CUDA( cudaEventSynchronize( event->event ) );
    return DPCU_OK;
  }

  int dpcuEventIsCompleted( dpcuEvent *event, int *completed )
  {
    DPCU_REQUIRE( event && completed, "NULL argument" );
    cudaError_t e = cudaEventQuery( event->event );
    if ( e == cudaErrorNotReady ) { cudaGetLastError(); *completed = 0; return DPCU_OK; }
    DPCU_CUDA( e );
    *completed = 1;
    return DPCU_OK;
  }

  int dpcuEventElapsedMs( dpcuEvent *start, dpcuEvent *stop, float *milliseconds )
  {
    DPCU_REQUIRE( start && stop && milliseconds, "NULL argument" );
    if ( ( start->flags | stop->flags ) & DPCU_EVENT_DISABLE_TIMING )
      return dpcu::fail( DPCU_ERR_INVALID_VALUE, "dpcuEventElapsedMs: event created with DPCU_EVENT_DISABLE_TIMING" );
    DPCU_CUDA( cudaEventElapsedTime( milliseconds, start->event, stop->event ) );
    return DPCU_OK;
  }

  // ------------------------------------------------------------------ IPC
  int dpcuIpcGetHandle( const void *devicePointer, unsigned char handle[DPCU_IPC_HANDLE_BYTES] )
  {
    DPCU_REQUIRE( devicePointer && handle, "NULL argument" );
    static_assert( sizeof( cudaIpcMemHandle_t ) == DPCU_IPC_HANDLE_BYTES, "ipc handle size" );
    cudaIpcMemHandle_t h;
    DPCU_CUDA( cudaIpcGetMemHandle( &h, const_cast<void *>( devicePointer ) ) );
    memcpy( handle, &h, sizeof h );
    return DPCU_OK;
  }

  int dpcuIpcOpen( const unsigned char handle[DPCU_IPC_HANDLE_BYTES], void **devicePointer )
  {
    DPCU_REQUIRE( devicePointer && handle, "NULL argument" );
    cudaIpcMemHandle_t h;
    memcpy( &h, handle, sizeof h );
    DPCU_CUDA( cudaIpcOpenMemHandle( devicePointer, h, cudaIpcMemLazyEnablePeerAccess ) );
    return DPCU_OK;
  }

  int dpcuIpcClose( void *devicePointer )
  {
    if ( !devicePointer ) return DPCU_OK;
    DPCU_CUDA( cudaIpcCloseMemHandle( devicePointer ) );
    return DPCU_OK;
  }
}
