// Transform layer of the C ABI: device mirror of dp::transform::Tree and K1, the level-by-level
// world-matrix propagation (dp/transform/src/Tree.cpp:133-166, data layout Tree.h:104-128).
//
// HBM layout:  local[numNodes], world[numNodes]   4 x float4 per node, 64-byte stride
//              entries[E]                         uint2 {parent, transform}, levels back to back
//              dirtyLocal / dirtyWorld            u32 bit words over node indices
//
// K1 runs one launch per level, in one of two forms:
//   treeLevelKernel      four threads per node, thread r computes row r of world[t] = local[t] *
//                        world[parent] with the reference's association order (a2); one short
//                        dependent chain per thread - used for small levels (latency bound anyway);
//   treeLevelWideKernel  one thread per node on a persistent grid, {parent, node} entries two tiles
//                        and dirty flags one tile ahead of the matrix loads, a full warp of dirty
//                        consecutive nodes moves its 2 KiB of locals / worlds with coalesced 16-byte
//                        accesses through a shared-memory transpose (tree_propagate.cuh) - the form
//                        for levels large enough to be HBM bound.
// Compiled with -fmad=false.
#include "cull_math.cuh"
#include "tree_propagate.cuh"
#include "dpcu_internal.h"
#include "dpcu_tree.h"

#include <new>
#include <vector>

namespace dpcu
{
  __device__ __forceinline__ bool testBit( uint32_t const *w, uint32_t i )
  {
    return ( w[i >> 5] >> ( i & 31u ) ) & 1u;
  }

  __global__ void __launch_bounds__( 256 ) treeLevelKernel( uint2 const *__restrict__ entries, uint32_t count,
                                                            float4 const *__restrict__ local, float4 *world,
                                                            uint32_t const *__restrict__ dirtyLocal, uint32_t *dirtyWorld )
  {
    // launched as a programmatic dependent of the previous level (treeComputeLevels): its CTAs are placed while that
    // level drains, and everything it reads of it is read behind this point
    cudaGridDependencySynchronize();
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t e = t >> 2, r = t & 3u;
    if ( e >= count ) return;
    const uint2 pe = __ldg( entries + e );
    const uint32_t parent = pe.x, node = pe.y;
    // Tree.cpp:155 - parent bits were set by earlier launches; this launch only ORs bits of its own level
    if ( !( testBit( dirtyWorld, parent ) || testBit( dirtyLocal, node ) ) ) return;
    const float4 a  = ldStream( local + 4ull * node + r );
    float4 const *pw = world + 4ull * parent;
    const float4 b0 = pw[0], b1 = pw[1], b2 = pw[2], b3 = pw[3];
    world[4ull * node + r] = vecMulMat( a, b0, b1, b2, b3 );      // Tree.cpp:157, Matmnt.h:1381-1415
    if ( r == 0 ) atomicOr( dirtyWorld + ( node >> 5 ), 1u << ( node & 31u ) );   // Tree.cpp:158
  }

  constexpr int      kWideThreads    = 256;

  __global__ void __launch_bounds__( kWideThreads ) treeLevelWideKernel( uint2 const *__restrict__ entries, uint32_t count,
                                                                        float4 const *local, float4 *world,
                                                                        uint32_t const *dirtyLocal, uint32_t *dirtyWorld )
  {
    __shared__ float4 sTranspose[kWideThreads / 32][2][128];     // per warp: locals in, worlds out (2 KiB each)
    cudaGridDependencySynchronize();                             // see treeLevelKernel
    const uint32_t lane   = threadIdx.x & 31u;
    const uint32_t stride = gridDim.x * kWideThreads;
    float4 *bufIn = sTranspose[threadIdx.x >> 5][0], *bufOut = sTranspose[threadIdx.x >> 5][1];
    uint32_t i = blockIdx.x * kWideThreads + threadIdx.x;
    uint2 ent = make_uint2( 0u, 0u ), entN = make_uint2( 0u, 0u );
    if ( i < count ) ent = __ldg( entries + i );
    if ( i + stride < count && i + stride >= i ) entN = __ldg( entries + i + stride );
    // Tree.cpp:155 - parent bits were set by earlier launches; this launch only ORs bits of its own level
    bool dirty = i < count && ( testBit( dirtyWorld, ent.x ) || testBit( dirtyLocal, ent.y ) );
    const uint32_t nTiles = ( count + kWideThreads - 1 ) / kWideThreads;
    for ( uint32_t tile = blockIdx.x; tile < nTiles; tile += gridDim.x, i += stride )
    {
      const bool     live = i < count;
      const uint32_t iN = i + stride, iNN = iN + stride;
      const bool     liveN = iN < count && iN >= i;
      uint2 entNN = make_uint2( 0u, 0u );
      if ( iNN < count && iNN >= iN && liveN ) entNN = __ldg( entries + iNN );
      const bool dirtyN = liveN && ( testBit( dirtyWorld, entN.x ) || testBit( dirtyLocal, entN.y ) );
      float4 w0, w1, w2, w3;
      const uint32_t node0 = __shfl_sync( 0xffffffffu, ent.y, 0 );
      if ( __all_sync( 0xffffffffu, live && dirty && ent.y == node0 + lane ) )
      {
        propagateWarpCoalesced( local, world, dirtyWorld, ent.x, node0, lane, bufIn, bufOut, w0, w1, w2, w3 );
      }
      else if ( live && dirty )
      {
        propagateNode( local, world, dirtyWorld, ent.x, ent.y, w0, w1, w2, w3 );
      }
      ent = entN; entN = entNN; dirty = dirtyN;
    }
  }

  __global__ void treeInitNodesKernel( float4 *local, float4 *world, uint32_t *dirtyLocal, uint32_t first, uint32_t count )
  {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if ( t >= count * 4u ) return;
    const uint32_t node = first + ( t >> 2 ), r = t & 3u;
    const float4 row = make_float4( r == 0 ? 1.f : 0.f, r == 1 ? 1.f : 0.f, r == 2 ? 1.f : 0.f, r == 3 ? 1.f : 0.f );
    local[4ull * node + r] = row;
    world[4ull * node + r] = row;
    if ( r == 0 ) atomicOr( dirtyLocal + ( node >> 5 ), 1u << ( node & 31u ) );
  }

  __global__ void treeMarkRangeKernel( uint32_t *dirtyLocal, uint32_t first, uint32_t count )
  {
    // one thread per word touched by [first, first+count)
    const uint32_t w0 = first >> 5;
    const uint32_t w  = w0 + blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t lo = uint64_t( w ) << 5, hi = lo + 32, end = uint64_t( first ) + count;
    if ( lo >= end ) return;
    uint32_t mask = ~0u;
    if ( first > lo ) mask &= ~0u << ( first - lo );
    if ( end < hi )   mask &= ~0u >> ( hi - end );
    atomicOr( dirtyLocal + w, mask );
  }

  __global__ void treeScatterLocalsKernel( uint32_t const *indices, float4 const *packed, uint32_t n, uint32_t numNodes,
                                           float4 *local, uint32_t *dirtyLocal )
  {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if ( t >= n * 4u ) return;
    const uint32_t node = indices[t >> 2], r = t & 3u;
    if ( node >= numNodes ) return;
    local[4ull * node + r] = packed[t];
    if ( r == 0 ) atomicOr( dirtyLocal + ( node >> 5 ), 1u << ( node & 31u ) );
  }
}


namespace dpcu
{
  // Tree::compute in three steps so that the culling layer can run the last level inside its own
  // kernel (dpcuCullRunWithTree): begin = ordering + drop the previous published dirty set,
  // levels [first, last), end = clear the local dirty bits (Tree.cpp:163) + fence.
  // world matrices of the nodes whose dirty-world bit is set, compacted (order irrelevant: every matrix travels with its
  // node index) - what dp::transform::cuda::Tree needs to refresh the host copy behind getWorldMatrices()
  __global__ void __launch_bounds__( 256 ) gatherDirtyWorldKernel( uint32_t const *dirtyWorld, float4 const *world, uint32_t nWords,
                                                                   uint32_t capacity, uint32_t *counter, uint32_t *outIndex, float4 *outMats )
  {
    const uint32_t lane = threadIdx.x & 31u;
    for ( uint32_t w0 = ( blockIdx.x * blockDim.x + threadIdx.x ) & ~31u; w0 < nWords; w0 += gridDim.x * blockDim.x )
    {
      const uint32_t w = w0 + lane;
      uint32_t bits = w < nWords ? dirtyWorld[w] : 0u;
      const uint32_t pc = __popc( bits );
      uint32_t incl = pc;
#pragma unroll
      for ( int d = 1; d < 32; d <<= 1 )
      {
        const uint32_t t = __shfl_up_sync( 0xffffffffu, incl, d );
        if ( lane >= d ) incl += t;
      }
      const uint32_t total = __shfl_sync( 0xffffffffu, incl, 31 );
      if ( !total ) continue;
      uint32_t base = 0;
      if ( lane == 0 ) base = atomicAdd( counter, total );
      base = __shfl_sync( 0xffffffffu, base, 0 );
      uint32_t slot = base + incl - pc;
      while ( bits )
      {
        const uint32_t node = ( w << 5 ) + uint32_t( __ffs( bits ) - 1 );
        bits &= bits - 1;
        if ( slot < capacity )
        {
          outIndex[slot] = node;
          float4 const *m = world + 4ull * node;
          float4 *o = outMats + 4ull * slot;
          o[0] = m[0]; o[1] = m[1]; o[2] = m[2]; o[3] = m[3];
        }
        ++slot;
      }
    }
  }

  int treeBeginCompute( dpcuTree *t, cudaStream_t s )
  {
    if ( s != t->stream )
    {
      DPCU_CUDA( t->uploads.record( t->stream ) );      // topology / local-matrix updates ran on the tree's stream
      DPCU_CUDA( t->uploads.orderBefore( s ) );
    }
    DPCU_CUDA( t->done.orderBefore( s ) );
    DPCU_CUDA( t->readers.orderBefore( s ) );          // a cull may still be reading the world matrices in place
    size_t words = divUp( t->numNodes, 32 );
    // the previous compute's published set is dropped now (the reference clears it right after notifying, Tree.cpp:163-164)
    DPCU_CUDA( cudaMemsetAsync( t->dirtyWorld.ptr, 0, words * 4, s ) );
    return DPCU_OK;
  }

  int treeComputeLevels( dpcuTree *t, cudaStream_t s, size_t firstLevel, size_t lastLevel )
  {
    for ( size_t l = firstLevel; l < lastLevel; ++l )
    {
      uint32_t first = t->levelOffsets[l], count = t->levelOffsets[l + 1] - first;
      if ( !count ) continue;
      uint2 const *entries = static_cast<uint2 const *>( t->entries.ptr ) + first;
      float4 const *local = static_cast<float4 const *>( t->local.ptr );
      float4 *world = static_cast<float4 *>( t->world.ptr );
      uint32_t const *dirtyLocal = static_cast<uint32_t const *>( t->dirtyLocal.ptr );
      uint32_t *dirtyWorld = static_cast<uint32_t *>( t->dirtyWorld.ptr );
      // Levels are a chain of small dependent launches (C3: 4096, 65 536 and 1 Mi nodes before the fused leaf level): each
      // is launched with programmatic stream serialization, so its CTAs are placed while the level before drains and wait
      // in cudaGridDependencySynchronize() - the launch latency of a level is off the frame's critical path.
      cudaLaunchConfig_t cfg = {};
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cfg.stream = s;
      if ( count >= t->wideMinNodes )
      {
        size_t grid = divUp( size_t( count ), size_t( kWideThreads ) );
        size_t cap  = size_t( t->smCount ) * t->wideCtasPerSm;
        cfg.gridDim  = dim3( unsigned( grid < cap ? grid : cap ) );
        cfg.blockDim = dim3( kWideThreads );
        DPCU_CUDA( cudaLaunchKernelEx( &cfg, treeLevelWideKernel, entries, count, local, world, dirtyLocal, dirtyWorld ) );
      }
      else
      {
        cfg.gridDim  = dim3( unsigned( divUp( size_t( count ) * 4, 256 ) ) );
        cfg.blockDim = dim3( 256 );
        DPCU_CUDA( cudaLaunchKernelEx( &cfg, treeLevelKernel, entries, count, local, world, dirtyLocal, dirtyWorld ) );
      }
      ++t->launches;
    }
    return DPCU_OK;
  }

  int treeEndCompute( dpcuTree *t, cudaStream_t s )
  {
    size_t words = divUp( t->numNodes, 32 );
    DPCU_CUDA( cudaMemsetAsync( t->dirtyLocal.ptr, 0, words * 4, s ) );   // Tree.cpp:163
    DPCU_CUDA( t->done.record( s ) );
    return DPCU_OK;
  }
}

extern "C"
{
  int dpcuTreeCreate( dpcuTree **out, int device )
  {
    DPCU_REQUIRE( out, "out is NULL" );
    *out = nullptr;
    DPCU_TRY( dpcu::requireDevice() );
    int count = 0;
    DPCU_CUDA( cudaGetDeviceCount( &count ) );
    DPCU_REQUIRE( device >= 0 && device < count, "device index out of range" );
    dpcu::DeviceGuard guard( device );
    dpcuTree *t = new ( std::nothrow ) dpcuTree;
    if ( !t ) return dpcu::fail( DPCU_ERR_OUT_OF_MEMORY, "dpcuTreeCreate: host allocation failed" );
    t->device = device;
    cudaDeviceGetAttribute( &t->smCount, cudaDevAttrMultiProcessorCount, device );
    cudaOccupancyMaxActiveBlocksPerMultiprocessor( &t->wideCtasPerSm, dpcu::treeLevelWideKernel, dpcu::kWideThreads, 0 );
    if ( t->smCount <= 0 ) t->smCount = 1;
    if ( t->wideCtasPerSm <= 0 ) t->wideCtasPerSm = 1;
    cudaError_t e = cudaStreamCreateWithFlags( &t->stream, cudaStreamNonBlocking );
    if ( e != cudaSuccess ) { delete t; return dpcu::failCuda( e, "cudaStreamCreateWithFlags", __FILE__, __LINE__ ); }
    *out = t;
    return DPCU_OK;
  }

  int dpcuTreeDestroy( dpcuTree *t )
  {
    if ( !t ) return DPCU_OK;
    dpcu::DeviceGuard guard( t->device );
    cudaStreamSynchronize( t->stream );
    t->done.hostWait();
    t->readers.hostWait();
    t->done.destroy();
    t->uploads.destroy();
    t->readers.destroy();
    t->gather.release();
    t->gatherHost.release();
    t->local.release(); t->world.release(); t->entries.release(); t->dirtyLocal.release(); t->dirtyWorld.release(); t->scratch.release();
    cudaStreamDestroy( t->stream );
    delete t;
    return DPCU_OK;
  }

  int dpcuTreeSetTopology( dpcuTree *t, const uint32_t *entries, const uint32_t *levelOffsets, int numLevels, size_t numNodes )
  {
    dpcu::Range nvtxRange( "dpcuTreeSetTopology" );
    DPCU_REQUIRE( t, "tree is NULL" );
    DPCU_REQUIRE( numLevels >= 0 && ( numLevels == 0 || ( entries && levelOffsets ) ), "NULL topology" );
    DPCU_REQUIRE( numNodes >= 1 && numNodes < ( size_t( 1 ) << 30 ), "numNodes must be in [1, 2^30)" );
    DPCU_REQUIRE( numNodes >= t->numNodes, "the node array never shrinks (Tree::resizeDataStructures only grows, Tree.cpp:117-126)" );
    size_t numEntries = numLevels ? levelOffsets[numLevels] : 0;
    for ( int l = 0; l < numLevels; ++l ) DPCU_REQUIRE( levelOffsets[l] <= levelOffsets[l + 1], "levelOffsets must be ascending" );
    for ( size_t e = 0; e < numEntries; ++e )
    {
      // Tree::addTransform throws for an invalid parent (Tree.cpp:56-59)
      if ( entries[2 * e] >= numNodes || entries[2 * e + 1] >= numNodes || entries[2 * e + 1] == 0 )
        return dpcu::fail( DPCU_ERR_INVALID_VALUE, "dpcuTreeSetTopology: entry %zu {%u,%u} out of range", e, entries[2 * e], entries[2 * e + 1] );
    }
    dpcu::DeviceGuard guard( t->device );
    cudaStream_t s = t->stream;
    DPCU_CUDA( t->done.orderBefore( s ) );            // a compute may still be running on a caller stream
    size_t oldNodes = t->numNodes;
    size_t oldWords = dpcu::divUp( oldNodes, 32 ), newWords = dpcu::divUp( numNodes, 32 );
    DPCU_TRY( t->local.reserve( numNodes * 64, true, s ) );
    DPCU_TRY( t->world.reserve( numNodes * 64, true, s ) );
    size_t oldDirtyCap = t->dirtyLocal.capacity;
    DPCU_TRY( t->dirtyLocal.reserve( newWords * 4, true, s ) );
    DPCU_TRY( t->dirtyWorld.reserve( newWords * 4, true, s ) );
    if ( t->dirtyLocal.capacity > oldDirtyCap )
    {
      size_t keep = oldDirtyCap < oldWords * 4 ? oldDirtyCap : oldWords * 4;
      DPCU_CUDA( cudaMemsetAsync( static_cast<char *>( t->dirtyLocal.ptr ) + keep, 0, t->dirtyLocal.capacity - keep, s ) );
      DPCU_CUDA( cudaMemsetAsync( static_cast<char *>( t->dirtyWorld.ptr ) + keep, 0, t->dirtyWorld.capacity - keep, s ) );
    }
    if ( numNodes > oldNodes )
    {
      uint32_t cnt = uint32_t( numNodes - oldNodes );
      dpcu::treeInitNodesKernel<<<unsigned( dpcu::divUp( size_t( cnt ) * 4, 256 ) ), 256, 0, s>>>(
        static_cast<float4 *>( t->local.ptr ), static_cast<float4 *>( t->world.ptr ), static_cast<uint32_t *>( t->dirtyLocal.ptr ),
        uint32_t( oldNodes ), cnt );
      DPCU_CUDA( cudaGetLastError() );
      ++t->launches;
    }
    DPCU_TRY( t->entries.reserve( ( numEntries ? numEntries : 1 ) * 8, false, s ) );
    if ( numEntries )
    {
      DPCU_CUDA( cudaMemcpyAsync( t->entries.ptr, entries, numEntries * 8, cudaMemcpyHostToDevice, s ) );
    }
    DPCU_CUDA( cudaStreamSynchronize( s ) );
    t->numNodes = numNodes;
    t->numEntries = numEntries;
    ++t->topologyVersion;
    t->levelOffsets.assign( levelOffsets, levelOffsets + ( numLevels ? numLevels + 1 : 0 ) );
    return DPCU_OK;
  }

  int dpcuTreeMarkDirty( dpcuTree *t, size_t first, size_t count )
  {
    DPCU_REQUIRE( t, "tree is NULL" );
    DPCU_REQUIRE( first + count <= t->numNodes, "range exceeds node count" );
    if ( !count ) return DPCU_OK;
    dpcu::DeviceGuard guard( t->device );
    size_t words = ( ( first + count - 1 ) >> 5 ) - ( first >> 5 ) + 1;
    DPCU_CUDA( t->done.orderBefore( t->stream ) );    // the previous compute clears the dirty bits at its end
    dpcu::treeMarkRangeKernel<<<unsigned( dpcu::divUp( words, 256 ) ), 256, 0, t->stream>>>(
      static_cast<uint32_t *>( t->dirtyLocal.ptr ), uint32_t( first ), uint32_t( count ) );
    DPCU_CUDA( cudaGetLastError() );
    ++t->launches;
    return DPCU_OK;
  }

  int dpcuTreeSetLocals( dpcuTree *t, size_t first, size_t count, const float *matrices, int memspace )
  {
    dpcu::Range nvtxRange( "dpcuTreeSetLocals" );
    DPCU_REQUIRE( t, "tree is NULL" );
    DPCU_REQUIRE( first + count <= t->numNodes, "range exceeds node count" );
    DPCU_REQUIRE( matrices || !count, "matrices is NULL" );
    DPCU_REQUIRE( memspace == DPCU_MEM_HOST || memspace == DPCU_MEM_DEVICE, "bad memspace" );
    if ( !count ) return DPCU_OK;
    dpcu::DeviceGuard guard( t->device );
    char *dst = static_cast<char *>( t->local.ptr ) + first * 64;
    DPCU_CUDA( t->done.orderBefore( t->stream ) );
    DPCU_CUDA( cudaMemcpyAsync( dst, matrices, count * 64, memspace == DPCU_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, t->stream ) );
    DPCU_TRY( dpcuTreeMarkDirty( t, first, count ) );
    if ( memspace == DPCU_MEM_HOST ) DPCU_CUDA( cudaStreamSynchronize( t->stream ) );
    return DPCU_OK;
  }

  int dpcuTreeUpdateLocals( dpcuTree *t, const uint32_t *indices, size_t n, const float *matrices, int memspace )
  {
    dpcu::Range nvtxRange( "dpcuTreeUpdateLocals" );
    DPCU_REQUIRE( t, "tree is NULL" );
    DPCU_REQUIRE( !n || ( indices && matrices ), "NULL argument" );
    DPCU_REQUIRE( memspace == DPCU_MEM_HOST, "scattered updates take host memory (indices and matrices)" );
    DPCU_REQUIRE( n < ( size_t( 1 ) << 30 ), "too many updates" );
    if ( !n ) return DPCU_OK;
    for ( size_t i = 0; i < n; ++i ) DPCU_REQUIRE( indices[i] < t->numNodes, "node index out of range" );
    dpcu::DeviceGuard guard( t->device );
    DPCU_CUDA( t->done.orderBefore( t->stream ) );
    DPCU_TRY( t->scratch.reserve( n * 68, false, t->stream ) );
    char *s = static_cast<char *>( t->scratch.ptr );
    DPCU_CUDA( cudaMemcpyAsync( s, matrices, n * 64, cudaMemcpyHostToDevice, t->stream ) );
    DPCU_CUDA( cudaMemcpyAsync( s + n * 64, indices, n * 4, cudaMemcpyHostToDevice, t->stream ) );
    dpcu::treeScatterLocalsKernel<<<unsigned( dpcu::divUp( n * 4, 256 ) ), 256, 0, t->stream>>>(
      reinterpret_cast<uint32_t const *>( s + n * 64 ), reinterpret_cast<float4 const *>( s ), uint32_t( n ), uint32_t( t->numNodes ),
      static_cast<float4 *>( t->local.ptr ), static_cast<uint32_t *>( t->dirtyLocal.ptr ) );
    DPCU_CUDA( cudaGetLastError() );
    ++t->launches;
    DPCU_CUDA( cudaStreamSynchronize( t->stream ) );
    return DPCU_OK;
  }

  int dpcuTreeCompute( dpcuTree *t, dpcuStream *stream )
  {
    dpcu::Range nvtxRange( "dpcuTreeCompute" );
    DPCU_REQUIRE( t, "tree is NULL" );
    DPCU_REQUIRE( t->numNodes >= 1, "no topology set" );
    dpcu::DeviceGuard guard( t->device );
    cudaStream_t s = stream ? stream->stream : t->stream;
    size_t levels = t->levelOffsets.empty() ? 0 : t->levelOffsets.size() - 1;
    DPCU_TRY( dpcu::treeBeginCompute( t, s ) );
    DPCU_TRY( dpcu::treeComputeLevels( t, s, 0, levels ) );
    DPCU_TRY( dpcu::treeEndCompute( t, s ) );
    return DPCU_OK;
  }

  int dpcuTreeWorldDevicePointer( dpcuTree *t, const float **deviceMatrices, size_t *numNodes )
  {
    DPCU_REQUIRE( t && deviceMatrices, "NULL argument" );
    *deviceMatrices = static_cast<float const *>( t->world.ptr );
    if ( numNodes ) *numNodes = t->numNodes;
    return DPCU_OK;
  }

  int dpcuTreeLocalDevicePointer( dpcuTree *t, float **deviceMatrices, size_t *numNodes )
  {
    DPCU_REQUIRE( t && deviceMatrices, "NULL argument" );
    *deviceMatrices = static_cast<float *>( t->local.ptr );
    if ( numNodes ) *numNodes = t->numNodes;
    return DPCU_OK;
  }

  int dpcuTreeGetWorld( dpcuTree *t, size_t first, size_t count, float *hostMatrices )
  {
    DPCU_REQUIRE( t && ( hostMatrices || !count ), "NULL argument" );
    DPCU_REQUIRE( first + count <= t->numNodes, "range exceeds node count" );
    if ( !count ) return DPCU_OK;
    dpcu::DeviceGuard guard( t->device );
    DPCU_CUDA( t->done.orderBefore( t->stream ) );
    DPCU_CUDA( cudaMemcpyAsync( hostMatrices, static_cast<char *>( t->world.ptr ) + first * 64, count * 64, cudaMemcpyDeviceToHost, t->stream ) );
    DPCU_CUDA( cudaStreamSynchronize( t->stream ) );
    return DPCU_OK;
  }

  int dpcuTreeGetDirtyWorld( dpcuTree *t, uint32_t *hostWords, size_t nWords )
  {
    DPCU_REQUIRE( t && ( hostWords || !nWords ), "NULL argument" );
    size_t have = dpcu::divUp( t->numNodes, 32 );
    DPCU_REQUIRE( nWords >= have, "nWords smaller than ceil(numNodes/32)" );
    dpcu::DeviceGuard guard( t->device );
    DPCU_CUDA( t->done.orderBefore( t->stream ) );
    DPCU_CUDA( cudaMemcpyAsync( hostWords, t->dirtyWorld.ptr, have * 4, cudaMemcpyDeviceToHost, t->stream ) );
    DPCU_CUDA( cudaStreamSynchronize( t->stream ) );
    return DPCU_OK;
  }

  int dpcuTreeGetWorldDirty( dpcuTree *t, float *hostWorld, size_t numNodes, size_t *updated )
  {
    dpcu::Range nvtxRange( "dpcuTreeGetWorldDirty" );
    DPCU_REQUIRE( t && hostWorld, "NULL argument" );
    DPCU_REQUIRE( numNodes >= t->numNodes, "hostWorld holds fewer matrices than the tree has nodes" );
    if ( updated ) *updated = 0;
    if ( !t->numNodes ) return DPCU_OK;
    dpcu::DeviceGuard guard( t->device );
    cudaStream_t s = t->stream;
    DPCU_CUDA( t->done.orderBefore( s ) );
    const size_t nWords = dpcu::divUp( t->numNodes, 32 );
    // first a count (one word back), then exactly that many {index, matrix} records through pinned staging
    DPCU_TRY( t->scratch.reserve( 256, false, s ) );
    uint32_t *counter = static_cast<uint32_t *>( t->scratch.ptr );
    // capacity grows with the dirty set: gather into what is there (the kernel counts everything, stores what fits), and
    // if it did not fit, size for the count and run once more
    for ( int attempt = 0; attempt < 2; ++attempt )
    {
      const size_t cap = t->gather.capacity / 72;
      DPCU_CUDA( cudaMemsetAsync( counter, 0, 4, s ) );
      unsigned grid = unsigned( dpcu::divUp( nWords, 256 ) );
      if ( grid > unsigned( t->smCount ) * 8u ) grid = unsigned( t->smCount ) * 8u;
      uint32_t *outIndex = static_cast<uint32_t *>( t->gather.ptr );
      float4 *outMats = reinterpret_cast<float4 *>( static_cast<char *>( t->gather.ptr ) + ( ( cap * 4 + 255 ) & ~size_t( 255 ) ) );
      dpcu::gatherDirtyWorldKernel<<<grid, 256, 0, s>>>( static_cast<uint32_t const *>( t->dirtyWorld.ptr ), static_cast<float4 const *>( t->world.ptr ),
                                                         uint32_t( nWords ), uint32_t( cap ), counter, outIndex, outMats );
      DPCU_CUDA( cudaGetLastError() );
      ++t->launches;
      uint32_t count = 0;
      DPCU_CUDA( cudaMemcpyAsync( &count, counter, 4, cudaMemcpyDeviceToHost, s ) );
      DPCU_CUDA( cudaStreamSynchronize( s ) );
      if ( count <= cap )
      {
        if ( count )
        {
          DPCU_TRY( t->gatherHost.reserve( size_t( count ) * 68 ) );
          uint32_t *hIndex = static_cast<uint32_t *>( t->gatherHost.ptr );
          float *hMats = reinterpret_cast<float *>( static_cast<char *>( t->gatherHost.ptr ) + size_t( count ) * 4 );
          DPCU_CUDA( cudaMemcpyAsync( hIndex, outIndex, size_t( count ) * 4, cudaMemcpyDeviceToHost, s ) );
          DPCU_CUDA( cudaMemcpyAsync( hMats, outMats, size_t( count ) * 64, cudaMemcpyDeviceToHost, s ) );
          DPCU_CUDA( cudaStreamSynchronize( s ) );
          for ( uint32_t k = 0; k < count; ++k ) memcpy( hostWorld + 16ull * hIndex[k], hMats + 16ull * k, 64 );
        }
        if ( updated ) *updated = count;
        return DPCU_OK;
      }
      size_t want = size_t( count ) + count / 4 + 64;
      if ( want > t->numNodes ) want = t->numNodes;
      DPCU_TRY( t->gather.reserve( want * 72 + 512, false, s ) );
    }
    return dpcu::fail( DPCU_ERR_NOT_READY, "dpcuTreeGetWorldDirty: the dirty set changed while it was being gathered" );
  }

  int dpcuTreeSetOption( dpcuTree *t, int option, size_t value )
  {
    DPCU_REQUIRE( t, "tree is NULL" );
    switch ( option )
    {
      case DPCU_TREE_OPT_WIDE_MIN_NODES: t->wideMinNodes = value; break;
      default: return dpcu::fail( DPCU_ERR_INVALID_VALUE, "dpcuTreeSetOption: unknown option %d", option );
    }
    return DPCU_OK;
  }

  int dpcuTreeGetLaunchCount( const dpcuTree *t, uint64_t *launches )
  {
    DPCU_REQUIRE( t && launches, "NULL argument" );
    *launches = t->launches;
    return DPCU_OK;
  }
}
