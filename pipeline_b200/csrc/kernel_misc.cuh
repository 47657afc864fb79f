// Compaction of the changed / visible lists, object and matrix bookkeeping kernels, K4 (group bounding box).
#pragma once

namespace dpcu
{
  // ------------------------------------------------------------------------------------------
  // Ordered changed list.  One CTA per 8192-object segment and view: the number of changes before
  // the segment comes from prefix[] (written by the cull kernel's last CTA), from the segment counters
  // (summed here) or from the flipped-bit words themselves (popcounted here) - CompactArgs::selfPrefix -
  // then the CTA block-scans the popcounts of its 256 flipped-bit words and expands them; ascending
  // group index order falls out of the layout (BitArray::traverseBits order, dp/util/BitArray.h:127-136).
  struct CompactArgs
  {
    uint32_t const *chg[DPCU_MAX_VIEWS];
    uint32_t       *prefix[DPCU_MAX_VIEWS];
    uint32_t       *changed[DPCU_MAX_VIEWS];
    uint32_t       *hostChanged[DPCU_MAX_VIEWS];    // optional mirror of the list in pinned host memory ...
    uint32_t       *hostCount[DPCU_MAX_VIEWS];      // ... and of its length
    uint32_t        hostCapacity[DPCU_MAX_VIEWS];
    uint32_t        nWords;
    uint32_t        nSegs;
    int             selfPrefix;   // prefix[] holds nothing yet: 1 = every CTA popcounts the words before its segment itself,
                                  // 2 = every CTA sums the segment counters before its own (and the last one to do so zeroes them)
    uint32_t       *seg[DPCU_MAX_VIEWS];
    uint32_t const *bits[DPCU_MAX_VIEWS];           // optional: the new visibility words ...
    uint32_t       *hostBits[DPCU_MAX_VIEWS];       // ... and their mirror in pinned host memory (stored by this kernel, segment by segment)
    uint32_t       *done;         // mode 2: ticket of the CTAs that have read the counters
  };

  __global__ void __launch_bounds__( 256 ) compactChangedKernel( const __grid_constant__ CompactArgs a )
  {
    const uint32_t v    = blockIdx.y;
    const uint32_t s    = blockIdx.x;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    // launched with programmatic stream serialization behind the cull kernel: wait for its results here
    // (a no-op when the launch carries no such dependency, e.g. behind segmentPopcountKernel)
    cudaGridDependencySynchronize();
    // the segment's indices are expanded into shared memory first, so that the list (and its host
    // mirror, over PCIe) is written as one contiguous, coalesced run per segment
    __shared__ uint32_t sIdx[1u << kSegObjectsLog2];
    __shared__ uint32_t sWarp[8];
    __shared__ uint32_t sBefore[8];
    const uint32_t w = s * kSegWords + threadIdx.x;
    uint32_t c = ( w < a.nWords ) ? a.chg[v][w] : 0u;
    // the bitset mirror of the forms that do not store whole lines themselves (direct, views, staged, fused leaf): the
    // segment's 1 KiB crosses PCIe from here, coalesced, instead of as a copy queued behind this kernel (15 us per step
    // at 1 Mi objects: copy-engine start-up and completion)
    if ( a.hostBits[v] && w < a.nWords ) a.hostBits[v][w] = __ldcg( a.bits[v] + w );
    const uint32_t pc = __popc( c );
    uint32_t base0, count;
    if ( a.selfPrefix )
    {
      // prefix[] holds nothing yet: this CTA finds the number of flips before its segment (`before`, summed over the
      // threads below) and in it (`mine`) itself.
      uint32_t before = 0, mine = 0;
      if ( a.selfPrefix == 2 )
      {
        // The cull kernel counted flips per segment (fire-and-forget atomics) and stopped there: no fence, no ticket, no
        // serial scan by its last CTA.  This CTA sums the counters before its own - one L2 round trip for up to 4096
        // segments, 16 loads per thread at most.
        uint32_t const *seg = a.seg[v];
        for ( uint32_t k = threadIdx.x; k <= s; k += 256 )
        {
          const uint32_t x = __ldcg( seg + k );
          if ( k < s ) before += x; else mine = x;
        }
      }
      else
      {
        // The cull kernel kept no counters at all (groups of at most 256 segments): the flips before this segment are the
        // popcount of s KiB of flipped-bit words, read from L2 as 16-byte vectors.
        uint4 const *q = reinterpret_cast<uint4 const *>( a.chg[v] );
        const uint32_t nVec = s * ( kSegWords / 4 );
#pragma unroll 16
        for ( uint32_t k = threadIdx.x; k < nVec; k += 256 )
        {
          const uint4 x = __ldcg( q + k );
          before += __popc( x.x ) + __popc( x.y ) + __popc( x.z ) + __popc( x.w );
        }
        mine = pc;
      }
#pragma unroll
      for ( int d = 16; d > 0; d >>= 1 )
      {
        before += __shfl_xor_sync( 0xffffffffu, before, d );
        mine   += __shfl_xor_sync( 0xffffffffu, mine, d );
      }
      if ( lane == 0 ) { sBefore[warp] = before; sWarp[warp] = mine; }
      __syncthreads();
      base0 = 0; count = 0;
#pragma unroll
      for ( int k = 0; k < 8; ++k ) { base0 += sBefore[k]; count += sWarp[k]; }
      __syncthreads();                               // sWarp is reused by the scan below
      if ( threadIdx.x == 0 )
      {
        a.prefix[v][s] = base0;                      // what the other consumers of the result read (count = prefix[nSegs])
        if ( s == a.nSegs - 1 )
        {
          a.prefix[v][a.nSegs] = base0 + count;
          if ( a.hostCount[v] ) *a.hostCount[v] = base0 + count;
        }
      }
      if ( a.selfPrefix == 2 )
      {
        // the last CTA to have read the counters (every thread's reads went through the barriers above) zeroes them, and
        // the ticket, for the next cull
        __shared__ uint32_t sLast;
        if ( threadIdx.x == 0 ) sLast = ( atomicAdd( a.done, 1u ) == gridDim.x * gridDim.y - 1 ) ? 1u : 0u;
        __syncthreads();
        if ( sLast )
        {
          for ( uint32_t vv = 0; vv < gridDim.y; ++vv )
            for ( uint32_t k = threadIdx.x; k < a.nSegs; k += 256 ) a.seg[vv][k] = 0u;
          if ( threadIdx.x == 0 ) *a.done = 0u;
        }
      }
    }
    else
    {
      base0 = a.prefix[v][s];
      count = a.prefix[v][s + 1] - base0;
      if ( s == 0 && threadIdx.x == 0 && a.hostCount[v] ) *a.hostCount[v] = a.prefix[v][a.nSegs];
    }
    if ( count == 0 ) return;                        // nothing changed in this segment
    uint32_t incl = pc;
#pragma unroll
    for ( int d = 1; d < 32; d <<= 1 )
    {
      uint32_t t = __shfl_up_sync( 0xffffffffu, incl, d );
      if ( lane >= d ) incl += t;
    }
    if ( lane == 31 ) sWarp[warp] = incl;
    __syncthreads();
    uint32_t off = incl - pc;
    for ( uint32_t k = 0; k < warp; ++k ) off += sWarp[k];
    const uint32_t base = w << 5;
    while ( c )
    {
      const uint32_t b = __ffs( c ) - 1;
      sIdx[off++] = base + b;
      c &= c - 1;
    }
    __syncthreads();
    uint32_t *out = a.changed[v] + base0;
    uint32_t *host = a.hostChanged[v];
    const uint32_t hostRoom = a.hostCapacity[v] > base0 ? a.hostCapacity[v] - base0 : 0u;
    for ( uint32_t k = threadIdx.x; k < count; k += 256 )
    {
      const uint32_t x = sIdx[k];
      out[k] = x;
      if ( host && k < hostRoom ) host[base0 + k] = x;
    }
  }

  // ------------------------------------------------------------------------------------------
  // Visible-instance list (the consumer of the result stays on the GPU).
  // Per-segment popcounts of the visibility words, scanned by the last CTA exactly like the
  // changed counts; compactChangedKernel then expands bits[] instead of chg[].
  __global__ void __launch_bounds__( kCullThreads ) segmentPopcountKernel( uint32_t const *bits, uint32_t nWords, uint32_t nSegs,
                                                                          uint32_t *seg, uint32_t *prefix, uint32_t *done )
  {
    __shared__ uint32_t sPart[kCullThreads / 32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for ( uint32_t s = blockIdx.x; s < nSegs; s += gridDim.x )
    {
      const uint32_t w = s * kSegWords + threadIdx.x;
      uint32_t pc = ( w < nWords ) ? __popc( bits[w] ) : 0u;
#pragma unroll
      for ( int d = 16; d > 0; d >>= 1 ) pc += __shfl_xor_sync( 0xffffffffu, pc, d );
      if ( lane == 0 ) sPart[warp] = pc;
      __syncthreads();
      if ( threadIdx.x == 0 )
      {
        uint32_t total = 0;
        for ( int k = 0; k < kCullThreads / 32; ++k ) total += sPart[k];
        seg[s] = total;
      }
      __syncthreads();
    }
    ViewOut out[1];
    out[0].seg = seg;
    out[0].prefix = prefix;
    scanSegmentsInLastCta<1>( out, nSegs, done );
  }

  // ------------------------------------------------------------------------------------------
  // object upload: pack transformIndex into lower.w, zero extent.w, track the largest index
  __global__ void packObjectsKernel( float4 const *lower, float4 const *extent, uint32_t const *tidx, uint32_t n,
                                     float4 *lowerIdx, float4 *extentOut, uint32_t *maxIndex )
  {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t m = 0;
    if ( i < n )
    {
      float4 lo = lower[i];
      float4 ex = extent[i];
      uint32_t t = tidx ? tidx[i] : __float_as_uint( lo.w );
      lo.w = __uint_as_float( t );
      ex.w = 0.0f;
      lowerIdx[i]  = lo;
      extentOut[i] = ex;
      m = t;
    }
#pragma unroll
    for ( int d = 16; d > 0; d >>= 1 ) m = max( m, __shfl_xor_sync( 0xffffffffu, m, d ) );
    if ( ( threadIdx.x & 31 ) == 0 && m ) atomicMax( maxIndex, m );
  }

  // batched edit of live objects (objectSetBoundingBox / objectSetTransformIndex, swap-removes, appends):
  // object indices[k] <- (lower[k], extent[k], tidx[k]); indices past the object count are skipped
  __global__ void scatterObjectsKernel( uint32_t const *indices, float4 const *lower, float4 const *extent, uint32_t const *tidx, uint32_t k,
                                        uint32_t n, float4 *lowerIdx, float4 *extentOut, uint32_t *maxIndex )
  {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t m = 0;
    if ( t < k )
    {
      const uint32_t i = indices[t];
      if ( i < n )
      {
        float4 lo = lower[t];
        float4 ex = extent[t];
        lo.w = __uint_as_float( tidx[t] );
        ex.w = 0.0f;
        lowerIdx[i]  = lo;
        extentOut[i] = ex;
        m = tidx[t];
      }
    }
#pragma unroll
    for ( int d = 16; d > 0; d >>= 1 ) m = max( m, __shfl_xor_sync( 0xffffffffu, m, d ) );
    if ( ( threadIdx.x & 31 ) == 0 && m ) atomicMax( maxIndex, m );
  }

  // exact largest transform index of the current objects (after overwrites / removals the running maximum may be stale)
  __global__ void maxIndexKernel( float4 const *lowerIdx, uint32_t n, uint32_t *maxIndex )
  {
    uint32_t m = 0;
    for ( uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x )
    {
      m = max( m, __ldg( reinterpret_cast<uint32_t const *>( lowerIdx + i ) + 3 ) );
    }
#pragma unroll
    for ( int d = 16; d > 0; d >>= 1 ) m = max( m, __shfl_xor_sync( 0xffffffffu, m, d ) );
    if ( ( threadIdx.x & 31 ) == 0 && m ) atomicMax( maxIndex, m );
  }

  // matrices[indices[k]] = packed[k]
  __global__ void scatterMatricesKernel( uint32_t const *indices, float4 const *packed, uint32_t n, float4 *mats )
  {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;   // one thread per matrix row
    if ( t < n * 4u ) mats[4ull * indices[t >> 2] + ( t & 3u )] = packed[t];
  }

  // strided device -> packed device copy (device-side groupSetMatrices with stride != 64)
  __global__ void gatherStridedKernel( char const *src, size_t stride, uint32_t count, float *dst )
  {
    size_t t = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;   // one thread per float
    if ( t < size_t( count ) * 16 ) dst[t] = *reinterpret_cast<float const *>( src + ( t >> 4 ) * stride + ( t & 15 ) * 4 );
  }

  // ResultBitSet incarnation step (dp/culling/src/ResultBitSet.cpp:65-79): bits of objects
  // [oldN, newN) become 1, bits >= newN become 0, older bits are kept.
  __global__ void resizeBitsKernel( uint32_t *bits, uint32_t oldN, uint32_t newN, uint32_t capWords )
  {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if ( w >= capWords ) return;
    const uint64_t lo = uint64_t( w ) << 5, hi = lo + 32;
    if ( hi <= oldN && hi <= newN ) return;
    uint32_t x = bits[w];
    uint32_t keep  = ( lo >= oldN ) ? 0u : ( hi <= oldN ? ~0u : ( ~0u >> ( 32 - ( oldN - lo ) ) ) );   // bits < oldN
    uint32_t valid = ( lo >= newN ) ? 0u : ( hi <= newN ? ~0u : ( ~0u >> ( 32 - ( newN - lo ) ) ) );   // bits < newN
    x = ( ( x & keep ) | ~keep ) & valid;
    bits[w] = x;
  }

  // ResultBitSet::onNotify (dp/culling/src/ResultBitSet.cpp:110-128)
  __global__ void moveBitKernel( uint32_t *bits, uint32_t size, uint32_t oldIndex, uint32_t newIndex, uint32_t *mirror )
  {
    if ( newIndex < size )
    {
      uint32_t value = 1u;
      if ( oldIndex < size ) value = ( bits[oldIndex >> 5] >> ( oldIndex & 31 ) ) & 1u;
      uint32_t w = bits[newIndex >> 5];
      w = value ? ( w | ( 1u << ( newIndex & 31 ) ) ) : ( w & ~( 1u << ( newIndex & 31 ) ) );
      bits[newIndex >> 5] = w;
      if ( mirror ) mirror[newIndex >> 5] = w;
    }
  }

  // a frame's bit moves applied on the host mirror, the touched words written back in one batch
  __global__ void scatterWordsKernel( uint32_t const *indices, uint32_t const *words, uint32_t k, uint32_t *bits, uint32_t *mirror )
  {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if ( t < k )
    {
      bits[indices[t]] = words[t];
      if ( mirror ) mirror[indices[t]] = words[t];
    }
  }

  // ------------------------------------------------------------------------------------------
  // K4: group bounding box, ManagerBitSet::calculateBoundingBox scalar branch
  // (dp/culling/src/ManagerBitSet.cpp:268-306).  min/max are order independent, so a tree
  // reduction gives the reference's sequential Box4f::update result bit for bit, except that
  // Boxnt::update skips NaN coordinates (both comparisons false) - fminf/fmaxf do the same.
  // Signed zeros: update() keeps the first of +0/-0 it met; the final box is compared with ==
  // semantics by every consumer, and the test-suite compares with np.array_equal (-0 == +0).
  struct BoxAcc
  {
    float lo[3], hi[3];
  };

  __device__ __forceinline__ void boxUpdate( BoxAcc &b, float4 p )
  {
    b.lo[0] = fminf( b.lo[0], p.x ); b.hi[0] = fmaxf( b.hi[0], p.x );
    b.lo[1] = fminf( b.lo[1], p.y ); b.hi[1] = fmaxf( b.hi[1], p.y );
    b.lo[2] = fminf( b.lo[2], p.z ); b.hi[2] = fmaxf( b.hi[2], p.z );
  }

  __device__ __forceinline__ uint32_t orderedKey( float f )
  {
    uint32_t u = __float_as_uint( f );
    return ( u & 0x80000000u ) ? ~u : ( u | 0x80000000u );
  }

  __global__ void __launch_bounds__( 256 ) boundingBoxKernel( float4 const *lowerIdx, float4 const *extent,
                                                              float4 const *mats, uint32_t n, uint32_t *keys /* 3 min, 3 max */ )
  {
    const float FMAX = 3.402823466e+38f;
    BoxAcc b = { { FMAX, FMAX, FMAX }, { -FMAX, -FMAX, -FMAX } };
    for ( uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x )
    {
      const float4 lo = ldStream( lowerIdx + i );
      const float4 ex = ldStream( extent + i );
      float4 const *m = mats + 4ull * __float_as_uint( lo.w );
      float4 m0, m1, m2, m3;
      ldMatrix( m, m0, m1, m2, m3 );
      const Obb o = makeObb( lo.x, lo.y, lo.z, ex.x, ex.y, ex.z, m0, m1, m2, m3 );
      const float4 v0 = o.pt;
      const float4 v1 = add4( v0, o.ax );
      const float4 v2 = add4( v0, o.ay );
      const float4 v3 = add4( v1, o.ay );
      boxUpdate( b, v0 ); boxUpdate( b, v1 ); boxUpdate( b, v2 ); boxUpdate( b, v3 );
      boxUpdate( b, add4( v0, o.az ) ); boxUpdate( b, add4( v1, o.az ) );
      boxUpdate( b, add4( v2, o.az ) ); boxUpdate( b, add4( v3, o.az ) );
    }
#pragma unroll
    for ( int k = 0; k < 3; ++k )
    {
#pragma unroll
      for ( int d = 16; d > 0; d >>= 1 )
      {
        b.lo[k] = fminf( b.lo[k], __shfl_xor_sync( 0xffffffffu, b.lo[k], d ) );
        b.hi[k] = fmaxf( b.hi[k], __shfl_xor_sync( 0xffffffffu, b.hi[k], d ) );
      }
    }
    if ( ( threadIdx.x & 31 ) == 0 )
    {
#pragma unroll
      for ( int k = 0; k < 3; ++k )
      {
        atomicMin( keys + k, orderedKey( b.lo[k] ) );
        atomicMax( keys + 3 + k, orderedKey( b.hi[k] ) );
      }
    }
  }

  // ------------------------------------------------------------------------------------------
  // Multi-GPU gather behind the fused leaf kernel (C3 sharded over several GPUs): this GPU's
  // finished bitset words into every peer's full bitset at the shard's word offset, 16 bytes per thread
  // where the alignment allows (shard offsets are whole 128-byte lines, pipeline_b200/sharding.py).
  struct PeerGatherArgs
  {
    uint32_t const *bits[8];
    uint32_t       *peer[8][kMaxPeers];
    int             nViews;
    uint32_t        nPeers, nWords, wordOffset;
  };

  __global__ void __launch_bounds__( 256 ) peerGatherKernel( const __grid_constant__ PeerGatherArgs a )
  {
    const int v = blockIdx.y;
    uint32_t const *src = a.bits[v];
    const bool wide = ( a.wordOffset & 3u ) == 0;
    const uint32_t nQuads = wide ? a.nWords >> 2 : 0u;
    for ( uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < nQuads; q += gridDim.x * blockDim.x )
    {
      const uint4 w = reinterpret_cast<uint4 const *>( src )[q];
      for ( uint32_t p = 0; p < a.nPeers; ++p )
      {
        if ( a.peer[v][p] ) reinterpret_cast<uint4 *>( a.peer[v][p] + a.wordOffset )[q] = w;
      }
    }
    for ( uint32_t k = ( nQuads << 2 ) + blockIdx.x * blockDim.x + threadIdx.x; k < a.nWords; k += gridDim.x * blockDim.x )
    {
      const uint32_t w = src[k];
      for ( uint32_t p = 0; p < a.nPeers; ++p )
      {
        if ( a.peer[v][p] ) a.peer[v][p][a.wordOffset + k] = w;
      }
    }
  }
}
