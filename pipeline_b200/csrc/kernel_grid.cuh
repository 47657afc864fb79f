// K2 for groups that fit the machine in one wave (BASELINE config C2: 1 Mi objects): ONE launch per cull.
//
// The direct / view-sequential kernels leave the ordered changed list to a second kernel (segment counters ->
// compactChangedKernel): at 1 Mi objects that second launch and the gap in front of it were a third of the step
// (profiles/r01_bench_c2.json: step 49 us, cull kernel 33 us).  The line-granular kernel builds the list itself,
// but its warps walk 1024 objects one 32-object step after the other - at this size that is a chain of dependent
// DRAM round trips on a fifth of the machine (57 us).  A decoupled look-back over CTA tiles does not fit either:
// all resident CTAs finish their first tile together, so the walk back to tile 0 is ~40 dependent L2 reads.
//
// This form keeps one thread per object and makes the whole grid co-resident (cooperative launch, grid =
// SMs x resident CTAs, every CTA owns a CONTIGUOUS run of 256-object tiles):
//   phase 1  cull the CTA's tiles: bitset words, flipped-bit words, the CTA's number of flips per view (no atomics);
//            the transform index of the thread's object in the NEXT tile is fetched one tile ahead, so a tile's six
//            16-byte loads go out together instead of as two dependent round trips;
//   barrier  one grid-wide barrier (arrive counter + acquire spin; the CTAs are co-resident by construction);
//   phase 2  every CTA sums the totals of the CTAs before it (<= 1184 values from L2), then expands its own
//            flipped-bit words into the list at that offset - ascending group index order falls out of the layout
//            (BitArray::traverseBits order, dp/util/BitArray.h:127-136) - and stores its bitset words / list run into
//            the host mirror and the peers as coalesced runs.  The last CTA writes the length.
// Replaces ResultBitSet::updateChanged (dp/culling/src/ResultBitSet.cpp:61-108) like the other list builders.
#pragma once

namespace dpcu
{
  // polls go to L2 and leave the SM's L1 alone: an acquire load is compiled to LDG + CCTL.IVALL, and that per-poll L1
  // invalidation slowed down the CTAs of the same SM that were still culling (measured: 45 us instead of 25 us per
  // 1 Mi-object cull); the one fence after the loop orders the reads of phase 2
  __device__ __forceinline__ uint32_t ldRelaxedGpu( uint32_t const *p )
  {
    uint32_t v;
    asm volatile( "ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"( v ) : "l"( p ) : "memory" );
    return v;
  }

  // all CTAs of a cooperative launch; `arrive` is zero at launch and zeroed again by gridDepart
  __device__ __forceinline__ void gridArriveAndWait( uint32_t *arrive )
  {
    __syncthreads();
    if ( threadIdx.x == 0 )
    {
      __threadfence();
      atomicAdd( arrive, 1u );
      while ( ldRelaxedGpu( arrive ) < gridDim.x ) __nanosleep( 100 );
      __threadfence();
    }
    __syncthreads();
  }
  __device__ __forceinline__ void gridDepart( uint32_t *arrive, uint32_t *depart )
  {
    if ( threadIdx.x == 0 && atomicAdd( depart, 1u ) == gridDim.x - 1 )
    {
      *arrive = 0u;                                          // everyone has passed the barrier: ready for the next cull
      *depart = 0u;
    }
  }

#ifndef DPCU_GRID_PIPE
#define DPCU_GRID_PIPE 1              // 1 view: the next tile's loads in flight during this tile's arithmetic
#endif
#ifndef DPCU_GRID_MIN_CTAS1
#define DPCU_GRID_MIN_CTAS1 ( DPCU_GRID_PIPE ? 4 : 6 )
#endif
  template <int NV>
  __global__ void __launch_bounds__( kCullThreads, NV == 1 ? DPCU_GRID_MIN_CTAS1 : 4 )
  cullGridKernel( const __grid_constant__ CullArgs<NV> a )
  {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    __shared__ f32x2 sP[NV * 8];
    __shared__ FilterScratch<NV> sScratch[NV > 1 ? kCullThreads / 32 : 1];
    __shared__ uint32_t sPart[kCullThreads / 32][NV];
    __shared__ uint32_t sBase;
    if ( NV > 1 ) fillViewTable<NV>( sP, a );

    // this CTA's contiguous run of tiles
    const uint32_t q = a.nTiles / gridDim.x, r = a.nTiles % gridDim.x;
    const uint32_t tile0 = blockIdx.x * q + min( blockIdx.x, r ), tile1 = tile0 + q + ( blockIdx.x < r ? 1u : 0u );

    // ---------------- phase 1
    uint32_t myFlips = 0;                                     // lane v < NV: flips of view v seen by this warp
    uint32_t idxNext = 0;
    {
      const uint32_t i0 = min( tile0 * kCullThreads + threadIdx.x, a.n - 1u );
      if ( tile0 < tile1 ) idxNext = __ldg( reinterpret_cast<uint32_t const *>( a.lowerIdx + i0 ) + 3 );
    }
    // One view: the kernel is lean enough (32 registers) to keep the NEXT tile's six loads in flight while this tile
    // is evaluated (register double buffer, 64 registers, 4 CTAs per SM).  At this size all CTAs of an SM start together
    // and run in step, so without it every tile is "everybody loads, then everybody computes".
    constexpr bool kPipe = ( NV == 1 ) && DPCU_GRID_PIPE;
    float4 nLo, nEx, nM0, nM1, nM2, nM3;
    uint32_t idxNext2 = 0;
    if ( kPipe )
    {
      const uint32_t i0 = min( tile0 * kCullThreads + threadIdx.x, a.n - 1u );
      const uint32_t i1 = min( i0 + kCullThreads, a.n - 1u );
      if ( tile0 + 1 < tile1 ) idxNext2 = __ldg( reinterpret_cast<uint32_t const *>( a.lowerIdx + i1 ) + 3 );
      float4 const *m = a.mats + 4ull * idxNext;
      nLo = ldStream( a.lowerIdx + i0 ); nEx = ldStream( a.extent + i0 );
      ldMatrix( m, nM0, nM1, nM2, nM3 );
      idxNext = idxNext2;
    }
#pragma unroll 1
    for ( uint32_t tile = tile0; tile < tile1; ++tile )
    {
      const uint32_t i        = tile * kCullThreads + threadIdx.x;
      const bool     live     = i < a.n;
      const bool     wordLive = ( i - lane ) < a.n;
      const uint32_t word     = i >> 5;
      uint32_t oldBits = 0;
      if ( lane < NV && wordLive ) oldBits = a.out[lane].bits[word];
      float4 lo, ex, m0, m1, m2, m3;
      if ( kPipe )
      {
        lo = nLo; ex = nEx; m0 = nM0; m1 = nM1; m2 = nM2; m3 = nM3;
        if ( tile + 1 < tile1 )
        {
          const uint32_t i1 = min( i + kCullThreads, a.n - 1u );
          float4 const *m = a.mats + 4ull * idxNext;
          nLo = ldStream( a.lowerIdx + i1 ); nEx = ldStream( a.extent + i1 );
          ldMatrix( m, nM0, nM1, nM2, nM3 );
        }
        if ( tile + 2 < tile1 ) idxNext = __ldg( reinterpret_cast<uint32_t const *>( a.lowerIdx + min( i + 2u * kCullThreads, a.n - 1u ) ) + 3 );
      }
      else
      {
        // lanes past the end re-read the last object; `live` drops their results
        const uint32_t ic = min( i, a.n - 1u );
        lo = ldStream( a.lowerIdx + ic );
        ex = ldStream( a.extent + ic );
        float4 const *m = a.mats + 4ull * idxNext;
        ldMatrix( m, m0, m1, m2, m3 );
        if ( tile + 1 < tile1 ) idxNext = __ldg( reinterpret_cast<uint32_t const *>( a.lowerIdx + min( i + kCullThreads, a.n - 1u ) ) + 3 );
      }
      const Obb obb = makeObb( lo.x, lo.y, lo.z, ex.x, ex.y, ex.z, m0, m1, m2, m3 );
      uint32_t myWord;
      if ( NV == 1 )
      {
        myWord = __ballot_sync( 0xffffffffu, obbVisible( obb, a.vp[0][0], a.vp[0][1], a.vp[0][2], a.vp[0][3] ) & live );
      }
      else
      {
        const bool affine = !live || ( obb.pt.w == 1.0f && obb.ax.w == 0.0f && obb.ay.w == 0.0f && obb.az.w == 0.0f );
        const bool fast   = __all_sync( 0xffffffffu, affine ) && a.vpFinite;
        if ( fast && a.useFilter )
        {
          myWord = cullViewsFiltered<NV>( obb, a.filter, sP, sScratch[NV > 1 ? warp : 0], a.onePair, live, lane );
        }
        else
        {
          const ObbPairs ob = broadcastObb( obb );
          myWord = fast ? cullViews<NV, true>( ob, a.vp, a.onePair, live, lane ) : cullViews<NV, false>( ob, a.vp, a.onePair, live, lane );
        }
      }
      if ( lane < NV && wordLive )
      {
        ViewOut const &o = a.out[lane];
        o.bits[word] = myWord;
        if ( a.buildChanged )
        {
          const uint32_t c = oldBits ^ myWord;
          o.chg[word] = c;
          myFlips += __popc( c );
        }
      }
    }
    const uint32_t word0 = tile0 * ( kCullThreads / 32 );
    const uint32_t nWords = ( a.n + 31u ) >> 5;
    const uint32_t word1 = min( tile1 * ( kCullThreads / 32 ), nWords );
    if ( a.buildChanged )
    {
      if ( lane < NV ) sPart[warp][lane] = myFlips;
      __syncthreads();
      if ( threadIdx.x < NV )
      {
        uint32_t total = 0;
#pragma unroll
        for ( int w = 0; w < kCullThreads / 32; ++w ) total += sPart[w][threadIdx.x];
        a.out[threadIdx.x].gridTotals[blockIdx.x] = total;
      }
      // ---------------- barrier: every CTA's totals (and words) are out
      gridArriveAndWait( a.done );
    }

    // ---------------- phase 2
#pragma unroll 1
    for ( int v = 0; v < NV; ++v )
    {
      ViewOut const &o = a.out[v];
      // whole-run stores of the finished words: host mirror (PCIe) and peers (NVLink)
      if ( o.mirror || a.nPeers )
      {
        __syncthreads();                                     // (without a list: the words of phase 1 are this CTA's own)
        for ( uint32_t w = word0 + threadIdx.x; w < word1; w += kCullThreads )
        {
          const uint32_t x = __ldcg( o.bits + w );
          if ( o.mirror ) o.mirror[w] = x;
          for ( uint32_t p = 0; p < a.nPeers; ++p )
          {
            if ( o.peer[p] ) o.peer[p][a.peerWordOffset + w] = x;
          }
        }
      }
      if ( !a.buildChanged ) continue;
      // changes before this CTA
      uint32_t part = 0;
      for ( uint32_t b = threadIdx.x; b < blockIdx.x; b += kCullThreads ) part += __ldcg( o.gridTotals + b );
#pragma unroll
      for ( int d = 16; d > 0; d >>= 1 ) part += __shfl_xor_sync( 0xffffffffu, part, d );
      __syncthreads();
      if ( lane == 0 ) sPart[warp][0] = part;
      __syncthreads();
      if ( threadIdx.x == 0 )
      {
        uint32_t s = 0;
#pragma unroll
        for ( int w = 0; w < kCullThreads / 32; ++w ) s += sPart[w][0];
        sBase = s;
      }
      __syncthreads();
      uint32_t run = sBase;
      // expand this CTA's flipped-bit words, 256 words per pass
      for ( uint32_t blk = word0; blk < word1; blk += kCullThreads )
      {
        const uint32_t w = blk + threadIdx.x;
        uint32_t c = ( w < word1 ) ? __ldcg( o.chg + w ) : 0u;
        const uint32_t pc = __popc( c );
        uint32_t incl = pc;
#pragma unroll
        for ( int d = 1; d < 32; d <<= 1 )
        {
          const uint32_t t = __shfl_up_sync( 0xffffffffu, incl, d );
          if ( lane >= d ) incl += t;
        }
        __syncthreads();
        if ( lane == 31 ) sPart[warp][0] = incl;
        __syncthreads();
        uint32_t off = run + incl - pc, passTotal = 0;
#pragma unroll
        for ( int k = 0; k < kCullThreads / 32; ++k )
        {
          const uint32_t t = sPart[k][0];
          if ( k < int( warp ) ) off += t;
          passTotal += t;
        }
        const uint32_t base = w << 5;
        while ( c )
        {
          o.changed[off++] = base + uint32_t( __ffs( c ) - 1 );
          c &= c - 1;
        }
        if ( o.hostChanged && passTotal )
        {
          // the pass's run again as coalesced stores into the pinned host mirror (the entries just written are in L2)
          __syncthreads();
          for ( uint32_t k = threadIdx.x; k < passTotal; k += kCullThreads )
          {
            if ( run + k < o.hostCap ) o.hostChanged[run + k] = __ldcg( o.changed + run + k );
          }
        }
        run += passTotal;
      }
      if ( blockIdx.x == gridDim.x - 1 && threadIdx.x == 0 )
      {
        o.prefix[a.nSegs] = run;                             // where the compaction path keeps the length of the list
        if ( o.hostCount ) *o.hostCount = run;
      }
    }
    if ( a.buildChanged ) gridDepart( a.done, a.done + 2 );
  }
}
