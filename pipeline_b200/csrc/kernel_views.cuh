// K2, view-sequential form for several views (arithmetic in cull_views.cuh / cull_filter.cuh).
#pragma once

namespace dpcu
{
  // ------------------------------------------------------------------------------------------
  // K2, view-sequential variant for V >= 2 views (cull_views.cuh): one thread per object, the
  // OBB is built once, then the views run one after the other through packed f32x2 arithmetic.
  // Lane v of each warp owns view v's epilogue (previous word, new word, flipped bits, segment
  // counter, peer stores), so the V epilogues of a warp are one divergent block instead of V.
#ifndef DPCU_VIEWS_MIN_CTAS
#define DPCU_VIEWS_MIN_CTAS 4
#endif
#ifndef DPCU_VIEWS_PREFETCH
#define DPCU_VIEWS_PREFETCH 1
#endif
  __device__ __forceinline__ void prefetchL2( void const *p )
  {
    asm volatile( "prefetch.global.L2 [%0];" :: "l"( p ) );
  }

  template <int NV, bool kCount>
  __global__ void __launch_bounds__( kCullThreads, DPCU_VIEWS_MIN_CTAS )
  cullViewsKernel( const __grid_constant__ CullArgs<NV> a )
  {
    const uint32_t lane = threadIdx.x & 31u;
    __shared__ f32x2 sP[NV * 8];
    __shared__ FilterScratch<NV> sScratch[kCullThreads / 32];
    fillViewTable<NV>( sP, a );
#if DPCU_VIEWS_PREFETCH
    // Two dependent DRAM round trips (object -> its matrix) head every tile and this kernel runs at
    // 8 warps per scheduler at most, so a third of the warp time was spent waiting on them (ncu:
    // long_scoreboard 2.05 warps per issue).  The transform index of this thread's object two tiles
    // ahead is fetched now (one register; it also pulls that tile's lowerIdx lines in), the index
    // fetched a tile ago turns into an L2 prefetch of the next tile's matrix and extent lines.
    // (Only for long runs of tiles per CTA: on the small groups AUTO gives this kernel now - below 2.6 - 7 M objects -
    // the look-ahead costs more than it hides: 1 Mi x 6 views 44.0 -> 42.0 us, 2 Mi x 3 views 54.2 -> 52.2 us without.)
    const uint32_t strideObjects = gridDim.x * kCullThreads;
    const bool     ahead = a.nTiles >= 32u * gridDim.x;
    uint32_t idxNext = 0;
    if ( ahead )
    {
      const uint32_t i1 = blockIdx.x * kCullThreads + threadIdx.x + strideObjects;
      if ( i1 < a.n && i1 >= strideObjects ) idxNext = __ldg( reinterpret_cast<uint32_t const *>( a.lowerIdx + i1 ) + 3 );
    }
#endif
    for ( uint32_t tile = blockIdx.x; tile < a.nTiles; tile += gridDim.x )
    {
      const uint32_t i        = tile * kCullThreads + threadIdx.x;
      const bool     live     = i < a.n;
      const bool     wordLive = ( i - lane ) < a.n;
      const uint32_t word     = i >> 5;
#if DPCU_VIEWS_PREFETCH
      uint32_t idxNext2 = 0;
      if ( ahead )
      {
        const uint32_t i1 = i + strideObjects, i2 = i1 + strideObjects;
        if ( i2 < a.n && i2 > i1 ) idxNext2 = __ldg( reinterpret_cast<uint32_t const *>( a.lowerIdx + i2 ) + 3 );
        if ( i1 < a.n && i1 > i )
        {
          prefetchL2( a.mats + 4ull * idxNext );
          if ( ( lane & 7u ) == 0 ) prefetchL2( a.extent + i1 );
        }
      }
#endif

      uint32_t oldBits = 0;
      if ( lane < NV && wordLive ) oldBits = a.out[lane].bits[word];

      Obb obb;
      obb.pt = obb.ax = obb.ay = obb.az = make_float4( 0.f, 0.f, 0.f, 0.f );
      if ( live )
      {
        const float4 lo = ldStream( a.lowerIdx + i );
        const float4 ex = ldStream( a.extent + i );
        float4 const *m = a.mats + 4ull * __float_as_uint( lo.w );
        float4 m0, m1, m2, m3;
        ldMatrix( m, m0, m1, m2, m3 );
        obb = makeObb( lo.x, lo.y, lo.z, ex.x, ex.y, ex.z, m0, m1, m2, m3 );
      }
      const bool affine = !live || ( obb.pt.w == 1.0f && obb.ax.w == 0.0f && obb.ay.w == 0.0f && obb.az.w == 0.0f );
      const bool fast   = __all_sync( 0xffffffffu, affine ) && a.vpFinite;
      uint32_t myWord;
      if ( fast && a.useFilter && kCount )
      {
        myWord = cullViewsFiltered<NV>( obb, a.filter, sP, sScratch[threadIdx.x >> 5], a.onePair, live, lane );
      }
      else
      {
        const ObbPairs ob = broadcastObb( obb );
        myWord = fast ? cullViews<NV, true, kCount>( ob, a.vp, a.onePair, live, lane ) : cullViews<NV, false, kCount>( ob, a.vp, a.onePair, live, lane );
      }

      if ( lane < NV && wordLive )
      {
        ViewOut const &o = a.out[lane];
        o.bits[word] = myWord;
        if ( a.buildChanged )
        {
          const uint32_t c = oldBits ^ myWord;
          o.chg[word] = c;
          if ( c && a.countSegs ) atomicAdd( o.seg + ( word >> ( kSegObjectsLog2 - 5 ) ), __popc( c ) );
        }
      }
#if DPCU_VIEWS_PREFETCH
      idxNext = idxNext2;
#endif
    }
    if ( a.buildChanged && a.countSegs == 1 ) scanSegmentsInLastCta<NV>( a.out, a.nSegs, a.done );
  }
}
