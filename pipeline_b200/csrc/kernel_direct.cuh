// K2, direct form.  Compiled twice: exact (-fmad=false) and as the reporting-only FMA variant (dpcu_cull_fma.cu).
#pragma once

namespace dpcu
{
  // ------------------------------------------------------------------------------------------
  // K2, direct variant: one thread per object, two 16-byte loads (AABB, index) and two 256-bit loads (its matrix),
  // persistent grid-stride tiles.
  template <int NV>
  __global__ void __launch_bounds__( kCullThreads )
  DPCU_KERNEL_NAME( cullDirectKernel )( const __grid_constant__ CullArgs<NV> a )
  {
    const uint32_t lane = threadIdx.x & 31u;
    for ( uint32_t tile = blockIdx.x; tile < a.nTiles; tile += gridDim.x )
    {
      const uint32_t i    = tile * kCullThreads + threadIdx.x;
      const bool     live = i < a.n;
      const uint32_t word = i >> 5;

      // previous visibility words are fetched early by the lane that will need them
      uint32_t oldBits[NV];
      if ( lane == 0 && live )
      {
#pragma unroll
        for ( int v = 0; v < NV; ++v ) oldBits[v] = a.out[v].bits[word];
      }

      bool vis[NV];
#pragma unroll
      for ( int v = 0; v < NV; ++v ) vis[v] = false;

      if ( live )
      {
        const float4 lo = ldStream( a.lowerIdx + i );
        const float4 ex = ldStream( a.extent + i );
        float4 const *m = a.mats + 4ull * __float_as_uint( lo.w );
        float4 m0, m1, m2, m3;
        ldMatrix( m, m0, m1, m2, m3 );
        const Obb obb = makeObb( lo.x, lo.y, lo.z, ex.x, ex.y, ex.z, m0, m1, m2, m3 );
#pragma unroll
        for ( int v = 0; v < NV; ++v )
        {
          vis[v] = obbVisible( obb, a.vp[v][0], a.vp[v][1], a.vp[v][2], a.vp[v][3] );
        }
      }

#pragma unroll
      for ( int v = 0; v < NV; ++v )
      {
        const uint32_t nw = __ballot_sync( 0xffffffffu, vis[v] );
        if ( lane == 0 && live )
        {
          ViewOut const &o = a.out[v];
          o.bits[word] = nw;
          if ( a.buildChanged )
          {
            const uint32_t c = oldBits[v] ^ nw;
            o.chg[word] = c;
            if ( c && a.countSegs )
            {
              // one counter per 8192 objects: concurrently running CTAs spread over ~40 addresses.
              // (A second, coarser level of counters was measured to serialise in L2: +0.5 ms at 64 Mi objects.)
              atomicAdd( o.seg + ( word >> ( kSegObjectsLog2 - 5 ) ), __popc( c ) );
            }
          }
        }
      }
    }
    if ( a.buildChanged && a.countSegs == 1 ) scanSegmentsInLastCta<NV>( a.out, a.nSegs, a.done );
  }
}
