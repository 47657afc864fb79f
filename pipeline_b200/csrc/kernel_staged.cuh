// K2, TMA-staged form (helpers in cull_stage.cuh).
#pragma once

namespace dpcu
{
  // ------------------------------------------------------------------------------------------
  // K2, staged variant.  Every warp is an independent persistent worker with its own
  // shared-memory rings over warp-tiles of 32 objects (one bitset word per view):
  //   P1(q+2)  lane 0: a bulk TMA copy (cp.async.bulk + mbarrier) brings the tile's lowerIdx[32]
  //            stream (boxes' lower corners + transform indices) into a three-deep ring;
  //   P2(q+1)  all lanes: wait for that tile's mbarrier, read the transform indices from shared
  //            memory and gather the matrix rows with 16-byte cp.async into a two-deep ring; lane 0
  //            adds the bulk copy of extent[32].  Four neighbouring lanes fetch the four rows of
  //            one matrix, so every global request covers whole 32-byte sectors; rows land
  //            XOR-swizzled so that both the copy and the later 128-bit reads are free of bank
  //            conflicts;
  //   C(q)     all lanes: the object from shared memory -> OBB -> views -> ballots -> epilogue.
  // Every load is issued at least one tile-time before its use without spending registers on
  // prefetching, no CTA-wide barrier exists, and tiles are handed out dynamically in chunks of 32
  // warp-tiles (1024 objects = one 128-byte line of each bitset) from a global counter, so all
  // SMs stay full to the end.  6.6 KiB of shared memory per warp -> 4 CTAs (32 warps) per SM.
  constexpr uint32_t kChunkTiles = 32;            // warp-tiles per claimed chunk
  constexpr uint32_t kNoTile     = 0xffffffffu;

  struct alignas( 128 ) WarpRing
  {
    float4   lo[3][32];        // ring 3: lowerIdx tiles
    float4   ex[2][32];        // ring 2: extent tiles
    float4   m[2][128];        // ring 2: matrix of object o at o*4, 16-byte chunks XOR-swizzled by (o>>1)&3
    uint64_t loFull[3];        // mbarriers: bytes of the bulk copies have landed
    uint64_t exFull[2];
  };

  template <int NV>
  __device__ __forceinline__ void storeWord( ViewOut const &o, CullArgs<NV> const &a, uint32_t word, uint32_t nw, uint32_t old )
  {
    o.bits[word] = nw;
    if ( a.buildChanged )
    {
      const uint32_t c = old ^ nw;
      o.chg[word] = c;
      if ( c && a.countSegs ) atomicAdd( o.seg + ( word >> ( kSegObjectsLog2 - 5 ) ), __popc( c ) );
    }
  }

  template <int NV>
  __global__ void __launch_bounds__( kCullThreads, 4 )
  cullStagedKernel( const __grid_constant__ CullArgs<NV> a )
  {
    extern __shared__ __align__( 128 ) unsigned char smemRaw[];
    __shared__ f32x2 sP[NV * 8];
    __shared__ FilterScratch<NV> sScratch[kCullThreads / 32];
    if ( NV > 1 ) fillViewTable<NV>( sP, a );
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    WarpRing &ring = reinterpret_cast<WarpRing *>( smemRaw )[warp];

    const uint32_t nTiles     = ( a.n + 31u ) >> 5;                            // warp-tiles
    const uint32_t nChunks    = ( nTiles + kChunkTiles - 1 ) / kChunkTiles;
    const uint32_t totalWarps = gridDim.x * ( kCullThreads / 32 );
    // tile sequence of this warp: chunks of kChunkTiles tiles, the first one static, the following
    // ones claimed from the global counter one chunk ahead of their use (lane 0 holds the claim)
    uint32_t curChunk = blockIdx.x * ( kCullThreads / 32 ) + warp, sub = 0, claimed = 0;
    auto claim = [&]() { if ( lane == 0 ) claimed = ( curChunk < nChunks ) ? totalWarps + atomicAdd( a.chunkCounter, 1u ) : nChunks; };
    auto nextTile = [&]() -> uint32_t
    {
      if ( sub == kChunkTiles )
      {
        curChunk = __shfl_sync( 0xffffffffu, claimed, 0 );
        claim();
        sub = 0;
      }
      const uint32_t t = curChunk * kChunkTiles + sub++;
      return ( curChunk < nChunks && t < nTiles ) ? t : kNoTile;
    };
    // P1: lowerIdx of `tile` (sequence index q) -> lo ring
    auto issueLower = [&]( uint32_t q, uint32_t tile )
    {
      if ( tile == kNoTile || lane != 0 ) return;
      const uint32_t first = tile << 5, bytes = min( 32u, a.n - first ) * 16u, s = q % 3u;
      mbarArriveExpectTx( &ring.loFull[s], bytes );
      tmaLoad1d( ring.lo[s], a.lowerIdx + first, bytes, &ring.loFull[s] );
    };
    // P2: extent and gathered matrices of `tile` (sequence index q); returns the previous
    // visibility word lane v will need in the epilogue of that tile
    auto issueGather = [&]( uint32_t q, uint32_t tile ) -> uint32_t
    {
      uint32_t old = 0;
      if ( tile != kNoTile )
      {
        const uint32_t first = tile << 5, s3 = q % 3u, s2 = q & 1u;
        if ( lane == 0 )
        {
          const uint32_t bytes = min( 32u, a.n - first ) * 16u;
          mbarArriveExpectTx( &ring.exFull[s2], bytes );
          tmaLoad1d( ring.ex[s2], a.extent + first, bytes, &ring.exFull[s2] );
        }
        if ( lane < NV ) old = a.out[lane].bits[tile];
        mbarWait( &ring.loFull[s3], ( q / 3u ) & 1u );
#pragma unroll
        for ( uint32_t j = 0; j < 4; ++j )
        {
          const uint32_t o = j * 8u + ( lane >> 2 ), r = lane & 3u;
          if ( first + o < a.n )
          {
            const uint32_t idx = __float_as_uint( ring.lo[s3][o].w );
            cpAsync16( &ring.m[s2][swizzledRow( o, r )], a.mats + 4ull * idx + r );
          }
        }
      }
      cpAsyncCommit();
      return old;
    };

    if ( lane == 0 )
    {
      for ( int s = 0; s < 3; ++s ) mbarInit( &ring.loFull[s], 1 );
      for ( int s = 0; s < 2; ++s ) mbarInit( &ring.exFull[s], 1 );
      mbarInitFence();
    }
    claim();
    __syncwarp();
    uint32_t tile0 = nextTile(), tile1 = nextTile();
    issueLower( 0, tile0 );
    issueLower( 1, tile1 );
    uint32_t old0 = issueGather( 0, tile0 );

    for ( uint32_t q = 0; tile0 != kNoTile; ++q )
    {
      __syncwarp();                                        // every lane has finished reading the ring slots of tile q-1
      const uint32_t tile2 = nextTile();
      issueLower( q + 2, tile2 );                          // P1(q+2)
      const uint32_t old1 = issueGather( q + 1, tile1 );   // P2(q+1)
      cpAsyncWait<1>();                                    // rows of tile q (committed one iteration ago) have landed ...
      mbarWait( &ring.exFull[q & 1u], ( q >> 1 ) & 1u );   // ... and so has its extent stream
      __syncwarp();                                        // ... for every lane of the warp that fetched them

      const bool live = ( tile0 << 5 ) + lane < a.n;
      Obb obb;
      obb.pt = obb.ax = obb.ay = obb.az = make_float4( 0.f, 0.f, 0.f, 0.f );
      if ( live )
      {
        const float4 lo = ring.lo[q % 3u][lane];
        const float4 ex = ring.ex[q & 1u][lane];
        float4 const *m = ring.m[q & 1u];
        obb = makeObb( lo.x, lo.y, lo.z, ex.x, ex.y, ex.z, m[swizzledRow( lane, 0 )], m[swizzledRow( lane, 1 )],
                       m[swizzledRow( lane, 2 )], m[swizzledRow( lane, 3 )] );
      }
      uint32_t myWord = 0;
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
          myWord = cullViewsFiltered<NV>( obb, a.filter, sP, sScratch[threadIdx.x >> 5], a.onePair, live, lane );
        }
        else
        {
          const ObbPairs ob = broadcastObb( obb );
          myWord = fast ? cullViews<NV, true>( ob, a.vp, a.onePair, live, lane ) : cullViews<NV, false>( ob, a.vp, a.onePair, live, lane );
        }
      }
      if ( lane < NV ) storeWord<NV>( a.out[lane], a, tile0, myWord, old0 );
      tile0 = tile1; tile1 = tile2;
      old0 = old1;
    }
    if ( a.buildChanged ) scanSegmentsInLastCta<NV>( a.out, a.nSegs, a.done );
  }
}
