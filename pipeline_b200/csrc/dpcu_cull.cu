// Culling layer of the C ABI: device mirror of a culling group and of its results, the
// frustum-cull kernels (K2), the ordered changed-list compaction and the group bounding box (K4).
//
// Replaces, on the GPU, the reference's hot loops B, C and D:
//   GroupCPU::updateOBBs          dp/culling/cpu/src/ManagerImpl.cpp:114-162
//   isVisible + visible.setBit    dp/culling/cpu/src/ManagerImpl.cpp:263-289,511-514
//   ResultBitSet::updateChanged   dp/culling/src/ResultBitSet.cpp:61-108
//
// HBM layout (all arrays 256-byte aligned cudaMalloc blocks):
//   lowerIdx[n]  float4  (box.lower.xyz, transformIndex as raw u32 bits)          16 B / object
//   extent[n]    float4  (box.upper - box.lower, 0)                               16 B / object
//   mats[m]      4 x float4 per matrix, row-major, 64-byte stride                 64 B / matrix
//   per result:  bits[ceil(n/32)] u32 (previous visibility, updated in place),
//                chg[ceil(n/32)]  u32 (bits that flipped in the last cull),
//                changed[n] u32 (ascending group indices), seg[] u32 changed-counts per 8192
//                objects (exclusive prefix after the cull; seg[nSegs] = total).
//
// One cull = memset(counters) -> cull kernel (one pass over the objects for up to 8 views:
// ballot -> word, XOR with the previous word, popc into the segment counters, last CTA scans the
// counters) -> compaction kernel (segment-ordered expansion of the flipped bits into the list).
//
// This translation unit is compiled with -fmad=false (exact mode).  dpcu_cull_fma.cu includes
// it again with DPCU_FMA_VARIANT defined and -fmad=true to provide the reporting-only fast mode.
#include "cull_math.cuh"
#include "cull_stage.cuh"
#include "cull_views.cuh"
#include "cull_filter.cuh"
#include "tree_propagate.cuh"
#include "dpcu_internal.h"
#include "dpcu_tree.h"

#include <cstddef>
#include <new>
#include <vector>

namespace dpcu
{
  constexpr int      kCullThreads    = 256;                 // objects per tile = threads per CTA
  constexpr uint32_t kSegObjectsLog2 = 13;                  // 8192 objects = 256 words per segment
  constexpr uint32_t kSegWords       = 1u << ( kSegObjectsLog2 - 5 );
    constexpr int      kMaxPeers       = 8;

  struct ViewOut
  {
    uint32_t *bits;      // in: previous visibility, out: new visibility
    uint32_t *chg;       // out: bits that flipped
    uint32_t *seg;       // += popc per 8192-object segment; read and zeroed again by the last CTA
    uint32_t *prefix;    // out: exclusive prefix of seg[] written by the last CTA, prefix[nSegs] = total
    uint32_t *mirror;    // optional: the result's bitset mirror in pinned host memory (line-granular kernel)
    // changed list built inside the line-granular kernel (decoupled look-back over 1024-object chunks)
    unsigned long long *look;   // per chunk: epoch << 34 | status << 32 | value
    uint32_t  epoch;            // this cull's tag: entries of earlier culls read as "not there yet", no reset needed
    uint32_t  hostCap;          // capacity of hostChanged
    uint32_t *changed;          // out: ascending group indices
    uint32_t *hostChanged;      // optional mirrors in pinned host memory
    uint32_t *hostCount;
    uint32_t *peer[kMaxPeers];   // optional: full bitsets on peer GPUs (NVLink stores)
  };

  template <int NV>
  struct CullArgs
  {
    float4 const *lowerIdx;
    float4 const *extent;
    float4 const *mats;
    uint32_t      n;
    uint32_t      nTiles;
    uint32_t      nPeers;
    uint32_t      peerWordOffset;
    int           buildChanged;
    uint32_t      nSegs;
    uint32_t     *done;      // CTA completion ticket (last CTA scans the segment counters)
    uint32_t     *chunkCounter;   // staged kernel: next unclaimed chunk of kChunkTiles tiles
    int           vpFinite;  // every view-projection entry is finite (enables the affine shortcut of cull_views.cuh)
    unsigned long long onePair;   // (1.0f, 1.0f): runtime multiplier of cull_views.cuh::addProd
    uint32_t      lineWords; // line-granular kernel: bitset words per warp (32 = one 128-byte line; 8 for mid-size groups)
    int           useFilter; // cull_filter.cuh: decide provable (object, view) pairs from centre and radius
    ViewOut       out[NV];
    float4        vp[NV][4];
    ViewFilter    filter[NV];
  };

  // the views' rows as packed pairs in shared memory, for lanes that evaluate different views (cull_filter.cuh)
  template <int NV>
  __device__ __forceinline__ void fillViewTable( f32x2 *sP, CullArgs<NV> const &a )
  {
    f32x2 const *src = reinterpret_cast<f32x2 const *>( &a.vp[0][0] );
    for ( uint32_t k = threadIdx.x; k < NV * 8u; k += blockDim.x ) sP[k] = src[k];
    __syncthreads();
  }

#ifndef DPCU_FMA_VARIANT
#define DPCU_KERNEL_NAME( name ) name
#else
#define DPCU_KERNEL_NAME( name ) name##_fma
#endif

  // The CTA that finishes last turns every view's per-segment changed counts into an exclusive
  // prefix (prefix[s] = number of changed objects before segment s, prefix[nSegs] = total) and
  // leaves the counters and the ticket zeroed for the next cull, so no memset runs between culls.
  // Replaces the XOR + traverseBits bookkeeping of ResultBitSet::updateChanged
  // (dp/culling/src/ResultBitSet.cpp:100-107) together with the compaction kernel below.
  template <int NV>
  __device__ __forceinline__ void scanSegmentsInLastCta( ViewOut const ( &out )[NV], uint32_t nSegs, uint32_t *done )
  {
    __shared__ uint32_t sLast;
    __shared__ uint32_t sPart[kCullThreads / 32];
    __threadfence();                       // this CTA's counter updates are visible before its ticket
    __syncthreads();
    if ( threadIdx.x == 0 ) sLast = ( atomicAdd( done, 1u ) == gridDim.x - 1 ) ? 1u : 0u;
    __syncthreads();
    if ( !sLast ) return;
    __threadfence();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t total = nSegs + 1;                              // the extra entry receives the grand total
    const uint32_t chunk = ( total + kCullThreads - 1 ) / kCullThreads;
    const uint32_t b = threadIdx.x * chunk, e = min( b + chunk, total );
#pragma unroll 1
    for ( int v = 0; v < NV; ++v )
    {
      uint32_t *seg = out[v].seg, *prefix = out[v].prefix;
      uint32_t sum = 0;
      for ( uint32_t k = b; k < e; ++k ) sum += __ldcg( seg + k );
      uint32_t incl = sum;
#pragma unroll
      for ( int d = 1; d < 32; d <<= 1 )
      {
        uint32_t t = __shfl_up_sync( 0xffffffffu, incl, d );
        if ( lane >= d ) incl += t;
      }
      if ( lane == 31 ) sPart[warp] = incl;
      __syncthreads();
      uint32_t run = incl - sum;
      for ( uint32_t w = 0; w < warp; ++w ) run += sPart[w];
      for ( uint32_t k = b; k < e; ++k )
      {
        const uint32_t c = __ldcg( seg + k );
        prefix[k] = run;
        seg[k] = 0u;
        run += c;
      }
      __syncthreads();
    }
    if ( threadIdx.x == 0 ) done[0] = done[1] = 0u;    // ticket and the staged kernel's chunk counter: ready for the next cull
  }

  // ------------------------------------------------------------------------------------------
  // K2, direct variant: one thread per object, six 16-byte loads, persistent grid-stride tiles.
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
        const float4 m0 = __ldg( m + 0 );
        const float4 m1 = __ldg( m + 1 );
        const float4 m2 = __ldg( m + 2 );
        const float4 m3 = __ldg( m + 3 );
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
            if ( c )
            {
              // one counter per 8192 objects: concurrently running CTAs spread over ~40 addresses.
              // (A second, coarser level of counters was measured to serialise in L2: +0.5 ms at 64 Mi objects.)
              atomicAdd( o.seg + ( word >> ( kSegObjectsLog2 - 5 ) ), __popc( c ) );
            }
          }
        }
      }
    }
    if ( a.buildChanged ) scanSegmentsInLastCta<NV>( a.out, a.nSegs, a.done );
  }


#ifndef DPCU_FMA_VARIANT
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
    const uint32_t strideObjects = gridDim.x * kCullThreads;
    uint32_t idxNext = 0;
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
        const float4 m0 = __ldg( m + 0 );
        const float4 m1 = __ldg( m + 1 );
        const float4 m2 = __ldg( m + 2 );
        const float4 m3 = __ldg( m + 3 );
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
          if ( c ) atomicAdd( o.seg + ( word >> ( kSegObjectsLog2 - 5 ) ), __popc( c ) );
        }
      }
#if DPCU_VIEWS_PREFETCH
      idxNext = idxNext2;
#endif
    }
    if ( a.buildChanged ) scanSegmentsInLastCta<NV>( a.out, a.nSegs, a.done );
  }

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
      if ( c ) atomicAdd( o.seg + ( word >> ( kSegObjectsLog2 - 5 ) ), __popc( c ) );
    }
  }

  template <int NV>
  __global__ void __launch_bounds__( kCullThreads, 4 )
  cullStagedKernel( const __grid_constant__ CullArgs<NV> a )
  {
    extern __shared__ __align__( 128 ) unsigned char smemRaw[];
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
        const ObbPairs ob = broadcastObb( obb );
        myWord = fast ? cullViews<NV, true>( ob, a.vp, a.onePair, live, lane ) : cullViews<NV, false>( ob, a.vp, a.onePair, live, lane );
      }
      if ( lane < NV ) storeWord<NV>( a.out[lane], a, tile0, myWord, old0 );
      tile0 = tile1; tile1 = tile2;
      old0 = old1;
    }
    if ( a.buildChanged ) scanSegmentsInLastCta<NV>( a.out, a.nSegs, a.done );
  }

  // ------------------------------------------------------------------------------------------
  // K2, line-granular variant - the multi-GPU form.  A warp owns 1024 consecutive objects (32
  // words = one 128-byte line of each bitset) and walks them in 32 steps of 32 objects; lane w
  // keeps the ballot of step w, so at the end lane l holds word l of the line.  Previous bits are
  // read and new bits / flipped bits are written as whole lines, the changed-count goes to the
  // segment counter once per line, and - the point of this form - the bitset all-gather of
  // SURVEY.md 8e is the same coalesced 128-byte store repeated into every peer's full bitset
  // over NVLink: whole lines on the wire, no barrier, no shared memory, no separate collective.
  // (Per-word 4-byte peer stores from the direct kernel were measured at 1.87 ms per 64 Mi-object
  // step on 8 GPUs; a shared-memory hand-over with two CTA barriers at 1.21 ms; the cull alone 0.98 ms.)
  //
  // kFuseList: the ordered changed list is built by this kernel as well.  Lines are claimed in
  // ascending order from a global counter, so a line's predecessors are always in flight or done
  // and a decoupled look-back (Merrill & Garland's single-pass scan) can hand every line the number
  // of changes before it: a warp publishes its line's count as an AGGREGATE, walks back over its
  // predecessors' entries 32 at a time until it meets an inclusive PREFIX, publishes its own prefix,
  // and expands its flipped bits straight into the list - no counters, no scan, no second kernel, and
  // with a host mirror the list crosses PCIe while the cull is still running instead of after it.
  constexpr uint32_t kLookAggregate = 1u, kLookPrefix = 2u;

  __device__ __forceinline__ unsigned long long lookPack( uint32_t epoch, uint32_t status, uint32_t value )
  {
    return ( static_cast<unsigned long long>( ( epoch << 2 ) | status ) << 32 ) | value;
  }
  __device__ __forceinline__ unsigned long long lookLoad( unsigned long long const *p )
  {
    unsigned long long v;
    asm volatile( "ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"( v ) : "l"( p ) : "memory" );
    return v;
  }
  __device__ __forceinline__ void lookStore( unsigned long long *p, unsigned long long v )
  {
    asm volatile( "st.relaxed.gpu.global.u64 [%0], %1;" :: "l"( p ), "l"( v ) : "memory" );
  }

  // exclusive prefix of line `line` (> 0): sum of the counts of lines 0 .. line-1
  __device__ __forceinline__ uint32_t lookBack( unsigned long long const *look, uint32_t epoch, uint32_t line, uint32_t lane )
  {
    uint32_t excl = 0;
    int32_t  base = int32_t( line ) - 1;                   // lane k inspects line base - k: lane 0 is the nearest predecessor
    uint32_t polls = 0;
    for ( ;; )
    {
      const int32_t idx = base - int32_t( lane );
      unsigned long long s = idx >= 0 ? lookLoad( look + idx ) : lookPack( epoch, kLookPrefix, 0u );
      const uint32_t tag    = uint32_t( s >> 32 );
      const bool     valid  = ( tag >> 2 ) == epoch;
      const bool     prefix = valid && ( tag & 3u ) == kLookPrefix;
      const uint32_t pmask  = __ballot_sync( 0xffffffffu, prefix );
      const uint32_t vmask  = __ballot_sync( 0xffffffffu, valid );
      const uint32_t first  = pmask ? uint32_t( __ffs( pmask ) - 1 ) : 31u;       // nearest line that already knows its prefix
      const uint32_t need   = first == 31u ? 0xffffffffu : ( ( 2u << first ) - 1u );
      if ( ( vmask & need ) != need )
      {
        // a predecessor in the window has not published yet: it is running on some other warp; poll again
        if ( ++polls > ( 1u << 24 ) ) __trap();            // fail loudly instead of hanging the device
        continue;
      }
      uint32_t v = ( lane <= first ) ? uint32_t( s ) : 0u;
#pragma unroll
      for ( int d = 16; d > 0; d >>= 1 ) v += __shfl_xor_sync( 0xffffffffu, v, d );
      excl += v;
      if ( pmask ) return excl;
      base -= 32;
    }
  }

  constexpr uint32_t kNoLine = 0xffffffffu;

  // Place line `line`'s changes in the list of every view: its flipped-bit words (written by this warp a
  // line ago, read back from L2), the number of changes before it by look-back, then the expansion.
  template <int NV>
  __device__ __forceinline__ void resolveLine( CullArgs<NV> const &a, uint32_t line, uint32_t nLines, uint32_t nWords, uint32_t lane )
  {
    const uint32_t myWord = line * a.lineWords + lane;
    const bool     mine   = lane < a.lineWords && myWord < nWords;
    __syncwarp();                                            // the warp's chg stores of that line are visible to all its lanes
#pragma unroll 1
    for ( int v = 0; v < NV; ++v )
    {
      ViewOut const &o = a.out[v];
      uint32_t c = mine ? __ldcg( o.chg + myWord ) : 0u;
      const uint32_t flips = __popc( c );
      uint32_t incl = flips;                                 // inclusive scan of the per-word counts across the line
#pragma unroll
      for ( int d = 1; d < 32; d <<= 1 )
      {
        const uint32_t t = __shfl_up_sync( 0xffffffffu, incl, d );
        if ( lane >= d ) incl += t;
      }
      const uint32_t total = __shfl_sync( 0xffffffffu, incl, 31 );
      uint32_t excl = 0;
      if ( line > 0 )
      {
        excl = lookBack( o.look, o.epoch, line, lane );
        if ( lane == 0 ) lookStore( o.look + line, lookPack( o.epoch, kLookPrefix, excl + total ) );
      }
      // expand: word l's flipped bits go to list[excl + (changes in words 0..l-1) ...], ascending
      uint32_t off = excl + incl - flips;
      const uint32_t base = myWord << 5;
      while ( c )
      {
        o.changed[off++] = base + uint32_t( __ffs( c ) - 1 );
        c &= c - 1;
      }
      if ( o.hostChanged )
      {
        // the line's run again as coalesced stores into the pinned host mirror (the entries just written
        // are in L2; __syncwarp orders the warp's writes before its reads)
        __syncwarp();
        for ( uint32_t k = lane; k < total; k += 32 )
        {
          if ( excl + k < o.hostCap ) o.hostChanged[excl + k] = __ldcg( o.changed + excl + k );
        }
      }
      if ( line == nLines - 1 && lane == 0 )
      {
        o.prefix[a.nSegs] = excl + total;                    // where the compaction path keeps the length of the list
        if ( o.hostCount ) *o.hostCount = excl + total;
      }
    }
  }

  __device__ __forceinline__ void rearmInLastCta( uint32_t *done )
  {
    __shared__ uint32_t sLastLines;
    __syncthreads();
    if ( threadIdx.x == 0 ) sLastLines = ( atomicAdd( done, 1u ) == gridDim.x - 1 ) ? 1u : 0u;
    __syncthreads();
    if ( sLastLines && threadIdx.x == 0 ) done[0] = done[1] = 0u;      // ticket and line counter: ready for the next cull
  }

  // measured at 64 Mi objects before the filter: 2 views 1.55 / 1.23 / 1.16 ms and 6 views 2.82 / 2.61 / 2.64 ms at
  // 2 / 3 / 4 CTAs per SM;
  // with the filter (fewer issue slots, more waiting on memory) 4 CTAs per SM win for every multi-view count:
  // 6 views 2.068 ms at 3 CTAs (80 registers) -> 1.937 ms at 4 (64 registers, ~70 bytes of spills)
  template <int NV, bool kFuseList>
  __global__ void __launch_bounds__( kCullThreads, NV == 1 ? 6 : 4 )
  cullLinesKernel( const __grid_constant__ CullArgs<NV> a )
  {
    const uint32_t lane   = threadIdx.x & 31u;
    const uint32_t W = a.lineWords;
    const uint32_t nWords = ( a.n + 31u ) >> 5, nLines = ( nWords + W - 1u ) / W;
    const uint32_t nWarps = gridDim.x * ( kCullThreads / 32 );
    uint32_t line = blockIdx.x * ( kCullThreads / 32 ) + ( threadIdx.x >> 5 );
    uint32_t pending = kNoLine;
    __shared__ f32x2 sP[NV * 8];
    __shared__ FilterScratch<NV> sScratch[kCullThreads / 32];
    if ( NV > 1 ) fillViewTable<NV>( sP, a );
    for ( ;; )
    {
      if ( kFuseList )
      {
        // ascending claims: every predecessor of a claimed line belongs to a warp that is already running
        uint32_t claimed = 0;
        if ( lane == 0 ) claimed = atomicAdd( a.chunkCounter, 1u );
        line = __shfl_sync( 0xffffffffu, claimed, 0 );
      }
      if ( line >= nLines ) break;
      const uint32_t word0 = line * W, myWord = word0 + lane;
      const bool     wordLive = lane < W && myWord < nWords;
      uint32_t old[NV], acc[NV];
#pragma unroll
      for ( int v = 0; v < NV; ++v )
      {
        old[v] = wordLive ? a.out[v].bits[myWord] : 0u;
        acc[v] = 0u;
      }
      const uint32_t steps = min( W, nWords - word0 );
      // With several views this kernel runs at 24-32 warps per SM and (since the filter) waits on memory more than
      // on the issue slots: the transform index two steps ahead is fetched now, the one fetched a step ago turns
      // into an L2 prefetch of the next step's matrices and extents (same scheme as cullViewsKernel).  Measured at
      // 64 Mi objects with the current filter: 2 views 1.174 -> 1.154 ms; 3 views 1.266 -> 1.276 ms, 4 views 1.441
      // -> 1.499 ms and 6 views unchanged (the two extra registers spill there), so only NV == 2 keeps it.
      constexpr bool kPrefetch = NV == 2;
      uint32_t idxNext = 0;
      if ( kPrefetch )
      {
        const uint32_t i1 = ( ( word0 + 1u ) << 5 ) + lane;
        if ( i1 < a.n ) idxNext = __ldg( reinterpret_cast<uint32_t const *>( a.lowerIdx + i1 ) + 3 );
      }
#pragma unroll 1      // measured: one step in flight at 48 warps per SM beats unroll 2 / 4 at lower occupancy
      for ( uint32_t w = 0; w < steps; ++w )
      {
        const uint32_t i    = ( ( word0 + w ) << 5 ) + lane;
        const bool     live = i < a.n;
        uint32_t idxNext2 = 0;
        if ( kPrefetch )
        {
          const uint32_t i1 = i + 32u, i2 = i + 64u;
          if ( i2 < a.n && i2 > i ) idxNext2 = __ldg( reinterpret_cast<uint32_t const *>( a.lowerIdx + i2 ) + 3 );
          if ( i1 < a.n && i1 > i )
          {
            prefetchL2( a.mats + 4ull * idxNext );
            if ( ( lane & 7u ) == 0 ) prefetchL2( a.extent + i1 );
          }
        }
        Obb obb;
        obb.pt = obb.ax = obb.ay = obb.az = make_float4( 0.f, 0.f, 0.f, 0.f );
        if ( live )
        {
          const float4 lo = ldStream( a.lowerIdx + i );
          const float4 ex = ldStream( a.extent + i );
          float4 const *m = a.mats + 4ull * __float_as_uint( lo.w );
          const float4 m0 = __ldg( m + 0 );
          const float4 m1 = __ldg( m + 1 );
          const float4 m2 = __ldg( m + 2 );
          const float4 m3 = __ldg( m + 3 );
          obb = makeObb( lo.x, lo.y, lo.z, ex.x, ex.y, ex.z, m0, m1, m2, m3 );
        }
        if ( NV == 1 )
        {
          const uint32_t b = __ballot_sync( 0xffffffffu, obbVisible( obb, a.vp[0][0], a.vp[0][1], a.vp[0][2], a.vp[0][3] ) & live );
          if ( lane == w ) acc[0] = b;
        }
        else
        {
          const bool affine = !live || ( obb.pt.w == 1.0f && obb.ax.w == 0.0f && obb.ay.w == 0.0f && obb.az.w == 0.0f );
          const bool fast   = __all_sync( 0xffffffffu, affine ) && a.vpFinite;
          uint32_t perView;
          if ( fast && a.useFilter )
          {
            perView = cullViewsFiltered<NV>( obb, a.filter, sP, sScratch[threadIdx.x >> 5], a.onePair, live, lane );
          }
          else
          {
            const ObbPairs ob = broadcastObb( obb );
            perView = fast ? cullViews<NV, true>( ob, a.vp, a.onePair, live, lane ) : cullViews<NV, false>( ob, a.vp, a.onePair, live, lane );
          }
#pragma unroll
          for ( int v = 0; v < NV; ++v )
          {
            const uint32_t b = __shfl_sync( 0xffffffffu, perView, v );     // lane v held view v's word
            if ( lane == w ) acc[v] = b;
          }
        }
        if ( kPrefetch ) idxNext = idxNext2;
      }
#pragma unroll
      for ( int v = 0; v < NV; ++v )
      {
        ViewOut const &o = a.out[v];
        uint32_t flips = 0;
        if ( wordLive )
        {
          o.bits[myWord] = acc[v];
          if ( o.mirror ) o.mirror[myWord] = acc[v];          // the same line over PCIe into pinned host memory
          for ( uint32_t p = 0; p < a.nPeers; ++p )
          {
            if ( o.peer[p] ) o.peer[p][a.peerWordOffset + myWord] = acc[v];
          }
          if ( a.buildChanged )
          {
            const uint32_t c = old[v] ^ acc[v];
            o.chg[myWord] = c;
            flips = __popc( c );
          }
        }
        if ( a.buildChanged && !kFuseList )
        {
#pragma unroll
          for ( int d = 16; d > 0; d >>= 1 ) flips += __shfl_xor_sync( 0xffffffffu, flips, d );
          if ( lane == 0 && flips ) atomicAdd( o.seg + ( word0 >> ( kSegObjectsLog2 - 5 ) ), flips );
        }
        if ( a.buildChanged && kFuseList )
        {
          // publish this line's count right away; its place in the list is resolved one line later (below)
#pragma unroll
          for ( int d = 16; d > 0; d >>= 1 ) flips += __shfl_xor_sync( 0xffffffffu, flips, d );
          if ( lane == 0 ) lookStore( o.look + line, lookPack( o.epoch, line == 0 ? kLookPrefix : kLookAggregate, flips ) );
        }
      }
      if ( kFuseList && a.buildChanged )
      {
        // The look-back of the PREVIOUS line runs now, a whole line of work after its count was published:
        // by then its predecessors have published theirs and the walk does not wait (resolving a line
        // immediately made every warp wait for its slowest recent predecessor: 1.11 ms instead of 1.01 ms).
        if ( pending != kNoLine ) resolveLine<NV>( a, pending, nLines, nWords, lane );
        pending = line;
      }
      if ( !kFuseList ) line += nWarps;
    }
    if ( kFuseList && a.buildChanged && pending != kNoLine ) resolveLine<NV>( a, pending, nLines, nWords, lane );
    if ( kFuseList ) rearmInLastCta( a.done );
    else if ( a.buildChanged ) scanSegmentsInLastCta<NV>( a.out, a.nSegs, a.done );
  }

  // ------------------------------------------------------------------------------------------
  // K1 + K2 fused: the last level of the transform tree is propagated inside the cull kernel
  // (the leaf world matrices are consumed from on-chip memory while still
  // being written out for the renderer, which removes their 64 B / object re-read).
  // Precondition, checked on the device by leafBindingKernel: object i is bound to the node of
  // the level's entry i (tidx[i] == entries[i].transform), i.e. one drawable per leaf transform
  // in tree order - the C3 layout.
  // Each thread computes world = local * world[parent] for its object's node exactly like
  // treeLevelKernel (same association order, same dirty protocol, Tree.cpp:153-160), stores the
  // four rows and culls straight out of its registers.
  struct LeafArgs
  {
    uint2 const    *entries;      // {parent, transform} of the fused level, entry i <-> object i
    float4 const   *local;
    float4         *world;
    uint32_t const *dirtyLocal;
    uint32_t       *dirtyWorld;
  };

  __device__ __forceinline__ bool leafTestBit( uint32_t const *w, uint32_t i )
  {
    return ( w[i >> 5] >> ( i & 31u ) ) & 1u;
  }

  // One thread per object / leaf node, persistent grid-stride tiles.  The {parent, node} entries
  // run two tiles ahead and the dirty test one tile ahead of the matrix loads, so every load a
  // tile issues (8 matrix rows, 2 AABB vectors, the next dirty words, the entry after next) is
  // independent of the others: one memory latency per tile instead of a chain of three.
  template <int NV>
  __global__ void __launch_bounds__( kCullThreads )
  cullFusedLeafKernel( const __grid_constant__ CullArgs<NV> a, const __grid_constant__ LeafArgs t )
  {
    __shared__ float4 sTranspose[kCullThreads / 32][2][128];     // per warp: locals in, worlds out (2 KiB each)
    const uint32_t lane   = threadIdx.x & 31u;
    const uint32_t stride = gridDim.x * kCullThreads;
    float4 *bufIn = sTranspose[threadIdx.x >> 5][0], *bufOut = sTranspose[threadIdx.x >> 5][1];
    uint32_t i = blockIdx.x * kCullThreads + threadIdx.x;
    // prologue of the software pipeline: entry of this tile and of the next, dirty flag of this tile
    uint2 ent  = make_uint2( 0u, 0u ), entN = make_uint2( 0u, 0u );
    if ( i < a.n ) ent = __ldg( t.entries + i );
    if ( i + stride < a.n && i + stride >= i ) entN = __ldg( t.entries + i + stride );
    bool dirty = i < a.n && ( leafTestBit( t.dirtyWorld, ent.x ) || leafTestBit( t.dirtyLocal, ent.y ) );

    for ( uint32_t tile = blockIdx.x; tile < a.nTiles; tile += gridDim.x, i += stride )
    {
      const bool     live = i < a.n;
      const uint32_t word = i >> 5;
      uint32_t oldBits = 0;
      if ( lane < NV && ( i - lane ) < a.n ) oldBits = a.out[lane].bits[word];

      // everything this tile needs from memory, issued together
      const uint32_t iN = i + stride, iNN = iN + stride;
      const bool liveN  = iN < a.n && iN >= i;
      uint2 entNN = make_uint2( 0u, 0u );
      if ( iNN < a.n && iNN >= iN && liveN ) entNN = __ldg( t.entries + iNN );
      // parent bits were set by the earlier level launches (never by this kernel); node bits are
      // set below, but only for nodes of other threads
      const bool dirtyN = liveN && ( leafTestBit( t.dirtyWorld, entN.x ) || leafTestBit( t.dirtyLocal, entN.y ) );
      float4 lo = make_float4( 0.f, 0.f, 0.f, 0.f ), ex = lo, w0 = lo, w1 = lo, w2 = lo, w3 = lo;
      if ( live )
      {
        lo = ldStream( a.lowerIdx + i );
        ex = ldStream( a.extent + i );
      }
      // (The two propagation paths below are the bodies of tree_propagate.cuh's propagateWarpCoalesced /
      // propagateNode written out in place: calling the shared helpers here measured 5 % slower - 0.529 ms
      // instead of 0.503 ms for the C3 leaf level - although the SASS differs only in scheduling.)
      // Warp-uniform fast path: a full warp of dirty nodes with consecutive indices (the usual
      // layout of a level).  The 32 local matrices are 2 KiB contiguous: four coalesced 16-byte
      // loads per lane bring them in, shared memory (XOR-swizzled, conflict-free both ways) turns
      // "row j*32+lane" into "my node's four rows", and the same trip backwards turns the world
      // matrices into four coalesced stores.  Any other warp uses strided per-thread accesses.
      const uint32_t node0 = __shfl_sync( 0xffffffffu, ent.y, 0 );
      if ( __all_sync( 0xffffffffu, live && dirty && ent.y == node0 + lane ) )
      {
        float4 const *ln = t.local + 4ull * node0;
        float4       *wn = t.world + 4ull * node0;
        float4 const *pw = t.world + 4ull * ent.x;
        const float4 r0 = ldStream( ln + lane ), r1 = ldStream( ln + 32 + lane ), r2 = ldStream( ln + 64 + lane ), r3 = ldStream( ln + 96 + lane );
        const float4 p0 = __ldg( pw + 0 ), p1 = __ldg( pw + 1 ), p2 = __ldg( pw + 2 ), p3 = __ldg( pw + 3 );
        const uint32_t oj = lane >> 2, rj = lane & 3u;          // row j*32+lane belongs to node j*8+oj, row rj
        bufIn[swizzledRow( oj, rj )]      = r0;
        bufIn[swizzledRow( 8 + oj, rj )]  = r1;
        bufIn[swizzledRow( 16 + oj, rj )] = r2;
        bufIn[swizzledRow( 24 + oj, rj )] = r3;
        __syncwarp();
        const float4 l0 = bufIn[swizzledRow( lane, 0 )], l1 = bufIn[swizzledRow( lane, 1 )];
        const float4 l2 = bufIn[swizzledRow( lane, 2 )], l3 = bufIn[swizzledRow( lane, 3 )];
        w0 = vecMulMat( l0, p0, p1, p2, p3 );                       // Tree.cpp:157, Matmnt.h:1381-1415
        w1 = vecMulMat( l1, p0, p1, p2, p3 );
        w2 = vecMulMat( l2, p0, p1, p2, p3 );
        w3 = vecMulMat( l3, p0, p1, p2, p3 );
        bufOut[swizzledRow( lane, 0 )] = w0;
        bufOut[swizzledRow( lane, 1 )] = w1;
        bufOut[swizzledRow( lane, 2 )] = w2;
        bufOut[swizzledRow( lane, 3 )] = w3;
        __syncwarp();
        wn[lane]      = bufOut[swizzledRow( oj, rj )];
        wn[32 + lane] = bufOut[swizzledRow( 8 + oj, rj )];
        wn[64 + lane] = bufOut[swizzledRow( 16 + oj, rj )];
        wn[96 + lane] = bufOut[swizzledRow( 24 + oj, rj )];
        if ( lane == 0 ) atomicOr( t.dirtyWorld + ( node0 >> 5 ), 0xffffffffu << ( node0 & 31u ) );          // Tree.cpp:158, 32 nodes
        if ( lane == 0 && ( node0 & 31u ) ) atomicOr( t.dirtyWorld + ( node0 >> 5 ) + 1, ~( 0xffffffffu << ( node0 & 31u ) ) );
      }
      else if ( live )
      {
        float4 *wn = t.world + 4ull * ent.y;
        if ( dirty )
        {
          float4 const *ln = t.local + 4ull * ent.y;
          float4 const *pw = t.world + 4ull * ent.x;
          const float4 l0 = __ldg( ln + 0 ), l1 = __ldg( ln + 1 ), l2 = __ldg( ln + 2 ), l3 = __ldg( ln + 3 );
          const float4 p0 = __ldg( pw + 0 ), p1 = __ldg( pw + 1 ), p2 = __ldg( pw + 2 ), p3 = __ldg( pw + 3 );
          w0 = vecMulMat( l0, p0, p1, p2, p3 );                     // Tree.cpp:157, Matmnt.h:1381-1415
          w1 = vecMulMat( l1, p0, p1, p2, p3 );
          w2 = vecMulMat( l2, p0, p1, p2, p3 );
          w3 = vecMulMat( l3, p0, p1, p2, p3 );
          wn[0] = w0; wn[1] = w1; wn[2] = w2; wn[3] = w3;
          atomicOr( t.dirtyWorld + ( ent.y >> 5 ), 1u << ( ent.y & 31u ) );   // Tree.cpp:158
        }
        else
        {
          w0 = wn[0]; w1 = wn[1]; w2 = wn[2]; w3 = wn[3];
        }
      }
      const Obb obb = makeObb( lo.x, lo.y, lo.z, ex.x, ex.y, ex.z, w0, w1, w2, w3 );
      uint32_t myWord = 0;
      if ( NV == 1 )
      {
        myWord = __ballot_sync( 0xffffffffu, obbVisible( obb, a.vp[0][0], a.vp[0][1], a.vp[0][2], a.vp[0][3] ) & live );
      }
      else
      {
        const bool affine = !live || ( obb.pt.w == 1.0f && obb.ax.w == 0.0f && obb.ay.w == 0.0f && obb.az.w == 0.0f );
        const bool fast   = __all_sync( 0xffffffffu, affine ) && a.vpFinite;
        const ObbPairs ob = broadcastObb( obb );
        myWord = fast ? cullViews<NV, true>( ob, a.vp, a.onePair, live, lane ) : cullViews<NV, false>( ob, a.vp, a.onePair, live, lane );
      }
      if ( lane < NV && ( i - lane ) < a.n ) storeWord<NV>( a.out[lane], a, word, myWord, oldBits );
      ent = entN; entN = entNN; dirty = dirtyN;
    }
    if ( a.buildChanged ) scanSegmentsInLastCta<NV>( a.out, a.nSegs, a.done );
  }

  // counts objects whose transform index is not the node of the level entry with the same index
  __global__ void leafBindingKernel( float4 const *lowerIdx, uint2 const *entries, uint32_t n, uint32_t *mismatches )
  {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool bad = i < n && __float_as_uint( lowerIdx[i].w ) != entries[i].y;
    const uint32_t m = __ballot_sync( 0xffffffffu, bad );
    if ( ( threadIdx.x & 31u ) == 0 && m ) atomicAdd( mismatches, __popc( m ) );
  }
#endif

#ifndef DPCU_FMA_VARIANT
  // ------------------------------------------------------------------------------------------
  // Ordered changed list.  One CTA per 8192-object segment and view: seg[] already holds the
  // exclusive prefix, so the CTA only block-scans the popcounts of its 256 flipped-bit words and
  // expands them; ascending group index order falls out of the layout (BitArray::traverseBits
  // order, dp/util/BitArray.h:127-136).
  struct CompactArgs
  {
    uint32_t const *chg[DPCU_MAX_VIEWS];
    uint32_t const *prefix[DPCU_MAX_VIEWS];
    uint32_t       *changed[DPCU_MAX_VIEWS];
    uint32_t       *hostChanged[DPCU_MAX_VIEWS];    // optional mirror of the list in pinned host memory ...
    uint32_t       *hostCount[DPCU_MAX_VIEWS];      // ... and of its length
    uint32_t        hostCapacity[DPCU_MAX_VIEWS];
    uint32_t        nWords;
    uint32_t        nSegs;
  };

  __global__ void __launch_bounds__( 256 ) compactChangedKernel( const __grid_constant__ CompactArgs a )
  {
    const uint32_t v    = blockIdx.y;
    const uint32_t s    = blockIdx.x;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t base0 = a.prefix[v][s];
    const uint32_t count = a.prefix[v][s + 1] - base0;
    if ( s == 0 && threadIdx.x == 0 && a.hostCount[v] ) *a.hostCount[v] = a.prefix[v][a.nSegs];
    if ( count == 0 ) return;                        // nothing changed in this segment

    // the segment's indices are expanded into shared memory first, so that the list (and its host
    // mirror, over PCIe) is written as one contiguous, coalesced run per segment
    __shared__ uint32_t sIdx[1u << kSegObjectsLog2];
    __shared__ uint32_t sWarp[8];
    const uint32_t w = s * kSegWords + threadIdx.x;
    uint32_t c = ( w < a.nWords ) ? a.chg[v][w] : 0u;
    const uint32_t pc = __popc( c );
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
      const Obb o = makeObb( lo.x, lo.y, lo.z, ex.x, ex.y, ex.z, __ldg( m ), __ldg( m + 1 ), __ldg( m + 2 ), __ldg( m + 3 ) );
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
#endif   // !DPCU_FMA_VARIANT
}   // namespace dpcu

#ifdef DPCU_FMA_VARIANT
// launcher used by the exact translation unit for DPCU_CULL_OPT_FMA = 1
namespace dpcu
{
  template <int NV>
  cudaError_t launchCullDirectFma( CullArgs<NV> const &args, int grid, cudaStream_t stream )
  {
    cullDirectKernel_fma<NV><<<grid, kCullThreads, 0, stream>>>( args );
    return cudaGetLastError();
  }
  template <int NV> int occupancyCullDirectFma()
  {
    int b = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor( &b, cullDirectKernel_fma<NV>, kCullThreads, 0 );
    return b;
  }
#define DPCU_INSTANTIATE( NV )                                                                        \
  template cudaError_t launchCullDirectFma<NV>( CullArgs<NV> const &, int, cudaStream_t );            \
  template int occupancyCullDirectFma<NV>();
  DPCU_INSTANTIATE( 1 ) DPCU_INSTANTIATE( 2 ) DPCU_INSTANTIATE( 3 ) DPCU_INSTANTIATE( 4 )
  DPCU_INSTANTIATE( 5 ) DPCU_INSTANTIATE( 6 ) DPCU_INSTANTIATE( 7 ) DPCU_INSTANTIATE( 8 )
}
#else

namespace dpcu
{
  template <int NV> cudaError_t launchCullDirectFma( CullArgs<NV> const &args, int grid, cudaStream_t stream );
  template <int NV> int occupancyCullDirectFma();
}

// =============================================================================================
// host side
struct dpcuCullResult
{
  dpcuCull *ctx = nullptr;
  dpcu::DeviceArray bits, chg, changed, counters;   // counters: done, chunk, -, - | seg[cap] | prefix[cap] (count = prefix[nSegs])
  size_t   n = 0;                // object count the stored bits are valid for (ResultBitSet::m_results size)
  size_t   capWords = 0;
  size_t   nSegsCap = 0;
  bool     ran = false;          // a changed list exists
  dpcu::StreamFence done;        // last cull / bit move submitted for this result
  uint32_t *peer[dpcu::kMaxPeers] = { nullptr };
  int      nPeers = 0;
  size_t   peerWordOffset = 0;
  dpcuCullResult *next = nullptr, *prev = nullptr;

  dpcu::DeviceArray look;                           // look-back entries of the fused list, one per 1024 objects
  size_t   lookCap = 0;
  uint32_t lookEpoch = 0;
  dpcu::DeviceArray visible, visCounters;           // dpcuCullResultBuildVisibleList: indices | done[4], seg[cap], prefix[cap]
  size_t   visSegsCap = 0, visSegs = 0;
  bool     visBuilt = false;
  // optional mirror in pinned host memory (dpcuCullResultSetHostMirror): h* = host addresses, d* = device aliases
  uint32_t *hBits = nullptr, *dBits = nullptr, *hChanged = nullptr, *dChanged = nullptr, *hCount = nullptr, *dCount = nullptr;
  size_t   hBitsWords = 0, hChangedCap = 0;

  size_t   nSegs = 0;            // segments of the last run; the changed count lives at prefix[nSegs]
  uint32_t *donePtr() const { return static_cast<uint32_t *>( counters.ptr ); }
  uint32_t *segPtr() const { return static_cast<uint32_t *>( counters.ptr ) + 4; }
  uint32_t *prefixPtr() const { return segPtr() + nSegsCap; }
  uint32_t *countPtr() const { return prefixPtr() + nSegs; }
};

struct dpcuCull
{
  int          device = 0;
  int          smCount = 0;
  cudaStream_t stream = nullptr;
  dpcu::DeviceArray lowerIdx, extent, mats, scratch, maxIndex;
  dpcu::PinnedArray staging;
  dpcu::StreamFence uploads;     // object / matrix updates submitted on `stream`
  dpcu::StreamFence lastRun;     // last cull submitted (possibly on a caller stream); updates are ordered after it
  float const *boundMats = nullptr;      // borrowed device matrices (dpcuCullBindMatrices)
  size_t       n = 0, nMats = 0;
  uint32_t     maxTransformIndex = 0;
  bool         maxIndexKnown = true;
  int          optKernel = 0, optFma = 0, optChanged = 1, optCtasPerSm = 0, optProfile = 0, optFuseLeaf = 1, optFuseList = 1, optFilter = 1, optLineWords = 0;
  int          lastKernel = 0;           // DPCU_KERNEL_* of the last cull launched (DPCU_CULL_OPT_LAST_KERNEL)
  uint64_t     objectsVersion = 0;       // bumped whenever objects are (re)uploaded
  dpcuTree    *leafTree = nullptr;       // cached answer of the leaf-binding check of dpcuCullRunWithTree
  uint64_t     leafTopologyVersion = 0, leafObjectsVersion = 0;
  bool         leafBound = false;
  uint64_t     launches = 0;
  std::vector<cudaEvent_t> profEvents;   // start/stop pairs of profiled cull-kernel launches
  size_t       profUsed = 0;
  dpcuCullResult *results = nullptr;

  float4 const *matsPtr() const { return boundMats ? reinterpret_cast<float4 const *>( boundMats ) : static_cast<float4 const *>( mats.ptr ); }
};

namespace dpcu
{
  static int uploadOrCopy( void *dst, void const *src, size_t bytes, int memspace, dpcuCull *ctx )
  {
    if ( !bytes ) return DPCU_OK;
    if ( memspace == DPCU_MEM_DEVICE )
    {
      DPCU_CUDA( cudaMemcpyAsync( dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream ) );
    }
    else
    {
      // pageable or pinned caller memory: the copy is complete (w.r.t. the host buffer) on return
      DPCU_CUDA( cudaMemcpyAsync( dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream ) );
      DPCU_CUDA( cudaStreamSynchronize( ctx->stream ) );
    }
    return DPCU_OK;
  }

  static int refreshMaxIndex( dpcuCull *ctx )
  {
    if ( !ctx->maxIndexKnown )
    {
      DPCU_CUDA( cudaMemcpyAsync( &ctx->maxTransformIndex, ctx->maxIndex.ptr, sizeof( uint32_t ), cudaMemcpyDeviceToHost, ctx->stream ) );
      DPCU_CUDA( cudaStreamSynchronize( ctx->stream ) );
      ctx->maxIndexKnown = true;
    }
    return DPCU_OK;
  }

  static int ensureResultCapacity( dpcuCullResult *r, size_t n, cudaStream_t stream )
  {
    dpcuCull *ctx = r->ctx;
    // words padded to whole segments so the compaction kernel can read full 256-word rows
    size_t nSegs = divUp( n, size_t( 1 ) << kSegObjectsLog2 );
    size_t words = ( nSegs ? nSegs : 1 ) * kSegWords;
    if ( words > r->capWords )
    {
      size_t oldBytes = r->capWords * 4;
      DPCU_TRY( r->bits.reserve( words * 4, true, stream ) );
      size_t newCapWords = r->bits.capacity / 4;
      DPCU_CUDA( cudaMemsetAsync( static_cast<char *>( r->bits.ptr ) + oldBytes, 0, r->bits.capacity - oldBytes, stream ) );
      DPCU_TRY( r->chg.reserve( newCapWords * 4, false, stream ) );
      DPCU_CUDA( cudaMemsetAsync( r->chg.ptr, 0, r->chg.capacity, stream ) );
      r->capWords = newCapWords;
      if ( r->nPeers == 0 ) { /* local only */ }
    }
    if ( ctx->optChanged ) DPCU_TRY( r->changed.reserve( ( n ? n : 1 ) * 4, false, stream ) );
    size_t nLines = divUp( divUp( n, 32 ), 8 );          // look-back entries for the finest line size
    if ( ctx->optChanged && nLines > r->lookCap )
    {
      DPCU_TRY( r->look.reserve( ( nLines + nLines / 2 + 64 ) * 8, false, stream ) );
      DPCU_CUDA( cudaMemsetAsync( r->look.ptr, 0, r->look.capacity, stream ) );
      r->lookCap = r->look.capacity / 8;
      r->lookEpoch = 0;
    }
    if ( nSegs + 1 > r->nSegsCap )
    {
      size_t cap = nSegs + 1 + nSegs / 2;
      cap = ( cap + 127 ) & ~size_t( 127 );
      size_t entries = 2 * cap + 8;
      DPCU_TRY( r->counters.reserve( entries * 4, false, stream ) );
      r->nSegsCap = cap;
      DPCU_CUDA( cudaMemsetAsync( r->counters.ptr, 0, r->counters.capacity, stream ) );
    }
    return DPCU_OK;
  }

  // ViewFilter of one view-projection (row-major, row-vector convention), in double precision; every bound is
  // rounded towards "less decisive" (cull_filter.cuh)
  static float roundUp( double x )
  {
    float f = static_cast<float>( x );
    return ( static_cast<double>( f ) < x ) ? nextafterf( f, INFINITY ) : f;
  }
  static void makeViewFilter( float const *vp, ViewFilter &f, double marginScale )
  {
    double P[4][4];
    for ( int r = 0; r < 4; ++r ) for ( int c = 0; c < 4; ++c ) P[r][c] = vp[4 * r + c];
    const double inflate = 1.0 + 1.0 / 1048576.0;
    for ( int a = 0; a < 3; ++a )
    {
      double nN[3], nP[3];
      for ( int r = 0; r < 3; ++r )
      {
        nN[r] = P[r][a] + P[r][3];          // N plane: x + w
        nP[r] = P[r][3] - P[r][a];          // P plane: w - x
      }
      f.rhoN[a] = roundUp( sqrt( nN[0] * nN[0] + nN[1] * nN[1] + nN[2] * nN[2] ) * inflate );
      f.rhoP[a] = roundUp( sqrt( nP[0] * nP[0] + nP[1] * nP[1] + nP[2] * nP[2] ) * inflate );
    }
    double q = 0.0;
    for ( int r = 0; r < 4; ++r ) q += fabs( P[r][0] ) + fabs( P[r][1] ) + fabs( P[r][2] ) + fabs( P[r][3] );
    f.q = roundUp( marginScale * q / 131072.0 );
    f.pad = 0.0f;
    memcpy( f.rows, vp, 64 );
  }

  template <int NV>
  static int launchCull( dpcuCull *ctx, dpcuCullResult *const *results, float const *vps, cudaStream_t stream, LeafArgs const *leaf,
                         bool *mirrorsWritten, bool *listBuilt )
  {
    CullArgs<NV> args;
    memset( &args, 0, sizeof args );
    args.lowerIdx = static_cast<float4 const *>( ctx->lowerIdx.ptr );
    args.extent   = static_cast<float4 const *>( ctx->extent.ptr );
    args.mats     = ctx->matsPtr();
    args.n        = uint32_t( ctx->n );
    args.nTiles   = uint32_t( divUp( ctx->n, size_t( kCullThreads ) ) );
    args.buildChanged = ctx->optChanged;
    args.nSegs    = uint32_t( divUp( ctx->n, size_t( 1 ) << kSegObjectsLog2 ) );
    args.done     = results[0]->donePtr();
    args.nPeers   = 0;
    for ( int v = 0; v < NV; ++v )
    {
      dpcuCullResult *r = results[v];
      args.out[v].bits  = static_cast<uint32_t *>( r->bits.ptr );
      args.out[v].chg   = static_cast<uint32_t *>( r->chg.ptr );
      args.out[v].seg   = r->segPtr();
      args.out[v].prefix = r->prefixPtr();
      args.out[v].mirror = r->dBits;
      for ( int p = 0; p < kMaxPeers; ++p ) args.out[v].peer[p] = p < r->nPeers ? r->peer[p] : nullptr;
      if ( uint32_t( r->nPeers ) > args.nPeers ) args.nPeers = uint32_t( r->nPeers );
      args.peerWordOffset = uint32_t( r->peerWordOffset );
      memcpy( args.vp[v], vps + 16 * v, 64 );
      // DPCU_CULL_OPT_FILTER 2 / 3 shrink the margin to 1/8 (the bound of the error analysis itself) / to zero:
      // diagnostics that measure the slack of the proof, never for production
      if ( NV > 1 ) makeViewFilter( vps + 16 * v, args.filter[v], ctx->optFilter == 2 ? 0.125 : ctx->optFilter == 3 ? 0.0 : 1.0 );
    }
    args.useFilter = ( NV > 1 && ctx->optFilter ) ? 1 : 0;
    // kernel choice: one view is an HBM-bound stream (direct kernel); with more views the cull is
    // issue-bound and the view-sequential packed kernel (cull_views.cuh) is the faster exact form
    const bool peers     = args.nPeers > 0;
    bool mirrors = false;
    for ( int v = 0; v < NV; ++v ) mirrors = mirrors || results[v]->dBits != nullptr;
    // Whole 128-byte lines are what NVLink peers and PCIe host mirrors want to see, and the line-granular form
    // builds the changed list in the same pass (no compaction kernel): measured at 64 Mi objects, step time with
    // an ordered changed list, lines vs the alternative - 1 view 1.003 vs 1.021 ms (direct + compaction), 2 views
    // 1.093 vs 1.149 ms, 3 views 1.285 vs 1.33 ms, 6 views 1.93 vs 2.26 ms (views + compaction).  A warp works
    // through whole 1024-object lines, so the kernel wants several lines per resident warp: with 1 view it loses
    // to direct + compaction below ~32 Mi objects (8 Mi: 0.192 vs 0.161 ms, 16 Mi: 0.298 vs 0.277 ms, 32 Mi: 0.530
    // vs 0.535 ms per step) - AUTO asks for four lines per resident warp there; with several views (where the
    // alternative is the views kernel) one line per resident warp is enough: 6 views, 8 Mi objects 0.280 vs 0.307 ms,
    // 16 Mi 0.502 vs 0.588 ms.  Below that it stays with one thread per object and serves a host mirror by a copy
    // queued behind the kernel.
    // (The same kernel with 8 or 16 bitset words per warp - DPCU_CULL_OPT_LINE_WORDS - fills the machine on
    // mid-size groups but does not beat direct + compaction there either: 1 / 2 / 4 Mi objects 57 / 81 / 114 us vs
    // 48 / 70 / 108 us per step.)
    const size_t wantedLines = size_t( ctx->smCount ) * ( NV == 1 ? 4u * 48u : 32u );
    const bool bigEnough = divUp( divUp( ctx->n, 32 ), 32 ) >= wantedLines;
    const bool autoLines = ctx->optKernel == DPCU_KERNEL_AUTO && !leaf && bigEnough
                        && ( mirrors || ( ctx->optChanged && ctx->optFuseList ) );
    args.lineWords = ctx->optLineWords ? uint32_t( ctx->optLineWords ) : 32u;
    const bool useLines  = !ctx->optFma && ( peers || ctx->optKernel == DPCU_KERNEL_LINES || autoLines );
    *mirrorsWritten = useLines && !leaf;
    const bool fuseList = useLines && !leaf && ctx->optChanged && ctx->optFuseList;
    *listBuilt = fuseList;
    if ( fuseList )
    {
      for ( int v = 0; v < NV; ++v )
      {
        dpcuCullResult *r = results[v];
        if ( ++r->lookEpoch >= ( 1u << 30 ) )
        {
          DPCU_CUDA( cudaMemsetAsync( r->look.ptr, 0, r->look.capacity, stream ) );     // tags wrap: start over
          r->lookEpoch = 1;
        }
        args.out[v].look        = static_cast<unsigned long long *>( r->look.ptr );
        args.out[v].epoch       = r->lookEpoch;
        args.out[v].changed     = static_cast<uint32_t *>( r->changed.ptr );
        args.out[v].hostChanged = r->dChanged;
        args.out[v].hostCount   = r->dCount;
        args.out[v].hostCap     = uint32_t( r->hChangedCap < 0xffffffffull ? r->hChangedCap : 0xffffffffull );
      }
    }
    if ( peers && ( ctx->optFma || leaf ) )
      return fail( DPCU_ERR_INVALID_VALUE, "dpcuCullRun: peer bitsets are served by the line-granular kernel only (not the FMA or fused-leaf forms)" );
    const bool useFused  = leaf != nullptr;
    const bool useStaged = !useFused && !useLines && !ctx->optFma && ctx->optKernel == DPCU_KERNEL_STAGED;
    const bool useChains = ctx->optKernel == DPCU_KERNEL_VIEWS_CHAINS;
    const bool useViews  = !useFused && !useLines && !ctx->optFma && ( ctx->optKernel == DPCU_KERNEL_VIEWS || useChains || ( ctx->optKernel == DPCU_KERNEL_AUTO && NV >= 2 ) );
    args.chunkCounter = results[0]->donePtr() + 1;
    // the last CTA's scan re-arms ticket and chunk counter; without a changed list nobody does
    if ( useStaged && !ctx->optChanged ) DPCU_CUDA( cudaMemsetAsync( results[0]->donePtr(), 0, 16, stream ) );
    const size_t stagedSmem = sizeof( WarpRing ) * ( kCullThreads / 32 );
    if ( useStaged ) DPCU_CUDA( cudaFuncSetAttribute( cullStagedKernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, int( stagedSmem ) ) );
    args.vpFinite = 1;
    args.onePair  = 0x3f8000003f800000ull;
    for ( int k = 0; k < 16 * NV; ++k )
    {
      uint32_t u;
      memcpy( &u, vps + k, 4 );
      if ( ( u & 0x7f800000u ) == 0x7f800000u ) args.vpFinite = 0;
    }
    int perSm = ctx->optCtasPerSm;
    if ( perSm <= 0 )
    {
      if ( useFused ) cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, cullFusedLeafKernel<NV>, kCullThreads, 0 );
      else if ( useLines && fuseList ) cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, cullLinesKernel<NV, true>, kCullThreads, 0 );
      else if ( useLines ) cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, cullLinesKernel<NV, false>, kCullThreads, 0 );
      else if ( ctx->optFma ) perSm = occupancyCullDirectFma<NV>();
      else if ( useStaged ) cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, cullStagedKernel<NV>, kCullThreads, stagedSmem );
      else if ( useViews && useChains ) cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, cullViewsKernel<NV, false>, kCullThreads, 0 );
      else if ( useViews ) cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, cullViewsKernel<NV, true>, kCullThreads, 0 );
      else cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, cullDirectKernel<NV>, kCullThreads, 0 );
      if ( perSm <= 0 ) perSm = 1;
    }
    int grid = ctx->smCount * perSm;
    if ( uint32_t( grid ) > args.nTiles ) grid = int( args.nTiles );
    if ( useStaged )
    {
      // one chunk (kChunkTiles warp-tiles = 1024 objects) per warp to start with
      const uint32_t nChunks = uint32_t( divUp( divUp( ctx->n, 32 ), kChunkTiles ) );
      const uint32_t ctasForChunks = uint32_t( divUp( nChunks, kCullThreads / 32 ) );
      if ( uint32_t( grid ) > ctasForChunks ) grid = int( ctasForChunks );
    }
    cudaEvent_t evStart = nullptr, evStop = nullptr;
    if ( ctx->optProfile )
    {
      while ( ctx->profEvents.size() < ctx->profUsed + 2 )
      {
        cudaEvent_t e;
        DPCU_CUDA( cudaEventCreate( &e ) );
        ctx->profEvents.push_back( e );
      }
      evStart = ctx->profEvents[ctx->profUsed];
      evStop  = ctx->profEvents[ctx->profUsed + 1];
      ctx->profUsed += 2;
      DPCU_CUDA( cudaEventRecord( evStart, stream ) );
    }
    if ( useFused )
    {
      cullFusedLeafKernel<NV><<<grid, kCullThreads, 0, stream>>>( args, *leaf );
      DPCU_CUDA( cudaGetLastError() );
    }
    else if ( useLines )
    {
      const uint32_t nLines = uint32_t( divUp( divUp( ctx->n, 32 ), args.lineWords ) );
      const uint32_t ctasForLines = uint32_t( divUp( nLines, kCullThreads / 32 ) );
      if ( uint32_t( grid ) > ctasForLines ) grid = int( ctasForLines );
      if ( fuseList ) cullLinesKernel<NV, true><<<grid, kCullThreads, 0, stream>>>( args );
      else            cullLinesKernel<NV, false><<<grid, kCullThreads, 0, stream>>>( args );
      DPCU_CUDA( cudaGetLastError() );
    }
    else if ( ctx->optFma )
    {
      DPCU_CUDA( launchCullDirectFma<NV>( args, grid, stream ) );
    }
    else if ( useStaged )
    {
      cullStagedKernel<NV><<<grid, kCullThreads, stagedSmem, stream>>>( args );
      DPCU_CUDA( cudaGetLastError() );
    }
    else if ( useViews )
    {
      if ( useChains ) cullViewsKernel<NV, false><<<grid, kCullThreads, 0, stream>>>( args );
      else             cullViewsKernel<NV, true><<<grid, kCullThreads, 0, stream>>>( args );
      DPCU_CUDA( cudaGetLastError() );
    }
    else
    {
      cullDirectKernel<NV><<<grid, kCullThreads, 0, stream>>>( args );
      DPCU_CUDA( cudaGetLastError() );
    }
    if ( evStop ) DPCU_CUDA( cudaEventRecord( evStop, stream ) );
    ++ctx->launches;
    ctx->lastKernel = useFused ? DPCU_KERNEL_FUSED_LEAF : useLines ? DPCU_KERNEL_LINES : ctx->optFma ? DPCU_KERNEL_DIRECT
                    : useStaged ? DPCU_KERNEL_STAGED : useViews ? ( useChains ? DPCU_KERNEL_VIEWS_CHAINS : DPCU_KERNEL_VIEWS ) : DPCU_KERNEL_DIRECT;
    return DPCU_OK;
  }
}

extern "C"
{
  int dpcuCullCreate( dpcuCull **out, int device )
  {
    DPCU_REQUIRE( out, "out is NULL" );
    *out = nullptr;
    DPCU_TRY( dpcu::requireDevice() );
    int count = 0;
    DPCU_CUDA( cudaGetDeviceCount( &count ) );
    DPCU_REQUIRE( device >= 0 && device < count, "device index out of range" );
    dpcu::DeviceGuard guard( device );
    dpcuCull *ctx = new ( std::nothrow ) dpcuCull;
    if ( !ctx ) return dpcu::fail( DPCU_ERR_OUT_OF_MEMORY, "dpcuCullCreate: host allocation failed" );
    ctx->device = device;
    cudaError_t e = cudaDeviceGetAttribute( &ctx->smCount, cudaDevAttrMultiProcessorCount, device );
    if ( e == cudaSuccess ) e = cudaStreamCreateWithFlags( &ctx->stream, cudaStreamNonBlocking );
    if ( e != cudaSuccess ) { delete ctx; return dpcu::failCuda( e, "dpcuCullCreate", __FILE__, __LINE__ ); }
    int rc = ctx->maxIndex.reserve( 256, false, ctx->stream );
    if ( rc != DPCU_OK ) { cudaStreamDestroy( ctx->stream ); delete ctx; return rc; }
    *out = ctx;
    return DPCU_OK;
  }

  int dpcuCullDestroy( dpcuCull *ctx )
  {
    if ( !ctx ) return DPCU_OK;
    if ( ctx->results )
      return dpcu::fail( DPCU_ERR_INVALID_VALUE, "dpcuCullDestroy: results of this context are still alive "
                                                 "(destroy results first; ResultBitSet detaches from its group, ResultBitSet.cpp:51-54)" );
    dpcu::DeviceGuard guard( ctx->device );
    cudaStreamSynchronize( ctx->stream );
    ctx->lowerIdx.release(); ctx->extent.release(); ctx->mats.release(); ctx->scratch.release(); ctx->maxIndex.release();
    ctx->staging.release();
    ctx->uploads.destroy();
    ctx->lastRun.hostWait();
    ctx->lastRun.destroy();
    for ( cudaEvent_t e : ctx->profEvents ) cudaEventDestroy( e );
    cudaStreamDestroy( ctx->stream );
    delete ctx;
    return DPCU_OK;
  }

  int dpcuCullSetObjectRange( dpcuCull *ctx, size_t first, size_t count, const float *lower4, const float *extent4,
                              const uint32_t *transformIndex, int memspace )
  {
    DPCU_REQUIRE( ctx, "ctx is NULL" );
    DPCU_REQUIRE( first + count <= ctx->n, "range exceeds object count" );
    DPCU_REQUIRE( !count || ( lower4 && extent4 ), "NULL object arrays" );
    DPCU_REQUIRE( memspace == DPCU_MEM_DEVICE || transformIndex || !count, "transformIndex may only be NULL for device memory" );
    if ( !count ) return DPCU_OK;
    dpcu::DeviceGuard guard( ctx->device );
    DPCU_CUDA( ctx->lastRun.orderBefore( ctx->stream ) );
    float4 const *lo = nullptr, *ex = nullptr;
    uint32_t const *ti = nullptr;
    if ( memspace == DPCU_MEM_HOST )
    {
      // stage through scratch: [lower | extent | tidx]
      size_t bytes = count * ( 16 + 16 + 4 );
      DPCU_TRY( ctx->scratch.reserve( bytes, false, ctx->stream ) );
      char *s = static_cast<char *>( ctx->scratch.ptr );
      DPCU_CUDA( cudaMemcpyAsync( s, lower4, count * 16, cudaMemcpyHostToDevice, ctx->stream ) );
      DPCU_CUDA( cudaMemcpyAsync( s + count * 16, extent4, count * 16, cudaMemcpyHostToDevice, ctx->stream ) );
      DPCU_CUDA( cudaMemcpyAsync( s + count * 32, transformIndex, count * 4, cudaMemcpyHostToDevice, ctx->stream ) );
      lo = reinterpret_cast<float4 const *>( s );
      ex = reinterpret_cast<float4 const *>( s + count * 16 );
      ti = reinterpret_cast<uint32_t const *>( s + count * 32 );
    }
    else
    {
      lo = reinterpret_cast<float4 const *>( lower4 );
      ex = reinterpret_cast<float4 const *>( extent4 );
      ti = transformIndex;
    }
    dpcu::packObjectsKernel<<<unsigned( dpcu::divUp( count, 256 ) ), 256, 0, ctx->stream>>>(
      lo, ex, ti, uint32_t( count ), static_cast<float4 *>( ctx->lowerIdx.ptr ) + first,
      static_cast<float4 *>( ctx->extent.ptr ) + first, static_cast<uint32_t *>( ctx->maxIndex.ptr ) );
    DPCU_CUDA( cudaGetLastError() );
    ++ctx->launches;
    ++ctx->objectsVersion;
    ctx->maxIndexKnown = false;
    if ( memspace == DPCU_MEM_HOST ) DPCU_CUDA( cudaStreamSynchronize( ctx->stream ) );
    return DPCU_OK;
  }

  int dpcuCullSetObjects( dpcuCull *ctx, const float *lower4, const float *extent4, const uint32_t *transformIndex,
                          size_t n, int memspace )
  {
    DPCU_REQUIRE( ctx, "ctx is NULL" );
    DPCU_REQUIRE( n < ( size_t( 1 ) << 32 ), "object count must fit 32 bits" );
    DPCU_REQUIRE( memspace == DPCU_MEM_HOST || memspace == DPCU_MEM_DEVICE, "bad memspace" );
    dpcu::DeviceGuard guard( ctx->device );
    DPCU_CUDA( ctx->lastRun.orderBefore( ctx->stream ) );
    DPCU_TRY( ctx->lowerIdx.reserve( ( n ? n : 1 ) * 16, false, ctx->stream ) );
    DPCU_TRY( ctx->extent.reserve( ( n ? n : 1 ) * 16, false, ctx->stream ) );
    ctx->n = n;
    DPCU_CUDA( cudaMemsetAsync( ctx->maxIndex.ptr, 0, 4, ctx->stream ) );
    ctx->maxTransformIndex = 0;
    ctx->maxIndexKnown = true;
    return dpcuCullSetObjectRange( ctx, 0, n, lower4, extent4, transformIndex, memspace );
  }

  int dpcuCullGetObjectCount( const dpcuCull *ctx, size_t *n )
  {
    DPCU_REQUIRE( ctx && n, "NULL argument" );
    *n = ctx->n;
    return DPCU_OK;
  }

  int dpcuCullSetMatrices( dpcuCull *ctx, const void *matrices, size_t count, size_t strideBytes, int memspace )
  {
    DPCU_REQUIRE( ctx, "ctx is NULL" );
    DPCU_REQUIRE( !count || matrices, "matrices is NULL" );
    DPCU_REQUIRE( strideBytes >= 64 && strideBytes % 4 == 0, "stride must be >= 64 and a multiple of 4" );
    DPCU_REQUIRE( count < ( size_t( 1 ) << 32 ), "matrix count must fit 32 bits" );
    DPCU_REQUIRE( memspace == DPCU_MEM_HOST || memspace == DPCU_MEM_DEVICE, "bad memspace" );
    dpcu::DeviceGuard guard( ctx->device );
    DPCU_CUDA( ctx->lastRun.orderBefore( ctx->stream ) );
    DPCU_TRY( ctx->mats.reserve( ( count ? count : 1 ) * 64, false, ctx->stream ) );
    ctx->boundMats = nullptr;
    ctx->nMats = count;
    if ( !count ) return DPCU_OK;
    if ( strideBytes == 64 ) return dpcu::uploadOrCopy( ctx->mats.ptr, matrices, count * 64, memspace, ctx );
    if ( memspace == DPCU_MEM_HOST )
    {
      DPCU_CUDA( cudaMemcpy2DAsync( ctx->mats.ptr, 64, matrices, strideBytes, 64, count, cudaMemcpyHostToDevice, ctx->stream ) );
      DPCU_CUDA( cudaStreamSynchronize( ctx->stream ) );
    }
    else
    {
      dpcu::gatherStridedKernel<<<unsigned( dpcu::divUp( count * 16, 256 ) ), 256, 0, ctx->stream>>>(
        static_cast<char const *>( matrices ), strideBytes, uint32_t( count ), static_cast<float *>( ctx->mats.ptr ) );
      DPCU_CUDA( cudaGetLastError() );
      ++ctx->launches;
    }
    return DPCU_OK;
  }

  int dpcuCullUpdateMatrices( dpcuCull *ctx, const uint32_t *indices, size_t n, const void *matrices, size_t strideBytes,
                              int memspace )
  {
    DPCU_REQUIRE( ctx, "ctx is NULL" );
    DPCU_REQUIRE( !n || ( indices && matrices ), "NULL argument" );
    DPCU_REQUIRE( strideBytes >= 64 && strideBytes % 4 == 0, "stride must be >= 64 and a multiple of 4" );
    DPCU_REQUIRE( !ctx->boundMats, "matrices are bound to external device memory; update them there" );
    DPCU_REQUIRE( memspace == DPCU_MEM_HOST, "only host source matrices are supported for batched updates" );
    if ( !n ) return DPCU_OK;
    dpcu::DeviceGuard guard( ctx->device );
    DPCU_CUDA( ctx->lastRun.orderBefore( ctx->stream ) );
    // pack [matrices | indices] into pinned staging, skipping indices past the matrix count
    // (markMatrixDirty ignores those, dp/culling/GroupBitSet.h:140-150)
    DPCU_TRY( ctx->staging.reserve( n * 68 ) );
    char *st = static_cast<char *>( ctx->staging.ptr );
    uint32_t *sidx = reinterpret_cast<uint32_t *>( st + n * 64 );
    size_t k = 0;
    for ( size_t i = 0; i < n; ++i )
    {
      if ( indices[i] < ctx->nMats )
      {
        memcpy( st + k * 64, static_cast<char const *>( matrices ) + size_t( indices[i] ) * strideBytes, 64 );
        sidx[k++] = indices[i];
      }
    }
    if ( !k ) return DPCU_OK;
    if ( k != n ) memmove( st + k * 64, sidx, k * 4 );
    DPCU_TRY( ctx->scratch.reserve( k * 68, false, ctx->stream ) );
    DPCU_CUDA( cudaMemcpyAsync( ctx->scratch.ptr, st, k * 68, cudaMemcpyHostToDevice, ctx->stream ) );
    dpcu::scatterMatricesKernel<<<unsigned( dpcu::divUp( k * 4, 256 ) ), 256, 0, ctx->stream>>>(
      reinterpret_cast<uint32_t const *>( static_cast<char *>( ctx->scratch.ptr ) + k * 64 ),
      static_cast<float4 const *>( ctx->scratch.ptr ), uint32_t( k ), static_cast<float4 *>( ctx->mats.ptr ) );
    DPCU_CUDA( cudaGetLastError() );
    ++ctx->launches;
    DPCU_CUDA( cudaStreamSynchronize( ctx->stream ) );   // staging is reused by the next call
    return DPCU_OK;
  }

  int dpcuCullBindMatrices( dpcuCull *ctx, const void *deviceMatrices, size_t count )
  {
    DPCU_REQUIRE( ctx, "ctx is NULL" );
    DPCU_REQUIRE( deviceMatrices || !count, "deviceMatrices is NULL" );
    DPCU_REQUIRE( ( reinterpret_cast<uintptr_t>( deviceMatrices ) & 15 ) == 0, "matrices must be 16-byte aligned" );
    DPCU_REQUIRE( count < ( size_t( 1 ) << 32 ), "matrix count must fit 32 bits" );
    ctx->boundMats = static_cast<float const *>( deviceMatrices );
    ctx->nMats = count;
    return DPCU_OK;
  }

  int dpcuCullGetMatrixCount( const dpcuCull *ctx, size_t *count )
  {
    DPCU_REQUIRE( ctx && count, "NULL argument" );
    *count = ctx->nMats;
    return DPCU_OK;
  }

  int dpcuCullResultCreate( dpcuCull *ctx, dpcuCullResult **out )
  {
    DPCU_REQUIRE( ctx && out, "NULL argument" );
    *out = nullptr;
    dpcuCullResult *r = new ( std::nothrow ) dpcuCullResult;
    if ( !r ) return dpcu::fail( DPCU_ERR_OUT_OF_MEMORY, "dpcuCullResultCreate: host allocation failed" );
    r->ctx = ctx;
    r->next = ctx->results;
    if ( ctx->results ) ctx->results->prev = r;
    ctx->results = r;
    *out = r;
    return DPCU_OK;
  }

  int dpcuCullResultDestroy( dpcuCullResult *r )
  {
    if ( !r ) return DPCU_OK;
    dpcuCull *ctx = r->ctx;
    dpcu::DeviceGuard guard( ctx->device );
    r->done.hostWait();
    r->done.destroy();
    r->bits.release(); r->chg.release(); r->changed.release(); r->counters.release();
    r->visible.release(); r->visCounters.release(); r->look.release();
    if ( r->prev ) r->prev->next = r->next; else ctx->results = r->next;
    if ( r->next ) r->next->prev = r->prev;
    delete r;
    return DPCU_OK;
  }

}   // extern "C"

  static int runCull( dpcuCull *ctx, dpcuCullResult *const *results, const float *viewProjections, int nViews, cudaStream_t s, dpcu::LeafArgs const *leaf )
  {
    DPCU_REQUIRE( ctx && results && viewProjections, "NULL argument" );
    DPCU_REQUIRE( nViews >= 1 && nViews <= DPCU_MAX_VIEWS, "nViews must be 1..DPCU_MAX_VIEWS" );
    for ( int v = 0; v < nViews; ++v )
    {
      DPCU_REQUIRE( results[v] && results[v]->ctx == ctx, "result does not belong to this context" );
      for ( int u = 0; u < v; ++u ) DPCU_REQUIRE( results[u] != results[v], "results must be distinct" );
    }
    if ( s != ctx->stream )
    {
      // uploads were submitted on the context stream: order this run after them on the device
      DPCU_CUDA( ctx->uploads.record( ctx->stream ) );
      DPCU_CUDA( ctx->uploads.orderBefore( s ) );
    }
    const size_t n = ctx->n;
    if ( n )
    {
      DPCU_TRY( dpcu::refreshMaxIndex( ctx ) );
      if ( ctx->maxTransformIndex >= ctx->nMats )
        return dpcu::fail( DPCU_ERR_INVALID_VALUE, "dpcuCullRun: transform index %u out of range (%zu matrices)",
                           ctx->maxTransformIndex, ctx->nMats );
    }
    const size_t nSegs = dpcu::divUp( n, size_t( 1 ) << dpcu::kSegObjectsLog2 );
    for ( int v = 0; v < nViews; ++v )
    {
      dpcuCullResult *r = results[v];
      DPCU_CUDA( r->done.orderBefore( s ) );      // after this result's previous cull / bit moves, wherever they ran
      DPCU_TRY( dpcu::ensureResultCapacity( r, n, s ) );
      if ( r->n != n )
      {
        size_t words = dpcu::divUp( r->n > n ? r->n : n, 32 );
        dpcu::resizeBitsKernel<<<unsigned( dpcu::divUp( words, 256 ) ), 256, 0, s>>>(
          static_cast<uint32_t *>( r->bits.ptr ), uint32_t( r->n ), uint32_t( n ), uint32_t( words ) );
        DPCU_CUDA( cudaGetLastError() );
        ++ctx->launches;
        r->n = n;
      }
      // ticket and segment counters are zero here: zeroed at allocation and again by every cull's last CTA
      r->nSegs = nSegs;
      r->ran = true;
      r->visBuilt = false;
      if ( r->hBits && r->hBitsWords < dpcu::divUp( n, 32 ) )
        return dpcu::fail( DPCU_ERR_INVALID_VALUE, "dpcuCullRun: host mirror holds %zu bitset words, %zu needed", r->hBitsWords, dpcu::divUp( n, 32 ) );
    }
    if ( !n )
    {
      for ( int v = 0; v < nViews; ++v )
      {
        if ( results[v]->dCount ) DPCU_CUDA( cudaMemsetAsync( results[v]->dCount, 0, 4, s ) );
        DPCU_CUDA( results[v]->done.record( s ) );
      }
      return DPCU_OK;
    }
    int rc = DPCU_OK;
    bool mirrorsWritten = false, listBuilt = false;
    switch ( nViews )
    {
      case 1: rc = dpcu::launchCull<1>( ctx, results, viewProjections, s, leaf, &mirrorsWritten, &listBuilt ); break;
      case 2: rc = dpcu::launchCull<2>( ctx, results, viewProjections, s, leaf, &mirrorsWritten, &listBuilt ); break;
      case 3: rc = dpcu::launchCull<3>( ctx, results, viewProjections, s, leaf, &mirrorsWritten, &listBuilt ); break;
      case 4: rc = dpcu::launchCull<4>( ctx, results, viewProjections, s, leaf, &mirrorsWritten, &listBuilt ); break;
      case 5: rc = dpcu::launchCull<5>( ctx, results, viewProjections, s, leaf, &mirrorsWritten, &listBuilt ); break;
      case 6: rc = dpcu::launchCull<6>( ctx, results, viewProjections, s, leaf, &mirrorsWritten, &listBuilt ); break;
      case 7: rc = dpcu::launchCull<7>( ctx, results, viewProjections, s, leaf, &mirrorsWritten, &listBuilt ); break;
      case 8: rc = dpcu::launchCull<8>( ctx, results, viewProjections, s, leaf, &mirrorsWritten, &listBuilt ); break;
    }
    DPCU_TRY( rc );
    if ( ctx->optChanged && !listBuilt )
    {
      dpcu::CompactArgs ca;
      memset( &ca, 0, sizeof ca );
      for ( int v = 0; v < nViews; ++v )
      {
        dpcuCullResult *r = results[v];
        ca.chg[v] = static_cast<uint32_t const *>( r->chg.ptr );
        ca.prefix[v] = r->prefixPtr();
        ca.changed[v] = static_cast<uint32_t *>( r->changed.ptr );
        ca.hostChanged[v]  = r->dChanged;
        ca.hostCount[v]    = r->dCount;
        ca.hostCapacity[v] = uint32_t( r->hChangedCap < 0xffffffffull ? r->hChangedCap : 0xffffffffull );
      }
      ca.nWords = uint32_t( dpcu::divUp( n, 32 ) );
      ca.nSegs = uint32_t( nSegs );
      dim3 grid( ca.nSegs, unsigned( nViews ) );
      dpcu::compactChangedKernel<<<grid, 256, 0, s>>>( ca );
      DPCU_CUDA( cudaGetLastError() );
      ++ctx->launches;
    }
    if ( !mirrorsWritten )
    {
      // kernel forms that do not store whole lines themselves (fused leaf level, explicitly chosen forms)
      for ( int v = 0; v < nViews; ++v )
      {
        dpcuCullResult *r = results[v];
        if ( r->hBits ) DPCU_CUDA( cudaMemcpyAsync( r->hBits, r->bits.ptr, dpcu::divUp( n, 32 ) * 4, cudaMemcpyDeviceToHost, s ) );
      }
    }
    for ( int v = 0; v < nViews; ++v ) DPCU_CUDA( results[v]->done.record( s ) );
    if ( s != ctx->stream ) DPCU_CUDA( ctx->lastRun.record( s ) );
    return DPCU_OK;
  }


extern "C"
{
  int dpcuCullRun( dpcuCull *ctx, dpcuCullResult *const *results, const float *viewProjections, int nViews, dpcuStream *stream )
  {
    DPCU_REQUIRE( ctx, "ctx is NULL" );
    dpcu::DeviceGuard guard( ctx->device );
    return runCull( ctx, results, viewProjections, nViews, stream ? stream->stream : ctx->stream, nullptr );
  }

  int dpcuCullRunWithTree( dpcuCull *ctx, dpcuTree *tree, dpcuCullResult *const *results, const float *viewProjections, int nViews,
                           dpcuStream *stream )
  {
    DPCU_REQUIRE( ctx && tree, "NULL argument" );
    DPCU_REQUIRE( tree->device == ctx->device, "tree and culling context live on different devices" );
    DPCU_REQUIRE( tree->numNodes >= 1, "no topology set" );
    dpcu::DeviceGuard guard( ctx->device );
    cudaStream_t s = stream ? stream->stream : ctx->stream;
    // the culler reads the tree's world matrices in place
    ctx->boundMats = static_cast<float const *>( tree->world.ptr );
    ctx->nMats     = tree->numNodes;
    const size_t levels = tree->levelOffsets.empty() ? 0 : tree->levelOffsets.size() - 1;
    // can the last level run inside the cull kernel?  object i <-> entry i of that level
    bool fuse = false;
    size_t lastFirst = 0;
    if ( levels >= 1 && ctx->n && ctx->optFuseLeaf && !ctx->optFma )
    {
      lastFirst = tree->levelOffsets[levels - 1];
      const size_t lastCount = tree->levelOffsets[levels] - lastFirst;
      if ( lastCount == ctx->n )
      {
        if ( ctx->leafTree != tree || ctx->leafTopologyVersion != tree->topologyVersion || ctx->leafObjectsVersion != ctx->objectsVersion )
        {
          DPCU_TRY( ctx->scratch.reserve( 256, false, ctx->stream ) );
          DPCU_CUDA( cudaMemsetAsync( ctx->scratch.ptr, 0, 4, ctx->stream ) );
          dpcu::leafBindingKernel<<<unsigned( dpcu::divUp( ctx->n, 256 ) ), 256, 0, ctx->stream>>>(
            static_cast<float4 const *>( ctx->lowerIdx.ptr ), static_cast<uint2 const *>( tree->entries.ptr ) + lastFirst,
            uint32_t( ctx->n ), static_cast<uint32_t *>( ctx->scratch.ptr ) );
          DPCU_CUDA( cudaGetLastError() );
          ++ctx->launches;
          uint32_t mismatches = 0;
          DPCU_CUDA( cudaMemcpyAsync( &mismatches, ctx->scratch.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream ) );
          DPCU_CUDA( cudaStreamSynchronize( ctx->stream ) );
          ctx->leafTree = tree;
          ctx->leafTopologyVersion = tree->topologyVersion;
          ctx->leafObjectsVersion  = ctx->objectsVersion;
          ctx->leafBound = mismatches == 0;
        }
        fuse = ctx->leafBound;
      }
    }
    DPCU_TRY( dpcu::treeBeginCompute( tree, s ) );
    DPCU_TRY( dpcu::treeComputeLevels( tree, s, 0, fuse ? levels - 1 : levels ) );
    dpcu::LeafArgs leaf;
    leaf.entries    = static_cast<uint2 const *>( tree->entries.ptr ) + lastFirst;
    leaf.local      = static_cast<float4 const *>( tree->local.ptr );
    leaf.world      = static_cast<float4 *>( tree->world.ptr );
    leaf.dirtyLocal = static_cast<uint32_t const *>( tree->dirtyLocal.ptr );
    leaf.dirtyWorld = static_cast<uint32_t *>( tree->dirtyWorld.ptr );
    int rc = runCull( ctx, results, viewProjections, nViews, s, fuse ? &leaf : nullptr );
    if ( fuse && rc == DPCU_OK ) ++tree->launches;        // the fused kernel is the tree's last level too
    // close the tree's compute even if the cull failed, so its dirty state stays consistent
    int rc2 = dpcu::treeEndCompute( tree, s );
    return rc != DPCU_OK ? rc : rc2;
  }

  int dpcuCullResultGetBits( dpcuCullResult *r, uint32_t *hostWords, size_t nWords )
  {
    DPCU_REQUIRE( r && ( hostWords || !nWords ), "NULL argument" );
    size_t have = dpcu::divUp( r->n, 32 );
    DPCU_REQUIRE( nWords >= have, "nWords smaller than ceil(n/32)" );
    if ( !have ) return DPCU_OK;
    dpcu::DeviceGuard guard( r->ctx->device );
    cudaStream_t s = r->ctx->stream;
    DPCU_CUDA( r->done.orderBefore( s ) );
    DPCU_CUDA( cudaMemcpyAsync( hostWords, r->bits.ptr, have * 4, cudaMemcpyDeviceToHost, s ) );
    DPCU_CUDA( cudaStreamSynchronize( s ) );
    return DPCU_OK;
  }

  int dpcuCullResultGetChangedCount( dpcuCullResult *r, size_t *count )
  {
    DPCU_REQUIRE( r && count, "NULL argument" );
    *count = 0;
    if ( !r->ran || !r->n ) return DPCU_OK;
    if ( !r->ctx->optChanged ) return dpcu::fail( DPCU_ERR_NOT_READY, "changed list disabled (DPCU_CULL_OPT_CHANGED_LIST = 0)" );
    dpcu::DeviceGuard guard( r->ctx->device );
    uint32_t c = 0;
    cudaStream_t s = r->ctx->stream;
    DPCU_CUDA( r->done.orderBefore( s ) );
    DPCU_CUDA( cudaMemcpyAsync( &c, r->countPtr(), 4, cudaMemcpyDeviceToHost, s ) );
    DPCU_CUDA( cudaStreamSynchronize( s ) );
    *count = c;
    return DPCU_OK;
  }

  int dpcuCullResultGetChanged( dpcuCullResult *r, uint32_t *hostIndices, size_t capacity, size_t *count )
  {
    DPCU_REQUIRE( r && count, "NULL argument" );
    DPCU_TRY( dpcuCullResultGetChangedCount( r, count ) );
    size_t c = *count < capacity ? *count : capacity;
    if ( c )
    {
      DPCU_REQUIRE( hostIndices, "hostIndices is NULL" );
      dpcu::DeviceGuard guard( r->ctx->device );
      cudaStream_t s = r->ctx->stream;     // already ordered after the cull by dpcuCullResultGetChangedCount
      DPCU_CUDA( cudaMemcpyAsync( hostIndices, r->changed.ptr, c * 4, cudaMemcpyDeviceToHost, s ) );
      DPCU_CUDA( cudaStreamSynchronize( s ) );
    }
    return DPCU_OK;
  }

  int dpcuCullResultIsVisible( dpcuCullResult *r, size_t groupIndex, int *visible )
  {
    DPCU_REQUIRE( r && visible, "NULL argument" );
    *visible = 1;                                  // ResultBitSet::isVisible: true when index >= size
    if ( groupIndex >= r->n ) return DPCU_OK;
    dpcu::DeviceGuard guard( r->ctx->device );
    if ( r->hBits && r->ran )
    {
      // the mirror holds the bits of the last cull (and of later bit moves): no device round trip per query
      r->done.hostWait();
      *visible = int( ( r->hBits[groupIndex >> 5] >> ( groupIndex & 31 ) ) & 1u );
      return DPCU_OK;
    }
    uint32_t w = 0;
    cudaStream_t s = r->ctx->stream;
    DPCU_CUDA( r->done.orderBefore( s ) );
    DPCU_CUDA( cudaMemcpyAsync( &w, static_cast<uint32_t *>( r->bits.ptr ) + ( groupIndex >> 5 ), 4, cudaMemcpyDeviceToHost, s ) );
    DPCU_CUDA( cudaStreamSynchronize( s ) );
    *visible = int( ( w >> ( groupIndex & 31 ) ) & 1u );
    return DPCU_OK;
  }

  int dpcuCullResultMoveBit( dpcuCullResult *r, size_t oldIndex, size_t newIndex )
  {
    DPCU_REQUIRE( r, "result is NULL" );
    if ( newIndex >= r->n ) return DPCU_OK;
    dpcu::DeviceGuard guard( r->ctx->device );
    uint32_t o = oldIndex < ( size_t( 1 ) << 32 ) ? uint32_t( oldIndex ) : 0xffffffffu;
    cudaStream_t s = r->ctx->stream;
    DPCU_CUDA( r->done.orderBefore( s ) );
    dpcu::moveBitKernel<<<1, 1, 0, s>>>( static_cast<uint32_t *>( r->bits.ptr ), uint32_t( r->n ), o, uint32_t( newIndex ),
                                         r->ran ? r->dBits : nullptr );
    DPCU_CUDA( cudaGetLastError() );
    DPCU_CUDA( r->done.record( s ) );
    ++r->ctx->launches;
    return DPCU_OK;
  }

  int dpcuCullResultDevicePointers( dpcuCullResult *r, const uint32_t **bits, size_t *nWords, const uint32_t **changedIndices,
                                    const uint32_t **changedCount )
  {
    DPCU_REQUIRE( r, "result is NULL" );
    if ( bits ) *bits = static_cast<uint32_t const *>( r->bits.ptr );
    if ( nWords ) *nWords = dpcu::divUp( r->n, 32 );
    if ( changedIndices ) *changedIndices = static_cast<uint32_t const *>( r->changed.ptr );
    if ( changedCount ) *changedCount = r->counters.ptr ? r->countPtr() : nullptr;
    return DPCU_OK;
  }

  int dpcuCullResultSetPeerBits( dpcuCullResult *r, uint32_t *const *peerBits, int nPeers, size_t wordOffset )
  {
    DPCU_REQUIRE( r, "result is NULL" );
    DPCU_REQUIRE( nPeers >= 0 && nPeers <= dpcu::kMaxPeers, "nPeers must be 0..8" );
    DPCU_REQUIRE( nPeers == 0 || peerBits, "peerBits is NULL" );
    DPCU_REQUIRE( wordOffset < ( size_t( 1 ) << 32 ), "wordOffset must fit 32 bits" );
    for ( int p = 0; p < dpcu::kMaxPeers; ++p ) r->peer[p] = p < nPeers ? peerBits[p] : nullptr;
    r->nPeers = nPeers;
    r->peerWordOffset = wordOffset;
    return DPCU_OK;
  }

  int dpcuCullResultBuildVisibleList( dpcuCullResult *r, dpcuStream *stream )
  {
    DPCU_REQUIRE( r, "result is NULL" );
    dpcuCull *ctx = r->ctx;
    dpcu::DeviceGuard guard( ctx->device );
    cudaStream_t s = stream ? stream->stream : ctx->stream;
    const size_t n = r->n;
    const size_t nSegs = dpcu::divUp( n, size_t( 1 ) << dpcu::kSegObjectsLog2 );
    DPCU_CUDA( r->done.orderBefore( s ) );
    if ( nSegs + 1 > r->visSegsCap )
    {
      size_t cap = ( nSegs + 1 + nSegs / 2 + 127 ) & ~size_t( 127 );
      DPCU_TRY( r->visCounters.reserve( ( 2 * cap + 8 ) * 4, false, s ) );
      DPCU_CUDA( cudaMemsetAsync( r->visCounters.ptr, 0, r->visCounters.capacity, s ) );
      r->visSegsCap = cap;
    }
    DPCU_TRY( r->visible.reserve( ( n ? n : 1 ) * 4, false, s ) );
    uint32_t *done = static_cast<uint32_t *>( r->visCounters.ptr ), *seg = done + 4, *prefix = seg + r->visSegsCap;
    r->visSegs = nSegs;
    if ( !n )
    {
      DPCU_CUDA( cudaMemsetAsync( prefix, 0, 4, s ) );
    }
    else
    {
      int grid = int( nSegs < size_t( ctx->smCount ) * 8 ? nSegs : size_t( ctx->smCount ) * 8 );
      dpcu::segmentPopcountKernel<<<grid, dpcu::kCullThreads, 0, s>>>( static_cast<uint32_t const *>( r->bits.ptr ), uint32_t( dpcu::divUp( n, 32 ) ),
                                                                       uint32_t( nSegs ), seg, prefix, done );
      DPCU_CUDA( cudaGetLastError() );
      dpcu::CompactArgs ca;
      memset( &ca, 0, sizeof ca );
      ca.chg[0]     = static_cast<uint32_t const *>( r->bits.ptr );
      ca.prefix[0]  = prefix;
      ca.changed[0] = static_cast<uint32_t *>( r->visible.ptr );
      ca.nWords = uint32_t( dpcu::divUp( n, 32 ) );
      ca.nSegs  = uint32_t( nSegs );
      dpcu::compactChangedKernel<<<dim3( ca.nSegs, 1 ), 256, 0, s>>>( ca );
      DPCU_CUDA( cudaGetLastError() );
      ctx->launches += 2;
    }
    DPCU_CUDA( r->done.record( s ) );
    r->visBuilt = true;
    return DPCU_OK;
  }

  int dpcuCullResultVisibleDevicePointers( dpcuCullResult *r, const uint32_t **indices, const uint32_t **count )
  {
    DPCU_REQUIRE( r, "result is NULL" );
    if ( !r->visBuilt ) return dpcu::fail( DPCU_ERR_NOT_READY, "dpcuCullResultVisibleDevicePointers: no visible list built since the last cull" );
    if ( indices ) *indices = static_cast<uint32_t const *>( r->visible.ptr );
    if ( count ) *count = static_cast<uint32_t const *>( r->visCounters.ptr ) + 4 + r->visSegsCap + r->visSegs;
    return DPCU_OK;
  }

  int dpcuCullResultGetVisible( dpcuCullResult *r, uint32_t *hostIndices, size_t capacity, size_t *count )
  {
    DPCU_REQUIRE( r && count, "NULL argument" );
    *count = 0;
    const uint32_t *dIdx = nullptr, *dCount = nullptr;
    DPCU_TRY( dpcuCullResultVisibleDevicePointers( r, &dIdx, &dCount ) );
    dpcu::DeviceGuard guard( r->ctx->device );
    cudaStream_t s = r->ctx->stream;
    uint32_t c = 0;
    DPCU_CUDA( r->done.orderBefore( s ) );
    DPCU_CUDA( cudaMemcpyAsync( &c, dCount, 4, cudaMemcpyDeviceToHost, s ) );
    DPCU_CUDA( cudaStreamSynchronize( s ) );
    *count = c;
    size_t take = c < capacity ? c : capacity;
    if ( take )
    {
      DPCU_REQUIRE( hostIndices, "hostIndices is NULL" );
      DPCU_CUDA( cudaMemcpyAsync( hostIndices, dIdx, take * 4, cudaMemcpyDeviceToHost, s ) );
      DPCU_CUDA( cudaStreamSynchronize( s ) );
    }
    return DPCU_OK;
  }

  int dpcuCullResultSetHostMirror( dpcuCullResult *r, uint32_t *hostBits, size_t nWords, uint32_t *hostChanged, size_t changedCapacity,
                                   uint32_t *hostChangedCount )
  {
    DPCU_REQUIRE( r, "result is NULL" );
    DPCU_REQUIRE( !hostBits || nWords, "hostBits given with nWords == 0" );
    DPCU_REQUIRE( !hostChanged || hostChangedCount, "hostChanged needs hostChangedCount" );
    dpcu::DeviceGuard guard( r->ctx->device );
    r->done.hostWait();                       // nothing in flight may still write through the old pointers
    void *dBits = nullptr, *dChanged = nullptr, *dCount = nullptr;
    if ( hostBits && cudaHostGetDevicePointer( &dBits, hostBits, 0 ) != cudaSuccess )
    {
      cudaGetLastError();
      return dpcu::fail( DPCU_ERR_INVALID_VALUE, "dpcuCullResultSetHostMirror: hostBits is not pinned, device-mapped host memory (use dpcuHostBufferCreate)" );
    }
    if ( hostChanged && cudaHostGetDevicePointer( &dChanged, hostChanged, 0 ) != cudaSuccess )
    {
      cudaGetLastError();
      return dpcu::fail( DPCU_ERR_INVALID_VALUE, "dpcuCullResultSetHostMirror: hostChanged is not pinned, device-mapped host memory" );
    }
    if ( hostChangedCount && cudaHostGetDevicePointer( &dCount, hostChangedCount, 0 ) != cudaSuccess )
    {
      cudaGetLastError();
      return dpcu::fail( DPCU_ERR_INVALID_VALUE, "dpcuCullResultSetHostMirror: hostChangedCount is not pinned, device-mapped host memory" );
    }
    r->hBits = hostBits;       r->dBits = static_cast<uint32_t *>( dBits );       r->hBitsWords = hostBits ? nWords : 0;
    r->hChanged = hostChanged; r->dChanged = static_cast<uint32_t *>( dChanged ); r->hChangedCap = hostChanged ? changedCapacity : 0;
    r->hCount = hostChangedCount; r->dCount = static_cast<uint32_t *>( dCount );
    return DPCU_OK;
  }

  int dpcuCullResultSynchronize( dpcuCullResult *r )
  {
    DPCU_REQUIRE( r, "result is NULL" );
    dpcu::DeviceGuard guard( r->ctx->device );
    if ( r->done.pending ) DPCU_CUDA( cudaEventSynchronize( r->done.event ) );
    return DPCU_OK;
  }

  int dpcuCullGetBoundingBox( dpcuCull *ctx, float *out6 )
  {
    DPCU_REQUIRE( ctx && out6, "NULL argument" );
    dpcu::DeviceGuard guard( ctx->device );
    const float FMAX = 3.402823466e+38f;
    float lo[3] = { FMAX, FMAX, FMAX }, hi[3] = { -FMAX, -FMAX, -FMAX };
    if ( ctx->n )
    {
      DPCU_TRY( dpcu::refreshMaxIndex( ctx ) );
      if ( ctx->maxTransformIndex >= ctx->nMats )
        return dpcu::fail( DPCU_ERR_INVALID_VALUE, "dpcuCullGetBoundingBox: transform index %u out of range (%zu matrices)",
                           ctx->maxTransformIndex, ctx->nMats );
      DPCU_TRY( ctx->scratch.reserve( 256, false, ctx->stream ) );
      uint32_t init[6] = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u };
      DPCU_CUDA( cudaMemcpyAsync( ctx->scratch.ptr, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream ) );
      int grid = int( dpcu::divUp( ctx->n, 256 ) );
      if ( grid > ctx->smCount * 8 ) grid = ctx->smCount * 8;
      dpcu::boundingBoxKernel<<<grid, 256, 0, ctx->stream>>>( static_cast<float4 const *>( ctx->lowerIdx.ptr ),
                                                              static_cast<float4 const *>( ctx->extent.ptr ), ctx->matsPtr(),
                                                              uint32_t( ctx->n ), static_cast<uint32_t *>( ctx->scratch.ptr ) );
      DPCU_CUDA( cudaGetLastError() );
      ++ctx->launches;
      uint32_t keys[6];
      DPCU_CUDA( cudaMemcpyAsync( keys, ctx->scratch.ptr, sizeof keys, cudaMemcpyDeviceToHost, ctx->stream ) );
      DPCU_CUDA( cudaStreamSynchronize( ctx->stream ) );
      for ( int k = 0; k < 6; ++k )
      {
        uint32_t u = ( keys[k] & 0x80000000u ) ? ( keys[k] & 0x7fffffffu ) : ~keys[k];
        float f;
        memcpy( &f, &u, 4 );
        ( k < 3 ? lo[k] : hi[k - 3] ) = f;
      }
    }
    // Box3f( lower, upper ) = init + update + update  (dp/math/Boxnt.h:215-221)
    float blo[3] = { FMAX, FMAX, FMAX }, bhi[3] = { -FMAX, -FMAX, -FMAX };
    for ( int pass = 0; pass < 2; ++pass )
    {
      float const *p = pass ? hi : lo;
      for ( int k = 0; k < 3; ++k )
      {
        if ( blo[k] > p[k] ) blo[k] = p[k];
        if ( bhi[k] < p[k] ) bhi[k] = p[k];
      }
    }
    for ( int k = 0; k < 3; ++k ) { out6[k] = blo[k]; out6[3 + k] = bhi[k]; }
    return DPCU_OK;
  }

  int dpcuCullSetOption( dpcuCull *ctx, int option, int value )
  {
    DPCU_REQUIRE( ctx, "ctx is NULL" );
    switch ( option )
    {
      case DPCU_CULL_OPT_KERNEL:       DPCU_REQUIRE( value >= 0 && value <= 5, "kernel must be 0..5" ); ctx->optKernel = value; break;
      case DPCU_CULL_OPT_FMA:          ctx->optFma = value ? 1 : 0; break;
      case DPCU_CULL_OPT_CHANGED_LIST: ctx->optChanged = value ? 1 : 0; break;
      case DPCU_CULL_OPT_CTAS_PER_SM:  DPCU_REQUIRE( value >= 0 && value <= 32, "ctas per SM must be 0..32" ); ctx->optCtasPerSm = value; break;
      case DPCU_CULL_OPT_PROFILE:      ctx->optProfile = value ? 1 : 0; break;
      case DPCU_CULL_OPT_FUSE_LEAF:    ctx->optFuseLeaf = value ? 1 : 0; break;
      case DPCU_CULL_OPT_FUSE_LIST:    ctx->optFuseList = value ? 1 : 0; break;
      case DPCU_CULL_OPT_FILTER:       DPCU_REQUIRE( value >= 0 && value <= 3, "filter must be 0..3" ); ctx->optFilter = value; break;
      case DPCU_CULL_OPT_LINE_WORDS:   DPCU_REQUIRE( value == 0 || value == 8 || value == 16 || value == 32, "line words must be 0 (auto), 8, 16 or 32" );
                                       ctx->optLineWords = value; break;
      default: return dpcu::fail( DPCU_ERR_INVALID_VALUE, "dpcuCullSetOption: unknown option %d", option );
    }
    return DPCU_OK;
  }

  int dpcuCullGetOption( const dpcuCull *ctx, int option, int *value )
  {
    DPCU_REQUIRE( ctx && value, "NULL argument" );
    switch ( option )
    {
      case DPCU_CULL_OPT_KERNEL:       *value = ctx->optKernel; break;
      case DPCU_CULL_OPT_FMA:          *value = ctx->optFma; break;
      case DPCU_CULL_OPT_CHANGED_LIST: *value = ctx->optChanged; break;
      case DPCU_CULL_OPT_CTAS_PER_SM:  *value = ctx->optCtasPerSm; break;
      case DPCU_CULL_OPT_PROFILE:      *value = ctx->optProfile; break;
      case DPCU_CULL_OPT_FUSE_LEAF:    *value = ctx->optFuseLeaf; break;
      case DPCU_CULL_OPT_FUSE_LIST:    *value = ctx->optFuseList; break;
      case DPCU_CULL_OPT_LAST_KERNEL:  *value = ctx->lastKernel; break;
      case DPCU_CULL_OPT_FILTER:       *value = ctx->optFilter; break;
      case DPCU_CULL_OPT_LINE_WORDS:   *value = ctx->optLineWords; break;
      default: return dpcu::fail( DPCU_ERR_INVALID_VALUE, "dpcuCullGetOption: unknown option %d", option );
    }
    return DPCU_OK;
  }

  int dpcuCullGetKernelTime( dpcuCull *ctx, double *totalMs, uint64_t *launches )
  {
    DPCU_REQUIRE( ctx && totalMs && launches, "NULL argument" );
    dpcu::DeviceGuard guard( ctx->device );
    double sum = 0.0;
    for ( size_t i = 0; i + 1 < ctx->profUsed; i += 2 )
    {
      DPCU_CUDA( cudaEventSynchronize( ctx->profEvents[i + 1] ) );
      float ms = 0.f;
      DPCU_CUDA( cudaEventElapsedTime( &ms, ctx->profEvents[i], ctx->profEvents[i + 1] ) );
      sum += ms;
    }
    *totalMs = sum;
    *launches = ctx->profUsed / 2;
    ctx->profUsed = 0;
    return DPCU_OK;
  }

  int dpcuCullGetKernelTimes( dpcuCull *ctx, float *perLaunchMs, size_t capacity, size_t *launches )
  {
    DPCU_REQUIRE( ctx && launches && ( perLaunchMs || !capacity ), "NULL argument" );
    dpcu::DeviceGuard guard( ctx->device );
    size_t n = 0;
    for ( size_t i = 0; i + 1 < ctx->profUsed; i += 2, ++n )
    {
      DPCU_CUDA( cudaEventSynchronize( ctx->profEvents[i + 1] ) );
      float ms = 0.f;
      DPCU_CUDA( cudaEventElapsedTime( &ms, ctx->profEvents[i], ctx->profEvents[i + 1] ) );
      if ( n < capacity ) perLaunchMs[n] = ms;
    }
    *launches = n;
    ctx->profUsed = 0;
    return DPCU_OK;
  }

  int dpcuDebugKernelArgLayout( int nViews, size_t *onePairOffset, size_t *viewProjectionOffset, size_t *filterOffset, size_t *totalBytes )
  {
    DPCU_REQUIRE( nViews >= 1 && nViews <= DPCU_MAX_VIEWS, "nViews must be 1..DPCU_MAX_VIEWS" );
    size_t one = 0, vp = 0, filter = 0, total = 0;
    switch ( nViews )
    {
#define DPCU_LAYOUT( NV ) case NV: one = offsetof( dpcu::CullArgs<NV>, onePair ); vp = offsetof( dpcu::CullArgs<NV>, vp ); \
                                   filter = offsetof( dpcu::CullArgs<NV>, filter ); total = sizeof( dpcu::CullArgs<NV> ); break;
      DPCU_LAYOUT( 1 ) DPCU_LAYOUT( 2 ) DPCU_LAYOUT( 3 ) DPCU_LAYOUT( 4 ) DPCU_LAYOUT( 5 ) DPCU_LAYOUT( 6 ) DPCU_LAYOUT( 7 ) DPCU_LAYOUT( 8 )
#undef DPCU_LAYOUT
    }
    if ( onePairOffset ) *onePairOffset = one;
    if ( viewProjectionOffset ) *viewProjectionOffset = vp;
    if ( filterOffset ) *filterOffset = filter;
    if ( totalBytes ) *totalBytes = total;
    return DPCU_OK;
  }

  int dpcuCullGetLaunchCount( const dpcuCull *ctx, uint64_t *launches )
  {
    DPCU_REQUIRE( ctx && launches, "NULL argument" );
    *launches = ctx->launches;
    return DPCU_OK;
  }
}
#endif   // !DPCU_FMA_VARIANT
