// K2, line-granular form for SEVERAL views: the structure of cullLinesKernel (kernel_lines.cuh: a warp owns a
// 1024-object line of every view's bitset, whole-line stores to the bitset / host mirror / NVLink peers, ordered
// changed list by decoupled look-back) around a leaner multi-view step:
//
//   * the (object, view) pairs are classified two views per packed instruction (cull_filter_pairs.cuh), ~25 warp
//     instructions per pair instead of ~50;
//   * the pairs the filter cannot decide (0.9 % on the cube-map scene: 1.8 of the 192 pairs of a 32-object step)
//     are not evaluated step by step - a step's one or two pairs used to cost a whole 32-lane pass of the reference
//     arithmetic (~200 instructions, 0.75 passes per step).  They are QUEUED in shared memory across the steps of the
//     line (the object's OBB once, a 2-byte tag per pair) and evaluated 32 pairs per pass, lane <-> pair, when 32 have
//     piled up and at the end of the line: ~2.5 passes per 32 steps;
//   * the line's ballot words live in shared memory (acc[view][step]; the exact passes OR their bits in with shared
//     atomics), and the previous visibility words are read at the END of the line, so neither costs registers during
//     the steps: no local-memory traffic (the earlier form ran at 64 registers with 33 LDL + 20 STL);
//   * every object the filter does not handle - non-affine world matrix, NaN / Inf, a view with a non-finite or
//     extreme view-projection, the filter switched off - is just an undecided pair: there is ONE copy of the
//     reference arithmetic in the kernel (the general, non-affine form of cull_views.cuh);
//   * the transform index of the NEXT step is fetched a step ahead (one register), so a step's six 16-byte loads
//     go out together: one DRAM round trip per step instead of two dependent ones (object -> its matrix).
//
// Measured on the way and NOT in this file any more (64 Mi objects x 6 views, this form: 1.245 ms):
// L2 prefetch of the next steps with ONE prefetch per matrix (half of its sectors, see the step loop; the full-sector form
// is in for up to four views), per lane (prefetch.global.L2) or by the TMA unit (cp.async.bulk.prefetch.L2, SASS
// UBLKPF) 1.25 - 1.47 ms; the next step's loads staged in shared memory with cp.async (LDGSTS) at 3 CTAs per SM 1.57 ms
// (long_scoreboard 6.0 -> 2.6 but the MIO queue saturates: short_scoreboard 0.9 -> 3.2); all six loads a step ahead in
// registers 1.94 ms (spills), also with the OBB parked in shared memory during the classification; 3 / 5 CTAs per SM
// 1.32 / 1.74 ms; undecided pairs queued view by view (a branch per view) 1.38 ms.
#pragma once

namespace dpcu
{
#ifndef DPCU_MV_PREFETCH_MAX_VIEWS
#define DPCU_MV_PREFETCH_MAX_VIEWS 4     // up to this many views the next step's sectors are prefetched into L2 (see the step loop)
#endif
#ifndef DPCU_MV_MIN_CTAS
#define DPCU_MV_MIN_CTAS 4          // 64 registers, 32 warps per SM, 45 KB of shared memory per CTA
#endif

  constexpr uint32_t kMvObjCap   = 64;     // queued OBBs: flushed when more than 32 are waiting, a step adds at most 32
  constexpr uint32_t kMvFlushAt  = 32;     // pairs

  template <int NV>
  struct MvWarp
  {
    float4   obb[4][kMvObjCap];            // pt, ax, ay, az of the queued objects
    uint32_t acc[NV][32];                  // ballot word of step w for view v
    uint16_t tag[kMvFlushAt + 32 * NV];    // undecided pairs: object slot << 3 | view
    uint16_t pos[kMvObjCap];               // step << 5 | lane of the queued object
  };

  // The reference arithmetic for the queued pairs, 32 per pass (lane <-> pair).
  template <int NV>
  __device__ __noinline__ void mvFlush( MvWarp<NV> &sh, f32x2 const *sP, f32x2 one, uint32_t nPairs, uint32_t lane )
  {
    __syncwarp();
    for ( uint32_t base = 0; base < nPairs; base += 32 )
    {
      const bool     mine = base + lane < nPairs;
      const uint32_t tag  = mine ? sh.tag[base + lane] : 0u;
      const uint32_t view = tag & 7u, slot = tag >> 3;
      Obb o;
      o.pt = sh.obb[0][slot]; o.ax = sh.obb[1][slot]; o.ay = sh.obb[2][slot]; o.az = sh.obb[3][slot];
      const uint32_t pos = sh.pos[slot];
      ViewPairs P;
#pragma unroll
      for ( int k = 0; k < 8; ++k ) P.p[k] = sP[view * 8 + k];
      const bool visible = cornersVisible<true>( clipVectors<false>( broadcastObb( o ), P, one ) );
      if ( mine && visible ) atomicOr( &sh.acc[view][pos >> 5], 1u << ( pos & 31u ) );
    }
    __syncwarp();
  }

  template <int NV, bool kFuseList>
  __global__ void __launch_bounds__( kCullThreads, DPCU_MV_MIN_CTAS )
  cullLinesMvKernel( const __grid_constant__ CullArgs<NV> a )
  {
    constexpr int kPairs = ( NV + 1 ) / 2;
    const uint32_t lane   = threadIdx.x & 31u;
#define below ( ( 1u << lane ) - 1u )
    // (recomputed from the kernel argument where needed: as variables they were what ptxas spilled inside the step loop)
#define nWords ( ( a.n + 31u ) >> 5 )
#define nLines ( ( ( ( a.n + 31u ) >> 5 ) + 31u ) >> 5 )
    const uint32_t nWarps = gridDim.x * ( kCullThreads / 32 );
    uint32_t line = blockIdx.x * ( kCullThreads / 32 ) + ( threadIdx.x >> 5 );
    uint32_t pending = kNoLine;
    __shared__ f32x2 sP[NV * 8];
    __shared__ MvWarp<NV> sWarp[kCullThreads / 32];
    MvWarp<NV> &sh = sWarp[threadIdx.x >> 5];
    fillViewTable<NV>( sP, a );
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
      const uint32_t word0 = line << 5, myWord = word0 + lane;
      const bool     wordLive = myWord < nWords;
      const uint32_t steps = min( 32u, nWords - word0 );
      uint32_t nObj = 0, nPairs = 0;                         // queue fill (warp-uniform)
      // The transform index of the NEXT step is fetched a step ahead (one register), so a step's six 16-byte loads
      // go out together: one DRAM round trip per step instead of two dependent ones (object -> its matrix), which
      // were 40 % of all stall samples of this kernel (first ncu capture of this kernel).
      uint32_t idxNext = 0;
      {
        const uint32_t i0 = min( ( word0 << 5 ) + lane, a.n - 1u );
        idxNext = __ldg( reinterpret_cast<uint32_t const *>( a.lowerIdx + i0 ) + 3 );
      }
      // Up to four views the step leaves issue slots free and the kernel waits on DRAM: every 32-byte sector of the NEXT
      // step's matrices and extents is then asked for a step ahead (prefetch.global.L2 fetches ONE sector - a single
      // prefetch per matrix, as tried before, covered half of it and the L2 hit rate stayed at 36 %), with the index that
      // names the matrices running two steps ahead.  64 Mi objects, 2 / 3 / 4 views: 1.012 / 1.049 / 1.079 -> 0.985 / 1.024 /
      // 1.054 ms; with 5 / 6 views the extra instructions cost more than the shorter waits give (1.170 / 1.232 -> 1.177 /
      // 1.255 ms), so those keep the plain form.
      constexpr bool kPrefetch = NV <= DPCU_MV_PREFETCH_MAX_VIEWS;
      uint32_t idxNext2 = 0;
      if ( kPrefetch )
      {
        const uint32_t i1 = min( ( word0 << 5 ) + lane + 32u, a.n - 1u );
        idxNext2 = __ldg( reinterpret_cast<uint32_t const *>( a.lowerIdx + i1 ) + 3 );
      }
#pragma unroll 1
      for ( uint32_t w = 0; w < steps; ++w )
      {
        const uint32_t i    = ( ( word0 + w ) << 5 ) + lane;
        const bool     live = i < a.n;
        const uint32_t liveMask = __ballot_sync( 0xffffffffu, live );
        // lanes past the end re-read the last object (no branch, no zero fill); liveMask drops their results
        const uint32_t ic = min( i, a.n - 1u );
        const uint32_t tidx = idxNext;
        const float4 lo = ldStream( a.lowerIdx + ic );
        const float4 ex = ldStream( a.extent + ic );
        float4 const *m = a.mats + 4ull * tidx;
        float4 m0, m1, m2, m3;
        ldMatrix( m, m0, m1, m2, m3 );
        if ( kPrefetch )
        {
          // (the index load two steps ahead also pulls in that step's lowerIdx sectors)
          if ( w + 1 < steps && a.l2Prefetch )                    // DPCU_CULL_OPT_L2_PREFETCH
          {
            float4 const *mn = a.mats + 4ull * idxNext2;
            prefetchL2( mn );
            prefetchL2( mn + 2 );
            if ( ( lane & 1u ) == 0 ) prefetchL2( a.extent + min( i + 32u, a.n - 1u ) );
          }
          idxNext = idxNext2;
          if ( w + 2 < steps ) idxNext2 = __ldg( reinterpret_cast<uint32_t const *>( a.lowerIdx + min( i + 64u, a.n - 1u ) ) + 3 );
        }
        else if ( w + 1 < steps ) idxNext = __ldg( reinterpret_cast<uint32_t const *>( a.lowerIdx + min( i + 32u, a.n - 1u ) ) + 3 );
        Obb obb = makeObb( lo.x, lo.y, lo.z, ex.x, ex.y, ex.z, m0, m1, m2, m3 );
        const ObbBall ball = makeBall( obb, a.filterHalf );
        // All views in one straight-line block (the three pair classifications interleave freely), each lane
        // collecting its object's undecided views as a bit mask; ONE branch per step then queues them.
        uint32_t openMask = 0;
#pragma unroll
        for ( int p = 0; p < kPairs; ++p )
        {
          bool vis[2], inv[2];
          classifyPair( ball, a.pairFilter[p], vis, inv );
#pragma unroll
          for ( int e = 0; e < 2; ++e )
          {
            const int v = 2 * p + e;
            if ( v >= NV ) break;
            const uint32_t bv = __ballot_sync( 0xffffffffu, vis[e] ) & liveMask;
            if ( lane == 0 ) sh.acc[v][w] = bv;
            if ( !vis[e] && !inv[e] ) openMask |= 1u << v;
          }
        }
        if ( !live ) openMask = 0;
        if ( __any_sync( 0xffffffffu, openMask != 0 ) )
        {
          // object slots by ballot; tag slots by an exclusive prefix of the per-lane counts, one ballot per count bit
          const uint32_t cnt = __popc( openMask );
          const uint32_t bh  = __ballot_sync( 0xffffffffu, openMask != 0 );
          uint32_t pre = 0, tot = 0;
          constexpr int kCountBits = NV >= 8 ? 4 : ( NV >= 4 ? 3 : 2 );
#pragma unroll
          for ( int bit = 0; bit < kCountBits; ++bit )
          {
            const uint32_t bb = __ballot_sync( 0xffffffffu, ( cnt >> bit ) & 1u );
            pre += __popc( bb & below ) << bit;
            tot += __popc( bb ) << bit;
          }
          if ( openMask )
          {
            const uint32_t slot = nObj + __popc( bh & below );
            sh.obb[0][slot] = obb.pt; sh.obb[1][slot] = obb.ax; sh.obb[2][slot] = obb.ay; sh.obb[3][slot] = obb.az;
            sh.pos[slot] = uint16_t( ( w << 5 ) | lane );
            uint32_t t = nPairs + pre, m = openMask;
            do
            {
              sh.tag[t++] = uint16_t( ( slot << 3 ) | uint32_t( __ffs( m ) - 1 ) );
              m &= m - 1u;
            } while ( m );
          }
          nObj   += __popc( bh );
          nPairs += tot;
        }
        if ( nPairs >= kMvFlushAt || nObj > kMvObjCap - 32u )
        {
          mvFlush<NV>( sh, sP, a.onePair, nPairs, lane );
          nObj = nPairs = 0;
        }
      }
      // end of the line: previous words (their latency hides behind the last exact pass), then the finished words
      uint32_t old[NV];
#pragma unroll
      for ( int v = 0; v < NV; ++v ) old[v] = wordLive ? __ldcg( a.out[v].bits + myWord ) : 0u;
      if ( nPairs ) mvFlush<NV>( sh, sP, a.onePair, nPairs, lane );
      else __syncwarp();
#pragma unroll
      for ( int v = 0; v < NV; ++v )
      {
        ViewOut const &o = a.out[v];
        uint32_t flips = 0;
        if ( wordLive )
        {
          const uint32_t acc = sh.acc[v][lane];
          o.bits[myWord] = acc;
          if ( o.mirror ) o.mirror[myWord] = acc;              // the same line over PCIe into pinned host memory
          for ( uint32_t p = 0; p < a.nPeers; ++p )
          {
            if ( o.peer[p] ) o.peer[p][a.peerWordOffset + myWord] = acc;
          }
          if ( a.buildChanged )
          {
            const uint32_t c = old[v] ^ acc;
            o.chg[myWord] = c;
            flips = __popc( c );
          }
        }
        if ( a.buildChanged )
        {
#pragma unroll
          for ( int d = 16; d > 0; d >>= 1 ) flips += __shfl_xor_sync( 0xffffffffu, flips, d );
          if ( kFuseList )
          {
            // publish this line's count right away; its place in the list is resolved one line later (below)
            if ( lane == 0 ) lookStore( o.look + line, lookPack( o.epoch, line == 0 ? kLookPrefix : kLookAggregate, flips ) );
          }
          else if ( lane == 0 && flips ) atomicAdd( o.seg + ( word0 >> ( kSegObjectsLog2 - 5 ) ), flips );
        }
      }
      __syncwarp();                                          // acc[][] is rewritten by the next line's first step
      if ( kFuseList && a.buildChanged )
      {
        if ( pending != kNoLine ) resolveLine<NV>( a, pending, nLines, nWords, lane );
        pending = line;
      }
      if ( !kFuseList ) line += nWarps;
    }
    if ( kFuseList && a.buildChanged && pending != kNoLine ) resolveLine<NV>( a, pending, nLines, nWords, lane );
    if ( kFuseList ) rearmInLastCta( a.done );
    else if ( a.buildChanged ) scanSegmentsInLastCta<NV>( a.out, a.nSegs, a.done );
#undef nWords
#undef below
#undef nLines
  }
}
