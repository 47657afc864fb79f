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
//   * loads of the NEXT steps are put in flight without registers: one lane asks the TMA unit to pull the next
//     512-byte runs of lowerIdx / extent into L2 (cp.async.bulk.prefetch.L2), and when the step's transform indices
//     are consecutive (object i <-> matrix i, the common layout) the 2 KiB of matrices behind them as well.
#pragma once

namespace dpcu
{
#ifndef DPCU_MV_MIN_CTAS
#define DPCU_MV_MIN_CTAS ( DPCU_MV_RING ? 3 : 4 )
#endif
#ifndef DPCU_MV_PIPE
#define DPCU_MV_PIPE 1              // 0: loads as they come, 1: transform index a step ahead, 2: 1 + L2 prefetch of the next
#endif                              // step's matrix  (all six loads a step ahead - a register double buffer - spilled: 1.9 ms)
#ifndef DPCU_MV_RING
#define DPCU_MV_RING 0              // 1: the next step's six 16-byte loads per lane go to shared memory with cp.async (LDGSTS) while
#endif                              // this step is classified - bytes in flight without registers, 3 CTAs per SM with 70 KB of shared
                                    // memory each.  Measured: long_scoreboard 6.0 -> 2.6 warps per issue, but the 6 LDGSTS + 6 LDS.128
                                    // per lane and step saturate the MIO queue (short_scoreboard 0.9 -> 3.2, mio_throttle 1.0) and the
                                    // L1 shrinks to 19 KB: 1.57 ms instead of 1.245 ms at 64 Mi x 6 views.  Kept as an experiment.
#ifndef DPCU_MV_RECOMPUTE_W
#define DPCU_MV_RECOMPUTE_W 0       // the OBB's w components are not kept in registers across the classification
#endif
#ifndef DPCU_MV_APPEND
#define DPCU_MV_APPEND 1            // 0: undecided pairs queued view by view (a branch per view), 1: one branch per step
#endif
#ifndef DPCU_MV_PREFETCH
#define DPCU_MV_PREFETCH 0          // 0: none, 1: per-lane prefetch.global.L2, 2: bulk L2 prefetch by one lane (TMA unit)
#endif
#ifndef DPCU_MV_PREFETCH_DIST
#define DPCU_MV_PREFETCH_DIST 2     // steps ahead
#endif

  constexpr uint32_t kMvObjCap   = 64;     // queued OBBs: flushed when more than 32 are waiting, a step adds at most 32
  constexpr uint32_t kMvFlushAt  = 32;     // pairs

  template <int NV>
  struct MvWarp
  {
#if DPCU_MV_RING
    float4   stage[6][32];                 // the NEXT step's inputs, landing while this step is classified: lowerIdx, extent, 4 matrix rows
#endif
    float4   obb[4][kMvObjCap];            // pt, ax, ay, az of the queued objects
    uint32_t acc[NV][32];                  // ballot word of step w for view v
    uint16_t tag[kMvFlushAt + 32 * NV];    // undecided pairs: object slot << 3 | view
    uint16_t pos[kMvObjCap];               // step << 5 | lane of the queued object
  };

  template <int NV> __host__ __device__ constexpr size_t mvViewTableBytes() { return ( size_t( NV ) * 8 * sizeof( f32x2 ) + 15 ) & ~size_t( 15 ); }
  template <int NV> __host__ __device__ constexpr size_t mvSharedBytes()      // dynamic shared memory of cullLinesMvKernel
  {
    return DPCU_MV_RING ? mvViewTableBytes<NV>() + sizeof( MvWarp<NV> ) * ( kCullThreads / 32 ) : 0;
  }

  // 16 bytes global -> shared without a register in between (LDGSTS); .ca keeps the line in L1 like the __ldg it replaces
  // (two matrix rows share a 32-byte sector)
  __device__ __forceinline__ void copyAsync16( void *smem, void const *gmem )
  {
    asm volatile( "cp.async.ca.shared.global [%0], [%1], 16;" :: "r"( uint32_t( __cvta_generic_to_shared( smem ) ) ), "l"( gmem ) : "memory" );
  }
  __device__ __forceinline__ void copyAsyncWaitAll()
  {
    asm volatile( "cp.async.wait_all;" ::: "memory" );
  }

  __device__ __forceinline__ void bulkPrefetchL2( void const *p, uint32_t bytes )
  {
    asm volatile( "cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"( p ), "r"( bytes ) : "memory" );
  }

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

#if DPCU_MV_RING
  // object `i` (already clamped) with transform index `tidx`: its six 16-byte pieces into the lane's stage slots
  template <int NV>
  __device__ __forceinline__ void mvStageStep( MvWarp<NV> &sh, CullArgs<NV> const &a, uint32_t i, uint32_t tidx, uint32_t lane )
  {
    float4 const *m = a.mats + 4ull * tidx;
    copyAsync16( &sh.stage[0][lane], a.lowerIdx + i );
    copyAsync16( &sh.stage[1][lane], a.extent + i );
    copyAsync16( &sh.stage[2][lane], m + 0 );
    copyAsync16( &sh.stage[3][lane], m + 1 );
    copyAsync16( &sh.stage[4][lane], m + 2 );
    copyAsync16( &sh.stage[5][lane], m + 3 );
  }
#endif

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
#if DPCU_MV_RING
    // dynamic shared memory (more than the 48 KB a static allocation may have): the view table, then one MvWarp per warp
    extern __shared__ __align__( 16 ) unsigned char sMvRaw[];
    f32x2 *sP = reinterpret_cast<f32x2 *>( sMvRaw );
    MvWarp<NV> &sh = reinterpret_cast<MvWarp<NV> *>( sMvRaw + mvViewTableBytes<NV>() )[threadIdx.x >> 5];
#else
    __shared__ f32x2 sP[NV * 8];
    __shared__ MvWarp<NV> sWarp[kCullThreads / 32];
    MvWarp<NV> &sh = sWarp[threadIdx.x >> 5];
#endif
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
#if DPCU_MV_PREFETCH
      {
        // the first steps of the line
        const uint32_t i0 = word0 << 5;
        const uint32_t cnt = min( a.n - i0, 32u * DPCU_MV_PREFETCH_DIST );
#if DPCU_MV_PREFETCH == 2
        if ( lane == 0 )
        {
          bulkPrefetchL2( a.lowerIdx + i0, cnt * 16u );
          bulkPrefetchL2( a.extent + i0, cnt * 16u );
        }
#else
        if ( lane < 8u * DPCU_MV_PREFETCH_DIST && lane * 8u < cnt ) prefetchL2( ( ( lane & 1u ) ? a.extent : a.lowerIdx ) + i0 + ( lane >> 1 ) * 8u );
#endif
      }
#endif
#if DPCU_MV_PIPE >= 1
      // The transform index of the NEXT step is fetched a step ahead (one register), so a step's six 16-byte loads
      // go out together: one DRAM round trip per step instead of two dependent ones (object -> its matrix), which
      // were 40 % of all stall samples of this kernel (profiles/r02_mv6_nopipe_*).
      uint32_t idxNext = 0;
      {
        const uint32_t i0 = min( ( word0 << 5 ) + lane, a.n - 1u );
        idxNext = __ldg( reinterpret_cast<uint32_t const *>( a.lowerIdx + i0 ) + 3 );
#if DPCU_MV_RING
        mvStageStep( sh, a, i0, idxNext, lane );               // step 0 of the line (two dependent round trips, once per line)
        if ( steps > 1 ) idxNext = __ldg( reinterpret_cast<uint32_t const *>( a.lowerIdx + min( i0 + 32u, a.n - 1u ) ) + 3 );
#endif
      }
#endif
#pragma unroll 1
      for ( uint32_t w = 0; w < steps; ++w )
      {
        const uint32_t i    = ( ( word0 + w ) << 5 ) + lane;
        const bool     live = i < a.n;
        const uint32_t liveMask = __ballot_sync( 0xffffffffu, live );
        // lanes past the end re-read the last object (no branch, no zero fill); liveMask drops their results
        const uint32_t ic = min( i, a.n - 1u );
#if DPCU_MV_RING
        // this step's inputs were copied into the lane's own stage slots during the previous step (or just now, for
        // the first step of the line): wait for the lane's copies, read them, and send the NEXT step's on their way
        copyAsyncWaitAll();
        const float4 lo = sh.stage[0][lane], ex = sh.stage[1][lane];
        const float4 m0 = sh.stage[2][lane], m1 = sh.stage[3][lane], m2 = sh.stage[4][lane], m3 = sh.stage[5][lane];
        const uint32_t tidx = __float_as_uint( lo.w );
#else
#if DPCU_MV_PIPE >= 1
        const uint32_t tidx = idxNext;
#endif
        const float4 lo = ldStream( a.lowerIdx + ic );
        const float4 ex = ldStream( a.extent + ic );
#if DPCU_MV_PIPE == 0
        const uint32_t tidx = __float_as_uint( lo.w );
#endif
        float4 const *m = a.mats + 4ull * tidx;
        const float4 m0 = __ldg( m + 0 );
        const float4 m1 = __ldg( m + 1 );
        const float4 m2 = __ldg( m + 2 );
        const float4 m3 = __ldg( m + 3 );
#if DPCU_MV_PIPE >= 1
        if ( w + 1 < steps ) idxNext = __ldg( reinterpret_cast<uint32_t const *>( a.lowerIdx + min( i + 32u, a.n - 1u ) ) + 3 );
#endif
#endif
        Obb obb = makeObb( lo.x, lo.y, lo.z, ex.x, ex.y, ex.z, m0, m1, m2, m3 );
#if DPCU_MV_RING
        if ( w + 1 < steps )
        {
          // (the stage slots were read into registers and consumed by makeObb above; asm volatile keeps the order)
          mvStageStep( sh, a, min( i + 32u, a.n - 1u ), idxNext, lane );
          if ( w + 2 < steps ) idxNext = __ldg( reinterpret_cast<uint32_t const *>( a.lowerIdx + min( i + 64u, a.n - 1u ) ) + 3 );
        }
#endif
#if DPCU_MV_PREFETCH
        {
          // step w + DIST of this line: object runs always; matrices when this step's indices are consecutive
          // (then the next steps' very likely continue the run; a wrong guess costs one useless L2 fill)
          const uint32_t wp = w + DPCU_MV_PREFETCH_DIST;
          const uint32_t ip = ( ( word0 + wp ) << 5 );
          const uint32_t t0 = __shfl_sync( 0xffffffffu, tidx, 0 );
          const bool     run = __all_sync( 0xffffffffu, tidx == t0 + lane );
          if ( wp < steps && ip + 32u <= a.n )
          {
#if DPCU_MV_PREFETCH == 2
            if ( lane == 0 )
            {
              bulkPrefetchL2( a.lowerIdx + ip, 512u );
              bulkPrefetchL2( a.extent + ip, 512u );
              if ( run && t0 + 32u * DPCU_MV_PREFETCH_DIST + 32u <= a.nMats )
                bulkPrefetchL2( a.mats + 4ull * ( t0 + 32u * DPCU_MV_PREFETCH_DIST ), 2048u );
            }
#else
            if ( lane < 8u ) prefetchL2( ( ( lane & 1u ) ? a.extent : a.lowerIdx ) + ip + ( lane >> 1 ) * 8u );
            else if ( lane < 24u && run && t0 + 32u * DPCU_MV_PREFETCH_DIST + 32u <= a.nMats )
              prefetchL2( a.mats + 4ull * ( t0 + 32u * DPCU_MV_PREFETCH_DIST ) + ( lane - 8u ) * 8u );
#endif
          }
        }
#endif
        const ObbBall ball = makeBall( obb, a.filterHalf );
#if DPCU_MV_APPEND == 0
        uint32_t mySlot = 0xffffffffu;
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
            const bool     open = !vis[e] && !inv[e];
            const uint32_t bo = __ballot_sync( 0xffffffffu, open ) & liveMask;
            if ( lane == 0 ) sh.acc[v][w] = bv;
            if ( bo )
            {
              // queue the undecided pairs of this view: the object's OBB once (first view that needs it), a tag per pair
              const bool     mine  = ( bo >> lane ) & 1u;
              const bool     fresh = mine && mySlot == 0xffffffffu;
              const uint32_t bf    = __ballot_sync( 0xffffffffu, fresh );
              if ( fresh )
              {
                mySlot = nObj + __popc( bf & below );
                sh.obb[0][mySlot] = obb.pt; sh.obb[1][mySlot] = obb.ax; sh.obb[2][mySlot] = obb.ay; sh.obb[3][mySlot] = obb.az;
                sh.pos[mySlot] = uint16_t( ( w << 5 ) | lane );
              }
              nObj += __popc( bf );
              if ( mine ) sh.tag[nPairs + __popc( bo & below )] = uint16_t( ( mySlot << 3 ) | uint32_t( v ) );
              nPairs += __popc( bo );
            }
          }
        }
#else
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
            // The w components of the OBB do not stay in registers across the classification (4 registers that made
            // the 6-view instantiation spill): an object the filter handles is affine - they are exactly 1, 0, 0, 0 -
            // and for the others (aw = +inf: projective, NaN / Inf or huge) they are computed again from the inputs,
            // with the operations of makeObb.
#if DPCU_MV_RECOMPUTE_W
            float4 wv = make_float4( 1.0f, 0.0f, 0.0f, 0.0f );
            if ( !( ball.aw < __int_as_float( 0x7f800000 ) ) )
            {
              const float4 lo = ldStream( a.lowerIdx + ic );
              const float4 ex = ldStream( a.extent + ic );
              float4 const *m = a.mats + 4ull * __float_as_uint( lo.w );
              const float w0 = __ldg( &m[0].w ), w1 = __ldg( &m[1].w ), w2 = __ldg( &m[2].w ), w3 = __ldg( &m[3].w );
              wv.x = ( ( lo.x * w0 + lo.y * w1 ) + lo.z * w2 ) + w3;
              wv.y = w0 * ex.x;
              wv.z = w1 * ex.y;
              wv.w = w2 * ex.z;
            }
            obb.pt.w = wv.x; obb.ax.w = wv.y; obb.ay.w = wv.z; obb.az.w = wv.w;
#endif
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
#endif
#if DPCU_MV_PIPE == 2
        if ( w + 1 < steps && i + 32u < a.n ) prefetchL2( a.mats + 4ull * idxNext );
#endif
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
