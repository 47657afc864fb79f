// K2, line-granular form: bitset lines, peer / host-mirror stores and the in-kernel changed list (look-back).
#pragma once

namespace dpcu
{
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
        // a predecessor in the window has not published yet: it is running on some other warp.  Back off (the waiting
        // warp must not take issue slots from the one it waits for) and poll again.  A predecessor can legitimately be
        // away for long - preemption under MPS / a debugger, a stalled PCIe or NVLink store - so the walk only gives up
        // after ~30 s of sleeping (2^25 polls of up to 1 us), where a trap is better than a silently hung device.
        ++polls;
        __nanosleep( polls < 8u ? 32u : ( polls < 64u ? 256u : 1024u ) );
        if ( polls > ( 1u << 25 ) ) __trap();
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
  // 6 views 2.068 ms at 3 CTAs (80 registers) -> 1.937 ms at 4 (64 registers, ~70 bytes of spills); the 2-view
  // instantiation is close to the HBM bound and prefers no spills: 1.154 ms at 4 CTAs, 1.103 ms at 3, 1.289 ms at 5
  template <int NV, bool kFuseList>
#ifndef DPCU_LINES1_MIN_CTAS
#define DPCU_LINES1_MIN_CTAS 5     // 47 registers without spills once the two look-ahead indices are live (6 CTAs: 40 registers, spills, no gain)
#endif
  __global__ void __launch_bounds__( kCullThreads, NV == 1 ? DPCU_LINES1_MIN_CTAS : ( NV == 2 ? 3 : 4 ) )
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
#ifndef DPCU_LINES1_AHEAD
#define DPCU_LINES1_AHEAD 1
#endif
      // One view (round 2): the scheme of the pair-filter kernel (kernel_lines_mv.cuh).  The transform index of THIS step
      // was fetched a step ago, so the step's loads go out together (one DRAM round trip instead of object -> matrix), and
      // the index of the NEXT step - fetched two steps ahead - turns into an L2 prefetch of every 32-byte sector of that
      // step's matrices and extents.
      constexpr bool kAhead1 = NV == 1 && DPCU_LINES1_AHEAD;
      uint32_t idxThis = 0, idxAfter = 0;
      if ( kAhead1 )
      {
        idxThis  = __ldg( reinterpret_cast<uint32_t const *>( a.lowerIdx + min( ( word0 << 5 ) + lane, a.n - 1u ) ) + 3 );
        idxAfter = __ldg( reinterpret_cast<uint32_t const *>( a.lowerIdx + min( ( word0 << 5 ) + lane + 32u, a.n - 1u ) ) + 3 );
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
        if ( kAhead1 && a.l2Prefetch && w + 1u < steps )               // DPCU_CULL_OPT_L2_PREFETCH
        {
          float4 const *mn = a.mats + 4ull * idxAfter;
          prefetchL2( mn );
          prefetchL2( mn + 2 );
          if ( ( lane & 1u ) == 0 ) prefetchL2( a.extent + min( i + 32u, a.n - 1u ) );
        }
        if ( live )
        {
          const float4 lo = ldStream( a.lowerIdx + i );
          const float4 ex = ldStream( a.extent + i );
          float4 const *m = a.mats + 4ull * ( kAhead1 ? idxThis : __float_as_uint( lo.w ) );
          float4 m0, m1, m2, m3;
          ldMatrix( m, m0, m1, m2, m3 );
          obb = makeObb( lo.x, lo.y, lo.z, ex.x, ex.y, ex.z, m0, m1, m2, m3 );
        }
        if ( kAhead1 )
        {
          idxThis = idxAfter;
          if ( w + 2u < steps ) idxAfter = __ldg( reinterpret_cast<uint32_t const *>( a.lowerIdx + min( i + 64u, a.n - 1u ) ) + 3 );
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
}
