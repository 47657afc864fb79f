// Kernel parameter blocks of the cull kernels (ViewOut, CullArgs), the shared-memory view table and the last-CTA
// scan of the per-segment changed counts.  Included by dpcu_cull.cu (and, through it, by the FMA variant).
#pragma once

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
    uint32_t *gridTotals;        // cullGridKernel: flips per CTA (the grid-wide prefix of phase 2)
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
    int           countSegs; // DPCU_CULL_OPT_LIST_OFFSETS: 1 = flips are counted per segment (atomics) and the last CTA scans
                             // the counters; 2 = counted, the compaction kernel sums them; 0 = not counted at all
    uint32_t      nSegs;
    uint32_t     *done;      // CTA completion ticket (last CTA scans the segment counters)
    uint32_t     *chunkCounter;   // staged kernel: next unclaimed chunk of kChunkTiles tiles
    int           vpFinite;  // every view-projection entry is finite (enables the affine shortcut of cull_views.cuh)
    unsigned long long onePair;   // (1.0f, 1.0f): runtime multiplier of cull_views.cuh::addProd
    uint32_t      lineWords; // line-granular kernel: bitset words per warp (32 = one 128-byte line; 8 for mid-size groups)
    int           useFilter; // cull_filter.cuh: decide provable (object, view) pairs from centre and radius
    uint32_t      nMats;     // matrices behind `mats` (bounds the speculative matrix prefetch of kernel_lines_mv.cuh)
    int           l2Prefetch; // line-granular kernels (1 view, 2-4 views): ask for the next step's sectors a step ahead (DPCU_CULL_OPT_L2_PREFETCH)
    ViewOut       out[NV];
    float4        vp[NV][4];
    ViewFilter    filter[NV];
    // cull_filter_pairs.cuh (line-granular multi-view kernel): the same filter, two views per packed instruction
    ViewPairFilter pairFilter[( NV + 1 ) / 2];
    float         filterHalf;   // 0.5f, as a filter constant (tests/test_sass.py)
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
}
