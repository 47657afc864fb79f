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
//                changed[n] u32 (ascending group indices), seg[] / prefix[] u32 changed-counts per
//                8192 objects and their exclusive prefix (prefix[nSegs] = total), look[] u64 look-back
//                entries per 1024-object line.
//
// One cull = ONE launch of the line-granular kernel for large groups (bitset lines, peer / host-mirror
// stores and the ordered changed list in one pass), otherwise cull kernel (ballot -> word, XOR with the
// previous word, popc into the segment counters) -> compaction kernel, launched as a programmatic dependent of
// the cull kernel: each of its CTAs sums the counters before its segment (groups beyond 32 Mi objects: the cull
// kernel's last CTA scans them instead), expands its segment's flipped bits into the list and stores the
// segment's part of a host mirror (DPCU_CULL_OPT_LIST_OFFSETS).
// The kernels live in kernel_*.cuh (arguments, direct, views, staged, lines, fused leaf, misc), their
// arithmetic in cull_math.cuh / cull_views.cuh / cull_filter.cuh; this file is the host side.
//
// This translation unit is compiled with -fmad=false (exact mode).  dpcu_cull_fma.cu includes
// it again with DPCU_FMA_VARIANT defined and -fmad=true to provide the reporting-only fast mode.
#include "cull_math.cuh"
#include "cull_stage.cuh"
#include "cull_views.cuh"
#include "cull_filter.cuh"
#include "cull_filter_pairs.cuh"
#include "tree_propagate.cuh"
#include "dpcu_internal.h"
#include "dpcu_tree.h"

#include <cmath>
#include <cstddef>
#include <new>
#include <vector>

#include "kernel_args.cuh"
#include "kernel_direct.cuh"
#ifndef DPCU_FMA_VARIANT
#include "kernel_views.cuh"
#include "kernel_staged.cuh"
#include "kernel_lines.cuh"
#include "kernel_lines_mv.cuh"
#include "kernel_grid.cuh"
#include "kernel_fused_leaf.cuh"
#include "kernel_misc.cuh"
#endif   // !DPCU_FMA_VARIANT

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
  dpcu::DeviceArray gridTotals;                     // cullGridKernel: flips per CTA
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
  dpcuTree    *boundTree = nullptr;      // ... of this tree (dpcuCullBindTree / dpcuCullRunWithTree): culls and its computes are ordered by events
  size_t       n = 0, nMats = 0;
  int          occupancy[11][DPCU_MAX_VIEWS] = {};   // resident CTAs per SM by kernel form and view count (0 = not asked yet)
  uint32_t     maxTransformIndex = 0;
  bool         maxIndexKnown = true;
  bool         maxIndexStale = false;    // objects were overwritten / removed since the running maximum was started
  dpcu::StreamFence stagingFree;         // the last upload out of `staging` (the next user waits for it, not the caller)
  int          optKernel = 0, optFma = 0, optChanged = 1, optCtasPerSm = 0, optProfile = 0, optFuseLeaf = 1, optFuseList = 1, optFilter = 1, optLineWords = 0, optListOffsets = 0, optL2Prefetch = 1;
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

  // every object's transform index must address a matrix (the reference would read out of bounds); the running
  // maximum is exact until objects are overwritten or removed, then it is recomputed before it may reject a cull
  static int checkIndexRange( dpcuCull *ctx, char const *who )
  {
    if ( !ctx->n ) return DPCU_OK;
    DPCU_TRY( refreshMaxIndex( ctx ) );
    if ( ctx->maxTransformIndex >= ctx->nMats && ctx->maxIndexStale )
    {
      DPCU_CUDA( cudaMemsetAsync( ctx->maxIndex.ptr, 0, 4, ctx->stream ) );
      int grid = int( divUp( ctx->n, 256 ) );
      if ( grid > ctx->smCount * 8 ) grid = ctx->smCount * 8;
      maxIndexKernel<<<grid, 256, 0, ctx->stream>>>( static_cast<float4 const *>( ctx->lowerIdx.ptr ), uint32_t( ctx->n ), static_cast<uint32_t *>( ctx->maxIndex.ptr ) );
      DPCU_CUDA( cudaGetLastError() );
      ++ctx->launches;
      ctx->maxIndexKnown = false;
      ctx->maxIndexStale = false;
      DPCU_TRY( refreshMaxIndex( ctx ) );
    }
    if ( ctx->maxTransformIndex >= ctx->nMats )
      return fail( DPCU_ERR_INVALID_VALUE, "%s: transform index %u out of range (%zu matrices)", who, ctx->maxTransformIndex, ctx->nMats );
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
    if ( ctx->optChanged && !r->gridTotals.ptr ) DPCU_TRY( r->gridTotals.reserve( ( size_t( ctx->smCount ) * 32 + 64 ) * 4, false, stream ) );
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
    double q3 = 0.0;
    for ( int r = 0; r < 3; ++r ) q3 += fabs( P[r][0] ) + fabs( P[r][1] ) + fabs( P[r][2] ) + fabs( P[r][3] );
    const double qw = fabs( P[3][0] ) + fabs( P[3][1] ) + fabs( P[3][2] ) + fabs( P[3][3] );
    const double q = q3 + qw;
    // q = +inf switches the filter off for this view (classify: `sane` is false): a non-finite entry, or a sum so small
    // that the relative margin would not cover the ABSOLUTE error of underflowing products (each up to 2^-150; with the
    // sum >= 2^-100 the margin is >= 2^-117), or so large that the margin arithmetic could overflow
    bool finite = true;
    for ( int k = 0; k < 16; ++k ) finite = finite && std::isfinite( vp[k] );
    const bool usable = finite && q >= 7.888609052210118e-31 /* 2^-100 */ && q <= 549755813888.0 /* 2^39 */;
    f.q  = usable ? roundUp( marginScale * q3 / 131072.0 ) : INFINITY;
    f.qw = usable ? roundUp( marginScale * qw / 131072.0 ) : 0.0f;
    memcpy( f.rows, vp, 64 );
  }

  // ViewPairFilter entry `half` (0: view u, 1: view v) of one view-projection (cull_filter_pairs.cuh).  `enabled`
  // false, a non-finite entry, or sum sum |P| outside [2^-100, 2^39] => q = +inf: the filter never decides this view
  // (below 2^-100 the margin would not cover the absolute error of underflowing products).
  static float roundAway( double x )
  {
    float f = static_cast<float>( x );
    return ( static_cast<double>( f ) > x ) ? nextafterf( f, -INFINITY ) : f;      // x <= 0: towards -inf
  }
  static void fillViewPairFilter( float const *vp, ViewPairFilter &f, int half, bool enabled, double marginScale )
  {
    double P[4][4];
    bool finite = true;
    for ( int r = 0; r < 4; ++r ) for ( int c = 0; c < 4; ++c )
    {
      P[r][c] = vp[4 * r + c];
      finite = finite && std::isfinite( vp[4 * r + c] );
      ( half ? f.k[4 * r + c].y : f.k[4 * r + c].x ) = vp[4 * r + c];
    }
    const double inflate = 1.0 + 1.0 / 1048576.0;
    double q3 = 0.0;
    for ( int r = 0; r < 3; ++r ) q3 += fabs( P[r][0] ) + fabs( P[r][1] ) + fabs( P[r][2] ) + fabs( P[r][3] );
    const double qw = fabs( P[3][0] ) + fabs( P[3][1] ) + fabs( P[3][2] ) + fabs( P[3][3] );
    const double q = q3 + qw;
    const bool usable = enabled && finite && q >= 7.888609052210118e-31 /* 2^-100 */ && q <= 549755813888.0 /* 2^39 */;
    for ( int a = 0; a < 3; ++a )
    {
      double nN[3], nP[3];
      for ( int r = 0; r < 3; ++r )
      {
        nN[r] = P[r][a] + P[r][3];          // N plane: x + w
        nP[r] = P[r][3] - P[r][a];          // P plane: w - x
      }
      const float rn = usable ? roundUp( sqrt( nN[0] * nN[0] + nN[1] * nN[1] + nN[2] * nN[2] ) * inflate ) : 0.0f;
      const float rp = usable ? roundAway( -sqrt( nP[0] * nP[0] + nP[1] * nP[1] + nP[2] * nP[2] ) * inflate ) : 0.0f;
      ( half ? f.rhoN[a].y : f.rhoN[a].x ) = rn;
      ( half ? f.nrhoP[a].y : f.nrhoP[a].x ) = rp;
    }
    const float rw = usable ? roundAway( -sqrt( P[0][3] * P[0][3] + P[1][3] * P[1][3] + P[2][3] * P[2][3] ) * inflate ) : 0.0f;
    ( half ? f.nrhoW.y : f.nrhoW.x ) = rw;
    ( half ? f.q.y : f.q.x )   = usable ? roundUp( marginScale * q3 / 131072.0 ) : INFINITY;
    ( half ? f.qw.y : f.qw.x ) = usable ? roundUp( marginScale * qw / 131072.0 ) : 0.0f;
  }

  constexpr int kAutoListOffsets = 3;

  template <int NV>
  static int launchCull( dpcuCull *ctx, dpcuCullResult *const *results, float const *vps, cudaStream_t stream, LeafArgs const *leaf,
                         bool *mirrorsWritten, bool *listBuilt, int *listOffsets )
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
      args.out[v].gridTotals = static_cast<uint32_t *>( r->gridTotals.ptr );
      for ( int p = 0; p < kMaxPeers; ++p ) args.out[v].peer[p] = p < r->nPeers ? r->peer[p] : nullptr;
      if ( uint32_t( r->nPeers ) > args.nPeers ) args.nPeers = uint32_t( r->nPeers );
      args.peerWordOffset = uint32_t( r->peerWordOffset );
      memcpy( args.vp[v], vps + 16 * v, 64 );
      // DPCU_CULL_OPT_FILTER 2 / 3 shrink the margin to 1/8 (the bound of the error analysis itself) / to zero:
      // diagnostics that measure the slack of the proof, never for production
      if ( NV > 1 ) makeViewFilter( vps + 16 * v, args.filter[v], ctx->optFilter == 2 ? 0.125 : ctx->optFilter == 3 ? 0.0 : 1.0 );
    }
    args.useFilter = ( NV > 1 && ctx->optFilter ) ? 1 : 0;
    args.nMats      = uint32_t( ctx->nMats );
    args.l2Prefetch = ctx->optL2Prefetch;
    args.filterHalf = 0.5f;
    if ( NV > 1 )
    {
      const double scale = ctx->optFilter == 2 ? 0.125 : ctx->optFilter == 3 ? 0.0 : 1.0;
      for ( int v = 0; v < 2 * ( ( NV + 1 ) / 2 ); ++v )
      {
        // an odd view count pads the last pair with a copy of the last view (its result is dropped)
        fillViewPairFilter( vps + 16 * ( v < NV ? v : NV - 1 ), args.pairFilter[v / 2], v & 1, ctx->optFilter != 0, scale );
      }
    }
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
    // (round 2, pair-filter form against views + compaction, step time with the L2 flushed: 6 views 1.5 Mi objects 79 vs 77 us,
    // 2 Mi 84 vs 95 us, 3 Mi 100 vs 128 us, 4 Mi 116 vs 157 us; 2 views 2 Mi 63 vs 62 us, 3 Mi 77 vs 79 us, 4 Mi 94 vs 97 us)
    // (round 2, after the compaction kernel took over the segment counters' prefix - DPCU_CULL_OPT_LIST_OFFSETS - the one thread
    // per object forms gained 2-14 us per step and the crossovers moved up; step time, L2 flushed, lines vs views / direct +
    // compaction: 6 views 2 Mi 79 vs 75 us, 3 Mi 97 vs 104 us; 4 views 3 Mi 83 vs 83 us, 4 Mi 97 vs 108 us; 3 views 3 Mi 79 vs
    // 75 us, 4 Mi 93 vs 95 us; 2 views 6 Mi 130 vs 120 us, 8 Mi 146 vs 155 us; 1 view 24 Mi 398 vs 388 us, 32 Mi 519 vs 515 us,
    // 40 Mi 641 vs 642 us, 64 Mi 1.011 vs 1.013 ms)
    // (one view again, after the line-granular kernel got its index look-ahead and full-sector L2 prefetch: lines vs direct +
    // compaction 4 Mi 85 vs 77 us, 8 Mi 140 vs 146 us, 16 Mi 252 vs 264 us, 32 Mi 480 vs 509 us, 64 Mi 0.979 vs 1.013 ms)
    const size_t wantedLines = size_t( ctx->smCount ) * ( NV == 1 ? 44u : NV == 2 ? 46u : NV == 3 ? 24u : NV == 4 ? 20u : 17u );
    const bool bigEnough = divUp( divUp( ctx->n, 32 ), 32 ) >= wantedLines;
    const bool autoLines = ctx->optKernel == DPCU_KERNEL_AUTO && !leaf && bigEnough
                        && ( mirrors || ( ctx->optChanged && ctx->optFuseList ) );
    args.lineWords = ctx->optLineWords ? uint32_t( ctx->optLineWords ) : 32u;
    // A resident warp that only sees four or five whole lines leaves a ragged tail (each line is 32 sequential steps):
    // half lines fill it in.  32 Mi objects, one view: 0.518 -> 0.508 ms (this is the per-GPU share of BASELINE C5 on 8 GPUs,
    // where peer bitsets make this kernel the only choice); at 64 Mi objects whole lines are as fast and store 128 bytes.
    if ( !ctx->optLineWords && NV == 1 && divUp( divUp( ctx->n, 32 ), 32 ) < size_t( ctx->smCount ) * 48u * 6u ) args.lineWords = 16u;
    const bool useLines  = !ctx->optFma && ( ( peers && !leaf ) || ctx->optKernel == DPCU_KERNEL_LINES || ctx->optKernel == DPCU_KERNEL_LINES_PAIRS || autoLines );
    // several views: the pair-filter form with the queued exact passes (kernel_lines_mv.cuh), unless the earlier form is asked for
    const bool useLinesMv = useLines && !leaf && NV >= 2 && ctx->optKernel != DPCU_KERNEL_LINES;
    if ( useLinesMv ) args.lineWords = 32u;
    // One thread per object on a co-resident grid with one grid-wide barrier between the cull and the list
    // (kernel_grid.cuh): one launch per cull for small groups.  Built for BASELINE config C2 and measured there against
    // direct + compaction (device time of the step, L2 flushed and swept clean): 37.8 us vs 36.0 us at 1 Mi objects, 83 vs
    // 53 us at 2 Mi - the barrier makes every CTA wait for the slowest one and phase 2 starts cold, which costs more than
    // the second launch it saves (that launch now overlaps the cull kernel's tail, see the compaction launch below).
    // So AUTO does not pick it; DPCU_KERNEL_GRID does.
    const bool useGrid = !useLines && !leaf && !ctx->optFma && ctx->optKernel == DPCU_KERNEL_GRID;
    *mirrorsWritten = ( useLines || useGrid ) && !leaf;
    const bool fuseList = ( ( useLines && ctx->optFuseList ) || useGrid ) && !leaf && ctx->optChanged;
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
    if ( peers && ctx->optFma )
      return fail( DPCU_ERR_INVALID_VALUE, "dpcuCullRun: peer bitsets are not served by the FMA reporting mode" );
    const bool useFused  = leaf != nullptr;
    const bool useStaged = !useFused && !useLines && !useGrid && !ctx->optFma && ctx->optKernel == DPCU_KERNEL_STAGED;
    const bool useChains = ctx->optKernel == DPCU_KERNEL_VIEWS_CHAINS;
    const bool useViews  = !useFused && !useLines && !useGrid && !ctx->optFma && ( ctx->optKernel == DPCU_KERNEL_VIEWS || useChains || ( ctx->optKernel == DPCU_KERNEL_AUTO && NV >= 2 ) );
    args.chunkCounter = results[0]->donePtr() + 1;
    // one thread per object forms (direct, views, fused leaf): who turns the flipped bits into list offsets
    // (DPCU_CULL_OPT_LIST_OFFSETS; *listOffsets = 1: last-CTA scan, 2: compaction popcounts words, 3: compaction sums counters)
    *listOffsets = 1;
    if ( !useLines && !useGrid && !useStaged && ctx->optChanged )
    {
      int want = ctx->optListOffsets ? ctx->optListOffsets : kAutoListOffsets;
      if ( want == 2 && args.nSegs > 256u ) want = 3;
      if ( want == 3 && args.nSegs > 4096u ) want = 1;
      *listOffsets = want;
    }
    args.countSegs = *listOffsets == 2 ? 0 : *listOffsets == 3 ? 2 : 1;
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
    // the occupancy query costs a microsecond or two of every cull on the host: asked once per kernel form and view count
    const int form = useFused ? 0 : ( useLinesMv && fuseList ) ? 1 : useLinesMv ? 2 : ( useLines && fuseList ) ? 3 : useLines ? 4 : useGrid ? 5 : ctx->optFma ? 6
                   : useStaged ? 7 : ( useViews && useChains ) ? 8 : useViews ? 9 : 10;
    if ( perSm <= 0 ) perSm = ctx->occupancy[form][NV - 1];
    if ( perSm <= 0 )
    {
      if ( useFused ) cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, cullFusedLeafKernel<NV>, kCullThreads, 0 );
      else if ( useLinesMv && fuseList ) cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, cullLinesMvKernel<NV, true>, kCullThreads, 0 );
      else if ( useLinesMv ) cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, cullLinesMvKernel<NV, false>, kCullThreads, 0 );
      else if ( useLines && fuseList ) cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, cullLinesKernel<NV, true>, kCullThreads, 0 );
      else if ( useLines ) cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, cullLinesKernel<NV, false>, kCullThreads, 0 );
      else if ( useGrid ) cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, cullGridKernel<NV>, kCullThreads, 0 );
      else if ( ctx->optFma ) perSm = occupancyCullDirectFma<NV>();
      else if ( useStaged ) cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, cullStagedKernel<NV>, kCullThreads, stagedSmem );
      else if ( useViews && useChains ) cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, cullViewsKernel<NV, false>, kCullThreads, 0 );
      else if ( useViews ) cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, cullViewsKernel<NV, true>, kCullThreads, 0 );
      else cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, cullDirectKernel<NV>, kCullThreads, 0 );
      if ( perSm <= 0 ) perSm = 1;
      ctx->occupancy[form][NV - 1] = perSm;
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
      // behind the tree's upper levels on the same stream: a programmatic dependent launch like theirs (dpcu_tree.cu)
      cudaLaunchConfig_t cfg = {};
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cfg.stream = stream;
      cfg.gridDim = dim3( unsigned( grid ) );
      cfg.blockDim = dim3( kCullThreads );
      DPCU_CUDA( cudaLaunchKernelEx( &cfg, cullFusedLeafKernel<NV>, args, *leaf ) );
    }
    else if ( useLines )
    {
      const uint32_t nLines = uint32_t( divUp( divUp( ctx->n, 32 ), args.lineWords ) );
      const uint32_t ctasForLines = uint32_t( divUp( nLines, kCullThreads / 32 ) );
      if ( uint32_t( grid ) > ctasForLines ) grid = int( ctasForLines );
      if ( useLinesMv && fuseList ) cullLinesMvKernel<NV, true><<<grid, kCullThreads, 0, stream>>>( args );
      else if ( useLinesMv )        cullLinesMvKernel<NV, false><<<grid, kCullThreads, 0, stream>>>( args );
      else if ( fuseList )          cullLinesKernel<NV, true><<<grid, kCullThreads, 0, stream>>>( args );
      else                          cullLinesKernel<NV, false><<<grid, kCullThreads, 0, stream>>>( args );
      DPCU_CUDA( cudaGetLastError() );
    }
    else if ( useGrid )
    {
      // cooperative launch: the grid-wide barrier needs every CTA resident, and the runtime checks that it is
      if ( ctx->optCtasPerSm > 0 )
      {
        int fit = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor( &fit, cullGridKernel<NV>, kCullThreads, 0 );
        if ( grid > ctx->smCount * fit ) grid = ctx->smCount * ( fit > 0 ? fit : 1 );
      }
      void *params[] = { &args };
      DPCU_CUDA( cudaLaunchCooperativeKernel( reinterpret_cast<void const *>( &cullGridKernel<NV> ), dim3( unsigned( grid ) ), dim3( kCullThreads ), params, 0, stream ) );
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
    ctx->lastKernel = useFused ? DPCU_KERNEL_FUSED_LEAF : useLinesMv ? DPCU_KERNEL_LINES_PAIRS : useLines ? DPCU_KERNEL_LINES : useGrid ? DPCU_KERNEL_GRID : ctx->optFma ? DPCU_KERNEL_DIRECT
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
    ctx->stagingFree.destroy();
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
    dpcu::Range nvtxRange( "dpcuCullSetObjects" );
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

  int dpcuCullSetObjectCount( dpcuCull *ctx, size_t n )
  {
    DPCU_REQUIRE( ctx, "ctx is NULL" );
    DPCU_REQUIRE( n < ( size_t( 1 ) << 32 ), "object count must fit 32 bits" );
    if ( n == ctx->n ) return DPCU_OK;
    dpcu::DeviceGuard guard( ctx->device );
    DPCU_CUDA( ctx->lastRun.orderBefore( ctx->stream ) );
    DPCU_TRY( ctx->lowerIdx.reserve( ( n ? n : 1 ) * 16, true, ctx->stream ) );
    DPCU_TRY( ctx->extent.reserve( ( n ? n : 1 ) * 16, true, ctx->stream ) );
    if ( n < ctx->n ) ctx->maxIndexStale = true;
    ctx->n = n;
    ++ctx->objectsVersion;
    return DPCU_OK;
  }

  int dpcuCullUpdateObjects( dpcuCull *ctx, const uint32_t *indices, size_t k, const float *lower4, const float *extent4,
                             const uint32_t *transformIndex )
  {
    dpcu::Range nvtxRange( "dpcuCullUpdateObjects" );
    DPCU_REQUIRE( ctx, "ctx is NULL" );
    DPCU_REQUIRE( !k || ( indices && lower4 && extent4 && transformIndex ), "NULL argument" );
    DPCU_REQUIRE( k < ( size_t( 1 ) << 32 ), "batch must fit 32 bits" );
    if ( !k ) return DPCU_OK;
    dpcu::DeviceGuard guard( ctx->device );
    DPCU_CUDA( ctx->lastRun.orderBefore( ctx->stream ) );
    // [lower | extent | tidx | indices] through the pinned staging buffer: one copy, one kernel, and the caller does not
    // wait for either - the NEXT user of the staging buffer does
    ctx->stagingFree.hostWait();
    DPCU_TRY( ctx->staging.reserve( k * 40 ) );
    char *st = static_cast<char *>( ctx->staging.ptr );
    memcpy( st, lower4, k * 16 );
    memcpy( st + k * 16, extent4, k * 16 );
    memcpy( st + k * 32, transformIndex, k * 4 );
    memcpy( st + k * 36, indices, k * 4 );
    DPCU_TRY( ctx->scratch.reserve( k * 40, false, ctx->stream ) );
    char *d = static_cast<char *>( ctx->scratch.ptr );
    DPCU_CUDA( cudaMemcpyAsync( d, st, k * 40, cudaMemcpyHostToDevice, ctx->stream ) );
    DPCU_CUDA( ctx->stagingFree.record( ctx->stream ) );
    dpcu::scatterObjectsKernel<<<unsigned( dpcu::divUp( k, 256 ) ), 256, 0, ctx->stream>>>(
      reinterpret_cast<uint32_t const *>( d + k * 36 ), reinterpret_cast<float4 const *>( d ), reinterpret_cast<float4 const *>( d + k * 16 ),
      reinterpret_cast<uint32_t const *>( d + k * 32 ), uint32_t( k ), uint32_t( ctx->n ), static_cast<float4 *>( ctx->lowerIdx.ptr ),
      static_cast<float4 *>( ctx->extent.ptr ), static_cast<uint32_t *>( ctx->maxIndex.ptr ) );
    DPCU_CUDA( cudaGetLastError() );
    ++ctx->launches;
    ++ctx->objectsVersion;
    ctx->maxIndexKnown = false;
    ctx->maxIndexStale = true;
    return DPCU_OK;
  }

  int dpcuCullGetObjectCount( const dpcuCull *ctx, size_t *n )
  {
    DPCU_REQUIRE( ctx && n, "NULL argument" );
    *n = ctx->n;
    return DPCU_OK;
  }

  int dpcuCullSetMatrices( dpcuCull *ctx, const void *matrices, size_t count, size_t strideBytes, int memspace )
  {
    dpcu::Range nvtxRange( "dpcuCullSetMatrices" );
    DPCU_REQUIRE( ctx, "ctx is NULL" );
    DPCU_REQUIRE( !count || matrices, "matrices is NULL" );
    DPCU_REQUIRE( strideBytes >= 64 && strideBytes % 4 == 0, "stride must be >= 64 and a multiple of 4" );
    DPCU_REQUIRE( count < ( size_t( 1 ) << 32 ), "matrix count must fit 32 bits" );
    DPCU_REQUIRE( memspace == DPCU_MEM_HOST || memspace == DPCU_MEM_DEVICE, "bad memspace" );
    dpcu::DeviceGuard guard( ctx->device );
    DPCU_CUDA( ctx->lastRun.orderBefore( ctx->stream ) );
    DPCU_TRY( ctx->mats.reserve( ( count ? count : 1 ) * 64, false, ctx->stream ) );
    ctx->boundMats = nullptr;
    ctx->boundTree = nullptr;
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
    dpcu::Range nvtxRange( "dpcuCullUpdateMatrices" );
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
    ctx->stagingFree.hostWait();
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
    DPCU_CUDA( ctx->stagingFree.record( ctx->stream ) );  // the next user of the staging buffer waits for this, the caller does not
    return DPCU_OK;
  }

  int dpcuCullBindMatrices( dpcuCull *ctx, const void *deviceMatrices, size_t count )
  {
    DPCU_REQUIRE( ctx, "ctx is NULL" );
    DPCU_REQUIRE( deviceMatrices || !count, "deviceMatrices is NULL" );
    DPCU_REQUIRE( ( reinterpret_cast<uintptr_t>( deviceMatrices ) & 31 ) == 0, "matrices must be 32-byte aligned (the kernels read them with 256-bit loads)" );
    DPCU_REQUIRE( count < ( size_t( 1 ) << 32 ), "matrix count must fit 32 bits" );
    ctx->boundMats = static_cast<float const *>( deviceMatrices );
    ctx->boundTree = nullptr;
    ctx->nMats = count;
    return DPCU_OK;
  }

  int dpcuCullBindTree( dpcuCull *ctx, dpcuTree *tree )
  {
    DPCU_REQUIRE( ctx && tree, "NULL argument" );
    DPCU_REQUIRE( tree->device == ctx->device, "tree and culling context live on different devices" );
    ctx->boundTree = tree;
    ctx->boundMats = static_cast<float const *>( tree->world.ptr );
    ctx->nMats     = tree->numNodes;
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
    r->visible.release(); r->visCounters.release(); r->look.release(); r->gridTotals.release();
    if ( r->prev ) r->prev->next = r->next; else ctx->results = r->next;
    if ( r->next ) r->next->prev = r->prev;
    delete r;
    return DPCU_OK;
  }

}   // extern "C"

  // Everything of a cull that can fail before a kernel is launched: arguments, index range, host mirror sizes, peer
  // configuration, result capacity.  dpcuCullRunWithTree runs it BEFORE it touches the tree's dirty state.
  static int prepareCull( dpcuCull *ctx, dpcuCullResult *const *results, const float *viewProjections, int nViews, cudaStream_t s )
  {
    DPCU_REQUIRE( ctx && results && viewProjections, "NULL argument" );
    DPCU_REQUIRE( nViews >= 1 && nViews <= DPCU_MAX_VIEWS, "nViews must be 1..DPCU_MAX_VIEWS" );
    for ( int v = 0; v < nViews; ++v )
    {
      DPCU_REQUIRE( results[v] && results[v]->ctx == ctx, "result does not belong to this context" );
      for ( int u = 0; u < v; ++u ) DPCU_REQUIRE( results[u] != results[v], "results must be distinct" );
      // one kernel stores every view's lines into the peers: all results of a run share one peer layout
      DPCU_REQUIRE( results[v]->nPeers == results[0]->nPeers && results[v]->peerWordOffset == results[0]->peerWordOffset,
                    "results of one run must agree on peer count and word offset (dpcuCullResultSetPeerBits)" );
    }
    if ( ctx->boundTree )
    {
      // bound in place: follow the tree (its arrays may have been reallocated by dpcuTreeSetTopology)
      ctx->boundMats = static_cast<float const *>( ctx->boundTree->world.ptr );
      ctx->nMats     = ctx->boundTree->numNodes;
    }
    if ( s != ctx->stream )
    {
      // uploads were submitted on the context stream: order this run after them on the device
      DPCU_CUDA( ctx->uploads.record( ctx->stream ) );
      DPCU_CUDA( ctx->uploads.orderBefore( s ) );
    }
    const size_t n = ctx->n;
    DPCU_TRY( dpcu::checkIndexRange( ctx, "dpcuCullRun" ) );
    for ( int v = 0; v < nViews; ++v )
    {
      dpcuCullResult *r = results[v];
      if ( r->hBits && r->hBitsWords < dpcu::divUp( n, 32 ) )
        return dpcu::fail( DPCU_ERR_INVALID_VALUE, "dpcuCullRun: host mirror holds %zu bitset words, %zu needed", r->hBitsWords, dpcu::divUp( n, 32 ) );
      DPCU_CUDA( r->done.orderBefore( s ) );      // after this result's previous cull / bit moves, wherever they ran
      DPCU_TRY( dpcu::ensureResultCapacity( r, n, s ) );
    }
    return DPCU_OK;
  }

  // prepareCull has succeeded for the same arguments
  static int runCull( dpcuCull *ctx, dpcuCullResult *const *results, const float *viewProjections, int nViews, cudaStream_t s, dpcu::LeafArgs const *leaf,
                      bool treeOnStream = false )
  {
    const size_t n = ctx->n;
    const size_t nSegs = dpcu::divUp( n, size_t( 1 ) << dpcu::kSegObjectsLog2 );
    // matrices bound to a tree: wait for the compute that produced them (wherever it ran)
    if ( ctx->boundTree && !treeOnStream ) DPCU_CUDA( ctx->boundTree->done.orderBefore( s ) );
    for ( int v = 0; v < nViews; ++v )
    {
      dpcuCullResult *r = results[v];
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
    int listOffsets = 1;
    switch ( nViews )
    {
      case 1: rc = dpcu::launchCull<1>( ctx, results, viewProjections, s, leaf, &mirrorsWritten, &listBuilt, &listOffsets ); break;
      case 2: rc = dpcu::launchCull<2>( ctx, results, viewProjections, s, leaf, &mirrorsWritten, &listBuilt, &listOffsets ); break;
      case 3: rc = dpcu::launchCull<3>( ctx, results, viewProjections, s, leaf, &mirrorsWritten, &listBuilt, &listOffsets ); break;
      case 4: rc = dpcu::launchCull<4>( ctx, results, viewProjections, s, leaf, &mirrorsWritten, &listBuilt, &listOffsets ); break;
      case 5: rc = dpcu::launchCull<5>( ctx, results, viewProjections, s, leaf, &mirrorsWritten, &listBuilt, &listOffsets ); break;
      case 6: rc = dpcu::launchCull<6>( ctx, results, viewProjections, s, leaf, &mirrorsWritten, &listBuilt, &listOffsets ); break;
      case 7: rc = dpcu::launchCull<7>( ctx, results, viewProjections, s, leaf, &mirrorsWritten, &listBuilt, &listOffsets ); break;
      case 8: rc = dpcu::launchCull<8>( ctx, results, viewProjections, s, leaf, &mirrorsWritten, &listBuilt, &listOffsets ); break;
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
        ca.seg[v] = r->segPtr();
        if ( !mirrorsWritten && r->dBits )
        {
          ca.bits[v] = static_cast<uint32_t const *>( r->bits.ptr );
          ca.hostBits[v] = r->dBits;
        }
        ca.changed[v] = static_cast<uint32_t *>( r->changed.ptr );
        ca.hostChanged[v]  = r->dChanged;
        ca.hostCount[v]    = r->dCount;
        ca.hostCapacity[v] = uint32_t( r->hChangedCap < 0xffffffffull ? r->hChangedCap : 0xffffffffull );
      }
      ca.nWords = uint32_t( dpcu::divUp( n, 32 ) );
      ca.nSegs = uint32_t( nSegs );
      ca.selfPrefix = listOffsets == 2 ? 1 : listOffsets == 3 ? 2 : 0;
      ca.done = results[0]->donePtr();
      // programmatic dependent launch: the compaction CTAs are placed while the cull kernel drains and wait in
      // cudaGridDependencySynchronize() for its memory - the launch latency of the second kernel is off the critical
      // path (it was a quarter of the step at 1 Mi objects)
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim  = dim3( ca.nSegs, unsigned( nViews ) );
      cfg.blockDim = dim3( 256 );
      cfg.stream   = s;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      DPCU_CUDA( cudaLaunchKernelEx( &cfg, dpcu::compactChangedKernel, ca ) );
      ++ctx->launches;
      mirrorsWritten = true;
    }
    if ( leaf && results[0]->nPeers > 0 )
    {
      // C3 on several GPUs: the fused leaf kernel writes this shard's words locally; one small kernel then stores
      // them as whole lines into every peer's full bitset (NVLink).  2 MB at 16 Mi objects: microseconds, against
      // the 1 GB re-read of the world matrices that un-fusing the leaf level would cost.
      dpcu::PeerGatherArgs ga;
      memset( &ga, 0, sizeof ga );
      ga.nViews = nViews;
      ga.nPeers = uint32_t( results[0]->nPeers );
      ga.nWords = uint32_t( dpcu::divUp( n, 32 ) );
      ga.wordOffset = uint32_t( results[0]->peerWordOffset );
      for ( int v = 0; v < nViews; ++v )
      {
        ga.bits[v] = static_cast<uint32_t const *>( results[v]->bits.ptr );
        for ( int p = 0; p < dpcu::kMaxPeers; ++p ) ga.peer[v][p] = p < results[v]->nPeers ? results[v]->peer[p] : nullptr;
      }
      unsigned blocks = unsigned( dpcu::divUp( dpcu::divUp( ga.nWords, 4 ), 256 ) );
      if ( blocks > unsigned( ctx->smCount ) * 4u ) blocks = unsigned( ctx->smCount ) * 4u;
      dpcu::peerGatherKernel<<<dim3( blocks ? blocks : 1u, unsigned( nViews ) ), 256, 0, s>>>( ga );
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
    // the tree's next compute must not overwrite the world matrices under this cull
    if ( ctx->boundTree && !treeOnStream ) DPCU_CUDA( ctx->boundTree->readers.record( s ) );
    return DPCU_OK;
  }


extern "C"
{
  int dpcuCullRun( dpcuCull *ctx, dpcuCullResult *const *results, const float *viewProjections, int nViews, dpcuStream *stream )
  {
    dpcu::Range nvtxRange( "dpcuCullRun" );
    DPCU_REQUIRE( ctx, "ctx is NULL" );
    dpcu::DeviceGuard guard( ctx->device );
    cudaStream_t s = stream ? stream->stream : ctx->stream;
    DPCU_TRY( prepareCull( ctx, results, viewProjections, nViews, s ) );
    return runCull( ctx, results, viewProjections, nViews, s, nullptr );
  }

  int dpcuCullRunWithTree( dpcuCull *ctx, dpcuTree *tree, dpcuCullResult *const *results, const float *viewProjections, int nViews,
                           dpcuStream *stream )
  {
    dpcu::Range nvtxRange( "dpcuCullRunWithTree" );
    DPCU_REQUIRE( ctx && tree, "NULL argument" );
    DPCU_REQUIRE( tree->device == ctx->device, "tree and culling context live on different devices" );
    DPCU_REQUIRE( tree->numNodes >= 1, "no topology set" );
    dpcu::DeviceGuard guard( ctx->device );
    cudaStream_t s = stream ? stream->stream : ctx->stream;
    // the culler reads the tree's world matrices in place
    ctx->boundTree = tree;
    ctx->boundMats = static_cast<float const *>( tree->world.ptr );
    ctx->nMats     = tree->numNodes;
    // everything that can fail without a launch fails HERE, before the tree's dirty state is touched: a fused cull that
    // never ran would otherwise leave the leaf level's dirty bits cleared and its world matrices stale
    DPCU_TRY( prepareCull( ctx, results, viewProjections, nViews, s ) );
    const size_t levels = tree->levelOffsets.empty() ? 0 : tree->levelOffsets.size() - 1;
    // can the last level run inside the cull kernel?  object i <-> entry i of that level
    bool fuse = false;
    size_t lastFirst = 0;
    // (peer bitsets - the multi-GPU gather - follow the fused kernel as one small copy kernel, see runCull)
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
    int rc = runCull( ctx, results, viewProjections, nViews, s, fuse ? &leaf : nullptr, true );
    if ( fuse && rc == DPCU_OK ) ++tree->launches;        // the fused kernel is the tree's last level too
    // a fused cull that failed at launch has not propagated the last level: do it now, so that closing the compute
    // (which clears the dirty bits) leaves no stale world matrix behind
    if ( fuse && rc != DPCU_OK ) dpcu::treeComputeLevels( tree, s, levels - 1, levels );
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

  int dpcuCullResultUpdateWords( dpcuCullResult *r, const uint32_t *indices, const uint32_t *words, size_t k )
  {
    DPCU_REQUIRE( r, "result is NULL" );
    DPCU_REQUIRE( !k || ( indices && words ), "NULL argument" );
    if ( !k ) return DPCU_OK;
    const size_t have = dpcu::divUp( r->n, 32 );
    for ( size_t i = 0; i < k; ++i ) DPCU_REQUIRE( indices[i] < have, "word index beyond the stored result" );
    dpcuCull *ctx = r->ctx;
    dpcu::DeviceGuard guard( ctx->device );
    cudaStream_t s = ctx->stream;
    DPCU_CUDA( r->done.orderBefore( s ) );
    ctx->stagingFree.hostWait();
    DPCU_TRY( ctx->staging.reserve( k * 8 ) );
    char *st = static_cast<char *>( ctx->staging.ptr );
    memcpy( st, indices, k * 4 );
    memcpy( st + k * 4, words, k * 4 );
    DPCU_TRY( ctx->scratch.reserve( k * 8, false, s ) );
    DPCU_CUDA( cudaMemcpyAsync( ctx->scratch.ptr, st, k * 8, cudaMemcpyHostToDevice, s ) );
    DPCU_CUDA( ctx->stagingFree.record( s ) );
    dpcu::scatterWordsKernel<<<unsigned( dpcu::divUp( k, 256 ) ), 256, 0, s>>>( static_cast<uint32_t const *>( ctx->scratch.ptr ),
      static_cast<uint32_t const *>( ctx->scratch.ptr ) + k, uint32_t( k ), static_cast<uint32_t *>( r->bits.ptr ), r->ran ? r->dBits : nullptr );
    DPCU_CUDA( cudaGetLastError() );
    DPCU_CUDA( r->done.record( s ) );
    ++ctx->launches;
    r->visBuilt = false;
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
    dpcu::Range nvtxRange( "dpcuCullResultBuildVisibleList" );
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
    dpcu::Range nvtxRange( "dpcuCullResultSynchronize" );
    DPCU_REQUIRE( r, "result is NULL" );
    dpcu::DeviceGuard guard( r->ctx->device );
    if ( r->done.pending ) DPCU_CUDA( cudaEventSynchronize( r->done.event ) );
    return DPCU_OK;
  }

  int dpcuCullGetBoundingBox( dpcuCull *ctx, float *out6 )
  {
    dpcu::Range nvtxRange( "dpcuCullGetBoundingBox" );
    DPCU_REQUIRE( ctx && out6, "NULL argument" );
    dpcu::DeviceGuard guard( ctx->device );
    const float FMAX = 3.402823466e+38f;
    float lo[3] = { FMAX, FMAX, FMAX }, hi[3] = { -FMAX, -FMAX, -FMAX };
    if ( ctx->n )
    {
      DPCU_TRY( dpcu::checkIndexRange( ctx, "dpcuCullGetBoundingBox" ) );
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
      case DPCU_CULL_OPT_KERNEL:       DPCU_REQUIRE( ( value >= 0 && value <= 5 ) || value == DPCU_KERNEL_LINES_PAIRS || value == DPCU_KERNEL_GRID, "kernel must be 0..5, 7 or 8" ); ctx->optKernel = value; break;
      case DPCU_CULL_OPT_FMA:          ctx->optFma = value ? 1 : 0; break;
      case DPCU_CULL_OPT_CHANGED_LIST: ctx->optChanged = value ? 1 : 0; break;
      case DPCU_CULL_OPT_CTAS_PER_SM:  DPCU_REQUIRE( value >= 0 && value <= 32, "ctas per SM must be 0..32" ); ctx->optCtasPerSm = value; break;
      case DPCU_CULL_OPT_PROFILE:      ctx->optProfile = value ? 1 : 0; break;
      case DPCU_CULL_OPT_FUSE_LEAF:    ctx->optFuseLeaf = value ? 1 : 0; break;
      case DPCU_CULL_OPT_FUSE_LIST:    ctx->optFuseList = value ? 1 : 0; break;
      case DPCU_CULL_OPT_FILTER:       DPCU_REQUIRE( value >= 0 && value <= 3, "filter must be 0..3" ); ctx->optFilter = value; break;
      case DPCU_CULL_OPT_LINE_WORDS:   DPCU_REQUIRE( value == 0 || value == 8 || value == 16 || value == 32, "line words must be 0 (auto), 8, 16 or 32" );
                                       ctx->optLineWords = value; break;
      case DPCU_CULL_OPT_LIST_OFFSETS: DPCU_REQUIRE( value >= 0 && value <= 3, "list offsets must be 0..3" ); ctx->optListOffsets = value; break;
      case DPCU_CULL_OPT_L2_PREFETCH:  ctx->optL2Prefetch = value ? 1 : 0; break;
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
      case DPCU_CULL_OPT_LIST_OFFSETS: *value = ctx->optListOffsets; break;
      case DPCU_CULL_OPT_L2_PREFETCH:  *value = ctx->optL2Prefetch; break;
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
