// K1 + K2 fused: the transform tree's last level inside the cull kernel.
#pragma once

namespace dpcu
{
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
    __shared__ f32x2 sP[NV > 1 ? NV * 8 : 1];
    __shared__ FilterScratch<NV> sScratch[NV > 1 ? kCullThreads / 32 : 1];
    if ( NV > 1 ) fillViewTable<NV>( sP, a );
    // launched as a programmatic dependent of the tree's last upper level (dpcuCullRunWithTree): everything read of it
    // (parents' world matrices, dirty bits) is read behind this point
    cudaGridDependencySynchronize();
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
        float4 p0, p1, p2, p3;
        ldMatrix( pw, p0, p1, p2, p3 );
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
          float4 l0, l1, l2, l3;
          ldMatrix( ln, l0, l1, l2, l3 );
          float4 p0, p1, p2, p3;
          ldMatrix( pw, p0, p1, p2, p3 );
          w0 = vecMulMat( l0, p0, p1, p2, p3 );                     // Tree.cpp:157, Matmnt.h:1381-1415
          w1 = vecMulMat( l1, p0, p1, p2, p3 );
          w2 = vecMulMat( l2, p0, p1, p2, p3 );
          w3 = vecMulMat( l3, p0, p1, p2, p3 );
          stMatrix( wn, w0, w1, w2, w3 );
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
      if ( lane < NV && ( i - lane ) < a.n ) storeWord<NV>( a.out[lane], a, word, myWord, oldBits );
      ent = entN; entN = entNN; dirty = dirtyN;
    }
    if ( a.buildChanged && a.countSegs == 1 ) scanSegmentsInLastCta<NV>( a.out, a.nSegs, a.done );
  }

  // counts objects whose transform index is not the node of the level entry with the same index
  __global__ void leafBindingKernel( float4 const *lowerIdx, uint2 const *entries, uint32_t n, uint32_t *mismatches )
  {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool bad = i < n && __float_as_uint( lowerIdx[i].w ) != entries[i].y;
    const uint32_t m = __ballot_sync( 0xffffffffu, bad );
    if ( ( threadIdx.x & 31u ) == 0 && m ) atomicAdd( mismatches, __popc( m ) );
  }
}
