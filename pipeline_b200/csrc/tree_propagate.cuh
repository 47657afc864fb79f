// world[node] = local[node] * world[parent] for the transform layer (K1, dpcu_tree.cu) and for the
// culling layer's fused leaf level (dpcu_cull.cu): the reference's operation order
// (dp/transform/src/Tree.cpp:157, dp/math/Matmnt.h:1381-1415) and dirty protocol (Tree.cpp:155-158).
#pragma once

#include "cull_math.cuh"

namespace dpcu
{
  // 16-byte chunk r of object o inside a 32-object x 4-row staging block, XOR-swizzled so that
  // "row j*32+lane" stores and "my object's four rows" loads are both free of bank conflicts
  __device__ __forceinline__ uint32_t swizzledRow( uint32_t o, uint32_t r )
  {
    return o * 4u + ( r ^ ( ( o >> 1 ) & 3u ) );
  }

  // One node per thread with strided 16-byte accesses (any node, any warp shape).
  __device__ __forceinline__ void propagateNode( float4 const *local, float4 *world, uint32_t *dirtyWorld, uint32_t parent, uint32_t node,
                                                 float4 &w0, float4 &w1, float4 &w2, float4 &w3 )
  {
    float4 const *ln = local + 4ull * node;
    float4 const *pw = world + 4ull * parent;
    float4       *wn = world + 4ull * node;
    float4 l0, l1, l2, l3;
    ldMatrix( ln, l0, l1, l2, l3 );
    float4 p0, p1, p2, p3;
    ldMatrix( pw, p0, p1, p2, p3 );
    w0 = vecMulMat( l0, p0, p1, p2, p3 );
    w1 = vecMulMat( l1, p0, p1, p2, p3 );
    w2 = vecMulMat( l2, p0, p1, p2, p3 );
    w3 = vecMulMat( l3, p0, p1, p2, p3 );
    stMatrix( wn, w0, w1, w2, w3 );
    atomicOr( dirtyWorld + ( node >> 5 ), 1u << ( node & 31u ) );
  }

  // A full warp of dirty nodes with consecutive indices node0 .. node0+31 (the usual layout of a
  // level).  The 32 local matrices are 2 KiB contiguous: four coalesced 16-byte loads per lane
  // bring them in, shared memory turns "row j*32+lane" into "my node's four rows", and the same
  // trip backwards turns the world matrices into four coalesced stores (strided 16-byte accesses
  // were measured to double the L2 traffic; strided 256-bit accesses - whole sectors, no shared memory - were measured
  // too: 0.61 instead of 0.50 ms for the fused C3 leaf level, 0.56 instead of 0.46 ms for Tree::compute).  bufIn / bufOut: 128 float4 each, private to the warp;
  // the two __syncwarp()s of a call also order its accesses against the next call's.
  __device__ __forceinline__ void propagateWarpCoalesced( float4 const *local, float4 *world, uint32_t *dirtyWorld, uint32_t parent,
                                                          uint32_t node0, uint32_t lane, float4 *bufIn, float4 *bufOut,
                                                          float4 &w0, float4 &w1, float4 &w2, float4 &w3 )
  {
    float4 const *ln = local + 4ull * node0;
    float4       *wn = world + 4ull * node0;
    float4 const *pw = world + 4ull * parent;
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
    w0 = vecMulMat( l0, p0, p1, p2, p3 );
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
    if ( lane == 0 ) atomicOr( dirtyWorld + ( node0 >> 5 ), 0xffffffffu << ( node0 & 31u ) );          // 32 node bits
    if ( lane == 0 && ( node0 & 31u ) ) atomicOr( dirtyWorld + ( node0 >> 5 ) + 1, ~( 0xffffffffu << ( node0 & 31u ) ) );
  }
}
