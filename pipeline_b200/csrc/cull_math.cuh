// Device arithmetic of the culling path, operation for operation in the reference's order
// (SURVEY.md section 8a "Exact per-object arithmetic").  Every function is written as plain
// binary32 expressions; the translation unit is compiled with -fmad=false so each * and +
// rounds once like the scalar CPU path, and denormals are kept (nvcc default -ftz=false).
// The same source compiled with -fmad=true is the reporting-only FMA fast mode.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

namespace dpcu
{
  struct Obb
  {
    float4 pt, ax, ay, az;
  };

  // dp/math/Matmnt.h:1371-1379 : r[j] = ((v0*m0j + v1*m1j) + v2*m2j) + v3*m3j
  __device__ __forceinline__ float4 vecMulMat( float4 v, float4 m0, float4 m1, float4 m2, float4 m3 )
  {
    float4 r;
    r.x = ( ( v.x * m0.x + v.y * m1.x ) + v.z * m2.x ) + v.w * m3.x;
    r.y = ( ( v.x * m0.y + v.y * m1.y ) + v.z * m2.y ) + v.w * m3.y;
    r.z = ( ( v.x * m0.z + v.y * m1.z ) + v.z * m2.z ) + v.w * m3.z;
    r.w = ( ( v.x * m0.w + v.y * m1.w ) + v.z * m2.w ) + v.w * m3.w;
    return r;
  }

  // same with v.w == 1.0f : 1.0f * m is exact, so the last product is elided, the add is not
  __device__ __forceinline__ float4 pointMulMat( float x, float y, float z, float4 m0, float4 m1, float4 m2, float4 m3 )
  {
    float4 r;
    r.x = ( ( x * m0.x + y * m1.x ) + z * m2.x ) + m3.x;
    r.y = ( ( x * m0.y + y * m1.y ) + z * m2.y ) + m3.y;
    r.z = ( ( x * m0.z + y * m1.z ) + z * m2.z ) + m3.z;
    r.w = ( ( x * m0.w + y * m1.w ) + z * m2.w ) + m3.w;
    return r;
  }

  __device__ __forceinline__ float4 scaleRow( float s, float4 m )
  {
    return make_float4( m.x * s, m.y * s, m.z * s, m.w * s );
  }

  __device__ __forceinline__ float4 add4( float4 a, float4 b )
  {
    return make_float4( a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w );
  }

  // GroupCPU::updateOBBs, dp/culling/cpu/src/ManagerImpl.cpp:127-153
  __device__ __forceinline__ Obb makeObb( float lx, float ly, float lz, float ex, float ey, float ez,
                                          float4 m0, float4 m1, float4 m2, float4 m3 )
  {
    Obb o;
    o.pt = pointMulMat( lx, ly, lz, m0, m1, m2, m3 );
    o.ax = scaleRow( ex, m0 );
    o.ay = scaleRow( ey, m1 );
    o.az = scaleRow( ez, m2 );
    return o;
  }

  // Running "all corners so far are outside plane k" flags = the six bits of cfa in
  // determineCullFlags (dp/culling/cpu/src/ManagerImpl.cpp:199-229):
  //   bit 0x01 : x <= -w            bit 0x02 : !(x <= -w) && w <= x      (the else-if)
  // and likewise for y, z.  NaN makes every comparison false, so no bit is set.
  struct OutsideAll
  {
    bool xn, xp, yn, yp, zn, zp;
  };

  __device__ __forceinline__ void accumulateCorner( OutsideAll &o, float4 p )
  {
    float nw = -p.w;
    o.xn = o.xn && ( p.x <= nw );
    o.xp = o.xp && ( p.w <= p.x ) && !( p.x <= nw );
    o.yn = o.yn && ( p.y <= nw );
    o.yp = o.yp && ( p.w <= p.y ) && !( p.y <= nw );
    o.zn = o.zn && ( p.z <= nw );
    o.zp = o.zp && ( p.w <= p.z ) && !( p.z <= nw );
  }

  // isVisible(projection, obb), dp/culling/cpu/src/ManagerImpl.cpp:263-289.
  // `!cfo || !cfa` reduces to cfa == 0 (cfo == 0 implies cfa == 0).
  // P is the view-projection as four rows.
  __device__ __forceinline__ bool obbVisible( Obb const &o, float4 p0, float4 p1, float4 p2, float4 p3 )
  {
    float4 v0 = vecMulMat( o.pt, p0, p1, p2, p3 );
    float4 x  = vecMulMat( o.ax, p0, p1, p2, p3 );
    float4 y  = vecMulMat( o.ay, p0, p1, p2, p3 );
    float4 z  = vecMulMat( o.az, p0, p1, p2, p3 );
    float4 v1 = add4( v0, x );
    float4 v2 = add4( v0, y );
    float4 v3 = add4( v1, y );
    float4 v4 = add4( v0, z );
    float4 v5 = add4( v1, z );
    float4 v6 = add4( v2, z );
    float4 v7 = add4( v3, z );
    OutsideAll o6 = { true, true, true, true, true, true };
    accumulateCorner( o6, v0 );
    accumulateCorner( o6, v1 );
    accumulateCorner( o6, v2 );
    accumulateCorner( o6, v3 );
    accumulateCorner( o6, v4 );
    accumulateCorner( o6, v5 );
    accumulateCorner( o6, v6 );
    accumulateCorner( o6, v7 );
    return !( o6.xn || o6.xp || o6.yn || o6.yp || o6.zn || o6.zp );
  }

  // 16-byte streaming loads: the object / matrix streams are read exactly once per cull
  // A 64-byte matrix as two 256-bit loads (sm_100: LDG.E.256) instead of four 128-bit ones.  A warp's matrix rows lie
  // 64 bytes apart, so every 128-byte line serves two lanes whatever the access size: the four row loads were 64 L1
  // wavefronts per warp and step, these are 32 - the L1 data pipe was 63 % busy in the six-view kernel (ncu) - and each
  // load is exactly one 32-byte sector.  Needs 32-byte aligned matrices (dpcuCullBindMatrices checks; own copies are).
  __device__ __forceinline__ void ldMatrix( float4 const *m, float4 &m0, float4 &m1, float4 &m2, float4 &m3 )
  {
#ifdef DPCU_NO_LD256       // A/B builds only (make variant)
    m0 = __ldg( m ); m1 = __ldg( m + 1 ); m2 = __ldg( m + 2 ); m3 = __ldg( m + 3 );
    return;
#endif
    asm volatile( "ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                  : "=f"( m0.x ), "=f"( m0.y ), "=f"( m0.z ), "=f"( m0.w ), "=f"( m1.x ), "=f"( m1.y ), "=f"( m1.z ), "=f"( m1.w ) : "l"( m ) );
    asm volatile( "ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                  : "=f"( m2.x ), "=f"( m2.y ), "=f"( m2.z ), "=f"( m2.w ), "=f"( m3.x ), "=f"( m3.y ), "=f"( m3.z ), "=f"( m3.w ) : "l"( m + 2 ) );
  }

  // ... and the store side (STG.E.256): a 64-byte matrix as two full 32-byte sectors
  __device__ __forceinline__ void stMatrix( float4 *m, float4 const &m0, float4 const &m1, float4 const &m2, float4 const &m3 )
  {
    asm volatile( "st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                  :: "l"( m ), "f"( m0.x ), "f"( m0.y ), "f"( m0.z ), "f"( m0.w ), "f"( m1.x ), "f"( m1.y ), "f"( m1.z ), "f"( m1.w ) : "memory" );
    asm volatile( "st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                  :: "l"( m + 2 ), "f"( m2.x ), "f"( m2.y ), "f"( m2.z ), "f"( m2.w ), "f"( m3.x ), "f"( m3.y ), "f"( m3.z ), "f"( m3.w ) : "memory" );
  }

  __device__ __forceinline__ float4 ldStream( float4 const *p )
  {
    float4 r;
    asm volatile( "ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                  : "=f"( r.x ), "=f"( r.y ), "=f"( r.z ), "=f"( r.w ) : "l"( p ) );
    return r;
  }
}
