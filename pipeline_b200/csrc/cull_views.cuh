// Multi-view form of the visibility test: the same binary32 operations in the same order as
// cull_math.cuh::obbVisible, arranged for Blackwell's issue limits.
//
// Why a second form: with V views the cull is no longer HBM-bound - per object and view the
// reference arithmetic is 64 mul + 76 add + 48 compares.  The direct kernel interleaves all
// views (6 x V live predicate chains -> predicate spills) and issues every operation as a
// scalar instruction.  Here
//   * views run one after the other (a rolled loop), so only six "all corners outside plane k"
//     chains are live;
//   * multiplies and adds are issued as packed pairs (mul.rn.f32x2 / add.rn.f32x2 -> FMUL2 /
//     FADD2): component pairs (x,y) and (z,w) of one vector share an instruction.  Each half is
//     an independent round-to-nearest binary32 operation - identical bits to the scalar form
//     (see addProd below for how contraction into FMA is kept out);
//   * when the object's computed OBB has pt.w == 1 and ax.w == ay.w == az.w == 0 (every affine
//     world matrix) and the view-projection is finite, the four products with those values are
//     skipped:  1*P[3][c] == P[3][c] exactly, and (+-0)*P[3][c] == +-0 only changes the sign of a
//     zero sum, which no later operation can observe (adds keep the value, the clip compares
//     treat -0 == +0).  The decision is made per warp so there is no divergence; any warp that
//     holds a non-affine object takes the general path.
#pragma once

#include "cull_math.cuh"

namespace dpcu
{
  typedef unsigned long long f32x2;    // two binary32 values in one 64-bit register pair (lo = first)

  __device__ __forceinline__ f32x2 pack2( float lo, float hi )
  {
    f32x2 r;
    asm( "mov.b64 %0, {%1, %2};" : "=l"( r ) : "f"( lo ), "f"( hi ) );
    return r;
  }
  __device__ __forceinline__ void unpack2( f32x2 v, float &lo, float &hi )
  {
    asm( "mov.b64 {%0, %1}, %2;" : "=f"( lo ), "=f"( hi ) : "l"( v ) );
  }
  __device__ __forceinline__ f32x2 mul2( f32x2 a, f32x2 b )
  {
    f32x2 r;
    asm( "mul.rn.f32x2 %0, %1, %2;" : "=l"( r ) : "l"( a ), "l"( b ) );
    return r;
  }
  __device__ __forceinline__ f32x2 add2( f32x2 a, f32x2 b )
  {
    f32x2 r;
    asm( "add.rn.f32x2 %0, %1, %2;" : "=l"( r ) : "l"( a ), "l"( b ) );
    return r;
  }

  // a vector as two packed halves: lo = (x, y), hi = (z, w)
  struct Vec4p
  {
    f32x2 lo, hi;
  };

  // one OBB, every scalar pre-broadcast into a pair (done once per object, reused by all views)
  struct ObbPairs
  {
    f32x2 pt[4], ax[4], ay[4], az[4];
  };

  __device__ __forceinline__ ObbPairs broadcastObb( Obb const &o )
  {
    ObbPairs b;
    b.pt[0] = pack2( o.pt.x, o.pt.x ); b.pt[1] = pack2( o.pt.y, o.pt.y ); b.pt[2] = pack2( o.pt.z, o.pt.z ); b.pt[3] = pack2( o.pt.w, o.pt.w );
    b.ax[0] = pack2( o.ax.x, o.ax.x ); b.ax[1] = pack2( o.ax.y, o.ax.y ); b.ax[2] = pack2( o.ax.z, o.ax.z ); b.ax[3] = pack2( o.ax.w, o.ax.w );
    b.ay[0] = pack2( o.ay.x, o.ay.x ); b.ay[1] = pack2( o.ay.y, o.ay.y ); b.ay[2] = pack2( o.ay.z, o.ay.z ); b.ay[3] = pack2( o.ay.w, o.ay.w );
    b.az[0] = pack2( o.az.x, o.az.x ); b.az[1] = pack2( o.az.y, o.az.y ); b.az[2] = pack2( o.az.z, o.az.z ); b.az[3] = pack2( o.az.w, o.az.w );
    return b;
  }

  // view-projection rows as pairs: row r = ( p[2r], p[2r+1] ) = ( (P[r][0],P[r][1]), (P[r][2],P[r][3]) )
  struct ViewPairs
  {
    f32x2 p[8];
  };

  // ptxas contracts mul.rn.f32x2 feeding add.rn.f32x2 into FFMA2 even under --fmad=false (checked
  // with CUDA 12.9: the explicit .rn does not protect the packed forms as it does the scalar
  // ones).  A contracted product is rounded once instead of twice, which breaks bit-exactness,
  // so every add that consumes a product is written as fma(x, one, y) with `one` = (1.0f, 1.0f)
  // arriving as a kernel argument: x*1 is exact, so the result is round(x + y) - the same bits as
  // add.rn - and ptxas cannot fold a multiplier it does not know.  tests/test_sass.py checks the
  // instruction mix of the built kernel for this.
  __device__ __forceinline__ f32x2 addProd( f32x2 x, f32x2 y, f32x2 one )
  {
    f32x2 r;
    asm( "fma.rn.f32x2 %0, %1, %2, %3;" : "=l"( r ) : "l"( x ), "l"( one ), "l"( y ) );
    return r;
  }

  // dp/math/Matmnt.h:1371-1379, all four components: ((v0*P0 + v1*P1) + v2*P2) + v3*P3
  __device__ __forceinline__ Vec4p vecMulMat4( f32x2 const ( &v )[4], ViewPairs const &P, f32x2 one )
  {
    Vec4p r;
    r.lo = addProd( addProd( addProd( mul2( v[0], P.p[0] ), mul2( v[1], P.p[2] ), one ), mul2( v[2], P.p[4] ), one ), mul2( v[3], P.p[6] ), one );
    r.hi = addProd( addProd( addProd( mul2( v[0], P.p[1] ), mul2( v[1], P.p[3] ), one ), mul2( v[2], P.p[5] ), one ), mul2( v[3], P.p[7] ), one );
    return r;
  }
  // v[3] == 1 : the last product is P3 itself
  __device__ __forceinline__ Vec4p vecMulMatW1( f32x2 const ( &v )[4], ViewPairs const &P, f32x2 one )
  {
    Vec4p r;
    r.lo = add2( addProd( addProd( mul2( v[0], P.p[0] ), mul2( v[1], P.p[2] ), one ), mul2( v[2], P.p[4] ), one ), P.p[6] );
    r.hi = add2( addProd( addProd( mul2( v[0], P.p[1] ), mul2( v[1], P.p[3] ), one ), mul2( v[2], P.p[5] ), one ), P.p[7] );
    return r;
  }
  // v[3] == +-0 and P finite: the last product is a zero, adding it keeps the value
  __device__ __forceinline__ Vec4p vecMulMatW0( f32x2 const ( &v )[4], ViewPairs const &P, f32x2 one )
  {
    Vec4p r;
    r.lo = addProd( addProd( mul2( v[0], P.p[0] ), mul2( v[1], P.p[2] ), one ), mul2( v[2], P.p[4] ), one );
    r.hi = addProd( addProd( mul2( v[0], P.p[1] ), mul2( v[1], P.p[3] ), one ), mul2( v[2], P.p[5] ), one );
    return r;
  }

  __device__ __forceinline__ Vec4p addp( Vec4p a, Vec4p b )
  {
    Vec4p r;
    r.lo = add2( a.lo, b.lo );
    r.hi = add2( a.hi, b.hi );
    return r;
  }

  // Per axis three running flags over the corners seen so far (determineCullFlags,
  // dp/culling/cpu/src/ManagerImpl.cpp:199-229):
  //   allN = every corner has  c <= -w          (bit 0x01 / 0x04 / 0x10 survives in cfa)
  //   anyN = some corner has   c <= -w
  //   allP = every corner has  w <= c           (with !anyN: bit 0x02 / 0x08 / 0x20 survives)
  // cfa keeps the "else if" bit only when no corner took the first branch, so
  //   outside(axis) = allN || ( allP && !anyN ).
  struct AxisFlags
  {
    bool allN, anyN, allP;
  };

  __device__ __forceinline__ void cornerFlags( AxisFlags &fx, AxisFlags &fy, AxisFlags &fz, Vec4p p )
  {
    float x, y, z, w;
    unpack2( p.lo, x, y );
    unpack2( p.hi, z, w );
    const float nw = -w;
    // & and | instead of && and ||: straight-line predicate chains (FSETP.AND / FSETP.OR), no branches
    fx.allN = fx.allN & ( x <= nw ); fx.anyN = fx.anyN | ( x <= nw ); fx.allP = fx.allP & ( w <= x );
    fy.allN = fy.allN & ( y <= nw ); fy.anyN = fy.anyN | ( y <= nw ); fy.allP = fy.allP & ( w <= y );
    fz.allN = fz.allN & ( z <= nw ); fz.anyN = fz.anyN | ( z <= nw ); fz.allP = fz.allP & ( w <= z );
  }

  // isVisible(projection, obb), dp/culling/cpu/src/ManagerImpl.cpp:263-289, from the four
  // clip-space vectors v0 = pt*P, X = ax*P, Y = ay*P, Z = az*P
  __device__ __forceinline__ bool cornersVisible( Vec4p v0, Vec4p X, Vec4p Y, Vec4p Z )
  {
    const Vec4p v1 = addp( v0, X );
    const Vec4p v2 = addp( v0, Y );
    const Vec4p v3 = addp( v1, Y );
    const Vec4p v4 = addp( v0, Z );
    const Vec4p v5 = addp( v1, Z );
    const Vec4p v6 = addp( v2, Z );
    const Vec4p v7 = addp( v3, Z );
    AxisFlags fx = { true, false, true }, fy = fx, fz = fx;
    cornerFlags( fx, fy, fz, v0 );
    cornerFlags( fx, fy, fz, v1 );
    cornerFlags( fx, fy, fz, v2 );
    cornerFlags( fx, fy, fz, v3 );
    cornerFlags( fx, fy, fz, v4 );
    cornerFlags( fx, fy, fz, v5 );
    cornerFlags( fx, fy, fz, v6 );
    cornerFlags( fx, fy, fz, v7 );
    const bool outside = fx.allN | ( fx.allP & !fx.anyN ) | fy.allN | ( fy.allP & !fy.anyN ) | fz.allN | ( fz.allP & !fz.anyN );
    return !outside;
  }

  // The same decision with one third fewer ALU-pipe instructions.  The kernel is bound by the ALU
  // pipe (one warp instruction per two cycles: 72 FSETP per view above), while the FMA pipe has
  // slack.  allN and anyN are two accumulations of the SAME compare, so the compare is
  // materialised once as 1.0f / 0.0f (FSET.BF) and COUNTED with packed adds on the FMA pipe:
  //   n = number of corners with c <= -w       allN <=> n == 8,  anyN <=> n > 0
  // (sums of at most eight ones are exact), allP stays a predicate chain.  Per view 24 FSET +
  // 24 FSETP + 8 FADD2 + 8 FADD instead of 72 FSETP; NaN compares are false in both forms.
  struct CountFlags
  {
    f32x2 nXY;              // (count for x, count for y)
    float nZ;
    bool  px, py, pz;       // every corner so far has w <= c
  };

  __device__ __forceinline__ void cornerCount( CountFlags &f, Vec4p p )
  {
    float x, y, z, w;
    unpack2( p.lo, x, y );
    unpack2( p.hi, z, w );
    const float nw = -w;
    const float qx = ( x <= nw ) ? 1.0f : 0.0f;
    const float qy = ( y <= nw ) ? 1.0f : 0.0f;
    f.nXY = add2( f.nXY, pack2( qx, qy ) );
    const float qz = ( z <= nw ) ? 1.0f : 0.0f;
    f.nZ  = f.nZ + qz;
    f.px = f.px & ( w <= x );
    f.py = f.py & ( w <= y );
    f.pz = f.pz & ( w <= z );
  }

  __device__ __forceinline__ bool cornersVisibleCounted( Vec4p v0, Vec4p X, Vec4p Y, Vec4p Z )
  {
    const Vec4p v1 = addp( v0, X );
    const Vec4p v2 = addp( v0, Y );
    const Vec4p v3 = addp( v1, Y );
    const Vec4p v4 = addp( v0, Z );
    const Vec4p v5 = addp( v1, Z );
    const Vec4p v6 = addp( v2, Z );
    const Vec4p v7 = addp( v3, Z );
    CountFlags f = { pack2( 0.0f, 0.0f ), 0.0f, true, true, true };
    cornerCount( f, v0 );
    cornerCount( f, v1 );
    cornerCount( f, v2 );
    cornerCount( f, v3 );
    cornerCount( f, v4 );
    cornerCount( f, v5 );
    cornerCount( f, v6 );
    cornerCount( f, v7 );
    float nx, ny;
    unpack2( f.nXY, nx, ny );
    const bool outside = ( nx == 8.0f ) | ( f.px & ( nx == 0.0f ) ) | ( ny == 8.0f ) | ( f.py & ( ny == 0.0f ) )
                       | ( f.nZ == 8.0f ) | ( f.pz & ( f.nZ == 0.0f ) );
    return !outside;
  }

  // The four clip-space vectors of one view; the view loop computes them one view ahead of the
  // corner tests so that each warp's instruction stream mixes FMA-pipe work (the products of the
  // next view) with ALU-pipe work (the 72 compares of the current one) instead of alternating
  // between long single-pipe phases.
  struct ClipVectors
  {
    Vec4p v0, X, Y, Z;
  };

  template <bool kAffine>
  __device__ __forceinline__ ClipVectors clipVectors( ObbPairs const &o, ViewPairs const &P, f32x2 one )
  {
    ClipVectors c;
    if ( kAffine )
    {
      c.v0 = vecMulMatW1( o.pt, P, one );
      c.X  = vecMulMatW0( o.ax, P, one );
      c.Y  = vecMulMatW0( o.ay, P, one );
      c.Z  = vecMulMatW0( o.az, P, one );
    }
    else
    {
      c.v0 = vecMulMat4( o.pt, P, one );
      c.X  = vecMulMat4( o.ax, P, one );
      c.Y  = vecMulMat4( o.ay, P, one );
      c.Z  = vecMulMat4( o.az, P, one );
    }
    return c;
  }

  template <bool kCount>
  __device__ __forceinline__ bool cornersVisible( ClipVectors const &c )
  {
    return kCount ? cornersVisibleCounted( c.v0, c.X, c.Y, c.Z ) : cornersVisible( c.v0, c.X, c.Y, c.Z );
  }

  __device__ __forceinline__ ViewPairs loadViewPairs( float4 const *vpRows )
  {
    ViewPairs P;
    f32x2 const *src = reinterpret_cast<f32x2 const *>( vpRows );
#pragma unroll
    for ( int k = 0; k < 8; ++k ) P.p[k] = src[k];
    return P;
  }

  // All NV views of one object, one after the other (a rolled loop: only the six flag chains of
  // one view are live).  vp = NV x 4 rows in kernel parameter space.  Returns, in lane v of the
  // warp, the ballot word of view v.
  template <int NV, bool kAffine, bool kCount = true>
  __device__ __forceinline__ uint32_t cullViews( ObbPairs const &ob, float4 const ( *vp )[4], f32x2 one, bool live, uint32_t lane )
  {
    uint32_t myWord = 0;
#ifndef DPCU_VIEWS_UNROLL
#define DPCU_VIEWS_UNROLL 3
#endif
    constexpr int kUnroll = DPCU_VIEWS_UNROLL;
#pragma unroll kUnroll
    for ( int v = 0; v < NV; ++v )
    {
      const ClipVectors c = clipVectors<kAffine>( ob, loadViewPairs( vp[v] ), one );
      const uint32_t b = __ballot_sync( 0xffffffffu, cornersVisible<kCount>( c ) & live );
      if ( lane == uint32_t( v ) ) myWord = b;
    }
    return myWord;
  }

}
