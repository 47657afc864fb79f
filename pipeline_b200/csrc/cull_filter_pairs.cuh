// The multi-view filter of cull_filter.cuh, second form: two views per packed instruction, fewer instructions
// per (object, view) pair, and a proof that also covers underflow.  Used by the line-granular multi-view kernel
// (kernel_lines_mv.cuh).  Same contract as cull_filter.cuh: a pair is DECIDED only when the reference's result
// (dp/culling/cpu/src/ManagerImpl.cpp:199-289) is proven from the OBB's centre and radius; everything else runs
// the reference arithmetic.
//
// With the plane functions fN_a = w + x_a and fP_a = w - x_a (a = x, y, z) of a view-projection P, (A_x, A_y, A_z, W)
// = centre * P the centre's clip coordinates, r the radius, rho the plane norms and m the margin (below):
//   (V)  W - max_a |A_a| > m                                  every plane function at the centre exceeds the error
//                                                             => per plane SOME float corner is inside => visible
//   (N)  min_a( A_a + r rhoN_a ) < -(W + m)                   every float corner has x_a + w <= 0 => bit N everywhere
//   (P)  max_a( A_a - r rhoP_a ) >  (W + m)  and  W - r rhoW > m
//                                                             every float corner has w <= x_a and w > 0; then
//                                                             x_a >= w > 0 > -w, the else-if is taken: bit P everywhere
// (P) differs from cull_filter.cuh, which asked for "fN_a - support > m" per axis: "the whole sphere is in front of
// the eye plane w = 0" gives the same conclusion with one test for all three axes; it leaves the few objects that
// straddle the eye plane AND poke through the near plane undecided (+0.2 % of the pairs on the cube-map scene).
//
// Error budget.  The reference forms a clip coordinate from <= 8 rounded operations on terms whose absolute values sum
// to at most S = sum_r (|p_r|+|a_r|+|b_r|+|c_r|) sum_c |P[r][c]|; a compare x <= -w involves two coordinates: <= 16u S,
// u = 2^-24.  The filter's centre (4 roundings), clip coordinates (3 fused steps on <= S) and the combinations with
// r rho and m add < 14u S.  For an affine OBB the w row of P only ever multiplies p_w = 1, so
//   S <= aw sum_{r<3} sum_c |P[r][c]| + sum_c |P[3][c]|      with
//   aw = max( |cx| + |cy| + |cz| + 3.25 r , 1 )  >=  max_{r<3} (|p_r|+|a_r|+|b_r|+|c_r|)
// (p = centre - (a+b+c)/2 and |a_r| <= |a|_2: each component sum is below |centre_r| + 1.5 (|a|+|b|+|c|) <= |centre_r|
// + 3 r).  Used: m = aw q + qw with q = 2^-17 sum_{r<3} sum_c |P[r][c]| and qw = 2^-17 sum_c |P[3][c]| (rounded up), so
// m >= 128u S: four times the sum of both errors.  (Keeping the translation row out of the product with aw matters: with
// the eye 150 units off the scene's centre that row sums to ~600, and charged against aw ~ 2000 it made the margin
// 40 times larger than needed - a quarter of all pairs undecided, 1.25 -> 1.8 ms per 64 Mi x 6 cull.)
// Underflow: every product may also lose up to 2^-150 absolutely.  The host disables the filter for a view whose
// sum sum |P| is below 2^-100 (q = +inf), so m >= q + qw >= 2^-117 always covers it; q is likewise +inf for a view-projection
// with a non-finite entry or a sum above 2^39, and aw is +inf for objects that are not affine, not finite or beyond
// 2^40 - then m = +inf, every comparison below is false and the pair is undecided.  (aw carries every NaN of the OBB:
// it is a SUM of the centre coordinates and the radius, and the select below uses !(aw < 2^40).)
#pragma once

#include "cull_views.cuh"

namespace dpcu
{
  // per pair of views (u, v), every entry = (value for u, value for v); computed on the host in double precision
  struct ViewPairFilter
  {
    float2 k[16];      // k[4 r + c] = P[r][c]: centre.x/y/z multiply rows 0..2, row 3 is the constant term
    float2 rhoN[3];    // |(col_a + col_w).xyz|_2, rounded up
    float2 nrhoP[3];   // -|(col_w - col_a).xyz|_2, rounded away from zero
    float2 nrhoW;      // -|col_w.xyz|_2, rounded away from zero
    float2 q;          // 2^-17 sum_{r<3} sum_c |P[r][c]|, rounded up; +inf: this view is never decided by the filter
    float2 qw;         // 2^-17 sum_c |P[3][c]|, rounded up
  };

  __device__ __forceinline__ f32x2 sub2( f32x2 a, f32x2 b )
  {
    f32x2 r;
    asm( "sub.rn.f32x2 %0, %1, %2;" : "=l"( r ) : "l"( a ), "l"( b ) );
    return r;
  }
  __device__ __forceinline__ f32x2 fmaPair( f32x2 a, f32x2 b, f32x2 c )
  {
    f32x2 r;
    asm( "fma.rn.f32x2 %0, %1, %2, %3;" : "=l"( r ) : "l"( a ), "l"( b ), "l"( c ) );
    return r;
  }
  __device__ __forceinline__ f32x2 asPair( float2 v )
  {
    return pack2( v.x, v.y );
  }
  __device__ __forceinline__ float sqrtApproxFtz( float x )
  {
    float r;
    asm( "sqrt.approx.ftz.f32 %0, %1;" : "=f"( r ) : "f"( x ) );    // relative error < 2^-22; flushed inputs are covered by the radius slack
    return r;
  }

  // what the filter needs of an object, once for all views
  struct ObbBall
  {
    float cx, cy, cz;      // centre
    float r;               // >= half diagonal
    float aw;              // error scale (see above); +inf when the filter must not decide this object
  };

  // `half` = 0.5f from the kernel arguments (a filter constant: tests/test_sass.py tells the filter's fused
  // multiply-adds from the reference arithmetic by their constant operands)
  __device__ __forceinline__ ObbBall makeBall( Obb const &o, float half )
  {
    ObbBall s;
    s.cx = fmaf( ( o.ax.x + o.ay.x ) + o.az.x, half, o.pt.x );
    s.cy = fmaf( ( o.ax.y + o.ay.y ) + o.az.y, half, o.pt.y );
    s.cz = fmaf( ( o.ax.z + o.ay.z ) + o.az.z, half, o.pt.z );
    const float la = sqrtApproxFtz( fmaf( o.ax.x, o.ax.x, fmaf( o.ax.y, o.ax.y, o.ax.z * o.ax.z ) ) );
    const float lb = sqrtApproxFtz( fmaf( o.ay.x, o.ay.x, fmaf( o.ay.y, o.ay.y, o.ay.z * o.ay.z ) ) );
    const float lc = sqrtApproxFtz( fmaf( o.az.x, o.az.x, fmaf( o.az.y, o.az.y, o.az.z * o.az.z ) ) );
    // 0.5 * (1 + 2^-20): rounded up past the sqrt / add errors; 2^-60: components whose squares underflowed
    s.r = ( ( la + lb ) + lc ) * 0.50000048f + 8.7e-19f;
    const float t = ( ( fabsf( s.cx ) + fabsf( s.cy ) ) + fabsf( s.cz ) ) + 3.25f * s.r;
    const bool affine = o.pt.w == 1.0f && o.ax.w == 0.0f && o.ay.w == 0.0f && o.az.w == 0.0f;
    const bool sane   = t < 1.0995116e12f;                 // 2^40; false for NaN
    s.aw = ( affine && sane ) ? fmaxf( t, 1.0f ) : __int_as_float( 0x7f800000 );
    return s;
  }

  // (V), (N), (P) for two views at once.  vis / inv : proven visible / proven invisible, [0] = view u, [1] = view v.
  __device__ __forceinline__ void classifyPair( ObbBall const &s, ViewPairFilter const &f, bool ( &vis )[2], bool ( &inv )[2] )
  {
    const f32x2 cx = pack2( s.cx, s.cx ), cy = pack2( s.cy, s.cy ), cz = pack2( s.cz, s.cz ), rr = pack2( s.r, s.r );
    f32x2 A[4];
#pragma unroll
    for ( int c = 0; c < 4; ++c )
    {
      A[c] = fmaPair( cx, asPair( f.k[c] ), fmaPair( cy, asPair( f.k[4 + c] ), fmaPair( cz, asPair( f.k[8 + c] ), asPair( f.k[12 + c] ) ) ) );
    }
    const f32x2 m  = fmaPair( pack2( s.aw, s.aw ), asPair( f.q ), asPair( f.qw ) );
    const f32x2 hi = add2( A[3], m );                      // W + m
    const f32x2 lo = sub2( A[3], m );                      // W - m
    const f32x2 fr = fmaPair( rr, asPair( f.nrhoW ), A[3] );   // W - r rhoW
    f32x2 tN[3], vP[3];
#pragma unroll
    for ( int a = 0; a < 3; ++a )
    {
      tN[a] = fmaPair( rr, asPair( f.rhoN[a] ), A[a] );
      vP[a] = fmaPair( rr, asPair( f.nrhoP[a] ), A[a] );
    }
    float ax[2], ay[2], az[2], n0[2], n1[2], n2[2], p0[2], p1[2], p2[2], mm[2], h[2], l[2], front[2];
    unpack2( A[0], ax[0], ax[1] ); unpack2( A[1], ay[0], ay[1] ); unpack2( A[2], az[0], az[1] );
    unpack2( tN[0], n0[0], n0[1] ); unpack2( tN[1], n1[0], n1[1] ); unpack2( tN[2], n2[0], n2[1] );
    unpack2( vP[0], p0[0], p0[1] ); unpack2( vP[1], p1[0], p1[1] ); unpack2( vP[2], p2[0], p2[1] );
    unpack2( m, mm[0], mm[1] ); unpack2( hi, h[0], h[1] ); unpack2( lo, l[0], l[1] ); unpack2( fr, front[0], front[1] );
#pragma unroll
    for ( int e = 0; e < 2; ++e )
    {
      // when m is finite every value here is finite (|value| <= 2^18 S), so no NaN hides behind fminf / fmaxf
      vis[e] = fmaxf( fabsf( ax[e] ), fmaxf( fabsf( ay[e] ), fabsf( az[e] ) ) ) < l[e];
      const bool n = fminf( n0[e], fminf( n1[e], n2[e] ) ) < -h[e];
      const bool p = ( fmaxf( p0[e], fmaxf( p1[e], p2[e] ) ) > h[e] ) & ( front[e] > mm[e] );
      inv[e] = n | p;
    }
  }
}
