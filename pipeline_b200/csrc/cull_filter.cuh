// Multi-view cull with an exact-by-construction filter in front of the reference arithmetic.
//
// cull_views.cuh evaluates the reference's eight-corner test for every (object, view) pair:
// ~160 warp instructions per pair, which makes a six-view cull issue bound at 0.4 of the HBM
// roofline.  Almost every pair is far from the frustum boundary, where the outcome can be PROVEN
// from the OBB's centre and radius with a few instructions; only the pairs that are not provable
// run the reference arithmetic.  The proof obligations (this is what keeps the result bit-exact):
//
//   The reference (dp/culling/cpu/src/ManagerImpl.cpp:199-289) computes, in binary32, the clip
//   coordinates (x_j, y_j, z_j, w_j) of the eight corners j of the OBB {p; a, b, c} and sets, per
//   axis, out-code bit N when x_j <= -w_j, else bit P when w_j <= x_j; the object is invisible
//   iff some bit is set for all eight corners.  For finite operands
//        x <= -w   <=>   x + w <= 0      and      w <= x   <=>   w - x <= 0      (real sums)
//   so with the plane functions  fN = x + w,  fP = w - x  (linear in the corner):
//     (V) if the real value of every plane function at the box centre exceeds the rounding error E
//         of the reference's corner computation, then for every plane SOME float corner has
//         f_j > 0 (the centre is the mean of the corners), i.e. no out-code bit survives all
//         corners: visible;
//     (N) if  fN(centre) + support_N < -E  every float corner has x_j + w_j < 0: bit N everywhere,
//         invisible;
//     (P) if  fP(centre) + support_P < -E  and  fN(centre) - support_N > E  every corner has
//         w_j <= x_j and not x_j <= -w_j: bit P everywhere, invisible.
//   E: the reference forms every clip coordinate from at most 8 rounded operations on terms whose
//   absolute values sum to at most  S = sum_r (|p_r| + |a_r| + |b_r| + |c_r|) * sum_c |P[r][c]|,
//   so its error is below 8u S / (1 - 8u), u = 2^-24 (two coordinates enter a compare: 16u S); the filter's own centre /
//   plane / support arithmetic adds less than another 14u S.  The margin used is 2^-17 S' with
//   S' = max_{r<3} (|p_r| + |a_r| + |b_r| + |c_r|) * sum_{r<3} sum_c |P[r][c]| + sum_c |P[3][c]| >= S  (the w row of P
//   only ever multiplies p_w = 1 of an affine OBB; four times the sum of the two errors; one fused multiply-add per
//   object and view); supports are bounded by Cauchy-Schwarz with radius and plane
//   norms rounded up.  The centre values come from the centre's clip coordinates (A_x, A_y, A_z, W) =
//   centre * P:  fN_a = W + A_a,  fP_a = W - A_a,  so (V) reads  W - max_a |A_a| > margin.
//   A decision is taken only when a comparison is TRUE and only when S' < 2^96 (nothing overflowed,
//   every intermediate is finite - so neither a NaN comparison nor a NaN dropped by fmin / fmax can
//   decide).  Everything else - and every warp that holds a non-affine object or a non-finite
//   view-projection - takes the reference arithmetic of cull_views.cuh.
//
// The undecided pairs of a warp's 32 objects x NV views are queued in shared memory and gathered into
// dense batches of 32 (pair -> lane); their OBBs come back from a per-warp shared-memory park (which
// also frees the OBB registers during classification), their view-projection rows from a small
// shared-memory table; they are evaluated with the exact packed arithmetic and scattered back into
// the per-view ballot words with warp OR-reductions.
#pragma once

#include "cull_views.cuh"

namespace dpcu
{
  // per view, computed on the host in double precision from the view-projection (see makeViewFilter)
  struct ViewFilter
  {
    float4 rows[4];    // the filter's own copy of the view-projection rows: its fused multiply-adds must be
                       // distinguishable (by constant-bank address) from the reference arithmetic, which never fuses
    float rhoN[3];     // per clip axis a: |(col_a + col_w).xyz|_2 of the view-projection, rounded up (N plane: x + w)
    float rhoP[3];     //                  |(col_w - col_a).xyz|_2, rounded up                        (P plane: w - x)
    float q;           // 2^-17 * sum_{r<3} sum_c |P[r][c]|: margin = q * max(aw, 1) + qw >= 2^-17 S
    float qw;          // 2^-17 * sum_c |P[3][c]|: the translation row only ever multiplies the point's w = 1
  };

  __device__ __forceinline__ f32x2 fma2( f32x2 a, f32x2 b, f32x2 c )
  {
    f32x2 r;
    asm( "fma.rn.f32x2 %0, %1, %2, %3;" : "=l"( r ) : "l"( a ), "l"( b ), "l"( c ) );
    return r;
  }

  __device__ __forceinline__ float sqrtApprox( float x )
  {
    float r;
    asm( "sqrt.approx.ftz.f32 %0, %1;" : "=f"( r ) : "f"( x ) );    // relative error < 2^-22, flushed inputs are covered by the radius slack
    return r;
  }

  // what the filter needs of an object, once for all views
  struct ObbSphere
  {
    float cx, cy, cz;      // centre
    float r;               // >= half diagonal
    float aw;              // max over components of |p| + |a| + |b| + |c|, and 1 (the w component of p)
  };

  __device__ __forceinline__ ObbSphere makeSphere( Obb const &o )
  {
    ObbSphere s;
    s.cx = o.pt.x + 0.5f * ( ( o.ax.x + o.ay.x ) + o.az.x );
    s.cy = o.pt.y + 0.5f * ( ( o.ax.y + o.ay.y ) + o.az.y );
    s.cz = o.pt.z + 0.5f * ( ( o.ax.z + o.ay.z ) + o.az.z );
    const float la = sqrtApprox( o.ax.x * o.ax.x + o.ax.y * o.ax.y + o.ax.z * o.ax.z );
    const float lb = sqrtApprox( o.ay.x * o.ay.x + o.ay.y * o.ay.y + o.ay.z * o.ay.z );
    const float lc = sqrtApprox( o.az.x * o.az.x + o.az.y * o.az.y + o.az.z * o.az.z );
    // 0.5 * (1 + 2^-20): rounded up past the sqrt / add errors; 2^-60: components whose squares underflowed
    s.r   = ( ( la + lb ) + lc ) * 0.50000048f + 8.7e-19f;
    const float awx = ( fabsf( o.pt.x ) + fabsf( o.ax.x ) ) + ( fabsf( o.ay.x ) + fabsf( o.az.x ) );
    const float awy = ( fabsf( o.pt.y ) + fabsf( o.ax.y ) ) + ( fabsf( o.ay.y ) + fabsf( o.az.y ) );
    const float awz = ( fabsf( o.pt.z ) + fabsf( o.ax.z ) ) + ( fabsf( o.ay.z ) + fabsf( o.az.z ) );
    // S = sum_r aw_r * rowsum_r <= max_r aw_r * sum_r rowsum_r; a NaN component must not get lost in fmaxf
    s.aw = fmaxf( fmaxf( awx, awy ), fmaxf( awz, 1.0f ) ) + 0.0f * ( ( awx + awy ) + awz );
    return s;
  }

  // (V), (N), (P) above for one view: visible / invisible proven, or neither.  The plane functions at the centre
  // come from its clip coordinates (A_x, A_y, A_z, W) = centre * P, one packed dot product per column pair:
  // fN_a = W + A_a, fP_a = W - A_a, so (V) is  W - max_a |A_a| > m.
  __device__ __forceinline__ void classify( ObbSphere const &s, ViewFilter const &f, bool &visible, bool &invisible )
  {
    const ViewPairs P = loadViewPairs( f.rows );
    const f32x2 cx = pack2( s.cx, s.cx ), cy = pack2( s.cy, s.cy ), cz = pack2( s.cz, s.cz );
    const f32x2 lo = fma2( cx, P.p[0], fma2( cy, P.p[2], fma2( cz, P.p[4], P.p[6] ) ) );
    const f32x2 hi = fma2( cx, P.p[1], fma2( cy, P.p[3], fma2( cz, P.p[5], P.p[7] ) ) );
    float A[3], W;
    unpack2( lo, A[0], A[1] );
    unpack2( hi, A[2], W );
    const float m = fmaf( s.aw, f.q, f.qw );
    const bool sane = m < 6.0e23f;                         // S < 2^96 (q carries the 2^-17): nothing overflowed; false for NaN
    // every value below is finite when m is (|value| <= S), so NaN cannot hide behind fminf / fmaxf
    visible = sane & ( W - fmaxf( fabsf( A[0] ), fmaxf( fabsf( A[1] ), fabsf( A[2] ) ) ) > m );
    const float kLow = -m - W, kHigh = m - W;              // f + support < -m  <=>  a-part < kLow ;  f - support > m  <=>  a-part > kHigh
    float tN[3];
    bool  inv = false;
#pragma unroll
    for ( int a = 0; a < 3; ++a )
    {
      tN[a] = fmaf( s.r, f.rhoN[a], A[a] );                                  // fN + support, without W
      const float uP = fmaf( s.r, f.rhoP[a], -A[a] );                        // fP + support, without W
      const float vN = fmaf( -s.r, f.rhoN[a], A[a] );                        // fN - support, without W
      inv = inv | ( ( uP < kLow ) & ( vN > kHigh ) );                        // (P)
    }
    inv = inv | ( fminf( tN[0], fminf( tN[1], tN[2] ) ) < kLow );            // (N)
    invisible = sane & !visible & inv;
  }

  // All NV views of the warp's 32 objects; every lane's object is affine and the view-projections are
  // finite (the caller checked).  sP: the views' rows as ViewPairs in shared memory (8 x f32x2 per view);
  // scratch: shared memory private to the warp.
  // Returns, in lane v, the ballot word of view v - the contract of cullViews.
  template <int NV>
  struct FilterScratch
  {
    float4  obb[3][32];        // the 32 OBBs (12 floats each: pt.xyz, ax.xyz, ay.xyz, az.xyz), parked here for the undecided pairs
    uint8_t slots[32 * NV];    // undecided pairs in (view, lane) order: view << 5 | lane
  };

  template <int NV>
  __device__ __forceinline__ uint32_t cullViewsFiltered( Obb const &obb, ViewFilter const ( &vf )[NV], f32x2 const *sP,
                                                         FilterScratch<NV> &scratch, f32x2 one, bool live, uint32_t lane )
  {
    const ObbSphere s = makeSphere( obb );
    const uint32_t below = ( 1u << lane ) - 1u;
    uint32_t myWord = 0, total = 0;                        // lane v: decided-visible ballot of view v; undecided pairs so far
    uint8_t *slots = scratch.slots;
    __syncwarp();                                          // the previous call's scratch reads are done
    // the OBB leaves the registers here: only centre, radius and the error scale stay live across the views
    scratch.obb[0][lane] = make_float4( obb.pt.x, obb.pt.y, obb.pt.z, obb.ax.x );
    scratch.obb[1][lane] = make_float4( obb.ax.y, obb.ax.z, obb.ay.x, obb.ay.y );
    scratch.obb[2][lane] = make_float4( obb.ay.z, obb.az.x, obb.az.y, obb.az.z );
#ifndef DPCU_FILTER_UNROLL
#define DPCU_FILTER_UNROLL 8
#endif
    constexpr int kUnroll = DPCU_FILTER_UNROLL;
#pragma unroll kUnroll
    for ( int v = 0; v < NV; ++v )
    {
      bool vis, inv;
      classify( s, vf[v], vis, inv );
      const bool open = live & !vis & !inv;
      const uint32_t bv = __ballot_sync( 0xffffffffu, vis & live );
      const uint32_t bo = __ballot_sync( 0xffffffffu, open );
      if ( lane == uint32_t( v ) ) myWord = bv;
      // undecided pairs queue up in (view, lane) order: slot -> (view, lane that owns the object)
      if ( open ) slots[total + __popc( bo & below )] = uint8_t( ( v << 5 ) | lane );
      total += __popc( bo );
    }
    __syncwarp();
    for ( uint32_t base = 0; base < total; base += 32 )
    {
      const bool     mine = base + lane < total;
      const uint32_t slot = mine ? slots[base + lane] : 0u;
      const uint32_t view = slot >> 5, src = slot & 31u;
      // the pair's OBB, parked by the lane that owns the object (affine: pt.w == 1, a.w == b.w == c.w == 0)
      const float4 o0 = scratch.obb[0][src], o1 = scratch.obb[1][src], o2 = scratch.obb[2][src];
      Obb o;
      o.pt = make_float4( o0.x, o0.y, o0.z, 1.0f );
      o.ax = make_float4( o0.w, o1.x, o1.y, 0.0f );
      o.ay = make_float4( o1.z, o1.w, o2.x, 0.0f );
      o.az = make_float4( o2.y, o2.z, o2.w, 0.0f );
      ViewPairs P;
#pragma unroll
      for ( int k = 0; k < 8; ++k ) P.p[k] = sP[view * 8 + k];
      const bool visible = cornersVisible<true>( clipVectors<true>( broadcastObb( o ), P, one ) );
      const uint32_t bit = ( mine && visible ) ? ( 1u << src ) : 0u;
#pragma unroll
      for ( int u = 0; u < NV; ++u )
      {
        const uint32_t m = __reduce_or_sync( 0xffffffffu, view == uint32_t( u ) ? bit : 0u );
        if ( lane == uint32_t( u ) ) myWord |= m;
      }
    }
    return myWord;
  }
}
