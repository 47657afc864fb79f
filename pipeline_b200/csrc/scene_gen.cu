// Bench utility: on-device replay of pipeline_b200/scenes.py::random_objects (SURVEY.md 8d,
// "the generator may run on-GPU from the same integer stream ... provided a host replay of any
// slice matches bit-for-bit").  Mirrors the numpy expressions operation for operation;
// compiled with -fmad=false, IEEE sqrt and division (nvcc defaults -prec-sqrt=true -prec-div=true).
#include "dpcu_internal.h"

namespace dpcu
{
  __device__ __forceinline__ float uniformAt( uint64_t seed, uint64_t k )
  {
    uint64_t z = seed + ( k + 1ull ) * 0x9E3779B97F4A7C15ull;
    z = ( z ^ ( z >> 30 ) ) * 0xBF58476D1CE4E5B9ull;
    z = ( z ^ ( z >> 27 ) ) * 0x94D049BB133111EBull;
    z = z ^ ( z >> 31 );
    return float( uint32_t( z >> 40 ) ) * 5.9604644775390625e-08f;   // 2^-24
  }

  __global__ void __launch_bounds__( 256 ) sceneGenerateKernel( uint64_t seed, uint64_t first, uint32_t count, uint32_t indexBase,
                                                                float4 *lower, float4 *extent, float4 *mats )
  {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if ( j >= count ) return;
    const uint64_t i = first + j;
    const uint64_t base = i * 16ull;
    float u[13];
#pragma unroll
    for ( int k = 0; k < 13; ++k ) u[k] = uniformAt( seed, base + k );

    const float hx = 0.5f + 4.5f * u[0], hy = 0.5f + 4.5f * u[1], hz = 0.5f + 4.5f * u[2];
    lower[j]  = make_float4( -hx, -hy, -hz, __uint_as_float( uint32_t( i ) - indexBase ) );
    extent[j] = make_float4( hx - ( -hx ), hy - ( -hy ), hz - ( -hz ), 0.0f );

    float qx = 2.0f * u[3] - 1.0f, qy = 2.0f * u[4] - 1.0f, qz = 2.0f * u[5] - 1.0f, qw = 2.0f * u[6] - 1.0f;
    float ln = sqrtf( ( ( qx * qx + qy * qy ) + qz * qz ) + qw * qw );
    const bool zero = ( ln == 0.0f );
    if ( zero ) ln = 1.0f;
    qx = qx / ln; qy = qy / ln; qz = qz / ln; qw = qw / ln;
    if ( zero ) qw = 1.0f;

    const float sx = 0.5f + 1.5f * u[7], sy = 0.5f + 1.5f * u[8], sz = 0.5f + 1.5f * u[9];
    const float tx = -1000.0f + 2000.0f * u[10], ty = -1000.0f + 2000.0f * u[11], tz = -1000.0f + 2000.0f * u[12];

    const float xx = qx * qx, yy = qy * qy, zz = qz * qz;
    const float xy = qx * qy, xz = qx * qz, yz = qy * qz;
    const float wx = qw * qx, wy = qw * qy, wz = qw * qz;

    float4 *m = mats + 4ull * j;
    m[0] = make_float4( sx * ( 1.0f - 2.0f * ( yy + zz ) ), sx * ( 2.0f * ( xy + wz ) ), sx * ( 2.0f * ( xz - wy ) ), 0.0f );
    m[1] = make_float4( sy * ( 2.0f * ( xy - wz ) ), sy * ( 1.0f - 2.0f * ( xx + zz ) ), sy * ( 2.0f * ( yz + wx ) ), 0.0f );
    m[2] = make_float4( sz * ( 2.0f * ( xz + wy ) ), sz * ( 2.0f * ( yz - wx ) ), sz * ( 1.0f - 2.0f * ( xx + yy ) ), 0.0f );
    m[3] = make_float4( tx, ty, tz, 1.0f );
  }
}

extern "C" int dpcuSceneGenerate( uint64_t seed, uint64_t first, size_t count, uint32_t indexBase, float *lower4Device,
                                  float *extent4Device, float *matricesDevice, dpcuStream *stream )
{
  DPCU_TRY( dpcu::requireDevice() );
  DPCU_REQUIRE( count < ( size_t( 1 ) << 32 ), "count must fit 32 bits" );
  DPCU_REQUIRE( !count || ( lower4Device && extent4Device && matricesDevice ), "NULL output" );
  if ( !count ) return DPCU_OK;
  dpcu::sceneGenerateKernel<<<unsigned( dpcu::divUp( count, 256 ) ), 256, 0, stream ? stream->stream : 0>>>(
    seed, first, uint32_t( count ), indexBase, reinterpret_cast<float4 *>( lower4Device ),
    reinterpret_cast<float4 *>( extent4Device ), reinterpret_cast<float4 *>( matricesDevice ) );
  DPCU_CUDA( cudaGetLastError() );
  return DPCU_OK;
}

// Bench support: read `bytes` of device memory once (streaming loads, result discarded).  After bench.py's L2 flush
// WRITE (a buffer larger than L2) this sweep over a second buffer leaves the L2 cold AND clean, so that the write-back
// of the flush's own dirty lines does not land inside the next timed step.
namespace dpcu
{
  __global__ void __launch_bounds__( 256 ) readSweepKernel( uint4 const *p, size_t n, uint32_t *sink )
  {
    uint32_t acc = 0;
    for ( size_t i = size_t( blockIdx.x ) * blockDim.x + threadIdx.x; i < n; i += size_t( gridDim.x ) * blockDim.x )
    {
      const uint4 v = __ldcg( p + i );
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if ( acc == 0x9e3779b9u ) *sink = acc;                   // keeps the loads alive; practically never taken
  }
}

extern "C" int dpcuDebugReadSweep( const void *deviceMemory, size_t bytes, dpcuStream *stream )
{
  DPCU_TRY( dpcu::requireDevice() );
  DPCU_REQUIRE( deviceMemory && bytes >= 16 && ( reinterpret_cast<uintptr_t>( deviceMemory ) & 15 ) == 0, "need a 16-byte aligned device buffer" );
  int device = 0, sms = 1;
  DPCU_CUDA( cudaGetDevice( &device ) );
  cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, device );
  uint32_t *sink = reinterpret_cast<uint32_t *>( const_cast<void *>( deviceMemory ) );
  dpcu::readSweepKernel<<<sms * 8, 256, 0, stream ? stream->stream : 0>>>( static_cast<uint4 const *>( deviceMemory ), bytes / 16, sink );
  DPCU_CUDA( cudaGetLastError() );
  return DPCU_OK;
}
