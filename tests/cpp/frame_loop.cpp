// TEST INFRASTRUCTURE - a frame loop written the way nvpro-pipeline's own code uses these classes
// (dp::sg::xbar::TransformTree::compute + CullingImpl::cull, dp/sg/xbar/src/TransformTree.cpp:92-100,
// dp/sg/xbar/culling/src/CullingImpl.cpp:126-177), run twice side by side:
//
//   reference stack:  dp::transform::Tree        -> groupSetMatrices( host world matrices ) -> dp::culling::cpu::Manager
//   new stack:        dp::transform::cuda::Tree  -> groupSetDeviceMatrices( device world )  -> dp::culling::cuda::Manager
//
// Every frame both stacks must report the same changed objects (same order), the same visibility for every
// object, the same bounding box and the same world matrices.  Exit code 0 = identical on all frames.
// Built by tests/cpp/Makefile against the reference headers where they lie; run by tests/test_dropin_manager.py.
#include <dp/culling/cpu/Manager.h>
#include <dp/culling/cuda/Manager.h>
#include <dp/transform/Tree.h>
#include <dp/transform/cuda/Tree.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

namespace
{
  struct Id : public dp::culling::Payload
  {
    explicit Id( size_t i ) : index( i ) {}
    size_t index;
  };

  uint64_t g_state = 0x5EED0003ull;
  float rnd()                                   // splitmix64 -> [0, 1)
  {
    uint64_t z = ( g_state += 0x9E3779B97F4A7C15ull );
    z = ( z ^ ( z >> 30 ) ) * 0xBF58476D1CE4E5B9ull;
    z = ( z ^ ( z >> 27 ) ) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return float( z >> 40 ) * ( 1.0f / 16777216.0f );
  }

  dp::math::Mat44f placement( float spread, float angle )
  {
    float c = std::cos( angle ), s = std::sin( angle );
    float m[16] = { c, s, 0, 0,   -s, c, 0, 0,   0, 0, 1, 0,
                    ( rnd() - 0.5f ) * spread, ( rnd() - 0.5f ) * spread, ( rnd() - 0.5f ) * spread, 1 };
    dp::math::Mat44f r;
    for ( int i = 0; i < 4; ++i ) for ( int j = 0; j < 4; ++j ) r[i][j] = m[4 * i + j];
    return r;
  }

  dp::math::Mat44f frustum( float eyeZ, float halfWidth )
  {
    // view: translate by -eyeZ along z; projection: symmetric frustum near 1, far 400 (dp/math/Matmnt.h:1324-1344)
    float n = 1.0f, f = 400.0f, w = halfWidth;
    float proj[16] = { n / w, 0, 0, 0,   0, n / w, 0, 0,   0, 0, -( f + n ) / ( f - n ), -1,   0, 0, -2 * f * n / ( f - n ), 0 };
    dp::math::Mat44f view = dp::math::cIdentity44f, p;
    view[3][2] = -eyeZ;
    for ( int i = 0; i < 4; ++i ) for ( int j = 0; j < 4; ++j ) p[i][j] = proj[4 * i + j];
    return view * p;
  }

  struct Stack
  {
    std::unique_ptr<dp::culling::Manager>     manager;
    dp::culling::GroupSharedPtr               group;
    dp::culling::ResultSharedPtr              result;
    std::vector<dp::culling::ObjectSharedPtr> objects;
  };
}

int main()
{
  try
  {
    dp::transform::Tree       hostTree;
    dp::transform::cuda::Tree deviceTree( 0 );

    // three levels: 8 / 64 / 4096 transforms; one object per leaf
    std::vector<dp::transform::Index> level0, level1, leaves;
    for ( int i = 0; i < 8; ++i )
    {
      dp::math::Mat44f m = placement( 120.0f, rnd() );
      dp::transform::Index a = hostTree.addTransform( hostTree.getRoot(), m ), b = deviceTree.addTransform( deviceTree.getRoot(), m );
      if ( a != b ) { std::printf( "index allocation differs\n" ); return 2; }
      level0.push_back( a );
    }
    for ( size_t p = 0; p < level0.size(); ++p ) for ( int i = 0; i < 8; ++i )
    {
      dp::math::Mat44f m = placement( 40.0f, rnd() );
      level1.push_back( hostTree.addTransform( level0[p], m ) );
      deviceTree.addTransform( level0[p], m );
    }
    for ( size_t p = 0; p < level1.size(); ++p ) for ( int i = 0; i < 64; ++i )
    {
      dp::math::Mat44f m = placement( 12.0f, rnd() );
      leaves.push_back( hostTree.addTransform( level1[p], m ) );
      deviceTree.addTransform( level1[p], m );
    }

    Stack ref, dev;
    ref.manager.reset( dp::culling::cpu::Manager::create() );
    dp::culling::cuda::Manager * cudaManager = dp::culling::cuda::Manager::create( 0 );
    dev.manager.reset( cudaManager );
    Stack * stacks[2] = { &ref, &dev };
    for ( Stack * s : stacks )
    {
      s->group = s->manager->groupCreate();
      for ( size_t i = 0; i < leaves.size(); ++i )
      {
        dp::culling::ObjectSharedPtr o = s->manager->objectCreate( std::make_shared<Id>( i ) );
        float h = 0.25f + float( i % 7 ) * 0.25f;
        s->manager->objectSetBoundingBox( o, dp::math::Box3f( dp::math::Vec3f( -h, -h, -h ), dp::math::Vec3f( h, h, h ) ) );
        s->manager->objectSetTransformIndex( o, leaves[i] );
        s->manager->groupAddObject( s->group, o );
        s->objects.push_back( o );
      }
      s->result = s->manager->groupCreateResult( s->group );
    }

    size_t totalChanged = 0;
    for ( int frame = 0; frame < 6; ++frame )
    {
      // animate: frame 1 a few inner transforms, frame 2 every leaf, frame 3 nothing, frame 4 one top-level transform
      std::vector<dp::transform::Index> touched;
      if ( frame == 1 ) for ( size_t i = 0; i < level1.size(); i += 9 ) touched.push_back( level1[i] );
      if ( frame == 2 ) touched = leaves;
      if ( frame == 4 ) touched.push_back( level0[3] );
      for ( size_t i = 0; i < touched.size(); ++i )
      {
        dp::math::Mat44f m = placement( frame == 2 ? 12.0f : 60.0f, 0.1f * float( frame ) + rnd() );
        hostTree.updateLocalMatrix( touched[i], m );
        deviceTree.updateLocalMatrix( touched[i], m );
      }
      hostTree.compute( dp::math::cIdentity44f );
      deviceTree.compute( dp::math::cIdentity44f );
      size_t const count = hostTree.getTransformCount();
      if ( memcmp( hostTree.getWorldMatrices(), deviceTree.getWorldMatrices(), ( 1 + 8 + 64 + 4096 ) * sizeof(dp::math::Mat44f) ) )
      {
        std::printf( "frame %d: world matrices differ\n", frame );
        return 3;
      }

      // the reference's culler keeps OBBs of unchanged matrices (GroupBitSet.h:140-150): flag what compute() recomputed
      ref.manager->groupSetMatrices( ref.group, hostTree.getWorldMatrices(), count, sizeof(dp::math::Mat44f) );
      if ( frame == 0 )
      {
        for ( size_t i = 0; i < count; ++i ) ref.manager->groupMatrixChanged( ref.group, i );
      }
      else
      {
        // (a superset of what changed: the touched transforms and everything that can hang below them)
        for ( size_t i = 0; i < level1.size(); ++i ) ref.manager->groupMatrixChanged( ref.group, level1[i] );
        for ( size_t i = 0; i < leaves.size(); ++i ) ref.manager->groupMatrixChanged( ref.group, leaves[i] );
      }
      cudaManager->groupSetDeviceMatrices( dev.group, deviceTree.getDeviceWorldMatrices(), count );

      dp::math::Mat44f vp = frustum( 150.0f - 20.0f * float( frame ), 0.35f + 0.05f * float( frame ) );
      ref.manager->cull( ref.group, ref.result, vp );
      dev.manager->cull( dev.group, dev.result, vp );

      std::vector<dp::culling::ObjectSharedPtr> const & a = ref.manager->resultGetChanged( ref.result );
      std::vector<dp::culling::ObjectSharedPtr> const & b = dev.manager->resultGetChanged( dev.result );
      if ( a.size() != b.size() ) { std::printf( "frame %d: %zu vs %zu changed objects\n", frame, a.size(), b.size() ); return 4; }
      for ( size_t i = 0; i < a.size(); ++i )
      {
        size_t ia = std::static_pointer_cast<Id>( ref.manager->objectGetUserData( a[i] ) )->index;
        size_t ib = std::static_pointer_cast<Id>( dev.manager->objectGetUserData( b[i] ) )->index;
        if ( ia != ib ) { std::printf( "frame %d: changed list differs at %zu (%zu vs %zu)\n", frame, i, ia, ib ); return 5; }
      }
      size_t visible = 0;
      for ( size_t i = 0; i < leaves.size(); ++i )
      {
        bool va = ref.manager->resultObjectIsVisible( ref.result, ref.objects[i] );
        bool vb = dev.manager->resultObjectIsVisible( dev.result, dev.objects[i] );
        if ( va != vb ) { std::printf( "frame %d: visibility of object %zu differs\n", frame, i ); return 6; }
        visible += va;
      }
      dp::math::Box3f ba = ref.manager->getBoundingBox( ref.group ), bb = dev.manager->getBoundingBox( dev.group );
      if ( memcmp( &ba, &bb, sizeof ba ) ) { std::printf( "frame %d: bounding boxes differ\n", frame ); return 7; }
      totalChanged += a.size();
      std::printf( "frame %d: %zu visible of %zu, %zu changed - identical\n", frame, visible, leaves.size(), a.size() );
    }
    if ( totalChanged == 0 ) { std::printf( "nothing ever changed: the scenario is not exercising the path\n" ); return 8; }
    ref.result.reset(); dev.result.reset();
    std::printf( "ok\n" );
    return 0;
  }
  catch ( std::exception const & e )
  {
    std::printf( "exception: %s\n", e.what() );
    return 1;
  }
}
