// TEST INFRASTRUCTURE - a frame loop through the PATCHED dp::sg::xbar::culling::CullingImpl
// (patches/0001-culling-cuda-backend.patch applied to a copy of the reference's CullingImpl.cpp / CullingImpl.h by
// tests/cpp/Makefile, compiled against stand-in SceneTree / GeoNode headers in tests/cpp/stubs), i.e. through the
// code path the apps take:  Culling::create( sceneTree, Mode )  ->  cull  ->  resultGetChangedIndices / resultIsVisible.
//
//   reference stack:  dp::transform::Tree        + Culling::create( scene, Mode::CPU )
//   new stack:        dp::transform::cuda::Tree  + Culling::create( scene, Mode::CUDA )   (case Mode::CUDA of the patch,
//                     device-resident matrix feed, TransformObserver attached)
//
// Scene edits go through SceneTree events (ADDED / REMOVED / CHANGED, CullingImpl.cpp:180-206).  Every frame both
// stacks must report the same changed ObjectTree indices (same order), the same visibility for every live object
// and the same bounding box.  Exit code 0 = identical on all frames.
#include <dp/sg/xbar/culling/Culling.h>
#include <dp/culling/opengl/Manager.h>
#include <dp/transform/Tree.h>
#include <dp/transform/cuda/Tree.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <vector>

// the GL backend is not part of this test; CullingImpl.cpp refers to its factory
dp::culling::opengl::Manager * dp::culling::opengl::Manager::create()
{
  throw std::runtime_error( "dp::culling::opengl is not linked into this test" );
}

namespace
{
  uint64_t g_state = 0x5EED00A5ull;
  float rnd()
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
    float n = 1.0f, f = 400.0f, w = halfWidth;
    float proj[16] = { n / w, 0, 0, 0,   0, n / w, 0, 0,   0, 0, -( f + n ) / ( f - n ), -1,   0, 0, -2 * f * n / ( f - n ), 0 };
    dp::math::Mat44f view = dp::math::cIdentity44f, p;
    view[3][2] = -eyeZ;
    for ( int i = 0; i < 4; ++i ) for ( int j = 0; j < 4; ++j ) p[i][j] = proj[4 * i + j];
    return view * p;
  }

  dp::math::Box3f box( float h )
  {
    return dp::math::Box3f( dp::math::Vec3f( -h, -h * 0.5f, -h ), dp::math::Vec3f( h, h * 1.5f, h * 0.75f ) );
  }
}

int main()
{
  using namespace dp::sg::xbar;
  try
  {
    dp::transform::Tree       hostTree;
    dp::transform::cuda::Tree deviceTree( 0 );
    dp::transform::Tree *     trees[2] = { &hostTree, &deviceTree };

    std::vector<dp::transform::Index> level0, level1, leaves;
    for ( int i = 0; i < 8; ++i )
    {
      dp::math::Mat44f m = placement( 120.0f, rnd() );
      level0.push_back( hostTree.addTransform( hostTree.getRoot(), m ) );
      deviceTree.addTransform( deviceTree.getRoot(), m );
    }
    for ( size_t p = 0; p < level0.size(); ++p ) for ( int i = 0; i < 8; ++i )
    {
      dp::math::Mat44f m = placement( 40.0f, rnd() );
      level1.push_back( hostTree.addTransform( level0[p], m ) );
      deviceTree.addTransform( level0[p], m );
    }
    for ( size_t p = 0; p < level1.size(); ++p ) for ( int i = 0; i < 32; ++i )
    {
      dp::math::Mat44f m = placement( 12.0f, rnd() );
      leaves.push_back( hostTree.addTransform( level1[p], m ) );
      deviceTree.addTransform( level1[p], m );
    }

    SceneTreeSharedPtr scenes[2] = { SceneTree::create( hostTree ), SceneTree::create( deviceTree ) };
    // the initial drawables exist before the culling object does: CullingImpl's constructor traverses the object tree
    std::vector<ObjectTreeIndex> alive;
    for ( size_t i = 0; i < leaves.size(); ++i )
    {
      ObjectTreeIndex a = scenes[0]->addDrawable( box( 0.25f + float( i % 7 ) * 0.25f ), leaves[i] );
      ObjectTreeIndex b = scenes[1]->addDrawable( box( 0.25f + float( i % 7 ) * 0.25f ), leaves[i] );
      if ( a != b ) { std::printf( "object tree indices differ\n" ); return 2; }
      alive.push_back( a );
    }

    culling::CullingSharedPtr cull[2] = { culling::Culling::create( scenes[0], dp::culling::Mode::CPU )
                                        , culling::Culling::create( scenes[1], dp::culling::Mode::CUDA ) };
    culling::ResultSharedPtr result[2] = { cull[0]->resultCreate(), cull[1]->resultCreate() };

    size_t totalChanged = 0;
    for ( int frame = 0; frame < 8; ++frame )
    {
      // ---- animation (dirty world matrices reach the culler through the TransformObserver the patch attaches)
      std::vector<dp::transform::Index> touched;
      if ( frame == 1 ) for ( size_t i = 0; i < level1.size(); i += 9 ) touched.push_back( level1[i] );
      if ( frame == 2 ) touched = leaves;
      if ( frame == 4 ) touched.push_back( level0[3] );
      if ( frame >= 5 ) for ( size_t i = frame; i < leaves.size(); i += 17 ) touched.push_back( leaves[i] );
      for ( size_t i = 0; i < touched.size(); ++i )
      {
        dp::math::Mat44f m = placement( frame == 2 ? 12.0f : 60.0f, 0.1f * float( frame ) + rnd() );
        for ( dp::transform::Tree * t : trees ) t->updateLocalMatrix( touched[i], m );
      }
      // ---- scene edits through SceneTree events
      if ( frame == 2 )
      {
        for ( int i = 0; i < 500; ++i )
        {
          dp::transform::Index leaf = leaves[size_t( rnd() * float( leaves.size() - 1 ) )];
          float h = 0.2f + rnd();
          ObjectTreeIndex a = scenes[0]->addDrawable( box( h ), leaf ), b = scenes[1]->addDrawable( box( h ), leaf );
          if ( a != b ) { std::printf( "object tree indices differ\n" ); return 2; }
          alive.push_back( a );
        }
      }
      if ( frame == 3 || frame == 6 )
      {
        for ( int i = 0; i < 300; ++i )
        {
          size_t k = size_t( rnd() * float( alive.size() - 1 ) );
          for ( int s = 0; s < 2; ++s ) scenes[s]->removeDrawable( alive[k] );
          alive[k] = alive.back();
          alive.pop_back();
        }
      }
      if ( frame == 5 || frame == 6 )
      {
        for ( int i = 0; i < 200; ++i )
        {
          size_t k = size_t( rnd() * float( alive.size() - 1 ) );
          float h = 0.5f + 4.0f * rnd();
          for ( int s = 0; s < 2; ++s ) scenes[s]->changeBoundingBox( alive[k], box( h ) );
        }
      }
      for ( dp::transform::Tree * t : trees ) t->compute( dp::math::cIdentity44f );

      dp::math::Mat44f vp = frustum( 150.0f - 15.0f * float( frame ), 0.35f + 0.04f * float( frame ) );
      for ( int s = 0; s < 2; ++s ) cull[s]->cull( result[s], vp );

      std::vector<ObjectTreeIndex> const & a = cull[0]->resultGetChangedIndices( result[0] );
      std::vector<ObjectTreeIndex> const & b = cull[1]->resultGetChangedIndices( result[1] );
      if ( a.size() != b.size() ) { std::printf( "frame %d: %zu vs %zu changed objects\n", frame, a.size(), b.size() ); return 4; }
      for ( size_t i = 0; i < a.size(); ++i )
      {
        if ( a[i] != b[i] ) { std::printf( "frame %d: changed list differs at %zu (%u vs %u)\n", frame, i, a[i], b[i] ); return 5; }
      }
      size_t visible = 0;
      for ( size_t i = 0; i < alive.size(); ++i )
      {
        bool va = cull[0]->resultIsVisible( result[0], alive[i] ), vb = cull[1]->resultIsVisible( result[1], alive[i] );
        if ( va != vb ) { std::printf( "frame %d: visibility of object %u differs\n", frame, alive[i] ); return 6; }
        visible += va;
      }
      dp::math::Box3f ba = cull[0]->getBoundingBox(), bb = cull[1]->getBoundingBox();
      if ( memcmp( &ba, &bb, sizeof ba ) ) { std::printf( "frame %d: bounding boxes differ\n", frame ); return 7; }
      totalChanged += a.size();
      std::printf( "frame %d: %zu visible of %zu, %zu changed - identical\n", frame, visible, alive.size(), a.size() );
    }
    if ( totalChanged == 0 ) { std::printf( "nothing ever changed: the scenario is not exercising the path\n" ); return 8; }
    for ( int s = 0; s < 2; ++s ) { result[s].reset(); cull[s].reset(); }
    std::printf( "ok\n" );
    return 0;
  }
  catch ( std::exception const & e )
  {
    std::printf( "exception: %s\n", e.what() );
    return 1;
  }
}
