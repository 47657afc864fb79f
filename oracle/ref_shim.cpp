// TEST INFRASTRUCTURE - not part of the product path.
//
// C shim over the UNMODIFIED reference classes (compiled from /root/reference where they
// lie; see oracle/Makefile).  It drives a dp::culling::Manager through its public virtual
// API (dp/culling/Manager.h:108-143) and a dp::transform::Tree (dp/transform/Tree.h:62-88)
// exactly the way dp::sg::xbar::culling::CullingImpl does (CullingImpl.cpp:126-177), so
// Python tests / bench.py can use the real reference as oracle and CPU baseline.
//
// The same file is compiled a second time with -DDPREF_WITH_CUDA by tests/cpp/Makefile;
// backend 1 then instantiates the new dp::culling::cuda::Manager behind the very same
// calls, which is what makes the drop-in test "same driver, two backends".
//
// Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
// load the resulting library.

#include <dp/culling/Manager.h>
#include <dp/culling/cpu/Manager.h>
#include <dp/culling/ObjectBitSet.h>
#include <dp/culling/GroupBitSet.h>
#include <dp/transform/Tree.h>
#include <dp/util/BitArray.h>

#ifdef DPREF_WITH_CUDA
#include <dp/culling/cuda/Manager.h>
#include <dp/transform/cuda/Tree.h>
#endif

#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

// dp/culling/src/Manager.cpp is not compiled (it pulls in the OpenGL backend,
// Manager.cpp:44-45); the only symbol it owns on this path is the empty destructor.
dp::culling::Manager::~Manager() {}

namespace
{
  // payload carrying the creation id of an object (CullingImpl uses the ObjectTreeIndex,
  // xbar/culling/inc/CullingImpl.h:75-99)
  class IdPayload : public dp::culling::Payload
  {
  public:
    explicit IdPayload( uint32_t id ) : m_id( id ) {}
    uint32_t m_id;
  };

  struct Session
  {
    std::unique_ptr<dp::culling::Manager>       manager;
    dp::culling::GroupSharedPtr                 group;
    std::vector<dp::culling::ResultSharedPtr>   results;
    uint32_t                                    nextId = 0;
    std::string                                 error;
  };

  // records the dirty world-matrix set a Tree publishes at the end of compute()
  // (Tree.cpp:161, event declared Tree.h:46-58) before the tree clears it
  class DirtyRecorder : public dp::util::Observer
  {
  public:
    void onNotify( dp::util::Event const & event, dp::util::Payload * ) override
    {
      auto const & e = static_cast<dp::transform::Tree::EventWorldMatricesChanged const &>( event );
      m_dirty = e.getDirtyWorldMatrices();
    }
    void onDestroyed( dp::util::Subject const &, dp::util::Payload * ) override {}
    dp::util::BitArray m_dirty;
  };

  // backend 0: dp::transform::Tree (the oracle); with DPREF_WITH_CUDA backend 1: dp::transform::cuda::Tree
  // edited through its own type (as xbar::TransformTree would hold it), backend 2: the same class edited
  // through a base-class reference (the non-virtual addTransform / removeTransform of the reference)
  struct TreeSession
  {
    std::unique_ptr<dp::transform::Tree> owner;
    dp::transform::Tree &                tree;
#ifdef DPREF_WITH_CUDA
    dp::transform::cuda::Tree *          cudaTree;
#endif
    bool                                 viaDerived;
    DirtyRecorder                        recorder;

    explicit TreeSession( dp::transform::Tree * t, bool derived = false )
      : owner( t ), tree( *t )
#ifdef DPREF_WITH_CUDA
      , cudaTree( nullptr )
#endif
      , viaDerived( derived )
    { tree.attach( &recorder ); }
    ~TreeSession() { tree.detach( &recorder ); }

    dp::transform::Index add( dp::transform::Index parent, dp::math::Mat44f const & m )
    {
#ifdef DPREF_WITH_CUDA
      if ( viaDerived ) return cudaTree->addTransform( parent, m );
#endif
      return tree.addTransform( parent, m );
    }
    void remove( dp::transform::Index index )
    {
#ifdef DPREF_WITH_CUDA
      if ( viaDerived ) { cudaTree->removeTransform( index ); return; }
#endif
      tree.removeTransform( index );
    }
  };

  inline dp::math::Mat44f toMat( float const * m )
  {
    dp::math::Mat44f r;
    for ( int i = 0; i < 4; ++i )
      for ( int j = 0; j < 4; ++j )
        r[i][j] = m[4*i+j];
    return r;
  }
}

static std::string g_createError;

#define SESSION( p ) ( static_cast<Session *>( p ) )
#define TREE( p )    ( static_cast<TreeSession *>( p ) )

extern "C"
{
  // backend: 0 = dp::culling::cpu::Manager (the oracle), 1 = dp::culling::cuda::Manager (only with DPREF_WITH_CUDA)
  void * dpref_cull_create( int backend )
  {
    try
    {
      std::unique_ptr<Session> s( new Session );
      switch ( backend )
      {
        case 0:
          s->manager.reset( dp::culling::cpu::Manager::create() );
          break;
#ifdef DPREF_WITH_CUDA
        case 1:
          s->manager.reset( dp::culling::cuda::Manager::create() );
          break;
#endif
        default:
          return nullptr;
      }
      s->group = s->manager->groupCreate();
      return s.release();
    }
    catch ( std::exception const & e )
    {
      g_createError = e.what();
      return nullptr;
    }
  }

  char const * dpref_create_error() { return g_createError.c_str(); }

  void dpref_cull_destroy( void * p )
  {
    Session * s = SESSION( p );
    if ( s )
    {
      s->results.clear();   // results detach from the group in their dtor; group must still exist (ResultBitSet.cpp:51-54)
      s->group.reset();
      delete s;
    }
  }

  char const * dpref_last_error( void * p ) { return SESSION( p )->error.c_str(); }

  // Append n objects.  Bounding box and transform index are set BEFORE groupAddObject,
  // see SURVEY.md section 7 "hard part 5" (ObjectBitSet::m_group is never set, so later
  // setters would not dirty the CPU backend's OBB cache).
  int dpref_cull_add_objects( void * p, size_t n, float const * lower3, float const * upper3, uint32_t const * transformIndex )
  {
    Session * s = SESSION( p );
    try
    {
      for ( size_t i = 0; i < n; ++i )
      {
        dp::culling::ObjectSharedPtr o = s->manager->objectCreate( std::make_shared<IdPayload>( s->nextId++ ) );
        dp::math::Box3f box( dp::math::Vec3f( lower3[3*i+0], lower3[3*i+1], lower3[3*i+2] )
                           , dp::math::Vec3f( upper3[3*i+0], upper3[3*i+1], upper3[3*i+2] ) );
        s->manager->objectSetBoundingBox( o, box );
        s->manager->objectSetTransformIndex( o, transformIndex[i] );
        s->manager->groupAddObject( s->group, o );
      }
      return 0;
    }
    catch ( std::exception const & e ) { s->error = e.what(); return 1; }
  }

  // Live edit of an object that is already in the group (the CullingImpl CHANGED event,
  // CullingImpl.cpp:198-202).  Because of the quirk above the caller has to force the OBB
  // cache dirty itself (any add / remove, or groupMatrixChanged, does).
  int dpref_cull_set_object( void * p, size_t groupIndex, float const * lower3, float const * upper3, uint32_t transformIndex )
  {
    Session * s = SESSION( p );
    try
    {
      dp::culling::ObjectSharedPtr o = s->manager->groupGetObject( s->group, groupIndex );
      dp::math::Box3f box( dp::math::Vec3f( lower3[0], lower3[1], lower3[2] ), dp::math::Vec3f( upper3[0], upper3[1], upper3[2] ) );
      s->manager->objectSetBoundingBox( o, box );
      s->manager->objectSetTransformIndex( o, transformIndex );
      return 0;
    }
    catch ( std::exception const & e ) { s->error = e.what(); return 1; }
  }

  // remove the object that currently sits at groupIndex (GroupBitSet.cpp:93-119 swaps the last one in)
  int dpref_cull_remove_object( void * p, size_t groupIndex )
  {
    Session * s = SESSION( p );
    try
    {
      dp::culling::ObjectSharedPtr o = s->manager->groupGetObject( s->group, groupIndex );
      s->manager->groupRemoveObject( s->group, o );
      return 0;
    }
    catch ( std::exception const & e ) { s->error = e.what(); return 1; }
  }

  // a frame's worth of edits in one call (what CullingImpl::onNotify does per CHANGED / REMOVED event, CullingImpl.cpp:186-202),
  // so that a test can time the backend instead of the ctypes call overhead
  int dpref_cull_set_objects_many( void * p, size_t n, uint32_t const * groupIndices, float const * lower3, float const * upper3, uint32_t const * transformIndex )
  {
    Session * s = SESSION( p );
    try
    {
      for ( size_t i = 0; i < n; ++i )
      {
        dp::culling::ObjectSharedPtr o = s->manager->groupGetObject( s->group, groupIndices[i] );
        dp::math::Box3f box( dp::math::Vec3f( lower3[3*i+0], lower3[3*i+1], lower3[3*i+2] )
                           , dp::math::Vec3f( upper3[3*i+0], upper3[3*i+1], upper3[3*i+2] ) );
        s->manager->objectSetBoundingBox( o, box );
        s->manager->objectSetTransformIndex( o, transformIndex[i] );
      }
      return 0;
    }
    catch ( std::exception const & e ) { s->error = e.what(); return 1; }
  }

  int dpref_cull_remove_objects_many( void * p, size_t n, uint32_t const * groupIndices )
  {
    Session * s = SESSION( p );
    try
    {
      for ( size_t i = 0; i < n; ++i )
      {
        dp::culling::ObjectSharedPtr o = s->manager->groupGetObject( s->group, groupIndices[i] );
        s->manager->groupRemoveObject( s->group, o );
      }
      return 0;
    }
    catch ( std::exception const & e ) { s->error = e.what(); return 1; }
  }

  size_t dpref_cull_count( void * p ) { return SESSION( p )->manager->groupGetCount( SESSION( p )->group ); }

  // creation id of the object now at groupIndex
  uint32_t dpref_cull_object_id( void * p, size_t groupIndex )
  {
    Session * s = SESSION( p );
    dp::culling::ObjectSharedPtr o = s->manager->groupGetObject( s->group, groupIndex );
    return std::static_pointer_cast<IdPayload>( s->manager->objectGetUserData( o ) )->m_id;
  }

  // matrices are BORROWED until the next cull returns (GroupBitSet.cpp:121-135)
  int dpref_cull_set_matrices( void * p, void const * matrices, size_t count, size_t strideBytes )
  {
    Session * s = SESSION( p );
    try { s->manager->groupSetMatrices( s->group, matrices, count, strideBytes ); return 0; }
    catch ( std::exception const & e ) { s->error = e.what(); return 1; }
  }

  int dpref_cull_matrix_changed( void * p, size_t index )
  {
    Session * s = SESSION( p );
    try { s->manager->groupMatrixChanged( s->group, index ); return 0; }
    catch ( std::exception const & e ) { s->error = e.what(); return 1; }
  }

  int dpref_cull_matrices_changed( void * p, uint32_t const * indices, size_t n )
  {
    Session * s = SESSION( p );
    try
    {
      for ( size_t i = 0; i < n; ++i ) s->manager->groupMatrixChanged( s->group, indices[i] );
      return 0;
    }
    catch ( std::exception const & e ) { s->error = e.what(); return 1; }
  }

  int dpref_cull_result_create( void * p )
  {
    Session * s = SESSION( p );
    try
    {
      s->results.push_back( s->manager->groupCreateResult( s->group ) );
      return int( s->results.size() ) - 1;
    }
    catch ( std::exception const & e ) { s->error = e.what(); return -1; }
  }

  int dpref_cull_run( void * p, int result, float const * viewProjection16 )
  {
    Session * s = SESSION( p );
    try
    {
      s->manager->cull( s->group, s->results[result], toMat( viewProjection16 ) );
      return 0;
    }
    catch ( std::exception const & e ) { s->error = e.what(); return 1; }
  }

  size_t dpref_cull_changed_count( void * p, int result )
  {
    Session * s = SESSION( p );
    return s->manager->resultGetChanged( s->results[result] ).size();
  }

  // changed objects as current group indices, in the order the reference reports them
  size_t dpref_cull_changed( void * p, int result, uint32_t * outGroupIndices, size_t capacity )
  {
    Session * s = SESSION( p );
    std::vector<dp::culling::ObjectSharedPtr> const & changed = s->manager->resultGetChanged( s->results[result] );
    size_t n = changed.size() < capacity ? changed.size() : capacity;
    for ( size_t i = 0; i < n; ++i )
    {
      outGroupIndices[i] = uint32_t( std::static_pointer_cast<dp::culling::ObjectBitSet>( changed[i] )->getGroupIndex() );
    }
    return changed.size();
  }

  // visibility of every object through resultObjectIsVisible, packed like BitArray
  // (bit i -> u32 word i/32, bit i%32; unused tail bits 0; BitArray.h:215-219,298-308)
  int dpref_cull_visible_bits( void * p, int result, uint32_t * words )
  {
    Session * s = SESSION( p );
    try
    {
      size_t n = s->manager->groupGetCount( s->group );
      memset( words, 0, ( ( n + 31 ) / 32 ) * sizeof(uint32_t) );
      for ( size_t i = 0; i < n; ++i )
      {
        if ( s->manager->resultObjectIsVisible( s->results[result], s->manager->groupGetObject( s->group, i ) ) )
        {
          words[i >> 5] |= 1u << ( i & 31 );
        }
      }
      return 0;
    }
    catch ( std::exception const & e ) { s->error = e.what(); return 1; }
  }

  int dpref_cull_bounding_box( void * p, float * out6 )
  {
    Session * s = SESSION( p );
    try
    {
      dp::math::Box3f b = s->manager->getBoundingBox( s->group );
      for ( int i = 0; i < 3; ++i ) { out6[i] = b.getLower()[i]; out6[3+i] = b.getUpper()[i]; }
      return 0;
    }
    catch ( std::exception const & e ) { s->error = e.what(); return 1; }
  }

#ifdef DPREF_WITH_CUDA
  // ---------------------------------------------------------------- cuda-only extensions of the new backend
  int dpref_cull_set_device_matrices( void * p, void const * deviceMatrices, size_t count )
  {
    Session * s = SESSION( p );
    try
    {
      dp::culling::cuda::Manager * m = dynamic_cast<dp::culling::cuda::Manager *>( s->manager.get() );
      if ( !m ) { s->error = "not a cuda manager"; return 1; }
      m->groupSetDeviceMatrices( s->group, deviceMatrices, count );
      return 0;
    }
    catch ( std::exception const & e ) { s->error = e.what(); return 1; }
  }

  int dpref_cull_run_multi( void * p, int const * results, int nViews, float const * viewProjections16 )
  {
    Session * s = SESSION( p );
    try
    {
      dp::culling::cuda::Manager * m = dynamic_cast<dp::culling::cuda::Manager *>( s->manager.get() );
      if ( !m ) { s->error = "not a cuda manager"; return 1; }
      std::vector<dp::culling::ResultSharedPtr> r;
      std::vector<dp::math::Mat44f> vp;
      for ( int v = 0; v < nViews; ++v )
      {
        r.push_back( s->results[results[v]] );
        vp.push_back( toMat( viewProjections16 + 16 * v ) );
      }
      m->cullMultiView( s->group, r, vp );
      return 0;
    }
    catch ( std::exception const & e ) { s->error = e.what(); return 1; }
  }
#endif

  // ---------------------------------------------------------------- dp::transform::Tree
  void * dpref_tree_create() { return new TreeSession( new dp::transform::Tree ); }
#ifdef DPREF_WITH_CUDA
  void * dpref_tree_create_backend( int backend )
  {
    try
    {
      if ( backend == 0 ) return dpref_tree_create();
      dp::transform::cuda::Tree * t = new dp::transform::cuda::Tree( 0 );
      TreeSession * s = new TreeSession( t, backend == 1 );
      s->cudaTree = t;
      return s;
    }
    catch ( std::exception const & e ) { g_createError = e.what(); return nullptr; }
  }
  // device pointer of the world matrices (for dpref_cull_set_device_matrices); NULL for the host tree
  void const * dpref_tree_device_world( void * p )
  {
    return TREE( p )->cudaTree ? TREE( p )->cudaTree->getDeviceWorldMatrices() : nullptr;
  }
  void dpref_tree_host_mirror( void * p, int enable )
  {
    if ( TREE( p )->cudaTree ) TREE( p )->cudaTree->setHostWorldMirror( !!enable );
  }
#endif
  void   dpref_tree_destroy( void * p ) { delete TREE( p ); }

  // returns the new index or -1 (Tree.cpp:54-77)
  int64_t dpref_tree_add( void * p, uint32_t parent, float const * local16 )
  {
    try { return TREE( p )->add( parent, toMat( local16 ) ); }
    catch ( std::exception const & ) { return -1; }
  }

  // bulk variant: parents[i] refers to tree indices; outIndices[i] receives the new index
  int dpref_tree_add_many( void * p, size_t n, uint32_t const * parents, float const * locals16, uint32_t * outIndices )
  {
    try
    {
      for ( size_t i = 0; i < n; ++i )
      {
        outIndices[i] = TREE( p )->add( parents[i], toMat( locals16 + 16 * i ) );
      }
      return 0;
    }
    catch ( std::exception const & ) { return 1; }
  }

  int dpref_tree_remove( void * p, uint32_t index )
  {
    try { TREE( p )->remove( index ); return 0; }
    catch ( std::exception const & ) { return 1; }
  }

  void dpref_tree_update_local( void * p, uint32_t index, float const * local16 )
  {
    TREE( p )->tree.updateLocalMatrix( index, toMat( local16 ) );
  }

  void dpref_tree_update_locals( void * p, size_t n, uint32_t const * indices, float const * locals16 )
  {
    for ( size_t i = 0; i < n; ++i ) TREE( p )->tree.updateLocalMatrix( indices[i], toMat( locals16 + 16 * i ) );
  }

  int dpref_tree_compute( void * p )
  {
    try { TREE( p )->tree.compute( dp::math::cIdentity44f ); return 0; }
    catch ( std::exception const & e ) { g_createError = e.what(); return 1; }
  }

  size_t        dpref_tree_count( void * p ) { return TREE( p )->tree.getTransformCount(); }
  float const * dpref_tree_world( void * p ) { return TREE( p )->tree.getWorldMatrices()->getPtr(); }

  // dirty world-matrix bits published by the last compute(), as u32 words (nWords = ceil(count/32))
  size_t dpref_tree_dirty_world( void * p, uint32_t * words, size_t nWords )
  {
    dp::util::BitArray const & d = TREE( p )->recorder.m_dirty;
    size_t have = ( d.getSize() + 31 ) / 32;
    size_t n = have < nWords ? have : nWords;
    if ( n ) memcpy( words, d.getBits(), n * sizeof(uint32_t) );
    return d.getSize();
  }
}
