// dp::culling::cuda - host side of the B200 culling backend.
//
// Derives ManagerBitSet exactly like the cpu and opengl backends do (they override objectCreate,
// groupCreate, groupCreateResult and cull: dp/culling/cpu/inc/ManagerImpl.h:67-77,
// dp/culling/opengl/inc/ManagerImpl.h:41-61) and keeps the group / object bookkeeping of
// GroupBitSet / ObjectBitSet.  The result type is its own (ResultCUDA) because the changed list
// and the bitset are produced on the GPU; resultGetChanged / resultObjectIsVisible are overridden
// accordingly.  Every GPU operation goes through the C ABI of include/dpcu.h; a failing call is
// rethrown as std::runtime_error like CUDA_VERIFY does (dp/cuda/Config.h:47-57).
#include <dp/culling/cuda/inc/ManagerImpl.h>
#include <dp/util/FrameProfiler.h>

#include <algorithm>
#include <cstring>
#include <stdexcept>

namespace dp
{
  namespace culling
  {
    namespace cuda
    {
      namespace
      {
        inline void verify( int status, char const * call )
        {
          if ( status != DPCU_OK )
          {
            throw std::runtime_error( std::string( call ) + ": " + dpcuGetLastError() );
          }
        }
#define DPCU_VERIFY( call ) verify( call, #call )
      }

      /************************************************************************/
      /* GroupCUDA                                                            */
      /************************************************************************/
      GroupCUDASharedPtr GroupCUDA::create( int device )
      {
        return( std::shared_ptr<GroupCUDA>( new GroupCUDA( device ) ) );
      }

      GroupCUDA::GroupCUDA( int device )
        : m_ctx( nullptr )
        , m_deviceMatrices( nullptr )
        , m_deviceMatricesCount( 0 )
        , m_deviceMatricesBound( false )
        , m_deviceObjects( ~size_t( 0 ) )
      {
        DPCU_VERIFY( dpcuCullCreate( &m_ctx, device ) );
      }

      GroupCUDA::~GroupCUDA()
      {
        dpcuCullDestroy( m_ctx );
      }

      void GroupCUDA::setDeviceMatrices( void const * deviceMatrices, size_t count )
      {
        m_deviceMatrices = deviceMatrices;
        m_deviceMatricesCount = count;
        m_deviceMatricesBound = false;
        setBoundingBoxDirty( true );
      }

      void GroupCUDA::markObjectEdited( size_t groupIndex )
      {
        m_editedObjects.push_back( static_cast<uint32_t>( groupIndex ) );
        setBoundingBoxDirty( true );
      }

      void GroupCUDA::update()
      {
        // ---- objects.  Every add, remove (GroupBitSet.cpp:93-119: the last object moves into the freed slot) and live
        // edit since the last update marked the group indices it touched; they go to the device as ONE batch
        // (dpcuCullSetObjectCount keeps what is there, dpcuCullUpdateObjects scatters the touched objects).  Only a group
        // that was never uploaded, or one in which more than a quarter of the objects changed, is uploaded whole.
        size_t const n = getObjectCount();
        if ( m_inputChanged || !m_editedObjects.empty() )
        {
          dp::util::ProfileEntry p( "cull::updateInputBuffer" );
          std::sort( m_editedObjects.begin(), m_editedObjects.end() );
          m_editedObjects.erase( std::unique( m_editedObjects.begin(), m_editedObjects.end() ), m_editedObjects.end() );
          while ( !m_editedObjects.empty() && m_editedObjects.back() >= n )
          {
            m_editedObjects.pop_back();     // indices that a later remove took away again
          }
          bool const whole = m_deviceObjects == ~size_t( 0 ) || 4 * m_editedObjects.size() > n;
          size_t const count = whole ? n : m_editedObjects.size();
          m_stageLower.resize( 4 * count );
          m_stageExtent.resize( 4 * count );
          m_stageIndex.resize( count );
          for ( size_t k = 0; k < count; ++k )
          {
            ObjectBitSetSharedPtr const & object = getObject( whole ? k : m_editedObjects[k] );
            memcpy( &m_stageLower[4 * k], object->getLowerLeft().getPtr(), 4 * sizeof(float) );
            memcpy( &m_stageExtent[4 * k], object->getExtent().getPtr(), 4 * sizeof(float) );
            m_stageIndex[k] = static_cast<uint32_t>( object->getTransformIndex() );
          }
          if ( whole )
          {
            DPCU_VERIFY( dpcuCullSetObjects( m_ctx, m_stageLower.data(), m_stageExtent.data(), m_stageIndex.data(), n, DPCU_MEM_HOST ) );
          }
          else
          {
            DPCU_VERIFY( dpcuCullSetObjectCount( m_ctx, n ) );
            DPCU_VERIFY( dpcuCullUpdateObjects( m_ctx, m_editedObjects.data(), count, m_stageLower.data(), m_stageExtent.data(), m_stageIndex.data() ) );
          }
          m_deviceObjects = n;
          m_inputChanged = false;
          m_editedObjects.clear();
        }

        // ---- matrices
        dp::util::ProfileEntry p( "cull::updateMatrices" );
        if ( m_deviceMatrices )
        {
          if ( !m_deviceMatricesBound )
          {
            DPCU_VERIFY( dpcuCullBindMatrices( m_ctx, m_deviceMatrices, m_deviceMatricesCount ) );
            m_deviceMatricesBound = true;
          }
          m_matricesChanged = false;
        }
        else if ( m_matricesChanged )
        {
          // pointer, count or stride changed: take everything (GroupBitSet.cpp:121-135)
          DPCU_VERIFY( dpcuCullSetMatrices( m_ctx, getMatrices(), getMatricesCount(), getMatricesStride(), DPCU_MEM_HOST ) );
          m_matricesChanged = false;
        }
        else
        {
          // only the matrices flagged through groupMatrixChanged (GroupBitSet.h:140-150)
          struct Collector
          {
            Collector( std::vector<uint32_t> & indices ) : m_indices( indices ) {}
            void operator()( size_t index ) { m_indices.push_back( static_cast<uint32_t>( index ) ); }
            std::vector<uint32_t> & m_indices;
          };
          m_stageIndex.clear();
          Collector collector( m_stageIndex );
          m_dirtyMatrices.traverseBits( collector );
          if ( !m_stageIndex.empty() )
          {
            DPCU_VERIFY( dpcuCullUpdateMatrices( m_ctx, m_stageIndex.data(), m_stageIndex.size(), getMatrices(), getMatricesStride(), DPCU_MEM_HOST ) );
          }
        }
        m_dirtyMatrices.clear();
        m_obbDirty = false;   // there is no OBB cache on the device: OBBs are rebuilt from the matrices in every cull
      }

      /************************************************************************/
      /* ResultCUDA                                                           */
      /************************************************************************/
      ResultCUDASharedPtr ResultCUDA::create( GroupCUDASharedPtr const & parentGroup )
      {
        return( std::shared_ptr<ResultCUDA>( new ResultCUDA( parentGroup ) ) );
      }

      ResultCUDA::ResultCUDA( GroupCUDASharedPtr const & parentGroup )
        : m_groupParent( parentGroup )
        , m_result( nullptr )
        , m_mirrorBits( nullptr )
        , m_mirrorChanged( nullptr )
        , m_mirrorCount( nullptr )
        , m_bits( nullptr )
        , m_changedIndices( nullptr )
        , m_changedCount( nullptr )
        , m_capacity( 0 )
        , m_size( 0 )
      {
        DP_ASSERT( m_groupParent );
        DPCU_VERIFY( dpcuCullResultCreate( m_groupParent->getContext(), &m_result ) );
        m_groupParent->attach( this );     // object index remap events, like ResultBitSet.cpp:41-49
      }

      ResultCUDA::~ResultCUDA()
      {
        m_groupParent->detach( this );
        dpcuCullResultDestroy( m_result );   // waits for work in flight before the mirror goes away
        dpcuHostBufferDestroy( m_mirrorBits );
        dpcuHostBufferDestroy( m_mirrorChanged );
        dpcuHostBufferDestroy( m_mirrorCount );
      }

      void ResultCUDA::flushMovedBits()
      {
        if ( m_touchedWords.empty() )
        {
          return;
        }
        std::sort( m_touchedWords.begin(), m_touchedWords.end() );
        m_touchedWords.erase( std::unique( m_touchedWords.begin(), m_touchedWords.end() ), m_touchedWords.end() );
        m_touchedValues.resize( m_touchedWords.size() );
        for ( size_t k = 0; k < m_touchedWords.size(); ++k )
        {
          m_touchedValues[k] = m_bits[m_touchedWords[k]];
        }
        DPCU_VERIFY( dpcuCullResultUpdateWords( m_result, m_touchedWords.data(), m_touchedValues.data(), m_touchedWords.size() ) );
        DPCU_VERIFY( dpcuCullResultSynchronize( m_result ) );   // the kernel also stores into the mirror: let it finish before the host edits again
        m_touchedWords.clear();
      }

      void ResultCUDA::prepare()
      {
        flushMovedBits();
        size_t const count = m_groupParent->getObjectCount();
        if ( count <= m_capacity && m_mirrorBits )
        {
          return;
        }
        // grow by half so that a scene that is being populated does not re-pin memory on every cull
        size_t capacity = count + count / 2;
        capacity = ( ( capacity ? capacity : 1 ) + 1023 ) & ~size_t( 1023 );
        dpcuHostBuffer * bits = nullptr, * changed = nullptr, * counter = m_mirrorCount;
        void * bitsPtr = nullptr, * changedPtr = nullptr, * countPtr = m_changedCount;
        DPCU_VERIFY( dpcuHostBufferCreate( &bits, capacity / 8, DPCU_HOST_MAPPED ) );
        DPCU_VERIFY( dpcuHostBufferPointer( bits, &bitsPtr ) );
        DPCU_VERIFY( dpcuHostBufferCreate( &changed, capacity * sizeof(uint32_t), DPCU_HOST_MAPPED ) );
        DPCU_VERIFY( dpcuHostBufferPointer( changed, &changedPtr ) );
        if ( !counter )
        {
          DPCU_VERIFY( dpcuHostBufferCreate( &counter, sizeof(uint32_t), DPCU_HOST_MAPPED ) );
          DPCU_VERIFY( dpcuHostBufferPointer( counter, &countPtr ) );
          *static_cast<uint32_t *>( countPtr ) = 0;
        }
        // the setter waits for work in flight; the old words stay readable for isVisible until the next cull
        DPCU_VERIFY( dpcuCullResultSetHostMirror( m_result, static_cast<uint32_t *>( bitsPtr ), capacity / 32
                                                , static_cast<uint32_t *>( changedPtr ), capacity, static_cast<uint32_t *>( countPtr ) ) );
        if ( m_bits )
        {
          memcpy( bitsPtr, m_bits, ( ( m_size + 31 ) / 32 ) * sizeof(uint32_t) );
        }
        dpcuHostBufferDestroy( m_mirrorBits );
        dpcuHostBufferDestroy( m_mirrorChanged );
        m_mirrorBits = bits;       m_bits = static_cast<uint32_t *>( bitsPtr );
        m_mirrorChanged = changed; m_changedIndices = static_cast<uint32_t *>( changedPtr );
        m_mirrorCount = counter;   m_changedCount = static_cast<uint32_t *>( countPtr );
        m_capacity = capacity;
      }

      void ResultCUDA::fetch()
      {
        dp::util::ProfileEntry p( "ResultBitSet::updateChanged" );   // same profiler key as the host diff it replaces

        // one wait, no copies: the cull kernel stored the visibility words and the compaction kernel the changed
        // list into the pinned mirror while they ran
        DPCU_VERIFY( dpcuCullResultSynchronize( m_result ) );
        size_t const changed = *m_changedCount;
        m_changedObjects.clear();
        m_changedObjects.reserve( changed );
        for ( size_t i = 0; i < changed; ++i )
        {
          m_changedObjects.push_back( m_groupParent->getObject( m_changedIndices[i] ) );   // ascending group index
        }
        m_size = m_groupParent->getObjectCount();
      }

      bool ResultCUDA::isVisible( ObjectBitSetSharedPtr const & object ) const
      {
        // ResultBitSet::isVisible, dp/culling/ResultBitSet.h:69-76
        size_t groupIndex = object->getGroupIndex();
        DP_ASSERT( groupIndex != ~0 );
        return ( groupIndex < m_size ) ? !!( m_bits[groupIndex >> 5] & ( 1u << ( groupIndex & 31 ) ) ) : true;
      }

      void ResultCUDA::onNotify( dp::util::Event const & event, dp::util::Payload * /*payload*/ )
      {
        // ResultBitSet::onNotify, dp/culling/src/ResultBitSet.cpp:110-128
        GroupBitSet::Event const & groupEvent = static_cast<GroupBitSet::Event const &>( event );
        size_t const oldIndex = groupEvent.getOldIndex(), newIndex = groupEvent.getNewIndex();
        if ( m_bits && m_size )
        {
          // the host mirror holds the bits of the last cull: move the bit there (same rule as the device kernel) and
          // remember the word; all words a frame's removals touched go back to the device in one batch before the next cull
          if ( newIndex < m_size )
          {
            uint32_t const value = ( oldIndex < m_size ) ? ( ( m_bits[oldIndex >> 5] >> ( oldIndex & 31 ) ) & 1u ) : 1u;
            uint32_t & word = m_bits[newIndex >> 5];
            word = value ? ( word | ( 1u << ( newIndex & 31 ) ) ) : ( word & ~( 1u << ( newIndex & 31 ) ) );
            m_touchedWords.push_back( static_cast<uint32_t>( newIndex >> 5 ) );
          }
        }
        else
        {
          DPCU_VERIFY( dpcuCullResultMoveBit( m_result, oldIndex, newIndex ) );
          DPCU_VERIFY( dpcuCullResultSynchronize( m_result ) );
        }
      }

      void ResultCUDA::onDestroyed( dp::util::Subject const & /*subject*/, dp::util::Payload * /*payload*/ )
      {
        // same contract as ResultBitSet::onDestroyed (ResultBitSet.cpp:130-133): the group must outlive its results
        throw std::runtime_error( "The method or operation is not implemented." );
      }

      /************************************************************************/
      /* ManagerImpl                                                          */
      /************************************************************************/
      Manager* Manager::create( int device )
      {
        return new ManagerImpl( device );
      }

      ManagerImpl::ManagerImpl( int device )
        : m_device( device )
      {
        int count = 0;
        DPCU_VERIFY( dpcuDeviceCount( &count ) );   // throws when there is no GPU: no silent CPU path
        if ( device < 0 || device >= count )
        {
          throw std::runtime_error( "dp::culling::cuda::Manager: device index out of range" );
        }
      }

      ManagerImpl::~ManagerImpl()
      {
      }

      ObjectSharedPtr ManagerImpl::objectCreate( PayloadSharedPtr const & userData )
      {
        return ObjectBitSet::create( userData );
      }

      void ManagerImpl::groupAddObject( GroupSharedPtr const & group, ObjectSharedPtr const & object )
      {
        ManagerBitSet::groupAddObject( group, object );
        // The reference never connects an object to its group (hard part 5); this
        // backend does, so that edits of live objects reach the device mirror.
        ObjectBitSetSharedPtr objectImpl = std::static_pointer_cast<ObjectBitSet>( object );
        objectImpl->setGroup( std::static_pointer_cast<GroupBitSet>( group ) );
        std::static_pointer_cast<GroupCUDA>( group )->markObjectEdited( objectImpl->getGroupIndex() );   // the new slot
      }

      void ManagerImpl::groupRemoveObject( GroupSharedPtr const & group, ObjectSharedPtr const & object )
      {
        size_t const freed = std::static_pointer_cast<ObjectBitSet>( object )->getGroupIndex();
        ManagerBitSet::groupRemoveObject( group, object );      // throws if the object is not in this group
        // the last object now lives in the freed slot (GroupBitSet.cpp:98-105): that slot is what changed on the device
        std::static_pointer_cast<GroupCUDA>( group )->markObjectEdited( freed );
      }

      void ManagerImpl::objectSetBoundingBox( ObjectSharedPtr const & object, dp::math::Box3f const & boundingBox )
      {
        ManagerBitSet::objectSetBoundingBox( object, boundingBox );
        ObjectBitSetSharedPtr objectImpl = std::static_pointer_cast<ObjectBitSet>( object );
        if ( GroupBitSetSharedPtr group = objectImpl->getGroup() )
        {
          std::static_pointer_cast<GroupCUDA>( group )->markObjectEdited( objectImpl->getGroupIndex() );
        }
      }

      void ManagerImpl::objectSetTransformIndex( ObjectSharedPtr const & object, size_t index )
      {
        ManagerBitSet::objectSetTransformIndex( object, index );
        ObjectBitSetSharedPtr objectImpl = std::static_pointer_cast<ObjectBitSet>( object );
        if ( GroupBitSetSharedPtr group = objectImpl->getGroup() )
        {
          std::static_pointer_cast<GroupCUDA>( group )->markObjectEdited( objectImpl->getGroupIndex() );
        }
      }

      GroupSharedPtr ManagerImpl::groupCreate()
      {
        return GroupCUDA::create( m_device );
      }

      ResultSharedPtr ManagerImpl::groupCreateResult( GroupSharedPtr const & group )
      {
        return ResultCUDA::create( std::static_pointer_cast<GroupCUDA>( group ) );
      }

      void ManagerImpl::groupSetDeviceMatrices( GroupSharedPtr const & group, void const * deviceMatrices, size_t numberOfMatrices )
      {
        std::static_pointer_cast<GroupCUDA>( group )->setDeviceMatrices( deviceMatrices, numberOfMatrices );
      }

      std::vector<ObjectSharedPtr> const & ManagerImpl::resultGetChanged( ResultSharedPtr const & result )
      {
        return( std::static_pointer_cast<ResultCUDA>( result )->getChangedObjects() );
      }

      bool ManagerImpl::resultObjectIsVisible( ResultSharedPtr const & result, ObjectSharedPtr const & object )
      {
        return( std::static_pointer_cast<ResultCUDA>( result )->isVisible( std::static_pointer_cast<ObjectBitSet>( object ) ) );
      }

      void ManagerImpl::cull( GroupSharedPtr const & group, ResultSharedPtr const & result, dp::math::Mat44f const & viewProjection )
      {
        dp::util::ProfileEntry p( "cull" );

        GroupCUDASharedPtr groupImpl = std::static_pointer_cast<GroupCUDA>( group );
        ResultCUDASharedPtr resultImpl = std::static_pointer_cast<ResultCUDA>( result );
        if ( resultImpl->getGroup() != groupImpl )
        {
          throw std::runtime_error( "result does not belong to this group" );
        }

        groupImpl->update();
        resultImpl->prepare();
        dpcuCullResult * handle = resultImpl->getHandle();
        DPCU_VERIFY( dpcuCullRun( groupImpl->getContext(), &handle, viewProjection.getPtr(), 1, nullptr ) );
        resultImpl->fetch();   // synchronous like the reference: the result is valid when cull returns
      }

      void ManagerImpl::cullMultiView( GroupSharedPtr const & group, std::vector<ResultSharedPtr> const & results
                                     , std::vector<dp::math::Mat44f> const & viewProjections )
      {
        dp::util::ProfileEntry p( "cull" );

        if ( results.size() != viewProjections.size() || results.empty() || results.size() > DPCU_MAX_VIEWS )
        {
          throw std::runtime_error( "cullMultiView: need 1..8 results and as many view-projection matrices" );
        }
        GroupCUDASharedPtr groupImpl = std::static_pointer_cast<GroupCUDA>( group );
        std::vector<dpcuCullResult *> handles( results.size() );
        std::vector<float> matrices( 16 * results.size() );
        for ( size_t v = 0; v < results.size(); ++v )
        {
          handles[v] = std::static_pointer_cast<ResultCUDA>( results[v] )->getHandle();
          memcpy( &matrices[16 * v], viewProjections[v].getPtr(), 16 * sizeof(float) );
        }
        groupImpl->update();
        for ( size_t v = 0; v < results.size(); ++v )
        {
          std::static_pointer_cast<ResultCUDA>( results[v] )->prepare();
        }
        DPCU_VERIFY( dpcuCullRun( groupImpl->getContext(), handles.data(), matrices.data(), static_cast<int>( results.size() ), nullptr ) );
        for ( size_t v = 0; v < results.size(); ++v )
        {
          std::static_pointer_cast<ResultCUDA>( results[v] )->fetch();
        }
      }

      dp::math::Box3f ManagerImpl::calculateBoundingBox( GroupSharedPtr const & group ) const
      {
        GroupCUDASharedPtr groupImpl = std::static_pointer_cast<GroupCUDA>( group );
        groupImpl->update();
        float box[6];
        DPCU_VERIFY( dpcuCullGetBoundingBox( groupImpl->getContext(), box ) );
        dp::math::Box3f result;
        // the C ABI already applied the Box3f( lower, upper ) construction of ManagerBitSet.cpp:304
        result.update( dp::math::Vec3f( box[0], box[1], box[2] ) );
        result.update( dp::math::Vec3f( box[3], box[4], box[5] ) );
        return result;
      }

    } // namespace cuda
  } // namespace culling
} // namespace dp
