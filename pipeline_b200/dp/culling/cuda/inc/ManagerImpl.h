// Private header of the CUDA culling backend (counterpart of dp/culling/cpu/inc/ManagerImpl.h:37-80
// and dp/culling/opengl/inc/{ManagerImpl,GroupImpl}.h).
#pragma once

#include <dp/culling/cuda/Manager.h>
#include <dp/culling/GroupBitSet.h>
#include <dp/culling/ObjectBitSet.h>
#include <dp/util/Observer.h>

#include <dpcu.h>

#include <cstdint>
#include <vector>

namespace dp
{
  namespace culling
  {
    namespace cuda
    {
      DEFINE_PTR_TYPES( GroupCUDA );
      DEFINE_PTR_TYPES( ResultCUDA );

      /** \brief Device mirror of a group: objects as SoA float4 streams + world matrices (dpcuCull). **/
      class GroupCUDA : public GroupBitSet
      {
      public:
        static GroupCUDASharedPtr create( int device );
        virtual ~GroupCUDA();

        /** \brief Bring the device mirror up to date: objects on m_inputChanged / edited objects,
                   matrices on m_matricesChanged or per dirty bit (protocol of
                   dp/culling/opengl/src/GroupImpl.cpp:75-183). **/
        void update();

        void setDeviceMatrices( void const * deviceMatrices, size_t count );
        void markObjectEdited( size_t groupIndex );

        dpcuCull * getContext() const { return m_ctx; }

      protected:
        GroupCUDA( int device );

      private:
        dpcuCull *            m_ctx;
        void const *          m_deviceMatrices;       // non-null: borrowed device array is bound
        size_t                m_deviceMatricesCount;
        bool                  m_deviceMatricesBound;
        size_t                m_deviceObjects;        // objects on the device (~0: never uploaded)
        std::vector<uint32_t> m_editedObjects;        // group indices touched by adds / removes / live edits since the last update
        std::vector<float>    m_stageLower, m_stageExtent;
        std::vector<uint32_t> m_stageIndex;
      };

      /** \brief Device mirror of a result: visibility bits of the last cull + ordered changed list
                 (dpcuCullResult), with a host copy for resultObjectIsVisible. **/
      class ResultCUDA : public Result, public dp::util::Observer
      {
      public:
        static ResultCUDASharedPtr create( GroupCUDASharedPtr const & parentGroup );
        virtual ~ResultCUDA();

        /** \brief Size the pinned host mirror for the group's current object count; call before the cull is queued. **/
        void prepare();
        /** \brief Bit moves of removed objects were applied on the host mirror: write the touched words back to the device. **/
        void flushMovedBits();
        /** \brief Wait for the cull that was just queued; its bitset and changed list are in the host mirror then. **/
        void fetch();

        std::vector<ObjectSharedPtr> const & getChangedObjects() const { return m_changedObjects; }
        bool isVisible( ObjectBitSetSharedPtr const & object ) const;

        dpcuCullResult * getHandle() const { return m_result; }
        GroupCUDASharedPtr const & getGroup() const { return m_groupParent; }

        virtual void onNotify( dp::util::Event const & event, dp::util::Payload * payload );
        virtual void onDestroyed( dp::util::Subject const & subject, dp::util::Payload * payload );

      protected:
        ResultCUDA( GroupCUDASharedPtr const & parentGroup );

      private:
        GroupCUDASharedPtr           m_groupParent;
        dpcuCullResult *             m_result;
        std::vector<ObjectSharedPtr> m_changedObjects;
        // pinned host mirror of the device result (dpcuCullResultSetHostMirror): the kernels write it over PCIe
        dpcuHostBuffer *             m_mirrorBits;
        dpcuHostBuffer *             m_mirrorChanged;
        dpcuHostBuffer *             m_mirrorCount;
        uint32_t *                   m_bits;            // visibility words, m_size objects valid
        uint32_t *                   m_changedIndices;  // ascending group indices
        uint32_t *                   m_changedCount;
        size_t                       m_capacity;        // objects the mirror has room for
        size_t                       m_size;
        std::vector<uint32_t>        m_touchedWords;    // words of m_bits edited by onNotify since the last flush
        std::vector<uint32_t>        m_touchedValues;
      };

      class ManagerImpl : public Manager
      {
      public:
        ManagerImpl( int device );
        virtual ~ManagerImpl();

        virtual ObjectSharedPtr objectCreate( PayloadSharedPtr const & userData );
        virtual void objectSetBoundingBox( ObjectSharedPtr const & object, dp::math::Box3f const & boundingBox );
        virtual void objectSetTransformIndex( ObjectSharedPtr const & object, size_t index );

        virtual GroupSharedPtr groupCreate();
        virtual void groupAddObject( GroupSharedPtr const & group, ObjectSharedPtr const & object );
        virtual void groupRemoveObject( GroupSharedPtr const & group, ObjectSharedPtr const & object );
        virtual ResultSharedPtr groupCreateResult( GroupSharedPtr const & group );
        virtual void groupSetDeviceMatrices( GroupSharedPtr const & group, void const * deviceMatrices, size_t numberOfMatrices );

        virtual std::vector<ObjectSharedPtr> const & resultGetChanged( ResultSharedPtr const & result );
        virtual bool resultObjectIsVisible( ResultSharedPtr const & result, ObjectSharedPtr const & object );

        virtual void cull( GroupSharedPtr const & group, ResultSharedPtr const & result, dp::math::Mat44f const & viewProjection );
        virtual void cullMultiView( GroupSharedPtr const & group, std::vector<ResultSharedPtr> const & results
                                  , std::vector<dp::math::Mat44f> const & viewProjections );

        virtual dp::math::Box3f calculateBoundingBox( GroupSharedPtr const & group ) const;

      private:
        int m_device;
      };

    } // namespace cuda
  } // namespace culling
} // namespace dp
